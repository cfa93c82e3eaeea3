// microbench_mem.cu -- HBM read bandwidth for the access patterns the Sinkhorn passes use
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// (1) plain grid-stride float4 read-reduce
__global__ void read_flat(const float4* __restrict__ p, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = p[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// (2) persistent CTAs, contiguous range per CTA, a warp reads one 1 KB row at a time (2 float4 per lane),
//     rows of a warp are 8 rows apart; UNR rows in flight per warp
template <int UNR>
__global__ void read_rows(const float* __restrict__ p, int64_t nrows, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t lo = nrows * blockIdx.x / gridDim.x, hi = nrows * (blockIdx.x + 1) / gridDim.x;
    float acc = 0.f;
    int64_t r = lo + warp;
    for (; r + (int64_t)(UNR - 1) * nw < hi; r += (int64_t)UNR * nw) {
        float4 a[UNR], b[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const float4* row = reinterpret_cast<const float4*>(p + (r + (int64_t)u * nw) * 256);
            a[u] = row[lane];
            b[u] = row[32 + lane];
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += a[u].x + a[u].w + b[u].y + b[u].z;
    }
    for (; r < hi; r += nw) {
        const float4* row = reinterpret_cast<const float4*>(p + r * 256);
        acc += row[lane].x + row[32 + lane].y;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int64_t nrows = 48 * 8192;          // the (M,B,K) table of BASELINE config 3: 403 MB
    const size_t bytes = (size_t)nrows * 1024;
    float* p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
    float* out; cudaMalloc(&out, 64);
    // a second buffer to flush L2 between variants is unnecessary: 403 MB >> 126 MB L2
    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
        float ms = time_ms([&] { read_flat<<<g, 256>>>((const float4*)p, bytes / 16, out); });
        printf("flat  grid %5d x256          : %.3f ms  %.0f GB/s\n", g, ms, bytes / ms / 1e6);
    }
    for (int cps : {2, 3, 4, 8}) {
        float m1 = time_ms([&] { read_rows<1><<<148 * cps, 256>>>(p, nrows, out); });
        float m2 = time_ms([&] { read_rows<2><<<148 * cps, 256>>>(p, nrows, out); });
        float m4 = time_ms([&] { read_rows<4><<<148 * cps, 256>>>(p, nrows, out); });
        float m8 = time_ms([&] { read_rows<8><<<148 * cps, 256>>>(p, nrows, out); });
        printf("rows  %d CTAs/SM x256, rows in flight/warp 1/2/4/8: %.0f / %.0f / %.0f / %.0f GB/s\n", cps,
               bytes / m1 / 1e6, bytes / m2 / 1e6, bytes / m4 / 1e6, bytes / m8 / 1e6);
    }
    return 0;
}
