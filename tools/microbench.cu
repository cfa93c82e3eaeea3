// microbench.cu -- pipe-rate probes used to set the compute ceilings quoted in DESIGN.md
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

// random 16-byte gathers from a 192 KB shared table (the ADC LUT access pattern)
template <int VEC>
__global__ void __launch_bounds__(1024) lds_gather_kernel(float* out, int iters, uint32_t seed) {
    extern __shared__ __align__(16) float tab[];
    const int nvec = 48 * 256;
    for (int i = threadIdx.x; i < nvec * VEC; i += blockDim.x) tab[i] = i * 1e-6f;
    __syncthreads();
    uint32_t codes[12];
    uint32_t s = seed ^ (threadIdx.x * 2654435761u) ^ (blockIdx.x * 40503u);
    for (int i = 0; i < 12; ++i) { s = s * 1664525u + 1013904223u; codes[i] = s ^ (s >> 13); }
    float acc[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t code = ((codes[i] >> (8 * b)) + it) & 0xffu;
                const int m = i * 4 + b;
                if (VEC == 4) {
                    const float4 v = reinterpret_cast<const float4*>(tab)[m * 256 + code];
                    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                } else if (VEC == 2) {
                    const float2 v = reinterpret_cast<const float2*>(tab)[m * 256 + code];
                    acc[0] += v.x; acc[1] += v.y;
                } else {
                    acc[0] += tab[m * 256 + code];
                }
            }
        }
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] == 12345.678f) out[0] = acc[0];
}

// conflict-free 16-byte gathers (the layout of the ADC filter scans): the 8 lanes of a quarter-warp read the 8
// consecutive entries of one random 128-byte line, so every LDS.128 phase is one wavefront -- the LSU data-pipe ceiling
__global__ void __launch_bounds__(1024) lds_cf_kernel(float* out, int iters, uint32_t seed) {
    extern __shared__ __align__(16) float tab[];
    const int nvec = 48 * 256;
    for (int i = threadIdx.x; i < nvec * 4; i += blockDim.x) tab[i] = i * 1e-6f;
    __syncthreads();
    const int j = threadIdx.x & 7;
    uint32_t codes[12];
    uint32_t s = seed ^ ((threadIdx.x >> 3) * 2654435761u) ^ (blockIdx.x * 40503u);
    for (int i = 0; i < 12; ++i) { s = s * 1664525u + 1013904223u; codes[i] = s ^ (s >> 13); }
    float acc[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t line = (((codes[i] >> (8 * b)) + it) & 0xffu) * 6 + ((i * 4 + b) % 6);   // 1536 lines
                const float4 v = reinterpret_cast<const float4*>(tab)[line * 8 + j];
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            }
        }
    }
    if (acc[0] + acc[1] + acc[2] + acc[3] == 12345.678f) out[0] = acc[0];
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d, max clock %.0f MHz\n", sms, clk / 1e3);
    double* dout; cudaMalloc(&dout, 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        float ms = time_ms([&] { dfma_kernel<8><<<sms, warps * 32>>>(dout, iters, 1.0000001, 1e-9); });
        double ops = (double)sms * warps * 32 * 8 * iters;
        printf("DFMA  ILP8 %2d warps/SM: %.3f ms  %.2f T DFMA/s  (%.1f /clk/SM at max clock)\n", warps, ms,
               ops / ms / 1e9, ops / ms / sms / (double)clk);
    }
    for (int warps : {16, 32}) {
        float ms = time_ms([&] { ffma_kernel<8><<<sms, warps * 32>>>((float*)dout, iters, 1.0000001f, 1e-9f); });
        double ops = (double)sms * warps * 32 * 8 * iters;
        printf("FFMA  ILP8 %2d warps/SM: %.3f ms  %.2f T FFMA/s  (%.1f /clk/SM at max clock)\n", warps, ms,
               ops / ms / 1e9, ops / ms / sms / (double)clk);
    }
    cudaFuncSetAttribute(lds_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 256 * 16);
    cudaFuncSetAttribute(lds_gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 256 * 8);
    cudaFuncSetAttribute(lds_gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 256 * 4);
    const int git = 400;
    for (int threads : {256, 512, 1024}) {
        float ms4 = time_ms([&] { lds_gather_kernel<4><<<sms, threads, 48 * 256 * 16>>>((float*)dout, git, 1u); });
        float ms2 = time_ms([&] { lds_gather_kernel<2><<<sms, threads, 48 * 256 * 8>>>((float*)dout, git, 1u); });
        float ms1 = time_ms([&] { lds_gather_kernel<1><<<sms, threads, 48 * 256 * 4>>>((float*)dout, git, 1u); });
        double g = (double)sms * threads * 48 * git;
        printf("LDS gather %4d thr/SM: LDS.128 %.2f T lookups/s (x4 queries), LDS.64 %.2f (x2), LDS.32 %.2f (x1)\n",
               threads, 4 * g / ms4 / 1e9, 2 * g / ms2 / 1e9, g / ms1 / 1e9);
    }
    cudaFuncSetAttribute(lds_cf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 256 * 16);
    for (int threads : {512, 1024}) {
        float ms = time_ms([&] { lds_cf_kernel<<<sms, threads, 48 * 256 * 16>>>((float*)dout, git, 1u); });
        const double wf = (double)sms * threads * 48 * git * 16.0 / 128.0;     // 128-byte wavefronts
        printf("LDS.128 conflict-free %4d thr/SM: %.3f ms  %.3f wavefronts/clk/SM at max clock (LSU data-pipe ceiling = 1)\n",
               threads, ms, wf / ms / sms / (double)clk);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
