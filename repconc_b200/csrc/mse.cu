// mse.cu -- decode (gather), its deterministic scatter-add backward, and the quantisation (MSE)
// loss + surrogate with closed-form gradients.
// Reference: modeling_repconc.py:168-184 (decode), finetune_repconc.py:367-374,389-396 (loss).
//
// HBM traffic per embedding row: forward reads x, g (and q, or 8*M bytes of codes when the decode
// is fused) once = 8..12*D bytes; backward reads the same and writes grad_x (+grad_q) = 4..8*D.
// The scatter-add into the (M,K,ds) centroid gradient is privatised in shared memory per
// (sub-vector, row-chunk) CTA -- every accumulator has exactly one owner thread that adds its rows
// in ascending order, so the result is deterministic (index_put_(accumulate=True) on CUDA is not).
#include "common.cuh"

namespace rc {



__device__ __forceinline__ int64_t load_code(const int64_t* codes, int64_t sb, int64_t sm, const uint8_t* u8,
                                             int64_t b, int m, int M) {
    return codes ? codes[b * sb + m * sm] : (int64_t)u8[b * M + m];
}

// ---------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_kernel(const int64_t* __restrict__ codes, int64_t sb, int64_t sm, const uint8_t* __restrict__ u8,
              const float* __restrict__ c, int64_t B, int M, int K, int ds, float* __restrict__ out,
              int32_t* __restrict__ flags) {
    const int64_t D = (int64_t)M * ds;
    const int64_t total = B * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        const int d = (int)(i - b * D);
        const int m = d / ds, j = d - m * ds;
        int64_t code = load_code(codes, sb, sm, u8, b, m, M);
        if (code < 0 || code >= K) {
            if (flags) atomicOr(flags, RC_FLAG_BADCODE);
            code = code < 0 ? 0 : K - 1;
        }
        out[i] = __ldg(c + ((int64_t)m * K + code) * ds + j);
    }
}

// ---------------------------------------------------------------------------------------------
// scatter-add backward: grad_c[m,k,:] = sum_{b: codes[b,m]==k} v[b,m,:]
//   SRC_GRADQ : v = grad_q[b, m*ds + j]
//   SRC_MSE   : v = gs*g - coef*(x - q),  q from the array or decoded from the centroids
// grid (M, nchunk); CTA stages its chunk's codes, thread (j, kr) owns accumulators (k in range kr, j)
// ---------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ROWS = 512;  // rows per chunk

struct ScatterSrc {
    const float* gq; int64_t ldgq;       // SRC_GRADQ
    const float* x; int64_t ldx;         // SRC_MSE
    const float* q; int64_t ldq;
    const float* g; int64_t ldg;
    const float* c;
    float coef, gs;
};

template <bool FROM_GRADQ>
__global__ void __launch_bounds__(SC_THREADS)
scatter_kernel(const int64_t* __restrict__ codes, int64_t sb, int64_t sm, ScatterSrc src, int64_t B, int M, int K,
               int ds, int64_t rows_per_chunk, float* __restrict__ dst /* (nchunk, M, K, ds) */) {
    extern __shared__ __align__(16) float acc_s[];  // K*ds floats, then rows_per_chunk int32 codes
    int* code_s = reinterpret_cast<int*>(acc_s + (size_t)K * ds);
    const int m = blockIdx.x;
    const int chunk = blockIdx.y;
    const int64_t b0 = (int64_t)chunk * rows_per_chunk;
    const int rows = (int)min(rows_per_chunk, B - b0);
    for (int i = threadIdx.x; i < K * ds; i += SC_THREADS) acc_s[i] = 0.0f;
    for (int r = threadIdx.x; r < rows; r += SC_THREADS) {
        int64_t code = codes[(b0 + r) * sb + m * sm];
        code_s[r] = (code < 0 || code >= K) ? -1 : (int)code;
    }
    __syncthreads();
    const int dsT = ds < SC_THREADS ? ds : SC_THREADS;   // threads across j
    const int nkr = SC_THREADS / dsT;                    // k ranges
    const int tj = threadIdx.x % dsT, kr = threadIdx.x / dsT;
    const int kspan = (K + nkr - 1) / nkr;
    if (kr < nkr) {
        const int klo = kr * kspan, khi = min(K, klo + kspan);
        for (int r = 0; r < rows; ++r) {
            const int code = code_s[r];
            if (code < klo || code >= khi) continue;
            const int64_t b = b0 + r;
            for (int j = tj; j < ds; j += dsT) {
                const int64_t col = (int64_t)m * ds + j;
                float v;
                if (FROM_GRADQ) {
                    v = src.gq[b * src.ldgq + col];
                } else {
                    const float xv = src.x[b * src.ldx + col];
                    const float qv = src.q ? src.q[b * src.ldq + col]
                                           : __ldg(src.c + ((int64_t)m * K + code) * ds + j);
                    const float gv = src.g ? src.g[b * src.ldg + col] : 0.0f;
                    v = fmaf(-src.coef, xv - qv, src.gs * gv);
                }
                acc_s[code * ds + j] += v;
            }
        }
    }
    __syncthreads();
    float* out = dst + ((int64_t)chunk * M + m) * K * ds;
    for (int i = threadIdx.x; i < K * ds; i += SC_THREADS) out[i] = acc_s[i];
}

// sum the chunk partials in chunk order
__global__ void __launch_bounds__(256)
scatter_reduce_kernel(const float* __restrict__ part, int64_t n, int nchunk, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = part[i];
    for (int c = 1; c < nchunk; ++c) s += part[(int64_t)c * n + i];
    out[i] = s;
}

static int scatter_chunks(int64_t B, int64_t* rows_per_chunk) {
    int64_t rpc = SC_ROWS;
    int64_t n = (B + rpc - 1) / rpc;
    if (n > 64) {  // bound the partial buffer; chunks grow instead
        n = 64;
        rpc = (B + n - 1) / n;
        n = (B + rpc - 1) / rpc;
    }
    if (n < 1) n = 1;
    *rows_per_chunk = rpc;
    return (int)n;
}

static size_t scatter_ws_bytes(int64_t B, int M, int K, int ds) {
    int64_t rpc;
    const int n = scatter_chunks(B, &rpc);
    return n > 1 ? align_up((size_t)n * M * K * ds * 4, 256) : 0;
}

template <bool FROM_GRADQ>
static int launch_scatter(const int64_t* codes, int64_t sb, int64_t sm, const ScatterSrc& src, int64_t B, int M,
                          int K, int ds, float* grad_c, void* ws, cudaStream_t st) {
    int64_t rpc;
    const int nchunk = scatter_chunks(B, &rpc);
    const size_t smem = (size_t)K * ds * 4 + (size_t)rpc * 4;
    if (smem > 227 * 1024) {
        set_error("scatter-add: K*ds=%d too large for shared-memory privatisation", K * ds);
        return RC_E_UNSUPPORTED;
    }
    auto kern = scatter_kernel<FROM_GRADQ>;
    if (smem > 48 * 1024) RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float* dst = nchunk > 1 ? (float*)ws : grad_c;
    if (nchunk > 1 && !ws) {
        set_error("scatter-add: workspace required for B=%lld", (long long)B);
        return RC_E_WORKSPACE;
    }
    dim3 grid((unsigned)M, (unsigned)nchunk);
    kern<<<grid, SC_THREADS, smem, st>>>(codes, sb, sm, src, B, M, K, ds, rpc, dst);
    RC_CHECK_LAUNCH("scatter_kernel");
    if (nchunk > 1) {
        const int64_t n = (int64_t)M * K * ds;
        scatter_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, n, nchunk, grad_c);
        RC_CHECK_LAUNCH("scatter_reduce_kernel");
    }
    return RC_OK;
}

// ---------------------------------------------------------------------------------------------
// loss forward: fp64 block partials -> deterministic final sum
// ---------------------------------------------------------------------------------------------
constexpr int MSE_THREADS = 256;

__global__ void __launch_bounds__(MSE_THREADS)
mse_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ q, int64_t ldq,
               const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ codes, int64_t sb, int64_t sm,
               const float* __restrict__ c, int64_t n, int M, int K, int ds, double* __restrict__ partial) {
    __shared__ double red[2][MSE_THREADS / 32];
    const int64_t D = (int64_t)M * ds;
    const int64_t total = n * D;
    double se = 0.0, su = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        const int d = (int)(i - b * D);
        const float xv = x[b * ldx + d];
        float qv;
        if (q) qv = q[b * ldq + d];
        else {
            const int m = d / ds, j = d - m * ds;
            int64_t code = codes[b * sb + m * sm];
            code = code < 0 ? 0 : (code >= K ? K - 1 : code);
            qv = __ldg(c + ((int64_t)m * K + code) * ds + j);
        }
        const float diff = qv - xv;
        se += (double)diff * (double)diff;
        if (g) {
            const float gv = g[b * ldg + d];
            su += (double)gv * (double)xv + (double)gv * (double)qv;
        }
    }
    se = warp_sum(se);
    su = warp_sum(su);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = se; red[1][warp] = su; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MSE_THREADS / 32; ++w) { se += red[0][w]; su += red[1][w]; }
        partial[2 * blockIdx.x] = se;
        partial[2 * blockIdx.x + 1] = su;
    }
}

__global__ void mse_final_kernel(const double* __restrict__ partial, int nblocks, double scale, float* __restrict__ out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double se = 0.0, su = 0.0;
        for (int i = 0; i < nblocks; ++i) { se += partial[2 * i]; su += partial[2 * i + 1]; }
        out2[0] = (float)(se * scale);
        out2[1] = (float)su;
    }
}

__global__ void __launch_bounds__(MSE_THREADS)
mse_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ q, int64_t ldq,
               const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ codes, int64_t sb, int64_t sm,
               const float* __restrict__ c, int64_t n, int M, int K, int ds, float coef, float gs,
               float* __restrict__ grad_x, float* __restrict__ grad_q) {
    const int64_t D = (int64_t)M * ds;
    const int64_t total = n * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        const int d = (int)(i - b * D);
        const float xv = x[b * ldx + d];
        float qv;
        if (q) qv = q[b * ldq + d];
        else {
            const int m = d / ds, j = d - m * ds;
            int64_t code = codes[b * sb + m * sm];
            code = code < 0 ? 0 : (code >= K ? K - 1 : code);
            qv = __ldg(c + ((int64_t)m * K + code) * ds + j);
        }
        const float gv = g ? gs * g[b * ldg + d] : 0.0f;
        const float e = coef * (xv - qv);
        if (grad_x) grad_x[i] = gv + e;
        if (grad_q) grad_q[i] = gv - e;
    }
}

static int mse_blocks(int64_t total) {
    int64_t nb = (total + MSE_THREADS * 4 - 1) / (MSE_THREADS * 4);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

}  // namespace rc

using namespace rc;

RC_API int rc_decode(const int64_t* codes, int64_t stride_b, int64_t stride_m, const uint8_t* codes_u8,
                     const float* centroids, int64_t B, int M, int K, int ds, float* out, int32_t* flags,
                     void* stream) {
    RC_REQUIRE((codes != nullptr) != (codes_u8 != nullptr), "rc_decode: pass exactly one of codes / codes_u8");
    RC_REQUIRE(centroids && out, "rc_decode: null pointer");
    RC_REQUIRE(B >= 0 && M >= 1 && K >= 1 && ds >= 1, "rc_decode: bad shape");
    if (B == 0) return RC_OK;
    const int64_t total = B * M * ds;
    int64_t nb = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (nb > cap) nb = cap;
    decode_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(codes, stride_b, stride_m, codes_u8, centroids, B,
                                                                  M, K, ds, out, flags);
    RC_CHECK_LAUNCH("decode_kernel");
    return RC_OK;
}

// counts[m,k] = #{b : codes[b,m] == k}: the code histogram of eval_balance (finetune_repconc.py:604-611) and the
// cluster sizes of the k-means warm-up.  Integer atomics on a shared-memory histogram -> deterministic.
__global__ void __launch_bounds__(256)
code_histogram_kernel(const int64_t* __restrict__ codes, int64_t sb, int64_t sm, const uint8_t* __restrict__ codes_u8,
                      int64_t B, int M, int K, int64_t rows_per_block, int32_t* __restrict__ counts,
                      int32_t* __restrict__ flags) {
    extern __shared__ int32_t hist[];
    const int m = blockIdx.y;
    for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.x * rows_per_block, b1 = min(B, b0 + rows_per_block);
    bool bad = false;
    for (int64_t b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        const int64_t k = codes_u8 ? (int64_t)codes_u8[b * M + m] : codes[b * sb + (int64_t)m * sm];
        if (k < 0 || k >= K) bad = true;
        else atomicAdd(&hist[k], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (hist[k]) atomicAdd(&counts[(int64_t)m * K + k], hist[k]);
    if (bad) atomicOr(flags, RC_FLAG_BADCODE);
}

RC_API int rc_code_histogram(const int64_t* codes, int64_t stride_b, int64_t stride_m, const uint8_t* codes_u8,
                             int64_t B, int M, int K, int32_t* counts, int32_t* flags, void* stream) {
    RC_REQUIRE((codes != nullptr) != (codes_u8 != nullptr), "rc_code_histogram: pass exactly one of codes / codes_u8");
    RC_REQUIRE(counts && flags && B >= 0 && M >= 1 && M <= 65535 && K >= 1 && K <= 8192,
               "rc_code_histogram: bad argument B=%lld M=%d K=%d", (long long)B, M, K);
    cudaStream_t st = (cudaStream_t)stream;
    RC_CUDA(cudaMemsetAsync(counts, 0, (size_t)M * K * sizeof(int32_t), st));
    if (B == 0) return RC_OK;
    const int64_t rows_per_block = 4096;
    dim3 grid((unsigned)((B + rows_per_block - 1) / rows_per_block), (unsigned)M);
    code_histogram_kernel<<<grid, 256, (size_t)K * sizeof(int32_t), st>>>(codes, stride_b, stride_m, codes_u8, B, M, K,
                                                                        rows_per_block, counts, flags);
    RC_CHECK_LAUNCH("code_histogram_kernel");
    return RC_OK;
}

RC_API size_t rc_decode_bwd_workspace_bytes(int64_t B, int M, int K, int ds) {
    if (B < 1) return 0;
    return scatter_ws_bytes(B, M, K, ds);
}

RC_API int rc_decode_bwd(const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* grad_q, int64_t ldg,
                         int64_t B, int M, int K, int ds, float* grad_c, void* workspace, void* stream) {
    RC_REQUIRE(codes && grad_q && grad_c, "rc_decode_bwd: null pointer");
    RC_REQUIRE(B >= 0 && M >= 1 && M <= 65535 && K >= 1 && ds >= 1, "rc_decode_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        RC_CUDA(cudaMemsetAsync(grad_c, 0, (size_t)M * K * ds * 4, st));
        return RC_OK;
    }
    ScatterSrc src{};
    src.gq = grad_q;
    src.ldgq = ldg;
    return launch_scatter<true>(codes, stride_b, stride_m, src, B, M, K, ds, grad_c, workspace, st);
}

RC_API size_t rc_mse_workspace_bytes(int64_t n, int M, int K, int ds) {
    if (n < 1) return 256;
    const size_t fwd = align_up((size_t)mse_blocks(n * M * ds) * 16, 256);
    const size_t bwd = scatter_ws_bytes(n, M, K, ds);
    return (fwd > bwd ? fwd : bwd) + 256;
}

RC_API int rc_mse_fwd(const float* x, int64_t ldx, const float* q, int64_t ldq, const float* g, int64_t ldg,
                      const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* centroids, int64_t n,
                      int M, int K, int ds, float w, float* out2, void* workspace, void* stream) {
    RC_REQUIRE(x && out2 && workspace, "rc_mse_fwd: null pointer");
    RC_REQUIRE(q || (codes && centroids), "rc_mse_fwd: need q or (codes, centroids)");
    RC_REQUIRE(n >= 1 && M >= 1 && K >= 1 && ds >= 1, "rc_mse_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = mse_blocks(n * M * ds);
    mse_fwd_kernel<<<nb, MSE_THREADS, 0, st>>>(x, ldx, q, ldq, g, ldg, codes, stride_b, stride_m, centroids, n, M,
                                               K, ds, (double*)workspace);
    RC_CHECK_LAUNCH("mse_fwd_kernel");
    mse_final_kernel<<<1, 32, 0, st>>>((const double*)workspace, nb, (double)w / (double)n, out2);
    RC_CHECK_LAUNCH("mse_final_kernel");
    return RC_OK;
}

RC_API int rc_mse_bwd(const float* x, int64_t ldx, const float* q, int64_t ldq, const float* g, int64_t ldg,
                      const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* centroids, int64_t n,
                      int M, int K, int ds, float w, float gm, float gs, float* grad_x, float* grad_q,
                      float* grad_c, void* workspace, void* stream) {
    RC_REQUIRE(x, "rc_mse_bwd: null x");
    RC_REQUIRE(q || (codes && centroids), "rc_mse_bwd: need q or (codes, centroids)");
    RC_REQUIRE(!grad_c || codes, "rc_mse_bwd: grad_c needs codes");
    RC_REQUIRE(n >= 1 && M >= 1 && M <= 65535 && K >= 1 && ds >= 1, "rc_mse_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const float coef = (float)(2.0 * (double)w * (double)gm / (double)n);
    if (grad_x || grad_q) {
        const int nb = mse_blocks(n * M * ds);
        mse_bwd_kernel<<<nb, MSE_THREADS, 0, st>>>(x, ldx, q, ldq, g, ldg, codes, stride_b, stride_m, centroids, n,
                                                   M, K, ds, coef, gs, grad_x, grad_q);
        RC_CHECK_LAUNCH("mse_bwd_kernel");
    }
    if (grad_c) {
        ScatterSrc src{};
        src.x = x; src.ldx = ldx;
        src.q = q; src.ldq = ldq;
        src.g = g; src.ldg = ldg;
        src.c = centroids;
        src.coef = coef;
        src.gs = gs;
        RC_REQUIRE(q || centroids, "rc_mse_bwd: need q or centroids for grad_c");
        return launch_scatter<false>(codes, stride_b, stride_m, src, n, M, K, ds, grad_c, workspace, st);
    }
    return RC_OK;
}
