// assign.cu -- NN assign, distance table + extrema, and the Sinkhorn uniform assignment
// (reference: src/repconc/models/repconc/modeling_repconc.py:47-85,137-165).
//
// Data layout in HBM
//   x          (B, D) fp32, D = M*ds                         caller's embeddings
//   centroids  (M, K, ds) fp32                               nn.Parameter of the module
//   table      (M, B, K) fp32, k fastest                     raw distances, centred IN PLACE by
//                                                            rc_sinkhorn_begin (never widened to fp64)
//   state      lu, lu_build, U (M,K) f64 | P (M,K) f64 (the all-reduce operand) | lv (M,B) f64 (dense pass) |
//              partial (G,S,K) f64 | drift (M,2) f64 + decision | survivor records: row-pair directory (8 B per
//              pair of rows) + pool of [32 x u16 lane header][E f64 x count] records (160 entries per row on average)
//
// Sinkhorn formulation.  The reference materialises Q = exp(-d~/eps) as (M,K,B) fp64 and divides it in place
// 4x per iteration.  Here Q_t[k,b] = 2^(a + lu[k] + lv[b]), a = -d~ * log2(e)/eps, is never stored.
//   dense pass   (any K): one pass over the fp32 table per iteration evaluates each element once, finishes the
//                column normalisation of iteration t inside a warp (a table row is one column of Q) and
//                accumulates the row sums that iteration t+1 needs: 4 B/element of HBM traffic instead of the
//                reference's ~80 B/element, and ONE exp2 per element.
//   sparse passes (K = 256, default): only the elements within 2^-72 of their column's maximum can change an
//                fp64 sum.  A selection pass finds them (fp32 filter, ballot compaction), evaluates them in
//                fp64 and emits them as per-row lists; list passes then iterate on the lists alone -- one
//                multiplication per survivor -- until lu has drifted by more than the selection slack.
//   Kernels: dist_table / nn_assign (packed fp32), sinkhorn_pass<BEGIN|STEP|FINISH> (dense), sinkhorn_step_sparse
//            (selection), sinkhorn_step_list (list), sinkhorn_reduce / sinkhorn_update / sinkhorn_reduce_update
//            (row sums, row scaling, drift, U, decision), sinkhorn_finish_sparse (fp32-filtered argmax),
//            sinkhorn_expand (Q for API parity).  Entry points: rc_sinkhorn_solve (one rank, one call) and
//            rc_sinkhorn_begin / step / finish (ranks exchange the row sums between the calls).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "sinkhorn_math.cuh"

namespace rc {

// =============================================================================================
// a2  NN assign
// =============================================================================================
constexpr int NN_THREADS = 256;

template <int DS>
__global__ void __launch_bounds__(NN_THREADS)
nn_assign_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ c, int64_t B, int M,
                 int K, int kchunk, int64_t* __restrict__ codes_mb, uint8_t* __restrict__ codes_u8) {
    extern __shared__ __align__(16) float cs[];  // kchunk * DS floats
    const int m = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * NN_THREADS + threadIdx.x;
    const bool live = b < B;
    float xr[DS];
    if (live) {
        const float* xp = x + b * ldx + (int64_t)m * DS;
#pragma unroll
        for (int j = 0; j < DS; ++j) xr[j] = __ldg(xp + j);
    } else {
#pragma unroll
        for (int j = 0; j < DS; ++j) xr[j] = 0.0f;
    }
    float best = 0.0f;
    int bi = 0;
    const float* cm = c + (int64_t)m * K * DS;
    for (int k0 = 0; k0 < K; k0 += kchunk) {
        const int kn = min(kchunk, K - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kn * DS; i += NN_THREADS) cs[i] = __ldg(cm + (int64_t)k0 * DS + i);
        __syncthreads();
        for (int kk = 0; kk < kn; ++kk) {
            const float* ck = cs + kk * DS;
            float d;
            if constexpr (DS % 8 == 0) {
                const f32x2_t* ck2 = reinterpret_cast<const f32x2_t*>(ck);
                d = sqdist_aten_order_packed<DS>([&](int p) { return pk2(xr[2 * p], xr[2 * p + 1]); },
                                                 [&](int p) { return ck2[p]; });
            } else {
                d = sum_aten_order<DS>([&](int j) { return sqdiff(xr[j], ck[j]); });
            }
            const int k = k0 + kk;
            // torch.argmin: first minimum, NaN counts as smallest
            if (k == 0 || d < best || (d != d && best == best)) {
                best = d;
                bi = k;
            }
        }
    }
    if (live) {
        if (codes_mb) codes_mb[(int64_t)m * B + b] = bi;
        if (codes_u8) codes_u8[b * M + m] = (uint8_t)bi;
    }
}

// any ds (run-time), slow: one thread per (b), reads straight from global / L1
__global__ void __launch_bounds__(NN_THREADS)
nn_assign_generic_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ c, int64_t B,
                         int M, int K, int ds, int64_t* __restrict__ codes_mb,
                         uint8_t* __restrict__ codes_u8) {
    const int m = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * NN_THREADS + threadIdx.x;
    if (b >= B) return;
    const float* xp = x + b * ldx + (int64_t)m * ds;
    const float* cm = c + (int64_t)m * K * ds;
    float best = 0.0f;
    int bi = 0;
    for (int k = 0; k < K; ++k) {
        const float* ck = cm + (int64_t)k * ds;
        const float d = sum_aten_order_rt([&](int j) { return sqdiff(__ldg(xp + j), __ldg(ck + j)); }, ds);
        if (k == 0 || d < best || (d != d && best == best)) {
            best = d;
            bi = k;
        }
    }
    if (codes_mb) codes_mb[(int64_t)m * B + b] = bi;
    if (codes_u8) codes_u8[b * M + m] = (uint8_t)bi;
}

template <int DS>
static int launch_nn(const float* x, int64_t ldx, const float* c, int64_t B, int M, int K, int64_t* mb,
                     uint8_t* u8, cudaStream_t st) {
    int kchunk = (48 * 1024) / (DS * 4);
    if (kchunk > K) kchunk = K;
    dim3 grid((unsigned)((B + NN_THREADS - 1) / NN_THREADS), (unsigned)M);
    nn_assign_kernel<DS><<<grid, NN_THREADS, (size_t)kchunk * DS * 4, st>>>(x, ldx, c, B, M, K, kchunk, mb, u8);
    RC_CHECK_LAUNCH("nn_assign_kernel");
    return RC_OK;
}

// =============================================================================================
// f3  fused corpus-encode epilogue: rotation + optional sub-vector L2 normalisation + NN assign + uint8 pack
//     (modeling_repconc.py:98-103 with use_constraint = False, evaluate_repconc.py:64-70)
// One CTA = 256 rows of the pooled encoder output x one sub-vector m.  Phase 1: y[b, j] = sum_d x[b, d] * R[m DS + j, d]
// (fp32 FMA, d ascending) as a register-tiled product (4 rows x DS/4 columns per thread), x and the DS rows of R
// staged through shared memory in chunks of 32 dimensions; the tile goes through shared memory once so that thread b
// ends with row b.  Phase 2 (COS metric): y /= max(||y||, 1e-12).  Phase 3: nn_assign_kernel's loop -- the squared
// distances in the reference's fp32 order against the 256 centroids of m broadcast from shared memory, first minimum.
// The rotated embeddings never travel to HBM unless the caller asks for them, the codes are written as the (B, M)
// uint8 rows GpuIndexPQ.add appends.  An identity rotation (the module's initial buffer) reproduces x bit for bit
// (products with 0 and 1 and sums with 0 are exact), hence the codes of rc_nn_assign exactly; for a learned rotation
// the result differs from a BLAS GEMM by summation order only (1e-6 relative), as two BLAS libraries differ.
// =============================================================================================
constexpr int EN_KC = 32;            // dimensions per staged chunk
constexpr int EN_XS = EN_KC + 4;     // row stride of the staged x chunk (floats): 16-byte aligned, rows 4 banks apart

template <int DS, int MB>
__global__ void __launch_bounds__(NN_THREADS)
encode_assign_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ rot, int D,
                     const float* __restrict__ c, int64_t B, int M, int K, int kchunk, int normalize,
                     float* __restrict__ y_out, int64_t ldy, int64_t* __restrict__ codes_mb,
                     uint8_t* __restrict__ codes_u8) {
    constexpr int NC = DS * MB;                          // output columns of the CTA (MB sub-vectors)
    static_assert(NC % 4 == 0 && NC <= 64, "4 column groups, at most 16 columns per thread");
    constexpr int CT = NC / 4;                           // columns of the thread's register tile (4 rows x CT)
    extern __shared__ __align__(16) float en_sh[];
    constexpr int STAGE_FLOATS = NN_THREADS * EN_XS + EN_KC * NC;   // one stage: x chunk [256 rows][EN_XS] + R chunk [EN_KC][NC]
    float* ys = en_sh;                                   // after the product: [NN_THREADS][NC + 1]
    constexpr size_t YS_FLOATS = (size_t)NN_THREADS * (NC + 1) > (size_t)2 * STAGE_FLOATS
                                     ? (size_t)NN_THREADS * (NC + 1) : (size_t)2 * STAGE_FLOATS;
    float* cs = en_sh + YS_FLOATS;                       // phase 3: kchunk * DS floats (behind the y tile)
    const int m0 = blockIdx.y * MB;
    const int64_t b0 = (int64_t)blockIdx.x * NN_THREADS;
    const int64_t b = b0 + threadIdx.x;
    const bool live = b < B;
    // phase 1: register-tiled product.  Thread (tr, tc) owns rows tr + 64 i (i < 4) x columns CT tc + cc: per 4
    // dimensions it reads 4 + CT 16-byte words for 16 CT FMAs; the 8 row addresses of a warp are 36 floats apart
    // (conflict-free), the column words are broadcast.  MB sub-vectors share one pass over x.  The chunks of 32
    // dimensions arrive through a two-stage cp.async pipeline (x: 16-byte copies, row-major; R: 4-byte copies,
    // transposed to dimension-major on the way in): chunk t + 1 is in flight while chunk t is multiplied.
    const int tr = threadIdx.x >> 2, tc = threadIdx.x & 3;
    constexpr bool PACKED = CT % 2 == 0;                 // column pairs: Blackwell FFMA2, two FMAs per lane
    constexpr int CP = PACKED ? CT / 2 : 1;
    f32x2_t acc2[4][CP];
    float acc1[4][CT];                                   // odd CT (unusual sub-vector sizes): scalar FMAs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) acc1[i][cc] = 0.0f;
#pragma unroll
        for (int cc = 0; cc < CP; ++cc) acc2[i][cc] = 0ull;
    }
    const int nchunks = (D + EN_KC - 1) / EN_KC;
    auto issue = [&](int t) {
        float* xs = en_sh + (size_t)(t & 1) * STAGE_FLOATS;
        float* rs = xs + NN_THREADS * EN_XS;
        const int d0 = t * EN_KC;
        // x: 256 rows x 8 segments of 4 floats
        for (int i = threadIdx.x; i < NN_THREADS * (EN_KC / 4); i += NN_THREADS) {
            const int r = i >> 3, sg = i & 7;
            const int64_t br = b0 + r;
            float* dst = xs + r * EN_XS + sg * 4;
            if (br < B && d0 + sg * 4 + 4 <= D) {
                cp_async_16(dst, x + br * ldx + d0 + sg * 4);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    dst[e] = (br < B && d0 + sg * 4 + e < D) ? __ldg(x + br * ldx + d0 + sg * 4 + e) : 0.0f;
            }
        }
        for (int i = threadIdx.x; i < NC * EN_KC; i += NN_THREADS) {
            const int j = i / EN_KC, dd = i - j * EN_KC;
            if (d0 + dd < D) cp_async_4(rs + dd * NC + j, rot + (int64_t)(m0 * DS + j) * D + d0 + dd);
            else rs[dd * NC + j] = 0.0f;
        }
        cp_async_commit();
    };
    issue(0);
    for (int t = 0; t < nchunks; ++t) {
        if (t + 1 < nchunks) {
            issue(t + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* xs = en_sh + (size_t)(t & 1) * STAGE_FLOATS;
        const float* rs = xs + NN_THREADS * EN_XS;
#pragma unroll 2
        for (int dd = 0; dd < EN_KC; dd += 4) {
            float4 xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (tr + 64 * i) * EN_XS + dd);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if constexpr (PACKED) {
                    f32x2_t rv2[CP];
#pragma unroll
                    for (int cc = 0; cc < CP; ++cc)
                        rv2[cc] = *reinterpret_cast<const f32x2_t*>(rs + (dd + q) * NC + CT * tc + 2 * cc);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float xq = q == 0 ? xv[i].x : q == 1 ? xv[i].y : q == 2 ? xv[i].z : xv[i].w;
                        const f32x2_t xx = pk2(xq, xq);
#pragma unroll
                        for (int cc = 0; cc < CP; ++cc) acc2[i][cc] = fma2_rn(xx, rv2[cc], acc2[i][cc]);
                    }
                } else {
                    float rv[CT];
#pragma unroll
                    for (int cc = 0; cc < CT; ++cc) rv[cc] = rs[(dd + q) * NC + CT * tc + cc];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float xq = q == 0 ? xv[i].x : q == 1 ? xv[i].y : q == 2 ? xv[i].z : xv[i].w;
#pragma unroll
                        for (int cc = 0; cc < CT; ++cc) acc1[i][cc] = fmaf(xq, rv[cc], acc1[i][cc]);
                    }
                }
            }
        }
        __syncthreads();                                 // the stage is refilled by the next iteration's issue
    }
    float acc[4][CT];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (PACKED) {
#pragma unroll
            for (int cc = 0; cc < CP; ++cc) unpk2(acc2[i][cc], acc[i][2 * cc], acc[i][2 * cc + 1]);
        } else {
#pragma unroll
            for (int cc = 0; cc < CT; ++cc) acc[i][cc] = acc1[i][cc];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) ys[(tr + 64 * i) * (NC + 1) + CT * tc + cc] = acc[i][cc];
    __syncthreads();
    for (int mm = 0; mm < MB; ++mm) {
        const int m = m0 + mm;
        float y[DS];
#pragma unroll
        for (int j = 0; j < DS; ++j) y[j] = ys[threadIdx.x * (NC + 1) + mm * DS + j];
        if (normalize) {
            float n2 = 0.0f;
#pragma unroll
            for (int j = 0; j < DS; ++j) n2 = fmaf(y[j], y[j], n2);
            const float inv = fmaxf(sqrtf(n2), 1e-12f);
#pragma unroll
            for (int j = 0; j < DS; ++j) y[j] = y[j] / inv;
        }
        if (y_out && live) {
            float* yp = y_out + b * ldy + (int64_t)m * DS;
#pragma unroll
            for (int j = 0; j < DS; ++j) yp[j] = y[j];
        }
        float best = 0.0f;
        int bi = 0;
        const float* cm = c + (int64_t)m * K * DS;
        for (int k0 = 0; k0 < K; k0 += kchunk) {
            const int kn = min(kchunk, K - k0);
            __syncthreads();
            for (int i = threadIdx.x; i < kn * DS; i += NN_THREADS) cs[i] = __ldg(cm + (int64_t)k0 * DS + i);
            __syncthreads();
            for (int kk = 0; kk < kn; ++kk) {
                const float* ck = cs + kk * DS;
                float d;
                if constexpr (DS % 8 == 0) {
                    const f32x2_t* ck2 = reinterpret_cast<const f32x2_t*>(ck);
                    d = sqdist_aten_order_packed<DS>([&](int p) { return pk2(y[2 * p], y[2 * p + 1]); },
                                                     [&](int p) { return ck2[p]; });
                } else {
                    d = sum_aten_order<DS>([&](int j) { return sqdiff(y[j], ck[j]); });
                }
                const int k = k0 + kk;
                if (k == 0 || d < best || (d != d && best == best)) {   // torch.argmin: first minimum, NaN smallest
                    best = d;
                    bi = k;
                }
            }
        }
        if (live) {
            if (codes_mb) codes_mb[(int64_t)m * B + b] = bi;
            if (codes_u8) codes_u8[b * M + m] = (uint8_t)bi;
        }
    }
}

template <int DS, int MB>
static int launch_encode_inst(const float* x, int64_t ldx, const float* rot, int D, const float* c, int64_t B, int M,
                              int K, int normalize, float* y, int64_t ldy, int64_t* mb, uint8_t* u8, cudaStream_t st) {
    constexpr int NC = DS * MB;
    const size_t ys_floats = std::max((size_t)NN_THREADS * (NC + 1), (size_t)2 * (NN_THREADS * EN_XS + EN_KC * NC));
    int kchunk = (16 * 1024) / (DS * 4);
    if (kchunk > K) kchunk = K;
    const size_t smem = (ys_floats + (size_t)kchunk * DS) * 4;
    auto kern = encode_assign_kernel<DS, MB>;
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen))
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    dim3 grid((unsigned)((B + NN_THREADS - 1) / NN_THREADS), (unsigned)(M / MB));
    kern<<<grid, NN_THREADS, smem, st>>>(x, ldx, rot, D, c, B, M, K, kchunk, normalize, y, ldy, mb, u8);
    RC_CHECK_LAUNCH("encode_assign_kernel");
    return RC_OK;
}

// sub-vectors per CTA: as many as fit 64 output columns and divide M
template <int DS>
static int launch_encode(const float* x, int64_t ldx, const float* rot, int D, const float* c, int64_t B, int M, int K,
                         int normalize, float* y, int64_t ldy, int64_t* mb, uint8_t* u8, cudaStream_t st) {
    if constexpr (DS * 4 <= 64) {
        if (M % 4 == 0) return launch_encode_inst<DS, 4>(x, ldx, rot, D, c, B, M, K, normalize, y, ldy, mb, u8, st);
    }
    if constexpr (DS * 2 <= 64) {
        if (M % 2 == 0) return launch_encode_inst<DS, 2>(x, ldx, rot, D, c, B, M, K, normalize, y, ldy, mb, u8, st);
    }
    return launch_encode_inst<DS, 1>(x, ldx, rot, D, c, B, M, K, normalize, y, ldy, mb, u8, st);
}

// =============================================================================================
// a1 + a3  distance table and extrema
// =============================================================================================
constexpr int TB_ROWS = 64;  // rows of x staged per CTA

// non-negative floats order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__device__ __forceinline__ void atomic_min_nonneg(float* addr, float v) {
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__global__ void minmax_init_kernel(float* minmax, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        minmax[i] = 0.0f;                         // max of non-negative distances
        minmax[M + i] = __int_as_float(0x7f800000);  // +inf
    }
}

template <int DS>
__global__ void __launch_bounds__(256)
dist_table_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ c, int64_t B, int M,
                  int K, float* __restrict__ table, float* __restrict__ minmax, int32_t* __restrict__ flags) {
    __shared__ __align__(16) float xs[TB_ROWS * DS];
    __shared__ float red_mx[8], red_mn[8];
    __shared__ int red_nan;
    const int m = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * TB_ROWS;
    const int rows = (int)min((int64_t)TB_ROWS, B - b0);
    if (threadIdx.x == 0) red_nan = 0;
    for (int i = threadIdx.x; i < rows * DS; i += blockDim.x) {
        const int r = i / DS, j = i - r * DS;
        xs[i] = __ldg(x + (b0 + r) * ldx + (int64_t)m * DS + j);
    }
    __syncthreads();
    float mx = 0.0f, mn = __int_as_float(0x7f800000);
    bool has_nan = false;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float cr[DS];
        const float* ck = c + ((int64_t)m * K + k) * DS;
#pragma unroll
        for (int j = 0; j < DS; ++j) cr[j] = __ldg(ck + j);
        float* out = table + ((int64_t)m * B + b0) * K + k;
        for (int r = 0; r < rows; ++r) {
            const float* xr = xs + r * DS;
            float d;
            if constexpr (DS % 8 == 0) {
                const f32x2_t* xr2 = reinterpret_cast<const f32x2_t*>(xr);
                d = sqdist_aten_order_packed<DS>([&](int p) { return xr2[p]; },
                                                 [&](int p) { return pk2(cr[2 * p], cr[2 * p + 1]); });
            } else {
                d = sum_aten_order<DS>([&](int j) { return sqdiff(xr[j], cr[j]); });
            }
            out[(int64_t)r * K] = d;
            if (d != d) has_nan = true;
            else {
                mx = fmaxf(mx, d);
                mn = fminf(mn, d);
            }
        }
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    if (__any_sync(0xffffffffu, has_nan) && (threadIdx.x & 31) == 0) red_nan = 1;
    if ((threadIdx.x & 31) == 0) {
        red_mx[threadIdx.x >> 5] = mx;
        red_mn[threadIdx.x >> 5] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) {
            mx = fmaxf(mx, red_mx[w]);
            mn = fminf(mn, red_mn[w]);
        }
        atomic_max_nonneg(minmax + m, mx);
        atomic_min_nonneg(minmax + M + m, mn);
        if (red_nan) atomicOr(flags, RC_FLAG_NONFINITE | RC_FLAG_AMPLITUDE);  // max() propagates NaN -> :83 fires
    }
}

__global__ void __launch_bounds__(256)
dist_table_generic_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ c, int64_t B,
                          int M, int K, int ds, float* __restrict__ table, float* __restrict__ minmax,
                          int32_t* __restrict__ flags) {
    const int m = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * TB_ROWS;
    const int rows = (int)min((int64_t)TB_ROWS, B - b0);
    float mx = 0.0f, mn = __int_as_float(0x7f800000);
    bool has_nan = false;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float* ck = c + ((int64_t)m * K + k) * ds;
        for (int r = 0; r < rows; ++r) {
            const float* xp = x + (b0 + r) * ldx + (int64_t)m * ds;
            const float d = sum_aten_order_rt([&](int j) { return sqdiff(__ldg(xp + j), __ldg(ck + j)); }, ds);
            table[((int64_t)m * B + b0 + r) * K + k] = d;
            if (d != d) has_nan = true;
            else {
                mx = fmaxf(mx, d);
                mn = fminf(mn, d);
            }
        }
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    has_nan = __any_sync(0xffffffffu, has_nan);
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(minmax + m, mx);
        atomic_min_nonneg(minmax + M + m, mn);
        if (has_nan) atomicOr(flags, RC_FLAG_NONFINITE | RC_FLAG_AMPLITUDE);
    }
}

template <int DS>
static int launch_table(const float* x, int64_t ldx, const float* c, int64_t B, int M, int K, float* table,
                        float* minmax, int32_t* flags, cudaStream_t st) {
    int threads = (K + 31) / 32 * 32;
    if (threads > 256) threads = 256;
    dim3 grid((unsigned)((B + TB_ROWS - 1) / TB_ROWS), (unsigned)M);
    dist_table_kernel<DS><<<grid, threads, 0, st>>>(x, ldx, c, B, M, K, table, minmax, flags);
    RC_CHECK_LAUNCH("dist_table_kernel");
    return RC_OK;
}

// =============================================================================================
// a3 + a4 + a5  Sinkhorn
// =============================================================================================
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_CTAS_PER_SM = 2;
constexpr int SK_DEPTH = 4;  // table rows in flight per warp (TMA ring)

// Programmatic dependent launch: the kernels of the solve loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are scheduled while this one
// drains.  Every such kernel calls pdl_wait() before it touches anything a predecessor wrote (or may still
// read) -- it returns once ALL prerequisite grids have completed and flushed -- and only then lets its own
// dependents be scheduled.  Both are no-ops under a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// one table row: global -> shared, 1-D bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load_row(float* dst_smem, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Work partition: the (m, row-tile) space is flattened (m slow) and cut into G equal contiguous
// ranges, one per persistent CTA -- balanced to one tile of SK_TILE rows whatever M and B are.  A tile is two
// rows per warp (b, b + SK_WARPS): the sparse passes keep their survivor records in such row pairs.
constexpr int SK_TILE = 2 * SK_WARPS;
struct SkPart {
    int64_t tpm;    // row tiles per sub-vector = ceil(B / SK_TILE)
    int64_t total;  // tiles of the range = (sub-vectors of the range) * tpm
    int G;          // CTAs
    int S;          // max distinct sub-vectors one CTA can touch inside the range (partial slots)
    int m0;         // first sub-vector of the range (a tile index t of the range belongs to sub-vector m0 + t / tpm)
    int slot0;      // partial rows of CTA g for this range: g * slots + slot0 + [0, S)
    int slots;      // partial rows per CTA (all ranges)
};

__host__ __device__ inline int64_t sk_lo(const SkPart& p, int g) { return (p.total * g) / p.G; }

constexpr int SK_MAX_CTAS_PER_SM = 4;  // partial buffer is sized for the largest grid any pass uses

static int sk_part_slots(int64_t total, int64_t tpm, int G) {
    const int64_t tpc = (total + G - 1) / G;
    const int S = (int)((tpc + tpm - 2) / tpm) + 1;
    return S < 2 ? 2 : S;
}

static SkPart sk_partition(int64_t B, int M, int ctas_per_sm = SK_CTAS_PER_SM) {
    SkPart p;
    p.tpm = (B + SK_TILE - 1) / SK_TILE;
    p.total = p.tpm * M;
    p.G = num_sms() * ctas_per_sm;
    p.S = sk_part_slots(p.total, p.tpm, p.G);
    p.m0 = 0;
    p.slot0 = 0;
    p.slots = p.S;
    return p;
}

// Partition of the ITERATION passes (sparse path): the sub-vectors are cut into two halves and every CTA owns a
// contiguous tile range in EACH half.  A CTA works through its range of half 0, then its range of half 1, pass
// after pass; inside the persistent kernel the row-sum reduction / exchange / update of a half's sub-vectors
// then runs while the CTAs are busy with the other half, instead of stalling them (a CTA's next pass over a
// sub-vector needs that sub-vector's update, and with one contiguous range per CTA every CTA would wait for
// the slowest CTA of its sub-vector plus the update latency in every pass).
struct SkPart2 {
    SkPart h[2];
};

// The split pays only inside the persistent kernel; the launch chain (the default) gains nothing from it and pays
// the per-segment overhead twice, so it runs with an empty first half (one contiguous range per CTA).
static bool persistent_single_rank();

static SkPart2 sk_partition2(int64_t B, int M, int ctas_per_sm) {
    SkPart2 q;
    const int64_t tpm = (B + SK_TILE - 1) / SK_TILE;
    const int G = num_sms() * ctas_per_sm;
    const int Ma = persistent_single_rank() ? M / 2 : 0, Mb = M - Ma;
    const int Sa = Ma > 0 ? sk_part_slots(tpm * Ma, tpm, G) : 0, Sb = sk_part_slots(tpm * Mb, tpm, G);
    q.h[0] = SkPart{tpm, tpm * Ma, G, Sa, 0, 0, Sa + Sb};
    q.h[1] = SkPart{tpm, tpm * Mb, G, Sb, Ma, Sa, Sa + Sb};
    return q;
}

struct SkState {
    double* lu;       // (M,K)   log2 row scaling
    double* P;        // (M,K)   row sums (all-reduce operand)
    double* lv;       // (M,B)   log2 column scaling
    double* partial;  // (G,S,K) per-CTA row-sum partials
    double* drift;    // (M,2)   {max_k, max_k - min_k} of lu - lu_build      (sparse pass)
    double* lu_build; // (M,K)   lu at the last survivor selection             (sparse pass)
    unsigned long long* cursor;  // control block (see dec / arrive / abort_w)
    unsigned int* dec;           // [0..1] done[h]: row-scaling updates completed in half h of the sub-vectors (all passes);
                                 // [2 + 2h + (u & 1)] trig[h]: u + 1 if some sub-vector of half h asked, in update u, for
                                 //     a new selection -> pass u re-selects the WHOLE half (everything else iterates on lists)
    unsigned int* arrive;        // (M) CTAs that finished the current pass on m          (persistent kernel)
    unsigned int* abort_w;       // (1) a bounded spin expired: every CTA leaves           (persistent kernel)
    uint2* csr;          // (M, tiles, SK_WARPS) row-pair directory: {pool record of the pair's first row in 16-byte
                         // units, sk_dir_word(survivors of row b, of row b + SK_WARPS)} (0: absent / did not fit);
                         // the second record follows the first one directly
    unsigned char* pool; // survivor records (layout below)
    double* U;           // (M,K)  2^(lu - lu_build - max_k(lu - lu_build)): per-column factor since selection
    unsigned long long* cta_ns;  // (G,4) diagnostics: ns per CTA in wait / selection / list / arrive+update
    int m_half1;         // first sub-vector of the second half of the iteration partition (SkPart2)
    uint32_t row_cap;    // test hook: a row with more survivors than this raises RC_FLAG_SPARSE_UNSAFE (256 = never)
    double slack;        // selection depth beyond SK_MARGIN (log2 units)
};

// Survivor record of one table row (pool, 16-byte units).  Lane l of the warp that owns the row looks after the
// 8 columns k(l,j) = (j>>2)*128 + 4*l + (j&3), j < 8 (the two float4 it loads from the row), and the record is
//   [0,64)    32 x u16, one per lane: low byte = keep mask over j, high byte = survivors in the lanes below
//   [64,..)   E = 2^(w - rowmax) of the survivors in (lane, j) order, padded to an even count
// so a lane of the list pass finds its own survivors contiguous, with their columns implied by the mask:
// no column index is stored, no per-column table is gathered and the row sums stay in registers.
// Every row pair owns a FIXED slot of the pool, sized for two full records (2 x 2112 B): the address of a pair's
// records follows from its index, nothing is allocated at run time, and the pool can never run out whatever the
// data (only the bytes of the actual survivors are ever written or read).
constexpr unsigned int SK_PAIR_BYTES = 2u * (64u + 8u * 256u);   // 4224
__host__ __device__ constexpr uint32_t sk_record_units(uint32_t cnt) { return 4u + ((cnt + 1u) >> 1); }
__device__ __forceinline__ int sp_col(int lane, int j) { return (j >> 2) * 128 + 4 * lane + (j & 3); }
// Row-pair directory word: bytes of the first record (0 = absent) | bytes of both records << 12 | odd-count
// bits (25, 26) so that the exact survivor counts can be recovered (diagnostics).  A record is
// 64 + 8 * (count rounded up to even) bytes <= 2112.
__host__ __device__ constexpr uint32_t sk_record_bytes(uint32_t cnt) { return cnt ? 64u + 8u * ((cnt + 1u) & ~1u) : 0u; }
__host__ __device__ constexpr uint32_t sk_dir_word(uint32_t cnt0, uint32_t cnt1) {
    return sk_record_bytes(cnt0) | ((sk_record_bytes(cnt0) + sk_record_bytes(cnt1)) << 12) | ((cnt0 & 1u) << 25) |
           ((cnt1 & 1u) << 26);
}
__host__ __device__ constexpr uint32_t sk_dir_len0(uint32_t w) { return w & 0xfffu; }
__host__ __device__ constexpr uint32_t sk_dir_total(uint32_t w) { return (w >> 12) & 0x1fffu; }
__host__ __device__ constexpr uint32_t sk_dir_cnt(uint32_t w, int h) {
    const uint32_t len = h ? sk_dir_total(w) - sk_dir_len0(w) : sk_dir_len0(w);
    return len ? (len - 64u) / 8u - ((w >> (25 + h)) & 1u) : 0u;
}

constexpr double SK_SLACK = 40.0;        // default extra selection depth = admissible drift of lu between selections
// (run-time tunable for experiments: env RC_SINKHORN_SLACK; deeper selection = longer lists, fewer re-selections;
//  the error bound depends on SK_MARGIN only)
static double sk_slack() {
    static double v = -1.0;
    if (v < 0.0) {
        const char* e = getenv("RC_SINKHORN_SLACK");
        v = (e && e[0]) ? atof(e) : SK_SLACK;
        if (!(v >= 1.0 && v <= 400.0)) v = SK_SLACK;
    }
    return v;
}

// test hook: pretend a row's record holds at most this many survivors (0 = no limit): a longer row raises
// RC_FLAG_SPARSE_UNSAFE exactly as an exhausted pool would -- on one rank only, if only that rank sets it
// (rc_sinkhorn_debug_pool_entries)
static int64_t g_pool_entries_override = 0;

static size_t sk_layout(int64_t B, int M, int K, const SkPart& p, void* base, SkState* s) {
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t n) {
        size_t o = off;
        off = align_up(off + n, 256);
        return o;
    };
    const size_t o_lu = take((size_t)M * K * 8);
    const size_t o_P = take((size_t)M * K * 8);
    const size_t o_lv = take((size_t)M * (size_t)B * 8);
    size_t pa = 0;
    for (int c = 1; c <= SK_MAX_CTAS_PER_SM; ++c) {
        const SkPart pc = sk_partition(B, M, c);
        const SkPart2 p2 = sk_partition2(B, M, c);
        pa = std::max(pa, (size_t)pc.G * std::max(pc.slots, p2.h[0].slots) * K * 8);
    }
    (void)p;
    const size_t o_pa = take(pa);
    const size_t o_drift = take((size_t)M * 16 + 16);
    const size_t o_lub = take((size_t)M * K * 8);
    const size_t o_cur = take((size_t)M * 16 + 64 + 32);      // (M x u64, unused) | dec (M x u32) | arrive (M x u32) | abort,
                                                         // diagnostics: selection / list segments, unsafe-mass, pool-full events
    const bool csr = (K == 256);
    const size_t pairs = (size_t)M * ((B + SK_TILE - 1) / SK_TILE) * SK_WARPS;
    const size_t o_csr = take(csr ? pairs * 8 : 0);
    const size_t o_U = take((size_t)M * K * 8);
    const size_t o_dbg = take((size_t)num_sms() * SK_MAX_CTAS_PER_SM * 4 * 8);   // per-CTA phase times (diagnostics)
    const size_t o_pool = take(csr ? pairs * SK_PAIR_BYTES : 0);
    if (s) {
        s->drift = (double*)(b + o_drift);
        s->lu_build = (double*)(b + o_lub);
        s->cursor = (unsigned long long*)(b + o_cur);
        s->dec = (unsigned int*)(b + o_cur + (size_t)M * 8);      // 8 words: done[2], trig[2][2]
        s->arrive = s->dec + 8;
        s->abort_w = s->arrive + M;
        s->csr = (uint2*)(b + o_csr);
        s->U = (double*)(b + o_U);
        s->cta_ns = (unsigned long long*)(b + o_dbg);
        s->pool = (unsigned char*)(b + o_pool);
        s->slack = sk_slack();
        s->m_half1 = sk_partition2(B, M, 2).h[1].m0;
        s->row_cap = (g_pool_entries_override > 0 && g_pool_entries_override < 256) ? (uint32_t)g_pool_entries_override : 256u;
        s->lu = (double*)(b + o_lu);
        s->P = (double*)(b + o_P);
        s->lv = (double*)(b + o_lv);
        s->partial = (double*)(b + o_pa);
    }
    return off;
}

enum { SK_BEGIN = 0, SK_STEP = 1, SK_FINISH = 2 };

// argmax tie window in log2 units: 0 = exact ties only (smallest k wins, like torch.argmax).  Duplicate
// centroids give bit-identical w and tie exactly here as in the reference.  The one configuration where
// the reference ties exactly but a log-domain evaluation does not -- a single column, B_global == 1,
// where every row normalises to exactly 1/K -- is handled by rc_sinkhorn_finish (all codes 0).
constexpr double SK_TIE_TOL_LOG2 = 0.0;

// One pass over the table.  A warp owns one table row (= one column of Q) at a time; lane l holds
// k = l, l+32, ... (KPL values, 128-byte coalesced loads).  The scaling vectors are in log2 units.
// TMA = true: every table row (K*4 bytes, contiguous) is brought in by one cp.async.bulk into a per-warp
// ring of SK_DEPTH shared-memory slots, completion on an mbarrier per slot -- SK_DEPTH rows in flight per
// warp (the 4 warps per scheduler this register-heavy kernel runs at cannot cover HBM latency with
// register prefetch alone).  TMA = false (K*4 not a multiple of 16): plain loads, one row ahead.
template <int MODE, int KPL, bool TMA, bool FULL>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM)
sinkhorn_pass_kernel(float* __restrict__ table, const float* __restrict__ minmax, int64_t B, double Bg, int M,
                     int K, double scale2 /* log2(e)/eps */, SkPart part, const double* __restrict__ lu_g,
                     double* __restrict__ lv_g, double* __restrict__ partial, int64_t* __restrict__ codes_mb,
                     uint8_t* __restrict__ codes_u8, int32_t* __restrict__ flags) {
    // per-warp ring of SK_DEPTH rows (TMA) -- its first half doubles as the warp's slice of the CTA
    // reduction buffer `red` once a segment's rows are consumed (SK_DEPTH * 4 >= 8 bytes per k)
    static_assert(SK_DEPTH >= 2, "red[] aliases the ring");
    __shared__ __align__(128) float ring[SK_WARPS * (TMA ? SK_DEPTH : 2) * KPL * 32];
    __shared__ __align__(8) unsigned long long bars[TMA ? SK_WARPS * SK_DEPTH : 1];
    constexpr int RING_W = (TMA ? SK_DEPTH : 2) * KPL * 32;  // floats per warp
    auto red = [&](int w, int k) -> double& { return reinterpret_cast<double*>(ring + w * RING_W)[k]; };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int64_t t_lo = sk_lo(part, g), t_hi = sk_lo(part, g + 1);
    if (t_lo >= t_hi) return;
    const int m_first = (int)(t_lo / part.tpm);
    int bad = 0;
    // per-warp ring state: `seq` counts rows consumed by this warp since kernel start
    float* my_ring = ring + warp * RING_W;
    const uint32_t my_bars = smem_u32(bars + (TMA ? warp * SK_DEPTH : 0));
    const uint32_t row_bytes = (uint32_t)K * 4u;
    uint32_t seq = 0;
    if (TMA) {
        if (lane == 0)
            for (int d = 0; d < SK_DEPTH; ++d) mbar_init(my_bars + 8 * d, 1);
        fence_barrier_init();
        __syncwarp();
    }

    int64_t t = t_lo;
    while (t < t_hi) {
        const int m = (int)(t / part.tpm);
        const int64_t t_end = min(t_hi, (int64_t)(m + 1) * part.tpm);
        // rows of this warp inside the segment: b = b_first + SK_WARPS * i, i < nrows
        const int64_t b_first = (t - (int64_t)m * part.tpm) * SK_TILE + warp;
        const int64_t b_stop = min(B, (t_end - (int64_t)m * part.tpm) * SK_TILE);
        const int64_t nrows = b_first < b_stop ? (b_stop - b_first + SK_WARPS - 1) / SK_WARPS : 0;

        double lu[KPL], acc[KPL];
        float middle = 0.0f, amplitude = 1.0f;
        if (MODE == SK_BEGIN) {
            const float mx = minmax[m], mn = minmax[M + m];
            middle = __fdiv_rn(__fadd_rn(mx, mn), 2.0f);                 // modeling_repconc.py:81
            amplitude = __fadd_rn(__fsub_rn(mx, middle), 1e-5f);         // :82
            if (!(amplitude > 0.0f)) bad |= RC_FLAG_AMPLITUDE;           // :83
        }
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = i * 32 + lane;
            lu[i] = (FULL || k < K) ? (MODE != SK_BEGIN ? lu_g[(int64_t)m * K + k] : 0.0) : RC_PAD_LOG2;
            acc[i] = 0.0;
        }
        float* tm = table + (int64_t)m * B * K;
        double* lvm = lv_g + (int64_t)m * B;

        float dv[KPL], nx[KPL];
        if (TMA) {
            if (lane == 0) {
                const int64_t pre = nrows < SK_DEPTH ? nrows : SK_DEPTH;
                fence_proxy_async();
                for (int64_t d = 0; d < pre; ++d) {
                    const uint32_t slot = (seq + (uint32_t)d) % SK_DEPTH;
                    bulk_load_row(my_ring + slot * KPL * 32, tm + (b_first + d * SK_WARPS) * K, row_bytes,
                                  my_bars + 8 * slot);
                }
            }
        } else if (nrows > 0) {
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = i * 32 + lane;
                dv[i] = k < K ? tm[b_first * K + k] : 0.0f;
            }
        }
        // deferred column-scaling updates: lane j keeps (z, lv, b) of the j-th row of the current group of
        // 32 rows, so the log2 costs one evaluation per 32 rows instead of one per row
        double zk = 1.0, lvk = 0.0;
        int64_t bk = -1;
        double lv_cur = 0.0, lv_nxt = 0.0;
        if (MODE == SK_STEP) lv_nxt = lane < nrows ? lvm[b_first + (int64_t)lane * SK_WARPS] : 0.0;
        for (int64_t r = 0; r < nrows; ++r) {
            const int64_t b = b_first + r * SK_WARPS;
            float* row = tm + b * K;
            if (TMA) {
                const uint32_t slot = seq % SK_DEPTH;
                mbar_wait(my_bars + 8 * slot, (seq / SK_DEPTH) & 1u);
                const float* src = my_ring + slot * KPL * 32;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = i * 32 + lane;
                    dv[i] = (FULL || k < K) ? src[k] : 0.0f;
                }
            } else if (r + 1 < nrows) {  // software prefetch of the next row
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = i * 32 + lane;
                    nx[i] = k < K ? row[(int64_t)SK_WARPS * K + k] : 0.0f;
                }
            }
            if (MODE == SK_BEGIN) {
                // centre in place: (d - middle) / amplitude in fp32 (:84), then Q0 = exp(-d~/eps) (:141)
                double w[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = i * 32 + lane;
                    const float dc = __fdiv_rn(__fsub_rn(dv[i], middle), amplitude);
                    if (FULL || k < K) row[k] = dc;
                    w[i] = fma(-(double)dc, scale2, lu[i]);
                    if (!(w[i] < 1024.0)) bad |= RC_FLAG_NONFINITE;      // exp overflow or NaN input
                }
                exp2_fast_batch<KPL, true>(w);
#pragma unroll
                for (int i = 0; i < KPL; ++i) acc[i] += w[i];
                if (lane == 0) lvm[b] = 0.0;
            } else if (MODE == SK_STEP) {
                // column scalings of a group of 32 rows are fetched together, one group ahead
                if ((r & 31) == 0) {
                    lv_cur = lv_nxt;
                    const int64_t rn = r + 32 + lane;
                    lv_nxt = rn < nrows ? lvm[b_first + rn * SK_WARPS] : 0.0;
                }
                const double lvb = __shfl_sync(0xffffffffu, lv_cur, (int)(r & 31));
                double q[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) q[i] = fma(-(double)dv[i], scale2, lu[i]) + lvb;
                exp2_fast_batch<KPL, false>(q);
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) s += q[i];
                s = warp_sum(s);                 // column sum of Q after the row normalisation (:162)
                const double z = Bg * s;         // Q /= sum; Q /= B  (:162-163)
                const double rz = __drcp_rn(z);
#pragma unroll
                for (int i = 0; i < KPL; ++i) acc[i] = fma(q[i], rz, acc[i]);   // next row sums (:155)
                if (lane == (int)(r & 31)) { zk = z; lvk = lvb; bk = b; }
                if ((r & 31) == 31 || r + 1 == nrows) {
                    if (bk >= 0) {
                        if (!(zk > 0.0) || !isfinite(zk)) bad |= RC_FLAG_NONFINITE;
                        lvm[bk] = lvk - log2(zk);
                    }
                    bk = -1;
                }
            } else {
                // argmax_k Q[m,b,k] == argmax_k (a + lu[k]); ties (see SK_TIE_TOL_LOG2) -> smallest k;
                // NaN counts as largest, as in torch.argmax (:63)
                double w[KPL];
                double best = -INFINITY;
                bool has_nan = false;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = i * 32 + lane;
                    w[i] = fma(-(double)dv[i], scale2, lu[i]);
                    if (k < K) {
                        if (w[i] != w[i]) has_nan = true;
                        else best = fmax(best, w[i]);
                        if (!(w[i] < 1024.0)) bad |= RC_FLAG_NONFINITE;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
                const bool any_nan = __any_sync(0xffffffffu, has_nan);
                const double cut = best - SK_TIE_TOL_LOG2;
                int bk2 = K;
#pragma unroll
                for (int i = KPL - 1; i >= 0; --i) {
                    const int k = i * 32 + lane;
                    if (k < K && (any_nan ? (w[i] != w[i]) : (w[i] >= cut))) bk2 = k;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bk2 = min(bk2, __shfl_xor_sync(0xffffffffu, bk2, o));
                if (bk2 >= K) bk2 = 0;  // all -inf: torch.argmax returns the first index
                if (lane == 0) {
                    if (codes_mb) codes_mb[(int64_t)m * B + b] = bk2;
                    if (codes_u8) codes_u8[b * M + m] = (uint8_t)bk2;
                }
            }
            if (TMA) {
                // the slot's values have been consumed by the arithmetic above: refill it SK_DEPTH rows ahead
                __syncwarp();
                if (lane == 0 && r + SK_DEPTH < nrows) {
                    const uint32_t slot = seq % SK_DEPTH;
                    fence_proxy_async();
                    bulk_load_row(my_ring + slot * KPL * 32, tm + (b + (int64_t)SK_DEPTH * SK_WARPS) * K, row_bytes,
                                  my_bars + 8 * slot);
                }
                ++seq;
            } else {
#pragma unroll
                for (int i = 0; i < KPL; ++i) dv[i] = nx[i];
            }
        }
        if (MODE != SK_FINISH) {
            // deterministic CTA reduction of the row-sum partials: warp 0..7 in order
#pragma unroll
            for (int i = 0; i < KPL; ++i) red(warp, i * 32 + lane) = acc[i];
            __syncthreads();
            double* dst = partial + ((int64_t)g * part.S + (m - m_first)) * K;
            for (int k = threadIdx.x; k < K; k += SK_THREADS) {
                double sum = red(0, k);
#pragma unroll
                for (int w = 1; w < SK_WARPS; ++w) sum += red(w, k);
                dst[k] = sum;
            }
            __syncthreads();
        }
        t = t_end;
    }
    if (bad) atomicOr(flags, bad);
}

// ---------------------------------------------------------------------------------------------
// Sparse passes (K == 256).  At the reference's eps (0.003) a column of Q spans hundreds of log2
// units: only ~10 % of a table row lies within 2^-72 of the row's largest element (probed: 28 of 256
// on synthetic data), and what lies below cannot change an fp64 column sum (dropped mass
// <= 256 * 2^-72 = 2^-64 of the sum) nor -- as long as every centroid keeps a sane share of the mass,
// which the row-scaling update verifies (RC_FLAG_SPARSE_UNSAFE) -- an fp64 row sum.
//
// SELECTION pass (sk_select_segment).  Lane l of the warp that owns a table row looks after the 8 columns
// k(l,j) = (j>>2)*128 + 4*l + (j&3) it loaded (two float4 from the TMA ring).  Per row the warp
//   1. evaluates log2 Q up to the column constant in fp32 for all 256 elements (8 FFMA per lane) and takes
//      the warp maximum,
//   2. keeps the elements within SK_MARGIN + slack + 0.5 of that maximum (the 0.5 covers the fp32 rounding,
//      <= 2e-4): an 8-bit keep mask per lane, an exclusive scan of the per-lane counts,
//   3. evaluates the kept elements in fp64 IN PLACE, lane by lane (w = a + lu[k] - max, 2^w by Estrin's scheme)
//      -- no cross-lane compaction: the values land in the survivor record directly in (lane, j) order, the
//      column sum is a butterfly over the per-lane sums and the row sums accumulate in REGISTERS (lane l owns
//      columns k(l,j) for the whole segment),
//   4. normalises by the column sum and adds Q / (B * sum) to its 8 row-sum accumulators.
// The column scaling lv is not needed at all here: each column is normalised by its own sum, and the
// row maximum keeps 2^w in range.  Accumulation order per (warp, k) is the row order -> deterministic.
// (The previous version compacted the survivors across lanes by ballot, evaluated them 32 at a time and then
//  re-ordered them into the lane-major record through shared memory: 560 instructions per row, a third of them
//  the re-ordering; profiles/r01_sinkhorn_step_sparse_ncu.txt.)
//
// Survivor lists are REUSED across iterations.  The selection keeps everything within SK_MARGIN + slack of the
// row maximum and emits the E = 2^(w - rowmax) of each row as a record (layout above).  Between selections only
// lu changes: an element left out had w - rowmax < -(MARGIN + slack) at selection time, and afterwards
// w - rowmax can grow by at most spread = max_k(dlu) - min_k(dlu), dlu = lu - lu_build.  The row-scaling update
// tracks that spread PER SUB-VECTOR; while spread_m <= slack the cheap LIST pass (sk_list_segment) iterates on
// the lists of sub-vector m alone -- no table read, no filter -- otherwise m is selected again.  The decision
// is one word per sub-vector (`dec[m]`), written by whoever updates the row scaling of m.
// The dense kernel above remains the path for K != 256, for dense = 1 and for the re-run after
// RC_FLAG_SPARSE_UNSAFE.
// ---------------------------------------------------------------------------------------------
constexpr double SK_MARGIN = 72.0;       // log2 units: dropped mass <= K * 2^-72 = 2^-64 of a column sum
constexpr double SK_UNSAFE_LOG2 = -8.0;  // a row that keeps < 2^-8 / K of mass voids the row-sum bound
constexpr int SP_K = 256;
constexpr int LOOP_CTAS_PER_SM = 2;

// shared memory of the iteration kernels: the list pass's ring of record slots and the selection pass's ring
// of table rows share the same bytes (a CTA runs one pass at a time and every copy in flight is consumed before
// the pass of a segment ends); the mbarriers are separate because the two rings have different depths
constexpr int LP_DEPTH = 3;                                         // list pass: slots (row pairs) in flight per warp
constexpr int LP_SLOT = 3072;                                       // bytes; a pair that does not fit is read from global
constexpr int LOOP_RING_BYTES = SK_WARPS * LP_DEPTH * LP_SLOT;      // 72 KB  (selection needs SK_WARPS*SK_DEPTH*1 KB = 32 KB)
static_assert(LOOP_RING_BYTES >= SK_WARPS * SK_DEPTH * SP_K * 4, "selection ring fits");
constexpr int LOOP_OFF_RED = LOOP_RING_BYTES;                       // red : 16 KB, [warp][k] f64
constexpr int LOOP_OFF_META = LOOP_OFF_RED + SK_WARPS * SP_K * 8;   // directory entry per list slot
constexpr int LOOP_OFF_LBAR = LOOP_OFF_META + SK_WARPS * LP_DEPTH * 8;
constexpr int LOOP_OFF_SBAR = LOOP_OFF_LBAR + SK_WARPS * LP_DEPTH * 8;
constexpr int LOOP_OFF_MISC = LOOP_OFF_SBAR + SK_WARPS * SK_DEPTH * 8;   // update scratch: 16 doubles + 8 ints
constexpr int LOOP_SMEM = LOOP_OFF_MISC + 16 * 8 + 8 * 8 + 32;

struct SkRings {       // per-warp ring positions, live for the whole kernel
    uint32_t sel_seq;                      // table rows consumed by the selection ring
    uint32_t c_slot, c_phase, i_slot;      // list ring: next slot to consume (and its phase), next slot to fill
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Selection pass over one segment (rows of sub-vector m inside tiles [t, t_end)) -- see the comment above.
// Writes the segment's row-sum partial (K doubles) to dst; ends with a block barrier.
// (tile0, tile1) = the segment's tile range INSIDE sub-vector m.
__device__ __forceinline__ void sk_select_segment(const float* __restrict__ table, int64_t B, double rBg, double scale2,
                                                  int64_t tpm, const SkState& st, int m, int64_t tile0, int64_t tile1,
                                                  unsigned char* smem, SkRings& rg, double* __restrict__ dst, int& bad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* my_ring = reinterpret_cast<float*>(smem) + warp * SK_DEPTH * SP_K;
    const uint32_t my_bars = smem_u32(smem + LOOP_OFF_SBAR) + warp * SK_DEPTH * 8;
    double* red_all = reinterpret_cast<double*>(smem + LOOP_OFF_RED);
    const int64_t b_first = tile0 * SK_TILE + warp;
    const int64_t b_stop = min(B, tile1 * SK_TILE);
    const int nrows = b_first < b_stop ? (int)((b_stop - b_first + SK_WARPS - 1) / SK_WARPS) : 0;
    const float* tm = table + ((int64_t)m * B + b_first) * SP_K;   // this warp's first row
    const int64_t pair0 = ((int64_t)m * tpm + tile0) * SK_WARPS + warp;   // index of this warp's first row pair
    uint2* dir = st.csr + pair0;
    const float scale32 = (float)scale2;
    const float sel_margin = (float)(SK_MARGIN + st.slack) + 0.5f;
    const uint32_t row_bytes = SP_K * 4u;
    uint32_t pair_cnt = 0u;

    double lu[8], A[8];
    float lu32[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        lu[j] = __ldcg(st.lu + (int64_t)m * SP_K + sp_col(lane, j));
        lu32[j] = (float)lu[j];
        A[j] = 0.0;
    }
    if (lane == 0) {
        const int pre = nrows < SK_DEPTH ? nrows : SK_DEPTH;
        fence_proxy_async();
        for (int d = 0; d < pre; ++d) {
            const uint32_t slot = (rg.sel_seq + (uint32_t)d) % SK_DEPTH;
            bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)d * SK_WARPS * SP_K, row_bytes, my_bars + 8 * slot);
        }
    }
    for (int r = 0; r < nrows; ++r) {
        const uint32_t slot = rg.sel_seq % SK_DEPTH;
        mbar_wait(my_bars + 8 * slot, (rg.sel_seq / SK_DEPTH) & 1u);
        const float* src = my_ring + slot * SP_K;
        const float4 d0 = reinterpret_cast<const float4*>(src)[lane];
        const float4 d1 = reinterpret_cast<const float4*>(src)[32 + lane];
        const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        ++rg.sel_seq;
        // 1. fp32 log2 Q (up to the column constant) and its warp maximum
        float wf[8];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            wf[j] = fmaf(-dv[j], scale32, lu32[j]);
            mx = fmaxf(mx, wf[j]);
        }
        mx = warp_max(mx);
        const float cutf = mx - sel_margin;
        // 2. keep mask, exclusive scan of the per-lane counts
        uint32_t keep_mask = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) keep_mask |= (wf[j] >= cutf ? 1u : 0u) << j;
        const int c = __popc(keep_mask);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - c;
        // every lane's values of this row have gone through the shuffles above, i.e. its shared-memory loads have
        // COMPLETED (issued is not enough: an asynchronous-proxy write must not overtake a generic read still queued
        // in the load/store unit): the slot can be refilled
        if (lane == 0 && r + SK_DEPTH < nrows)
            bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)(r + SK_DEPTH) * SK_WARPS * SP_K, row_bytes,
                          my_bars + 8 * slot);
        // the record: rows r (even) and r + 1 of this warp form a pair whose two records are contiguous in the
        // pair's fixed slot of the pool (segments are tile-aligned: local parity = global parity)
        const bool first = (r & 1) == 0;
        const int64_t pair = pair0 + (int64_t)(r >> 1) * SK_WARPS;
        unsigned char* rec = st.pool + (size_t)pair * SK_PAIR_BYTES + (first ? 0u : sk_record_bytes(pair_cnt));
        double* pe_base = reinterpret_cast<double*>(rec + 64);
        double* pe = pe_base + excl;
        const bool fits = total > 0;
        if (fits) {
            reinterpret_cast<uint16_t*>(rec)[lane] = (uint16_t)(((uint32_t)excl << 8) | keep_mask);
            if (lane == 0 && (total & 1)) pe_base[total] = 0.0;     // padding entry (copied, never used)
        }
        if ((uint32_t)total > st.row_cap) {                          // (test hook: simulated record capacity)
            bad |= RC_FLAG_SPARSE_UNSAFE;
            if (lane == 0) atomicAdd(st.abort_w + 4, 1u);
        }
        // 3. fp64 evaluation of this lane's survivors, shifted by the row maximum, straight into the record
        // (staging the record in shared memory and writing it out with coalesced 16-byte vectors was measured:
        //  no faster -- the pass is not bound by the partial-sector stores)
        const double shift = (double)mx;
        double e[8];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = 0.0;
            if ((keep_mask >> j) & 1u) {
                e[j] = exp2_fast_estrin(fma(-(double)dv[j], scale2, lu[j]) - shift);
                s += e[j];
                if (fits) *pe++ = e[j];
            }
        }
        // pair directory: written with the first row, completed with the second
        {
            uint32_t word;
            if (first) {
                pair_cnt = (uint32_t)total;
                word = sk_dir_word((uint32_t)total, 0u);
            } else {
                word = sk_dir_word(pair_cnt, (uint32_t)total);
            }
            if (lane == 0) dir[(int64_t)(r >> 1) * SK_WARPS] = make_uint2((uint32_t)pair, word);
        }
        s = warp_sum(s);
        if (!(s >= 0.5) || !isfinite(s)) bad |= RC_FLAG_NONFINITE;   // the maximum itself contributes ~1
        // 4. Q / (B_global * column sum) into the row sums (:155, :162-163)
        const double rz = __drcp_rn(s) * rBg;
#pragma unroll
        for (int j = 0; j < 8; ++j) A[j] = fma(e[j], rz, A[j]);
    }
    // the records are read back through the async proxy (cp.async.bulk) in later passes, possibly microseconds
    // from now: the generic-proxy stores above must have been performed (fence) and ordered against that proxy
    __threadfence();
    fence_proxy_async_all();
    // row sums of the segment: warp 0..7 in order (deterministic)
    {
        double* my_red = red_all + warp * SP_K;
#pragma unroll
        for (int j = 0; j < 8; ++j) my_red[sp_col(lane, j)] = A[j];
        __syncthreads();
        double sum = red_all[threadIdx.x];
#pragma unroll
        for (int w = 1; w < SK_WARPS; ++w) sum += red_all[w * SP_K + threadIdx.x];
        __stcg(dst + threadIdx.x, sum);
        __syncthreads();
    }
}

// FINISH for K == 256: argmax_k (a + lu[k]) per table row with an fp32 pre-filter.  The warp evaluates the 256
// candidates in fp32 (error < 2e-4 log2 units: |a| <= 1.45/eps, |lu| of the same order), keeps those within
// FS_TOL of the fp32 maximum and -- only if more than one is left -- evaluates the kept ones in fp64 exactly as
// the dense FINISH pass does (exact ties -> smallest k).  A row whose fp32 maximum is not an ordinary number
// (NaN anywhere, overflow range) takes the full fp64 evaluation with the dense pass's NaN / flag semantics.
// HBM-bound: one read of the table.
constexpr int FS_CTAS_PER_SM = 4;
constexpr float FS_TOL = 0.01f;

__global__ void __launch_bounds__(SK_THREADS, FS_CTAS_PER_SM)
sinkhorn_finish_sparse_kernel(const float* __restrict__ table, int64_t B, int M, double scale2, SkPart part,
                              const double* __restrict__ lu_g, int64_t* __restrict__ codes_mb,
                              uint8_t* __restrict__ codes_u8, int32_t* __restrict__ flags) {
    __shared__ __align__(128) float ring[SK_WARPS * SK_DEPTH * SP_K];
    __shared__ double lu_s[SP_K];
    __shared__ __align__(8) unsigned long long bars[SK_WARPS * SK_DEPTH];
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* my_ring = ring + warp * SK_DEPTH * SP_K;
    const uint32_t my_bars = smem_u32(bars + warp * SK_DEPTH);
    const int g = blockIdx.x;
    const int64_t t_lo = sk_lo(part, g), t_hi = sk_lo(part, g + 1);
    if (t_lo >= t_hi) return;
    int bad = 0;
    uint32_t seq = 0;
    if (lane == 0)
        for (int d = 0; d < SK_DEPTH; ++d) mbar_init(my_bars + 8 * d, 1);
    fence_barrier_init();
    __syncwarp();
    const float scale32 = (float)scale2;
    const uint32_t row_bytes = SP_K * 4u;

    int64_t t = t_lo;
    while (t < t_hi) {
        const int m = (int)(t / part.tpm);
        const int64_t t_end = min(t_hi, (int64_t)(m + 1) * part.tpm);
        const int64_t b_first = (t - (int64_t)m * part.tpm) * SK_TILE + warp;
        const int64_t b_stop = min(B, (t_end - (int64_t)m * part.tpm) * SK_TILE);
        const int nrows = b_first < b_stop ? (int)((b_stop - b_first + SK_WARPS - 1) / SK_WARPS) : 0;
        const float* tm = table + ((int64_t)m * B + b_first) * SP_K;
        __syncthreads();                                     // the previous segment's readers of lu_s are done
        lu_s[threadIdx.x] = lu_g[(int64_t)m * SP_K + threadIdx.x];
        __syncthreads();
        float lu32[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lu32[j] = (float)lu_s[sp_col(lane, j)];
        if (lane == 0) {
            const int pre = nrows < SK_DEPTH ? nrows : SK_DEPTH;
            for (int d = 0; d < pre; ++d) {
                const uint32_t slot = (seq + (uint32_t)d) % SK_DEPTH;
                bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)d * SK_WARPS * SP_K, row_bytes, my_bars + 8 * slot);
            }
        }
        for (int r = 0; r < nrows; ++r) {
            const uint32_t slot = seq % SK_DEPTH;
            mbar_wait(my_bars + 8 * slot, (seq / SK_DEPTH) & 1u);
            const float* src = my_ring + slot * SP_K;
            const float4 d0 = reinterpret_cast<const float4*>(src)[lane];
            const float4 d1 = reinterpret_cast<const float4*>(src)[32 + lane];
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            ++seq;
            float wf[8];
            float mx = -INFINITY;
            bool odd = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                wf[j] = fmaf(-dv[j], scale32, lu32[j]);
                odd |= wf[j] != wf[j];
                mx = fmaxf(mx, wf[j]);
            }
            mx = warp_max(mx);
            // every lane's values have gone through the reduction, i.e. the shared-memory loads have COMPLETED (an
            // asynchronous-proxy write must not overtake a generic read still queued in the load/store unit): refill
            if (lane == 0 && r + SK_DEPTH < nrows)
                bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)(r + SK_DEPTH) * SK_WARPS * SP_K, row_bytes,
                              my_bars + 8 * slot);
            const bool slow = __any_sync(0xffffffffu, odd) || !(mx < 1000.0f) || !(mx > -3.0e38f);
            uint32_t cand = 0xffu;
            int code = -1;
            if (!slow) {
                const float cut = mx - FS_TOL;
                cand = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) cand |= (wf[j] >= cut ? 1u : 0u) << j;
                const uint32_t bal = __ballot_sync(0xffffffffu, cand != 0u);
                const int owner = __ffs(bal) - 1;
                const int mine = __popc(cand) == 1 ? sp_col(lane, __ffs(cand) - 1) : -1;
                if ((bal & (bal - 1u)) == 0u) code = __shfl_sync(0xffffffffu, mine, owner);   // one lane has candidates
            }
            if (code < 0) {                                  // (warp-uniform) several candidates, or an odd row
                double w[8];
                double best = -INFINITY;
                bool has_nan = false;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    w[j] = -INFINITY;
                    if ((cand >> j) & 1u) {
                        w[j] = fma(-(double)dv[j], scale2, lu_s[sp_col(lane, j)]);
                        if (w[j] != w[j]) has_nan = true;
                        else best = fmax(best, w[j]);
                        if (!(w[j] < 1024.0)) bad |= RC_FLAG_NONFINITE;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
                const bool any_nan = __any_sync(0xffffffffu, has_nan);
                int bk = SP_K;
#pragma unroll
                for (int j = 7; j >= 0; --j)
                    if (((cand >> j) & 1u) && (any_nan ? (w[j] != w[j]) : (w[j] >= best))) bk = sp_col(lane, j);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bk = min(bk, __shfl_xor_sync(0xffffffffu, bk, o));
                code = bk >= SP_K ? 0 : bk;                   // all -inf: torch.argmax returns the first index
            }
            if (lane == 0) {
                const int64_t b = b_first + (int64_t)r * SK_WARPS;
                if (codes_mb) codes_mb[(int64_t)m * B + b] = code;
                if (codes_u8) codes_u8[b * M + m] = (uint8_t)code;
            }
        }
        t = t_end;
    }
    if (bad) atomicOr(flags, bad);
}

// Iteration on the survivor lists alone (see the comment above sk_select_segment).
// A record carries E = 2^(w - rowmax) as evaluated (fp64) by the selection pass; since then only lu moved, so
// the element's current value is E * 2^(lu[k] - lu_build[k]) -- up to a per-row constant that the column
// normalisation removes.  U[k] = 2^(dlu[k] - max_k dlu) is evaluated once per iteration by the row-scaling
// update (256 values per sub-vector).
// A warp owns a table row; lane l owns the 8 columns k(l,j) of the record layout, so its factors U[k(l,j)]
// and its 8 row-sum accumulators live in REGISTERS for a whole sub-vector segment: per survivor the pass
// issues one shared-memory load (the staged E) and two DFMAs -- no column gather, no read-modify-write of a
// shared row-sum array (the kernel this replaces did three random 8-byte shared accesses per survivor and
// sat at 78 % of the LSU wavefront limit, half of the wavefronts bank conflicts; profiles/r01_sinkhorn_list_*).
// Records come in through a per-warp ring of cp.async.bulk copies (one per row pair, LP_DEPTH slots in flight,
// an mbarrier per slot).  Two rows are processed together so that the cross-lane reduction (a transposing
// butterfly: 5 exchanges for both rows) and the reciprocal are shared.  The row sums are accumulated as
// sum_b E * rz_b and multiplied by U[k] once per segment.  Every sum has a fixed order.

// this lane's survivors of one row: e[j] = E of column k(lane,j) or 0; returns sum_j e[j] * U[j]
__device__ __forceinline__ double lp_gather_row(const unsigned char* rec, uint32_t hm, const double (&U)[8],
                                                double (&e)[8]) {
    const double* p = reinterpret_cast<const double*>(rec + 64) + (hm >> 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        e[j] = 0.0;
        if ((hm >> j) & 1u) e[j] = *p++;
    }
    double sa = e[0] * U[0], sb = e[1] * U[1];
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        sa = fma(e[j], U[j], sa);
        sb = fma(e[j + 1], U[j + 1], sb);
    }
    return sa + sb;
}

__device__ __forceinline__ void sk_list_segment(int64_t B, double rBg, int64_t tpm, const SkState& st, int m,
                                                int64_t tile0, int64_t tile1, unsigned char* smem, SkRings& rg,
                                                double* __restrict__ dst, int& bad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my_ring = smem + warp * LP_DEPTH * LP_SLOT;
    double* red_all = reinterpret_cast<double*>(smem + LOOP_OFF_RED);
    uint2* my_meta = reinterpret_cast<uint2*>(smem + LOOP_OFF_META) + warp * LP_DEPTH;
    const uint32_t my_bars = smem_u32(smem + LOOP_OFF_LBAR) + warp * LP_DEPTH * 8;
    // rows of this warp inside the segment: b = b_first + SK_WARPS * i, i < nrows (the rows this same warp wrote
    // back to back into its pool chunks when it selected the segment)
    const int64_t b_first = tile0 * SK_TILE + warp;
    const int64_t b_stop = min(B, tile1 * SK_TILE);
    const int nrows = b_first < b_stop ? (int)((b_stop - b_first + SK_WARPS - 1) / SK_WARPS) : 0;
    const int npairs = (nrows + 1) >> 1;
    const int64_t pair0 = ((int64_t)m * tpm + tile0) * SK_WARPS + warp;   // this warp's first row pair
    const uint2* dir = st.csr + pair0;
    double U[8], A[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        U[j] = __ldcg(st.U + (int64_t)m * SP_K + sp_col(lane, j));
        A[j] = 0.0;
    }
    // lane 0 holds the directory entry of the next pair to issue, fetched one pair ahead
    // (no proxy fence before a refill: the slot was only READ through the generic proxy, and those loads have
    //  delivered their values to the arithmetic before the __syncwarp that precedes the refill)
    uint2 dn = make_uint2(0u, 0u);
    auto fetch_meta = [&](int pair) {
        if (lane == 0) dn = pair < npairs ? dir[(int64_t)pair * SK_WARPS] : make_uint2(0u, 0u);
    };
    auto issue = [&](int pair) {
        const uint32_t slot = rg.i_slot;
        if (lane == 0) {
            const uint32_t tot = sk_dir_total(dn.y);
            const uint32_t bar = my_bars + 8 * slot;
            my_meta[slot] = make_uint2((uint32_t)pair, dn.y);
            const bool staged = tot <= (uint32_t)LP_SLOT;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(staged ? tot : 0u)
                         : "memory");
            if (staged && tot)
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(my_ring + slot * LP_SLOT)),
                    "l"(st.pool + (size_t)(pair0 + (int64_t)pair * SK_WARPS) * SK_PAIR_BYTES), "r"(tot), "r"(bar)
                    : "memory");
        }
        fetch_meta(pair + 1);
        rg.i_slot = rg.i_slot + 1 == LP_DEPTH ? 0u : rg.i_slot + 1;
    };
    if (lane == 0) fence_proxy_async();          // the ring bytes may last have been touched by generic accesses
    fetch_meta(0);
    for (int q = 0; q < npairs && q < LP_DEPTH; ++q) issue(q);
    __syncwarp();

    for (int q = 0; q < npairs; ++q) {
        const uint32_t slot = rg.c_slot;
        mbar_wait(my_bars + 8 * slot, rg.c_phase);
        const uint2 mt = my_meta[slot];                     // {record of the pair, sk_dir_word}
        const uint32_t len0 = sk_dir_len0(mt.y), tot = sk_dir_total(mt.y);
        const uint32_t cnt0 = len0, cnt1 = tot - len0;      // (only their being zero / non-zero is used below)
        uint32_t hm0 = 0u, hm1 = 0u;
        double e0[8], e1[8], s0, s1;
        if (tot <= (uint32_t)LP_SLOT) {                     // (two code paths: shared / global address space)
            const unsigned char* rec0 = my_ring + slot * LP_SLOT;
            const unsigned char* rec1 = rec0 + len0;
            if (cnt0) hm0 = reinterpret_cast<const uint16_t*>(rec0)[lane];
            if (cnt1) hm1 = reinterpret_cast<const uint16_t*>(rec1)[lane];
            s0 = lp_gather_row(rec0, hm0, U, e0);
            s1 = lp_gather_row(rec1, hm1, U, e1);
        } else {                                            // oversized pair (rare): straight from the pool
            const unsigned char* rec0 = st.pool + (size_t)(pair0 + (int64_t)mt.x * SK_WARPS) * SK_PAIR_BYTES;
            const unsigned char* rec1 = rec0 + len0;
            if (cnt0) hm0 = reinterpret_cast<const uint16_t*>(rec0)[lane];
            if (cnt1) hm1 = reinterpret_cast<const uint16_t*>(rec1)[lane];
            s0 = lp_gather_row(rec0, hm0, U, e0);
            s1 = lp_gather_row(rec1, hm1, U, e1);
        }
        // column sums of both rows: lanes 0-15 end up with row 0's, lanes 16-31 with row 1's
        const bool hi = lane & 16;
        double a = hi ? s1 : s0;
        a += __shfl_xor_sync(0xffffffffu, hi ? s0 : s1, 16);
        // every lane's partial sums -- hence every value it loaded from the slot -- went through the exchange above:
        // the shared-memory loads have COMPLETED (issued is not enough: an asynchronous-proxy write must not overtake
        // a generic read still queued in the load/store unit), so the slot can be refilled LP_DEPTH pairs ahead
        if (q + LP_DEPTH < npairs) issue(q + LP_DEPTH);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        const bool live = (hi ? cnt1 : cnt0) != 0u;
        if (live && (!(a > 0.0) || !isfinite(a))) bad |= RC_FLAG_NONFINITE;
        const double rz = live ? __drcp_rn(a) * rBg : 0.0;      // Q / (B_global * column sum)  (:162-163)
        const double rz0 = __shfl_sync(0xffffffffu, rz, 0), rz1 = __shfl_sync(0xffffffffu, rz, 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            A[j] = fma(e0[j], rz0, A[j]);
            A[j] = fma(e1[j], rz1, A[j]);
        }
        rg.c_slot = rg.c_slot + 1 == LP_DEPTH ? 0u : rg.c_slot + 1;
        rg.c_phase ^= rg.c_slot == 0 ? 1u : 0u;
    }
    // row sums of the segment: warp 0..7 in order (deterministic)
    {
        double* my_red = red_all + warp * SP_K;
#pragma unroll
        for (int j = 0; j < 8; ++j) my_red[sp_col(lane, j)] = U[j] * A[j];
        __syncthreads();
        double sum = red_all[threadIdx.x];
#pragma unroll
        for (int w = 1; w < SK_WARPS; ++w) sum += red_all[w * SP_K + threadIdx.x];
        __stcg(dst + threadIdx.x, sum);
        __syncthreads();
    }
}

// P[m,k] = sum over the CTAs that touched sub-vector m, in CTA order (deterministic); executed by one block.
// The rows of `partial` that hold sub-vector m are located once (they belong to a contiguous range of CTAs),
// then every thread sums its columns over that list with independent loads (L2: the partials come from other
// SMs, possibly within the same kernel).
constexpr int SK_RED_LIST = 256;
__device__ __forceinline__ void sk_reduce_m(int m, const double* __restrict__ partial, const SkPart& part, int K,
                                            double* __restrict__ P, int* s_off) {
    const int ml = m - part.m0;                              // sub-vector index inside the range
    const int64_t m_lo = (int64_t)ml * part.tpm, m_hi = m_lo + part.tpm;
    // only CTAs whose tile range [total*g/G, total*(g+1)/G) can touch [m_lo, m_hi)
    int g_first = (int)((m_lo * part.G) / part.total) - 1;
    int g_last = (int)((m_hi * part.G) / part.total) + 1;
    if (g_first < 0) g_first = 0;
    if (g_last > part.G - 1) g_last = part.G - 1;
    double sum[2] = {0.0, 0.0};                              // K <= 512, blockDim == 256: two columns per thread
    for (int g0 = g_first; g0 <= g_last; g0 += SK_RED_LIST) {
        const int n = min(SK_RED_LIST, g_last - g0 + 1);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int g = g0 + i;
            const int64_t lo = sk_lo(part, g), hi = sk_lo(part, g + 1);
            const bool hit = lo < hi && hi > m_lo && lo < m_hi;
            s_off[i] = hit ? g * part.slots + part.slot0 + (ml - (int)(lo / part.tpm)) : -1;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int k = threadIdx.x + c * 256;
            if (k < K) {
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const int o = s_off[i];
                    if (o >= 0) sum[c] += __ldcg(partial + (int64_t)o * K + k);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int k = threadIdx.x + c * 256;
        if (k < K) __stcg(P + (int64_t)m * K + k, sum[c]);
    }
}

// number of CTAs of the partition whose tile range touches sub-vector m (the arrivals its reduction waits for)
__device__ __forceinline__ int sk_ctas_of_m(const SkPart& part, int m) {
    const int64_t m_lo = (int64_t)(m - part.m0) * part.tpm, m_hi = m_lo + part.tpm;
    int g_first = (int)((m_lo * part.G) / part.total) - 1;
    int g_last = (int)((m_hi * part.G) / part.total) + 1;
    if (g_first < 0) g_first = 0;
    if (g_last > part.G - 1) g_last = part.G - 1;
    int n = 0;
    for (int g = g_first; g <= g_last; ++g) {
        const int64_t lo = sk_lo(part, g), hi = sk_lo(part, g + 1);
        n += (lo < hi && hi > m_lo && lo < m_hi) ? 1 : 0;
    }
    return n;
}

// pass `it` on half h selects (true) or iterates on the lists (false); valid once every update `it` of the half is done
__device__ __forceinline__ bool sk_pass_selects(const SkState& st, int h, unsigned int it) {
    return it == 0u || __ldcg(st.dec + 2 + 2 * h + (it & 1u)) == it + 1u;
}

// Row normalisation of sub-vector m in log2 form: lu[m,k] -= log2(K * P[m,k])   (Q /= sum_of_rows; Q /= K, :158-159),
// executed by one block; u = index of the update (0 follows BEGIN, u follows pass u - 1).  For the sparse passes it
// also
//   * remembers lu as lu_build when the pass that just ran on m was a selection,
//   * measures how far lu has moved since the survivors of m were selected (spread = max_k - min_k of lu - lu_build)
//     and, if that exceeds the selection slack, asks for a new selection: the request goes to the sub-vector's HALF
//     (trig[h], see SkPart2) -- all sub-vectors of a half re-select in the same pass, so that the (4x longer)
//     selection passes run side by side instead of each stalling everybody who waits for its sub-vector,
//   * evaluates the per-column factors U = 2^(dlu - max dlu) the list pass needs,
//   * checks that every centroid kept a sane share of the mass through the last column normalisation (the premise
//     of the sparse pass's row-sum bound),
//   * and counts itself done (release): pass u on the half may start once all its sub-vectors are.
__device__ __forceinline__ void sk_update_m(int m, const SkState& st, int K, int check_mass, int sparse, int h,
                                            unsigned int u, int32_t* __restrict__ flags, double* s_red /* >= 18 */) {
    const bool was_sel = sparse && u >= 1u && sk_pass_selects(st, h, u - 1u);
    double dmax = -INFINITY, dmin = INFINITY;
    int bad = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int64_t i = (int64_t)m * K + k;
        const double z = (double)K * __ldcg(st.P + i);
        if (!(z > 0.0) || !isfinite(z)) bad |= RC_FLAG_NONFINITE;
        const double dl = -log2(z);
        if (check_mass && dl > -SK_UNSAFE_LOG2) {                               // K*P[k] < 2^-8
            bad |= RC_FLAG_SPARSE_UNSAFE;
            atomicAdd(st.abort_w + 3, 1u);
        }
        const double lo = __ldcg(st.lu + i);
        const double nl = lo + dl;
        __stcg(st.lu + i, nl);
        if (sparse) {
            double lb;
            if (was_sel) {
                lb = lo;                                  // the selection that just ran used this lu
                __stcg(st.lu_build + i, lb);
            } else {
                lb = __ldcg(st.lu_build + i);
            }
            const double dv = nl - lb;
            dmax = fmax(dmax, dv);
            dmin = fmin(dmin, dv);
        }
    }
    if (bad) atomicOr(flags, bad);
    if (!sparse) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = dmax; s_red[8 + (threadIdx.x >> 5)] = dmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            dmax = fmax(dmax, s_red[w]);
            dmin = fmin(dmin, s_red[8 + w]);
        }
        const double spread = dmax - dmin;
        st.drift[2 * m] = dmax;
        st.drift[2 * m + 1] = spread;
        s_red[16] = dmax;
        if (u == 0u || !(spread <= st.slack)) atomicMax(st.dec + 2 + 2 * h + (u & 1u), u + 1u);     // also NaN
    }
    __syncthreads();
    const double dm = s_red[16];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int64_t i = (int64_t)m * K + k;
        __stcg(st.U + i, exp2(__ldcg(st.lu + i) - __ldcg(st.lu_build + i) - dm));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();          // (cumulative: the block's writes above are ordered before the barrier)
        atomicAdd(st.dec + h, 1u);
    }
}

// ---- step-wise kernels (ranks exchange P between reduce and update through the caller's all-reduce) ----------
__global__ void __launch_bounds__(256)
sinkhorn_reduce_kernel(const double* __restrict__ partial, SkPart pa, SkPart pb, int K, double* __restrict__ P) {
    __shared__ int s_off[SK_RED_LIST];
    pdl_wait();
    pdl_launch_dependents();
    const int m = blockIdx.x;
    sk_reduce_m(m, partial, m < pb.m0 ? pa : pb, K, P, s_off);      // (one range: pa == pb, m0 == 0)
}

__global__ void __launch_bounds__(256)
sinkhorn_update_kernel(SkState st, int K, int check_mass, int sparse, int m_half1, unsigned int u,
                       int32_t* __restrict__ flags) {
    __shared__ double s_red[18];
    pdl_wait();
    pdl_launch_dependents();
    sk_update_m(blockIdx.x, st, K, check_mass, sparse, (int)blockIdx.x >= m_half1 ? 1 : 0, u, flags, s_red);
}

// reduce + update in one launch: a single rank has no exchange between them.  Same functions, same order as the
// two kernels above and as the persistent kernel.
__global__ void __launch_bounds__(256)
sinkhorn_reduce_update_kernel(SkPart pa, SkPart pb, SkState st, int K, int check_mass, int sparse, int m_half1,
                              unsigned int u, int32_t* __restrict__ flags) {
    __shared__ int s_off[SK_RED_LIST];
    __shared__ double s_red[18];
    pdl_wait();
    pdl_launch_dependents();
    const int m = blockIdx.x;
    sk_reduce_m(m, st.partial, m < pb.m0 ? pa : pb, K, st.P, s_off);
    __syncthreads();
    sk_update_m(m, st, K, check_mass, sparse, m >= m_half1 ? 1 : 0, u, flags, s_red);
}

// one sparse pass over the table / the lists: per segment the selection or the list pass, as the half's decision says
__global__ void __launch_bounds__(SK_THREADS, LOOP_CTAS_PER_SM)
sinkhorn_step_kernel(const float* __restrict__ table, int64_t B, double rBg, int M, double scale2, SkPart2 part,
                     SkState st, unsigned int it, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char lp_smem[];
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    SkRings rg = {0u, 0u, 0u, 0u};
    if (lane == 0) {
        const uint32_t lb = smem_u32(lp_smem + LOOP_OFF_LBAR) + warp * LP_DEPTH * 8;
        const uint32_t sb = smem_u32(lp_smem + LOOP_OFF_SBAR) + warp * SK_DEPTH * 8;
        for (int d = 0; d < LP_DEPTH; ++d) mbar_init(lb + 8 * d, 1);
        for (int d = 0; d < SK_DEPTH; ++d) mbar_init(sb + 8 * d, 1);
    }
    fence_barrier_init();
    __syncthreads();
    int bad = 0;
    for (int h = 0; h < 2; ++h) {
        const SkPart& ph = part.h[h];
        const int64_t t_lo = sk_lo(ph, g), t_hi = sk_lo(ph, g + 1);
        if (t_lo >= t_hi) continue;
        const int ml_first = (int)(t_lo / ph.tpm);
        const bool resel = sk_pass_selects(st, h, it);
        int64_t t = t_lo;
        while (t < t_hi) {
            const int ml = (int)(t / ph.tpm), m = ph.m0 + ml;
            const int64_t t_end = min(t_hi, (int64_t)(ml + 1) * ph.tpm);
            const int64_t tile0 = t - (int64_t)ml * ph.tpm, tile1 = t_end - (int64_t)ml * ph.tpm;
            double* dst = st.partial + ((int64_t)g * ph.slots + ph.slot0 + (ml - ml_first)) * SP_K;
            if (threadIdx.x == 0) atomicAdd(st.abort_w + (resel ? 1 : 2), 1u);
            if (resel) sk_select_segment(table, B, rBg, scale2, ph.tpm, st, m, tile0, tile1, lp_smem, rg, dst, bad);
            else sk_list_segment(B, rBg, ph.tpm, st, m, tile0, tile1, lp_smem, rg, dst, bad);
            t = t_end;
        }
    }
    if (bad) atomicOr(flags, bad);
}

// ---- the persistent kernel: the whole iteration loop of one assignment in ONE launch -------------------------
// Grid = LOOP_CTAS_PER_SM x SMs co-resident CTAs (cooperative launch), the static partition of the step-wise
// kernels.  For every pass `it` and every segment (sub-vector m) of its range a CTA
//   waits until the row scaling of m for this pass is published (dec[m], acquire),
//   runs the selection or the list pass on the segment and writes its row-sum partial,
//   arrives on m's counter; the LAST CTA to arrive on m reduces the partials of m in CTA order, exchanges the
//   256 sums with the peer ranks (W > 1: publish in the symmetric buffer, raise a flag per peer with a
//   system-scope release store, wait for the W flags, sum the W vectors in rank order straight from peer
//   memory over NVLink -- the reference's `dist.all_reduce(sum_of_rows)`, modeling_repconc.py:156-157, fused
//   into the pass that produces the sums), updates lu / drift / U / the decision of m and publishes it.
// Dependencies are per sub-vector, so the reduction + exchange + update of one sub-vector overlap the passes of
// the others, and nothing ever returns to the host between iterations (the step-wise sequence costs three
// launches and, between ranks, one collective per iteration).  No deadlock: every CTA executes its (pass,
// segment) items in lexicographic order and a wait only ever targets an item that precedes the waiter's own.
// Every spin is bounded (globaltimer): on expiry RC_FLAG_PEER_TIMEOUT is raised and all CTAs leave.
constexpr int PEERX_MAX_W = 16;
struct SkPeer {
    unsigned char* base[PEERX_MAX_W];   // rank p's symmetric buffer as mapped into this process
    int rank, W;
    uint32_t seq_base;                  // exchange u of this call carries sequence number seq_base + u + 1
    unsigned long long timeout_ns;
};
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64x(const double* p) {
    double v;
    asm("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
// symmetric buffer of the fused exchange: [parity 2][sub-vector M][source rank 16][k 256] 16-byte words
__host__ __device__ constexpr size_t sk_peer_buffer_bytes(int M) { return (size_t)2 * M * PEERX_MAX_W * SP_K * 16; }

// reduce + (exchange) + update of sub-vector m by one block; u = index of the update (0 = after BEGIN)
template <int W_T>
__device__ __forceinline__ bool sk_reduce_update_m(int m, unsigned int u, const SkPart& part, const SkState& st, int M,
                                                   int32_t* __restrict__ flags, const SkPeer& peer, int* s_off,
                                                   double* s_red, int* s_flag) {
#ifdef RC_XCHG_PROFILE
    unsigned long long xt[6];
    auto xlap = [&](int i) { if (m == 0 && threadIdx.x == 0) xt[i] = global_timer_ns(); };
    xlap(0);
#else
    auto xlap = [&](int) {};
#endif
    sk_reduce_m(m, st.partial, part, SP_K, st.P, s_off);
    xlap(1);
    if (W_T != 1) {
        // Exchange without fences or separate flags (the idea of NCCL's LL protocol): a rank PUSHES its 256 sums into
        // every peer's buffer as 16-byte words {low half, seq, high half, seq} -- each 8-byte half carries the sequence
        // number of this exchange and 8-byte stores are single-copy atomic, so a reader that sees `seq` in a half sees
        // that half's data -- and every thread polls ITS element of each peer in local memory.  One NVLink store
        // latency per exchange; the publish + system fence + flag round trip + remote reads it replaces cost 14 us.
        // Slot reuse: parity of seq; a rank sends exchange s+1 only after it finished reading exchange s, and its
        // peers cannot send s+2 before they received its s+1.
        const int W = W_T > 0 ? W_T : peer.W;
        const uint32_t seq = peer.seq_base + u + 1u;
        const size_t par = seq & 1u;
        const double mine = __ldcg(st.P + (int64_t)m * SP_K + threadIdx.x);   // (each thread re-reads its own store)
        const uint32_t lo = (uint32_t)__double2loint(mine), hi = (uint32_t)__double2hiint(mine);
        const size_t slot_m = (par * M + m) * PEERX_MAX_W;
        if (threadIdx.x == 0) *s_flag = 0;
#pragma unroll
        for (int p = 0; p < PEERX_MAX_W; ++p)
            if (p < W && p != peer.rank)
                st_volatile_v4(reinterpret_cast<uint4*>(peer.base[p]) + (slot_m + peer.rank) * SP_K + threadIdx.x,
                               make_uint4(lo, seq, hi, seq));
        __syncthreads();
        xlap(2);
        double v[PEERX_MAX_W];
        const uint4* inbox = reinterpret_cast<const uint4*>(peer.base[peer.rank]) + slot_m * SP_K + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        uint32_t pending = 0u;
#pragma unroll
        for (int p = 0; p < PEERX_MAX_W; ++p)
            if (p < W && p != peer.rank) pending |= 1u << p;
        while (pending) {
#pragma unroll
            for (int p = 0; p < PEERX_MAX_W; ++p) {
                if ((pending >> p) & 1u) {
                    const uint4 w = ld_volatile_v4(inbox + (size_t)p * SP_K);
                    if (w.y == seq && w.w == seq) {
                        v[p] = __hiloint2double((int)w.z, (int)w.x);
                        pending &= ~(1u << p);
                    }
                }
            }
            if (pending && (++spins & 255u) == 0u &&
                (global_timer_ns() - t0 > peer.timeout_ns || *reinterpret_cast<volatile unsigned int*>(st.abort_w))) {
                *s_flag = 1;
                break;
            }
        }
        __syncthreads();
        xlap(3);
        if (*s_flag) {
            if (threadIdx.x == 0) {
                atomicOr(flags, RC_FLAG_PEER_TIMEOUT);
                atomicExch(st.abort_w, 1u);
            }
            return false;
        }
        // sum in rank order: every rank holds the bitwise identical result
        double sum = 0.0;
#pragma unroll
        for (int p = 0; p < PEERX_MAX_W; ++p)
            if (p < W) sum += p == peer.rank ? mine : v[p];
        __stcg(st.P + (int64_t)m * SP_K + threadIdx.x, sum);
    }
    __syncthreads();
    xlap(4);
    sk_update_m(m, st, SP_K, u > 0 ? 1 : 0, 1, m >= st.m_half1 ? 1 : 0, u, flags, s_red);
#ifdef RC_XCHG_PROFILE
    xlap(5);
    if (m == 0 && threadIdx.x == 0 && W_T != 1) {
        for (int i = 0; i < 5; ++i) atomicAdd(st.cta_ns + i, xt[i + 1] - xt[i]);
        atomicAdd(st.cta_ns + 5, 1ull);
    }
#endif
    return true;
}

// The exchange fused into the row-sum kernel of the launch chain (W ranks): block m reduces the partials of m,
// publishes the 256 sums to the peers, waits for theirs, sums in rank order from peer memory and updates the row
// scaling -- the three kernels reduce / all-reduce / update of the step-wise sequence in one launch, the
// reference's `dist.all_reduce(sum_of_rows)` (modeling_repconc.py:156-157) without a collective call.
template <int W_T>
__global__ void __launch_bounds__(256)
sinkhorn_reduce_exchange_update_kernel(SkPart pa, SkPart pb, SkState st, int M, unsigned int u,
                                       int32_t* __restrict__ flags, SkPeer peer) {
    __shared__ int s_off[SK_RED_LIST];
    __shared__ double s_red[18];
    __shared__ int s_flag;
    pdl_wait();
    pdl_launch_dependents();
    const int m = blockIdx.x;
    sk_reduce_update_m<W_T>(m, u, m < pb.m0 ? pa : pb, st, M, flags, peer, s_off, s_red, &s_flag);
}

template <int W_T>
__global__ void __launch_bounds__(SK_THREADS, LOOP_CTAS_PER_SM)
sinkhorn_loop_kernel(const float* __restrict__ table, int64_t B, double rBg, int M, double scale2, SkPart part_begin,
                     SkPart2 part, SkState st, int n_pass, int32_t* __restrict__ flags, SkPeer peer) {
    extern __shared__ __align__(128) unsigned char lp_smem[];
    int* s_ctl = reinterpret_cast<int*>(lp_smem + LOOP_OFF_MISC + 18 * 8 + 8);   // {decision word, last, abort}
    int* l_off = reinterpret_cast<int*>(lp_smem + LOOP_OFF_RED);          // (the reduction scratch is free between passes)
    double* l_red = reinterpret_cast<double*>(lp_smem + LOOP_OFF_MISC);
    int* l_flag = reinterpret_cast<int*>(lp_smem + LOOP_OFF_MISC + 18 * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    // the row scaling after BEGIN (update 0; BEGIN's partials follow its own single-range partition), dealt out
    // over the CTAs
    for (int m = g; m < M; m += gridDim.x)
        if (!sk_reduce_update_m<W_T>(m, 0u, part_begin, st, M, flags, peer, l_off, l_red, l_flag)) return;
    if (n_pass <= 0) return;
    SkRings rg = {0u, 0u, 0u, 0u};
    if (lane == 0) {
        const uint32_t lb = smem_u32(lp_smem + LOOP_OFF_LBAR) + warp * LP_DEPTH * 8;
        const uint32_t sb = smem_u32(lp_smem + LOOP_OFF_SBAR) + warp * SK_DEPTH * 8;
        for (int d = 0; d < LP_DEPTH; ++d) mbar_init(lb + 8 * d, 1);
        for (int d = 0; d < SK_DEPTH; ++d) mbar_init(sb + 8 * d, 1);
    }
    fence_barrier_init();
    __syncthreads();
    int bad = 0;
    // time spent waiting / selecting / iterating on lists / arriving + reducing + updating, summed over CTAs
    // (thread 0's clock; diagnostics read back through rc_sinkhorn_list_stats)
    unsigned long long prof[4] = {0ull, 0ull, 0ull, 0ull};
    unsigned long long tp = global_timer_ns();
    auto lap = [&](int what) {
#ifdef RC_LOOP_PROFILE
        if (threadIdx.x == 0) {
            const unsigned long long now = global_timer_ns();
            prof[what] += now - tp;
            tp = now;
        }
#else
        (void)what;
        (void)tp;
#endif
    };
    for (int it = 0; it < n_pass; ++it) {
        for (int h = 0; h < 2; ++h) {
            const SkPart& ph = part.h[h];
            const int64_t t_lo = sk_lo(ph, g), t_hi = sk_lo(ph, g + 1);
            if (t_lo >= t_hi) continue;
            const int ml_first = (int)(t_lo / ph.tpm);
            // pass `it` on this half may start once update `it` of ALL its sub-vectors is in place; they were
            // released while this CTA worked on the other half
            if (threadIdx.x == 0) {
                const unsigned int want = (unsigned int)(it + 1) * (unsigned int)(ph.total / ph.tpm);
                const unsigned long long t0 = global_timer_ns();
                unsigned int spins = 0;
                int ab = 0;
                // (pass 0 overwrites partial rows that BEGIN laid out differently: every initial update, of either
                //  half, must have consumed them)
                const unsigned int want_other = it == 0 ? (unsigned int)(part.h[1 - h].total / ph.tpm) : 0u;
                while (ld_acquire_gpu_u32(st.dec + h) < want || ld_acquire_gpu_u32(st.dec + 1 - h) < want_other) {
                    if ((++spins & 255u) == 0u) {
                        if (*reinterpret_cast<volatile unsigned int*>(st.abort_w)) { ab = 1; break; }
                        if (global_timer_ns() - t0 > peer.timeout_ns) {
                            atomicOr(flags, RC_FLAG_PEER_TIMEOUT);
                            atomicExch(st.abort_w, 1u);
                            ab = 1;
                            break;
                        }
                    }
                }
                s_ctl[0] = sk_pass_selects(st, h, (unsigned int)it) ? 1 : 0;
                s_ctl[2] = ab;
            }
            __syncthreads();
            if (s_ctl[2]) {
                if (bad) atomicOr(flags, bad);
                return;
            }
            const bool resel = s_ctl[0] != 0;
            lap(0);
            int own_m[2], n_own = 0;
            // wait until every CTA of m has arrived, then reduce + exchange + update m (this CTA arrived first)
            auto settle = [&](int m) -> bool {
                if (threadIdx.x == 0) {
                    const unsigned int n = (unsigned int)sk_ctas_of_m(ph, m);
                    const unsigned long long t0 = global_timer_ns();
                    unsigned int spins = 0;
                    int ab = 0;
                    while (ld_acquire_gpu_u32(st.arrive + m) < n) {
                        if ((++spins & 255u) == 0u) {
                            if (*reinterpret_cast<volatile unsigned int*>(st.abort_w)) { ab = 1; break; }
                            if (global_timer_ns() - t0 > peer.timeout_ns) {
                                atomicOr(flags, RC_FLAG_PEER_TIMEOUT);
                                atomicExch(st.abort_w, 1u);
                                ab = 1;
                                break;
                            }
                        }
                    }
                    st.arrive[m] = 0u;          // nobody touches the counter again before the update below is published
                    __threadfence();
                    s_ctl[2] = ab;
                }
                __syncthreads();
                if (s_ctl[2]) return false;
                return sk_reduce_update_m<W_T>(m, (unsigned int)(it + 1), ph, st, M, flags, peer, l_off, l_red, l_flag);
            };
            int64_t t = t_lo;
            while (t < t_hi) {
                const int ml = (int)(t / ph.tpm), m = ph.m0 + ml;
                const int64_t t_end = min(t_hi, (int64_t)(ml + 1) * ph.tpm);
                const int64_t tile0 = t - (int64_t)ml * ph.tpm, tile1 = t_end - (int64_t)ml * ph.tpm;
                if (threadIdx.x == 0) atomicAdd(st.abort_w + (resel ? 1 : 2), 1u);

                double* dst = st.partial + ((int64_t)g * ph.slots + ph.slot0 + (ml - ml_first)) * SP_K;
                if (resel) sk_select_segment(table, B, rBg, scale2, ph.tpm, st, m, tile0, tile1, lp_smem, rg, dst, bad);
                else sk_list_segment(B, rBg, ph.tpm, st, m, tile0, tile1, lp_smem, rg, dst, bad);
                lap(resel ? 1 : 2);
                // arrive on m (the pass functions end with a block barrier after the partial is written: thread 0's
                // fence + atomic publishes the whole block's writes).  The FIRST CTA to arrive owns the follow-up
                // (reduce, exchange, update) -- the work lands on a CTA that has time to spare, not on the slowest one
                // (which would then be late for its next segment, stay the slowest, and stretch every pass by the
                // update) -- and it does so after the LAST segment of its range in this half, when the other CTAs of
                // m have normally arrived too.
                if (bad) {
                    atomicOr(flags, bad);
                    bad = 0;
                }
                if (threadIdx.x == 0) {
                    __threadfence();
                    s_ctl[1] = atomicAdd(st.arrive + m, 1u) == 0u ? 1 : 0;
                }
                __syncthreads();
                if (s_ctl[1]) {
                    if (n_own == 2) {          // (more than two sub-vectors in one range: settle the oldest now)
                        if (!settle(own_m[0])) return;
                        own_m[0] = own_m[1];
                        n_own = 1;
                    }
                    own_m[n_own++] = m;
                }
                __syncthreads();               // s_ctl[1] is rewritten by the next arrival
                lap(3);
                t = t_end;
            }
            for (int i = 0; i < n_own; ++i)
                if (!settle(own_m[i])) return;
            lap(3);
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) st.cta_ns[4 * g + i] = prof[i];
    }
}

// Transport plan Q (M,K,B) fp64 from the row scaling alone (API parity with sinkhorn_algorithm's return
// value, modeling_repconc.py:164-165): Q[m,k,b] = 2^(a + lu[k]) / sum_k' 2^(a + lu[k']), columns sum to 1.
// One warp per table row; not on the training path.
__global__ void __launch_bounds__(256)
sinkhorn_expand_kernel(const float* __restrict__ table, int64_t B, int M, int K, double scale2,
                       const double* __restrict__ lu_g, double* __restrict__ Q) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.y;
    const int64_t b = (int64_t)blockIdx.x * 8 + warp;
    if (b >= B) return;
    const float* row = table + ((int64_t)m * B + b) * K;
    const double* lu = lu_g + (int64_t)m * K;
    double mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, fma(-(double)row[k], scale2, lu[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += exp2(fma(-(double)row[k], scale2, lu[k]) - mx);
    s = warp_sum(s);
    for (int k = lane; k < K; k += 32)
        Q[((int64_t)m * K + k) * B + b] = exp2(fma(-(double)row[k], scale2, lu[k]) - mx) / s;
}

__global__ void list_stats_kernel(const uint2* __restrict__ csr, int64_t pairs, unsigned long long* out,
                                  const unsigned int* __restrict__ counters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4) out[36 + i] = counters[1 + i];
    if (i >= pairs) return;
    const uint32_t cc = csr[i].y;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const unsigned int c = sk_dir_cnt(cc, h);
        if (c == 0) continue;
        atomicAdd(out + 0, (unsigned long long)c);
        atomicMax(out + 1, (unsigned long long)c);
        atomicAdd(out + 2, 1ull);
        atomicAdd(out + 3 + min(c >> 3, 32u), 1ull);
    }
}

__global__ void fill_f64_kernel(double* p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static bool use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RC_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// <<<grid, block, smem, st>>> with programmatic stream serialization (see pdl_wait)
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                                Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int MODE>
static int launch_pass(float* table, const float* minmax, int64_t B, double Bg, int M, int K, double eps,
                       const SkPart& p, const SkState& s, int64_t* mb, uint8_t* u8, int32_t* flags,
                       cudaStream_t st) {
    const double inv_eps = RC_LOG2E / eps;  // the passes work in base 2
    const int kpl = (K + 31) / 32;
    const bool tma = (K % 4 == 0) && (((uintptr_t)table & 15) == 0);
#define RC_SK_LAUNCH(KPL)                                                                                       \
    do {                                                                                                        \
        if (tma && K == KPL * 32)                                                                               \
            sinkhorn_pass_kernel<MODE, KPL, true, true><<<p.G, SK_THREADS, 0, st>>>(                                \
                table, minmax, B, Bg, M, K, inv_eps, p, s.lu, s.lv, s.partial, mb, u8, flags);                      \
        else if (tma)                                                                                           \
            sinkhorn_pass_kernel<MODE, KPL, true, false><<<p.G, SK_THREADS, 0, st>>>(                               \
                table, minmax, B, Bg, M, K, inv_eps, p, s.lu, s.lv, s.partial, mb, u8, flags);                      \
        else                                                                                                    \
            sinkhorn_pass_kernel<MODE, KPL, false, false><<<p.G, SK_THREADS, 0, st>>>(                              \
                table, minmax, B, Bg, M, K, inv_eps, p, s.lu, s.lv, s.partial, mb, u8, flags);                      \
    } while (0)
    if (kpl <= 2) RC_SK_LAUNCH(2);
    else if (kpl <= 4) RC_SK_LAUNCH(4);
    else if (kpl <= 8) RC_SK_LAUNCH(8);
    else if (kpl <= 16) {
        sinkhorn_pass_kernel<MODE, 16, false, false><<<p.G, SK_THREADS, 0, st>>>(table, minmax, B, Bg, M, K, inv_eps, p, s.lu,
                                                                         s.lv, s.partial, mb, u8, flags);
    } else {
        set_error("sinkhorn: K=%d > 512 is not supported", K);
        return RC_E_UNSUPPORTED;
    }
#undef RC_SK_LAUNCH
    RC_CHECK_LAUNCH("sinkhorn_pass_kernel");
    return RC_OK;
}

// pa / pb: the partition the pass that wrote the partials worked with (one range: pa == pb; the sparse passes: the
// two halves)
static int launch_reduce(const SkPart& pa, const SkPart& pb, const SkState& s, int M, int K, cudaStream_t st) {
    RC_CUDA(launch_chain(sinkhorn_reduce_kernel, (unsigned)M, 256u, 0, st, (const double*)s.partial, pa, pb, K, s.P));
    RC_CHECK_LAUNCH("sinkhorn_reduce_kernel");
    return RC_OK;
}

// update u: u = 0 follows BEGIN (its row sums are those of the unnormalised Q0: no mass check, and the first
// sparse pass always selects)
static int launch_update(const SkState& s, int M, int K, int sparse, unsigned int u, int32_t* flags, cudaStream_t st) {
    RC_CUDA(launch_chain(sinkhorn_update_kernel, (unsigned)M, 256u, 0, st, s, K, (sparse && u > 0) ? 1 : 0, sparse,
                         s.m_half1, u, flags));
    RC_CHECK_LAUNCH("sinkhorn_update_kernel");
    return RC_OK;
}

static int launch_reduce_update(const SkPart& pa, const SkPart& pb, const SkState& s, int M, int K, int sparse,
                                unsigned int u, int32_t* flags, cudaStream_t st) {
    RC_CUDA(launch_chain(sinkhorn_reduce_update_kernel, (unsigned)M, 256u, 0, st, pa, pb, s, K, (sparse && u > 0) ? 1 : 0,
                         sparse, s.m_half1, u, flags));
    RC_CHECK_LAUNCH("sinkhorn_reduce_update_kernel");
    return RC_OK;
}

static int launch_reduce_exchange_update(const SkPart& pa, const SkPart& pb, const SkState& s, int M, unsigned int u,
                                         int32_t* flags, const SkPeer& peer, cudaStream_t st) {
    void (*kern)(SkPart, SkPart, SkState, int, unsigned int, int32_t*, SkPeer);
    switch (peer.W) {
        case 2: kern = sinkhorn_reduce_exchange_update_kernel<2>; break;
        case 4: kern = sinkhorn_reduce_exchange_update_kernel<4>; break;
        case 8: kern = sinkhorn_reduce_exchange_update_kernel<8>; break;
        default: kern = sinkhorn_reduce_exchange_update_kernel<0>; break;
    }
    RC_CUDA(launch_chain(kern, (unsigned)M, 256u, 0, st, pa, pb, s, M, u, flags, peer));
    RC_CHECK_LAUNCH("sinkhorn_reduce_exchange_update_kernel");
    return RC_OK;
}

// RC_SINKHORN_PERSISTENT=1 / 0 forces the persistent kernel on / off for a single rank (default: off -- measured on
// B200: two launches per iteration with programmatic dependent launch cost ~14 us of overhead per iteration, the
// persistent kernel's arrivals, waits and in-kernel updates ~22 us -- at one rank and at two.  Between ranks both
// carry the exchange inside a kernel: the chain in its row-sum kernel, the persistent kernel in its update step)
static bool persistent_single_rank() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RC_SINKHORN_PERSISTENT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static int loop_smem_attr() {
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_loop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_loop_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_loop_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_loop_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOOP_SMEM));
    }
    return RC_OK;
}

// One sparse pass (K == 256) as its own launch: per segment the selection or the list pass, as dec[m] says.
static int launch_sparse_step(const float* table, int64_t B, int64_t B_global, int M, double eps, const SkPart2& p2,
                              const SkState& s, unsigned int it, int32_t* flags, cudaStream_t st) {
    int rc = loop_smem_attr();
    if (rc) return rc;
    const double rBg = 1.0 / (double)B_global, scale2 = RC_LOG2E / eps;
    RC_CUDA(launch_chain(sinkhorn_step_kernel, (unsigned)p2.h[1].G, (unsigned)SK_THREADS, (size_t)LOOP_SMEM, st, table, B,
                         rBg, M, scale2, p2, s, it, flags));
    RC_CHECK_LAUNCH("sinkhorn_step_kernel");
    return RC_OK;
}

static unsigned long long peer_timeout_ns() {
    static unsigned long long v = 0;
    if (v == 0) {
        const char* e = getenv("RC_PEER_TIMEOUT_MS");
        double ms = (e && e[0]) ? atof(e) : 30000.0;      // a stalled host on one rank (GC, checkpoint) must not kill training
        if (!(ms >= 1.0)) ms = 30000.0;
        v = (unsigned long long)(ms * 1e6);
    }
    return v;
}

// The whole iteration loop in one cooperative launch (see sinkhorn_loop_kernel).  Returns RC_E_UNSUPPORTED when the
// grid cannot be co-resident (the caller then runs the step-wise sequence).
static int launch_loop(const float* table, int64_t B, int64_t B_global, int M, double eps, int n_pass, const SkPart& p,
                       const SkPart2& p2, const SkState& s, int32_t* flags, const SkPeer& peer, cudaStream_t st) {
    int rc = loop_smem_attr();
    if (rc) return rc;
    void (*kern)(const float*, int64_t, double, int, double, SkPart, SkPart2, SkState, int, int32_t*, SkPeer);
    switch (peer.W) {
        case 1: kern = sinkhorn_loop_kernel<1>; break;
        case 2: kern = sinkhorn_loop_kernel<2>; break;
        case 4: kern = sinkhorn_loop_kernel<4>; break;
        case 8: kern = sinkhorn_loop_kernel<8>; break;
        default: kern = sinkhorn_loop_kernel<0>; break;
    }
    int per_sm = 0;
    RC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SK_THREADS, (size_t)LOOP_SMEM));
    if (per_sm * num_sms() < p2.h[1].G) {
        set_error("sinkhorn loop: %d CTAs cannot be co-resident (%d per SM)", p2.h[1].G, per_sm);
        return RC_E_UNSUPPORTED;
    }
    const double rBg = 1.0 / (double)B_global, scale2 = RC_LOG2E / eps;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p2.h[1].G);
    cfg.blockDim = dim3((unsigned)SK_THREADS);
    cfg.dynamicSmemBytes = (size_t)LOOP_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of the whole grid is what the spins rely on
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RC_CUDA(cudaLaunchKernelEx(&cfg, kern, table, B, rBg, M, scale2, p, p2, s, n_pass, flags, peer));
    RC_CHECK_LAUNCH("sinkhorn_loop_kernel");
    return RC_OK;
}

// Process-wide DEFAULT of the pass selection (env RC_SINKHORN_DENSE / rc_sinkhorn_set_dense: debugging, A/B runs).
// The choice that matters is per call: every entry point takes a `dense` argument, so a dense re-run on one
// stream / thread never changes what another one executes.
static int g_sinkhorn_dense = -1;  // -1: read RC_SINKHORN_DENSE from the environment on first use
static bool sinkhorn_dense_default() {
    if (g_sinkhorn_dense < 0) {
        const char* e = getenv("RC_SINKHORN_DENSE");
        g_sinkhorn_dense = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return g_sinkhorn_dense == 1;
}
static bool sk_use_sparse(const void* table, int K, int dense) {
    return K == SP_K && (((uintptr_t)table) & 15) == 0 && !dense && !sinkhorn_dense_default();
}

static int launch_finish(const float* table, int64_t B, int M, int K, double eps, const SkPart& p, const SkState& s,
                         int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags, int dense, cudaStream_t st) {
    if (sk_use_sparse(table, K, dense)) {
        const SkPart pf = sk_partition(B, M, FS_CTAS_PER_SM);
        RC_CUDA(launch_chain(sinkhorn_finish_sparse_kernel, (unsigned)pf.G, (unsigned)SK_THREADS, 0, st, table, B, M,
                             RC_LOG2E / eps, pf, (const double*)s.lu, codes_mb, codes_u8, flags));
        RC_CHECK_LAUNCH("sinkhorn_finish_sparse_kernel");
        return RC_OK;
    }
    return launch_pass<SK_FINISH>(const_cast<float*>(table), nullptr, B, (double)B, M, K, eps, p, s, codes_mb, codes_u8,
                                  flags, st);
}

}  // namespace rc

using namespace rc;

RC_API int rc_sinkhorn_set_dense(int dense) {
    const int old = sinkhorn_dense_default() ? 1 : 0;
    g_sinkhorn_dense = dense ? 1 : 0;
    return old;
}

RC_API int64_t rc_sinkhorn_debug_pool_entries(int64_t entries_per_row) {
    const int64_t old = g_pool_entries_override;
    g_pool_entries_override = entries_per_row;
    return old;
}

#define RC_DS_DISPATCH(ds, CALL, ...)     \
    switch (ds) {                         \
        case 1: CALL(1); break;           \
        case 2: CALL(2); break;           \
        case 3: CALL(3); break;           \
        case 4: CALL(4); break;           \
        case 5: CALL(5); break;           \
        case 6: CALL(6); break;           \
        case 8: CALL(8); break;           \
        case 12: CALL(12); break;         \
        case 16: CALL(16); break;         \
        case 24: CALL(24); break;         \
        case 32: CALL(32); break;         \
        case 48: CALL(48); break;         \
        case 64: CALL(64); break;         \
        case 96: CALL(96); break;         \
        default: __VA_ARGS__; break;      \
    }

RC_API int rc_nn_assign(const float* x, int64_t ldx, const float* centroids, int64_t B, int M, int K, int ds,
                        int64_t* codes_mb, uint8_t* codes_u8, void* stream) {
    RC_REQUIRE(x && centroids && (codes_mb || codes_u8), "rc_nn_assign: null pointer");
    RC_REQUIRE(B >= 0 && M >= 1 && M <= 65535 && K >= 1 && ds >= 1 && ldx >= (int64_t)M * ds,
               "rc_nn_assign: bad shape B=%lld M=%d K=%d ds=%d ldx=%lld", (long long)B, M, K, ds, (long long)ldx);
    RC_REQUIRE(!codes_u8 || K <= 256, "rc_nn_assign: uint8 codes need K <= 256 (K=%d)", K);
    if (B == 0) return RC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = RC_OK;
#define CALL(DS) rc = launch_nn<DS>(x, ldx, centroids, B, M, K, codes_mb, codes_u8, st)
    RC_DS_DISPATCH(ds, CALL, {
        dim3 grid((unsigned)((B + NN_THREADS - 1) / NN_THREADS), (unsigned)M);
        nn_assign_generic_kernel<<<grid, NN_THREADS, 0, st>>>(x, ldx, centroids, B, M, K, ds, codes_mb, codes_u8);
        RC_CHECK_LAUNCH("nn_assign_generic_kernel");
    })
#undef CALL
    return rc;
}

RC_API int rc_encode_assign(const float* pooled, int64_t ld, const float* rotation, const float* centroids, int64_t B,
                            int M, int K, int ds, int normalize, float* rotated_out, int64_t ld_out,
                            int64_t* codes_mb, uint8_t* codes_u8, void* stream) {
    RC_REQUIRE(pooled && rotation && centroids && (codes_mb || codes_u8), "rc_encode_assign: null pointer");
    RC_REQUIRE(B >= 0 && M >= 1 && K >= 1 && ds >= 1 && ld >= (int64_t)M * ds, "rc_encode_assign: bad shape");
    RC_REQUIRE(!codes_u8 || K <= 256, "rc_encode_assign: uint8 codes need K <= 256 (K=%d)", K);
    RC_REQUIRE(!rotated_out || ld_out >= (int64_t)M * ds, "rc_encode_assign: bad output stride");
    RC_REQUIRE(((uintptr_t)pooled & 15) == 0 && (ld & 3) == 0 && ((uintptr_t)rotation & 3) == 0,
               "rc_encode_assign: pooled rows must be 16-byte aligned (pointer and row stride)");
    if (B == 0) return RC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = M * ds;
    switch (ds) {
#define RC_EN(DS) case DS: return launch_encode<DS>(pooled, ld, rotation, D, centroids, B, M, K, normalize, rotated_out, ld_out, codes_mb, codes_u8, st)
        RC_EN(4); RC_EN(8); RC_EN(12); RC_EN(16); RC_EN(24); RC_EN(32);
#undef RC_EN
        default: break;
    }
    set_error("rc_encode_assign: sub-vector dimension %d is not compiled in (4, 8, 12, 16, 24, 32)", ds);
    return RC_E_UNSUPPORTED;
}

RC_API int rc_minmax_init(float* minmax, int M, void* stream) {
    RC_REQUIRE(minmax && M >= 1, "rc_minmax_init: bad argument");
    minmax_init_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(minmax, M);
    RC_CHECK_LAUNCH("minmax_init_kernel");
    return RC_OK;
}

RC_API int rc_dist_table(const float* x, int64_t ldx, const float* centroids, int64_t B, int M, int K, int ds,
                         float* table, float* minmax, int32_t* flags, void* stream) {
    RC_REQUIRE(x && centroids && table && minmax && flags, "rc_dist_table: null pointer");
    RC_REQUIRE(B >= 1 && M >= 1 && M <= 65535 && K >= 1 && ds >= 1 && ldx >= (int64_t)M * ds,
               "rc_dist_table: bad shape B=%lld M=%d K=%d ds=%d ldx=%lld", (long long)B, M, K, ds, (long long)ldx);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = RC_OK;
#define CALL(DS) rc = launch_table<DS>(x, ldx, centroids, B, M, K, table, minmax, flags, st)
    RC_DS_DISPATCH(ds, CALL, {
        dim3 grid((unsigned)((B + TB_ROWS - 1) / TB_ROWS), (unsigned)M);
        dist_table_generic_kernel<<<grid, 256, 0, st>>>(x, ldx, centroids, B, M, K, ds, table, minmax, flags);
        RC_CHECK_LAUNCH("dist_table_generic_kernel");
    })
#undef CALL
    return rc;
}

static int sk_reset(const SkState& s, int64_t B, int M, int K, cudaStream_t st) {
    const size_t n = (size_t)M * K * 8;
#ifdef RC_XCHG_PROFILE
    RC_CUDA(cudaMemsetAsync(s.cta_ns, 0, 64, st));
#endif
    if (K == SP_K)   // pair directory: entries of rows past B stay empty
        RC_CUDA(cudaMemsetAsync(s.csr, 0, (size_t)M * ((B + SK_TILE - 1) / SK_TILE) * SK_WARPS * sizeof(uint2), st));
    RC_CUDA(cudaMemsetAsync(s.lu, 0, n, st));              // +0.0
    RC_CUDA(cudaMemsetAsync(s.lu_build, 0, n, st));
    RC_CUDA(cudaMemsetAsync(s.cursor, 0, (size_t)M * 16 + 64 + 32, st));   // done / trig, arrival counters, abort, diagnostics
    return RC_OK;
}

RC_API size_t rc_sinkhorn_state_bytes(int64_t B, int M, int K) {
    if (B < 1 || M < 1 || K < 1) return 0;
    const SkPart p = sk_partition(B, M);
    return sk_layout(B, M, K, p, nullptr, nullptr);
}

RC_API double* rc_sinkhorn_rowsum_ptr(void* state, int64_t B, int M, int K) {
    if (!state || B < 1 || M < 1 || K < 1) return nullptr;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    return s.P;
}

static int sk_args(const void* table, int64_t B, int M, int K, double eps, void* state, int32_t* flags) {
    RC_REQUIRE(table && state && flags, "sinkhorn: null pointer");
    RC_REQUIRE(B >= 1 && M >= 1 && K >= 1 && K <= 512, "sinkhorn: bad shape B=%lld M=%d K=%d", (long long)B, M, K);
    RC_REQUIRE(eps > 0.0, "sinkhorn: eps must be > 0");
    RC_REQUIRE(((uintptr_t)state & 255) == 0, "sinkhorn: state must be 256-byte aligned");
    return RC_OK;
}

RC_API int rc_sinkhorn_begin(float* table, const float* minmax, int64_t B, int M, int K, double eps, void* state,
                             int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(minmax, "rc_sinkhorn_begin: null minmax");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    rc = sk_reset(s, B, M, K, st);
    if (rc) return rc;
    rc = launch_pass<SK_BEGIN>(table, minmax, B, (double)B, M, K, eps, p, s, nullptr, nullptr, flags, st);
    if (rc) return rc;
    return launch_reduce(p, p, s, M, K, st);
}

// Sinkhorn + argmax of one assignment in one call.  Sparse passes (K == 256, dense == 0): BEGIN, ONE persistent
// kernel for the whole iteration loop (with the row sums exchanged over peer memory inside it when W > 1), FINISH.
// Otherwise (any K, or the dense re-run) and W == 1: the kernels of the step-wise entry points back to back.
static int sinkhorn_solve_impl(float* table, const float* minmax, int64_t B, int64_t B_global, int M, int K, double eps,
                               int iters, int dense, void* state, const SkPeer& peer, int64_t* codes_mb,
                               uint8_t* codes_u8, int32_t* flags, cudaStream_t st) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(minmax, "rc_sinkhorn_solve: null minmax");
    RC_REQUIRE(iters >= 0, "rc_sinkhorn_solve: iters < 0");
    RC_REQUIRE(B_global >= B, "rc_sinkhorn_solve: B_global < B");
    RC_REQUIRE(codes_mb || codes_u8, "rc_sinkhorn_solve: no output");
    RC_REQUIRE(!codes_u8 || K <= 256, "rc_sinkhorn_solve: uint8 codes need K <= 256");
    const SkPart p = sk_partition(B, M);
    const SkPart2 p2 = sk_partition2(B, M, LOOP_CTAS_PER_SM);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    const bool sparse = sk_use_sparse(table, K, dense);
    if (peer.W > 1 && !sparse) {
        set_error("rc_sinkhorn_solve_peer: the fused exchange exists for the sparse passes (K == 256, dense == 0) only");
        return RC_E_UNSUPPORTED;
    }
    rc = sk_reset(s, B, M, K, st);
    if (rc) return rc;
    rc = launch_pass<SK_BEGIN>(table, minmax, B, (double)B, M, K, eps, p, s, nullptr, nullptr, flags, st);
    if (rc) return rc;
    if (iters >= 1 && B_global == 1) {
        // one column: every entry is exactly 1/K after the row normalisation (see rc_sinkhorn_finish)
        if (codes_mb) RC_CUDA(cudaMemsetAsync(codes_mb, 0, (size_t)M * B * sizeof(int64_t), st));
        if (codes_u8) RC_CUDA(cudaMemsetAsync(codes_u8, 0, (size_t)M * B, st));
        return RC_OK;
    }
    if (iters >= 1) {
        bool looped = false;
        if (sparse && persistent_single_rank()) {
            rc = launch_loop(table, B, B_global, M, eps, iters - 1, p, p2, s, flags, peer, st);
            if (rc == RC_OK) looped = true;
            else if (rc != RC_E_UNSUPPORTED) return rc;
        }
        if (!looped) {
            for (int it = 0; it < iters; ++it) {
                const bool two = sparse && it > 0;       // which partition wrote the partials
                if (peer.W > 1)
                    rc = launch_reduce_exchange_update(two ? p2.h[0] : p, two ? p2.h[1] : p, s, M, (unsigned int)it, flags,
                                                       peer, st);
                else
                    rc = launch_reduce_update(two ? p2.h[0] : p, two ? p2.h[1] : p, s, M, K, sparse ? 1 : 0,
                                              (unsigned int)it, flags, st);
                if (rc) return rc;
                if (it == iters - 1) break;
                if (sparse) rc = launch_sparse_step(table, B, B_global, M, eps, p2, s, (unsigned int)it, flags, st);
                else rc = launch_pass<SK_STEP>(table, nullptr, B, (double)B_global, M, K, eps, p, s, nullptr, nullptr, flags, st);
                if (rc) return rc;
            }
        }
    }
    return launch_finish(table, B, M, K, eps, p, s, codes_mb, codes_u8, flags, dense, st);
}

RC_API int rc_sinkhorn_solve(float* table, const float* minmax, int64_t B, int M, int K, double eps, int iters,
                             int dense, void* state, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                             void* stream) {
    SkPeer peer{};
    peer.W = 1;
    peer.timeout_ns = peer_timeout_ns();
    return sinkhorn_solve_impl(table, minmax, B, B, M, K, eps, iters, dense, state, peer, codes_mb, codes_u8, flags,
                               (cudaStream_t)stream);
}

RC_API size_t rc_sinkhorn_peer_buffer_bytes(int M, int K) {
    if (M < 1 || K < 1) return 0;
    (void)K;
    return sk_peer_buffer_bytes(M);
}

RC_API int rc_sinkhorn_solve_peer(float* table, const float* minmax, int64_t B, int64_t B_global, int M, int K,
                                  double eps, int iters, void* state, const uint64_t* peer_buffers_host, int rank,
                                  int W, uint32_t seq_base, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                                  void* stream) {
    RC_REQUIRE(peer_buffers_host, "rc_sinkhorn_solve_peer: null peer buffers");
    RC_REQUIRE(W >= 1 && W <= PEERX_MAX_W && rank >= 0 && rank < W, "rc_sinkhorn_solve_peer: bad rank %d / world %d", rank, W);
    RC_REQUIRE(K == SP_K, "rc_sinkhorn_solve_peer: K must be 256 (K=%d)", K);
    SkPeer peer{};
    for (int p = 0; p < W; ++p) {
        RC_REQUIRE(peer_buffers_host[p] != 0, "rc_sinkhorn_solve_peer: null peer buffer %d", p);
        peer.base[p] = reinterpret_cast<unsigned char*>(peer_buffers_host[p]);
    }
    peer.rank = rank;
    peer.W = W;
    peer.seq_base = seq_base;
    peer.timeout_ns = peer_timeout_ns();
    return sinkhorn_solve_impl(table, minmax, B, B_global, M, K, eps, iters, 0, state, peer, codes_mb, codes_u8, flags,
                               (cudaStream_t)stream);
}

RC_API int rc_sinkhorn_step(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                            int step_index, int dense, void* state, int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(B_global >= B, "rc_sinkhorn_step: B_global < B");
    RC_REQUIRE(step_index >= 0, "rc_sinkhorn_step: step_index < 0");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    // The sparse pass needs rows (centroids) that kept their share of the mass through the previous column
    // normalisation; the update checks that from the second STEP on (the first update sees the row sums of the
    // unnormalised Q0, which say nothing about it).
    const bool sparse = sk_use_sparse(table, K, dense);
    rc = launch_update(s, M, K, sparse ? 1 : 0, (unsigned int)step_index, flags, st);
    if (rc) return rc;
    if (sparse) {
        const SkPart2 p2 = sk_partition2(B, M, LOOP_CTAS_PER_SM);
        rc = launch_sparse_step(table, B, B_global, M, eps, p2, s, (unsigned int)step_index, flags, st);
        if (rc) return rc;
        return launch_reduce(p2.h[0], p2.h[1], s, M, K, st);
    }
    rc = launch_pass<SK_STEP>(const_cast<float*>(table), nullptr, B, (double)B_global, M, K, eps, p, s, nullptr, nullptr,
                              flags, st);
    if (rc) return rc;
    return launch_reduce(p, p, s, M, K, st);
}

RC_API int rc_sinkhorn_debug_cta_times(void* state, int64_t B, int M, int K, int64_t* out_host, int max_ctas) {
    RC_REQUIRE(state && out_host && B >= 1 && M >= 1 && K >= 1, "rc_sinkhorn_debug_cta_times: bad argument");
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    const int G = std::min(max_ctas, num_sms() * LOOP_CTAS_PER_SM);
    RC_CUDA(cudaDeviceSynchronize());
    RC_CUDA(cudaMemcpy(out_host, s.cta_ns, (size_t)G * 4 * 8, cudaMemcpyDeviceToHost));
    return G;
}

RC_API int rc_sinkhorn_debug_drift(void* state, int64_t B, int M, int K, double* out_host) {
    RC_REQUIRE(state && out_host && B >= 1 && M >= 1 && K >= 1, "rc_sinkhorn_debug_drift: bad argument");
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    RC_CUDA(cudaDeviceSynchronize());
    RC_CUDA(cudaMemcpy(out_host, s.drift, (size_t)M * 2 * 8, cudaMemcpyDeviceToHost));
    return RC_OK;
}

RC_API int rc_sinkhorn_list_stats(void* state, int64_t B, int M, int K, int64_t* out, void* stream) {
    RC_REQUIRE(state && out && B >= 1 && M >= 1, "rc_sinkhorn_list_stats: bad argument");
    RC_REQUIRE(K == SP_K, "rc_sinkhorn_list_stats: survivor lists exist for K == 256 only (K=%d)", K);
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    RC_CUDA(cudaMemsetAsync(out, 0, 40 * sizeof(int64_t), st));
    const int64_t pairs = (int64_t)M * p.tpm * SK_WARPS;
    list_stats_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(s.csr, pairs, (unsigned long long*)out, s.abort_w);
    RC_CHECK_LAUNCH("list_stats_kernel");
    return RC_OK;
}

RC_API int rc_sinkhorn_expand(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                              int apply_rowsum, void* state, double* Q, int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(Q, "rc_sinkhorn_expand: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    if (apply_rowsum) {
        rc = launch_update(s, M, K, 0, 0u, flags, st);
        if (rc) return rc;
    }
    dim3 grid((unsigned)((B + 7) / 8), (unsigned)M);
    sinkhorn_expand_kernel<<<grid, 256, 0, st>>>(table, B, M, K, RC_LOG2E / eps, s.lu, Q);
    RC_CHECK_LAUNCH("sinkhorn_expand_kernel");
    return RC_OK;
}

RC_API int rc_sinkhorn_finish(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                              int apply_rowsum, int steps_done, int dense, void* state, int64_t* codes_mb,
                              uint8_t* codes_u8, int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(codes_mb || codes_u8, "rc_sinkhorn_finish: no output");
    RC_REQUIRE(!codes_u8 || K <= 256, "rc_sinkhorn_finish: uint8 codes need K <= 256");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    if (apply_rowsum && B_global == 1) {
        // one column: after the row normalisation every entry is exactly 1/K in the reference, so its
        // argmax is index 0 for every sub-vector (modeling_repconc.py:158-159,63)
        if (codes_mb) RC_CUDA(cudaMemsetAsync(codes_mb, 0, (size_t)M * B * sizeof(int64_t), st));
        if (codes_u8) RC_CUDA(cudaMemsetAsync(codes_u8, 0, (size_t)M * B, st));
        return RC_OK;
    }
    if (apply_rowsum) {
        // the row sums of the last sparse pass get the same mass check as every other one (steps_done >= 1: they
        // come from a STEP pass, not from BEGIN's unnormalised Q0 -- see rc_sinkhorn_step)
        rc = launch_update(s, M, K, sk_use_sparse(table, K, dense) ? 1 : 0, (unsigned int)(steps_done > 0 ? steps_done : 0),
                           flags, st);
        if (rc) return rc;
    }
    return launch_finish(table, B, M, K, eps, p, s, codes_mb, codes_u8, flags, dense, st);
}
