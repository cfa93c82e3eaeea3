// assign.cu -- NN assign, distance table + extrema, and the Sinkhorn uniform assignment
// (reference: src/repconc/models/repconc/modeling_repconc.py:47-85,137-165).
//
// Data layout in HBM
//   x          (B, D) fp32, D = M*ds                         caller's embeddings
//   centroids  (M, K, ds) fp32                               nn.Parameter of the module
//   table      (M, B, K) fp32, k fastest                     raw distances, centred IN PLACE by
//                                                            rc_sinkhorn_begin (never widened to fp64)
//   state      lu (M,K) f64 | P (M,K) f64 | lv (M,B) f64 | partial (G, S, K) f64
//
// Sinkhorn formulation.  The reference materialises Q = exp(-d~/eps) as (M,K,B) fp64 and divides
// it in place 4x per iteration.  Here Q_t = exp(a + lu[k] + lv[b]) with a = -d~/eps is never
// stored: one pass over the fp32 table per iteration evaluates each element once, finishes the
// column normalisation of iteration t inside a warp (a table row is one column of Q) and
// accumulates the row sums that iteration t+1 needs.  Per iteration: 4 B/element of HBM traffic
// instead of the reference's ~80 B/element, and ONE exp per element.
#include <math.h>

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
            const float d = sum_aten_order<DS>([&](int j) { return sqdiff(xr[j], ck[j]); });
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
            const float d = sum_aten_order<DS>([&](int j) { return sqdiff(xr[j], cr[j]); });
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

// Work partition: the (m, row-tile) space is flattened (m slow) and cut into G equal contiguous
// ranges, one per persistent CTA -- balanced to one tile of SK_WARPS rows whatever M and B are.
struct SkPart {
    int64_t tpm;    // row tiles per sub-vector = ceil(B / SK_WARPS)
    int64_t total;  // M * tpm
    int G;          // CTAs
    int S;          // max distinct sub-vectors one CTA can touch (partial slots)
};

__host__ __device__ inline int64_t sk_lo(const SkPart& p, int g) { return (p.total * g) / p.G; }

static SkPart sk_partition(int64_t B, int M) {
    SkPart p;
    p.tpm = (B + SK_WARPS - 1) / SK_WARPS;
    p.total = p.tpm * M;
    p.G = num_sms() * SK_CTAS_PER_SM;
    const int64_t tpc = (p.total + p.G - 1) / p.G;
    p.S = (int)((tpc + p.tpm - 2) / p.tpm) + 1;
    if (p.S < 2) p.S = 2;
    return p;
}

struct SkState {
    double* lu;       // (M,K)
    double* P;        // (M,K)
    double* lv;       // (M,B)
    double* partial;  // (G,S,K)
};

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
    const size_t o_pa = take((size_t)p.G * p.S * K * 8);
    if (s) {
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
template <int MODE, int KPL>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM)
sinkhorn_pass_kernel(float* __restrict__ table, const float* __restrict__ minmax, int64_t B, double Bg, int M,
                     int K, double scale2 /* log2(e)/eps */, SkPart part, const double* __restrict__ lu_g,
                     double* __restrict__ lv_g, double* __restrict__ partial, int64_t* __restrict__ codes_mb,
                     uint8_t* __restrict__ codes_u8, int32_t* __restrict__ flags) {
    __shared__ double red[SK_WARPS][KPL * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int64_t t_lo = sk_lo(part, g), t_hi = sk_lo(part, g + 1);
    if (t_lo >= t_hi) return;
    const int m_first = (int)(t_lo / part.tpm);
    int bad = 0;

    int64_t t = t_lo;
    while (t < t_hi) {
        const int m = (int)(t / part.tpm);
        const int64_t t_end = min(t_hi, (int64_t)(m + 1) * part.tpm);
        // rows of this warp inside the segment: b = b_first + SK_WARPS * i, i < nrows
        const int64_t b_first = (t - (int64_t)m * part.tpm) * SK_WARPS + warp;
        const int64_t b_stop = min(B, (t_end - (int64_t)m * part.tpm) * SK_WARPS);
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
            lu[i] = k < K ? (MODE != SK_BEGIN ? lu_g[(int64_t)m * K + k] : 0.0) : RC_PAD_LOG2;
            acc[i] = 0.0;
        }
        float* tm = table + (int64_t)m * B * K;
        double* lvm = lv_g + (int64_t)m * B;

        float dv[KPL], nx[KPL];
        if (nrows > 0) {
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
        for (int64_t r = 0; r < nrows; ++r) {
            const int64_t b = b_first + r * SK_WARPS;
            float* row = tm + b * K;
            if (r + 1 < nrows) {  // software prefetch of the next row
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
                    if (k < K) row[k] = dc;
                    w[i] = fma(-(double)dc, scale2, lu[i]);
                    if (!(w[i] < 1024.0)) bad |= RC_FLAG_NONFINITE;      // exp overflow or NaN input
                }
#pragma unroll
                for (int i = 0; i < KPL; ++i) acc[i] += exp2_fast(w[i]);
                if (lane == 0) lvm[b] = 0.0;
            } else if (MODE == SK_STEP) {
                const double lvb = lvm[b];
                double q[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) q[i] = fma(-(double)dv[i], scale2, lu[i]) + lvb;
#pragma unroll
                for (int i = 0; i < KPL; ++i) q[i] = exp2_fast(q[i]);
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
#pragma unroll
            for (int i = 0; i < KPL; ++i) dv[i] = nx[i];
        }
        if (MODE != SK_FINISH) {
            // deterministic CTA reduction of the row-sum partials: warp 0..7 in order
#pragma unroll
            for (int i = 0; i < KPL; ++i) red[warp][i * 32 + lane] = acc[i];
            __syncthreads();
            double* dst = partial + ((int64_t)g * part.S + (m - m_first)) * K;
            for (int k = threadIdx.x; k < K; k += SK_THREADS) {
                double sum = red[0][k];
#pragma unroll
                for (int w = 1; w < SK_WARPS; ++w) sum += red[w][k];
                dst[k] = sum;
            }
            __syncthreads();
        }
        t = t_end;
    }
    if (bad) atomicOr(flags, bad);
}

// P[m,k] = sum over the CTAs that touched sub-vector m, in CTA order (deterministic).
__global__ void __launch_bounds__(256)
sinkhorn_reduce_kernel(const double* __restrict__ partial, SkPart part, int K, double* __restrict__ P) {
    const int m = blockIdx.x;
    const int64_t m_lo = (int64_t)m * part.tpm, m_hi = m_lo + part.tpm;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double sum = 0.0;
        for (int g = 0; g < part.G; ++g) {
            const int64_t lo = sk_lo(part, g), hi = sk_lo(part, g + 1);
            if (lo >= hi || hi <= m_lo || lo >= m_hi) continue;
            const int slot = m - (int)(lo / part.tpm);
            sum += partial[((int64_t)g * part.S + slot) * K + k];
        }
        P[(int64_t)m * K + k] = sum;
    }
}

// row normalisation in log2 form: lu[m,k] -= log2(K * P[m,k])     (Q /= sum_of_rows; Q /= K, :158-159)
__global__ void sinkhorn_update_kernel(double* __restrict__ lu, const double* __restrict__ P, int64_t n, double Kd,
                                       int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double z = Kd * P[i];
    if (!(z > 0.0) || !isfinite(z)) atomicOr(flags, RC_FLAG_NONFINITE);
    lu[i] -= log2(z);
}

__global__ void fill_f64_kernel(double* p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int MODE>
static int launch_pass(float* table, const float* minmax, int64_t B, double Bg, int M, int K, double eps,
                       const SkPart& p, const SkState& s, int64_t* mb, uint8_t* u8, int32_t* flags,
                       cudaStream_t st) {
    const double inv_eps = RC_LOG2E / eps;  // the passes work in base 2
    const int kpl = (K + 31) / 32;
#define RC_SK_LAUNCH(KPL)                                                                                  \
    sinkhorn_pass_kernel<MODE, KPL><<<p.G, SK_THREADS, 0, st>>>(table, minmax, B, Bg, M, K, inv_eps, p, s.lu, \
                                                                s.lv, s.partial, mb, u8, flags)
    if (kpl <= 2) RC_SK_LAUNCH(2);
    else if (kpl <= 4) RC_SK_LAUNCH(4);
    else if (kpl <= 8) RC_SK_LAUNCH(8);
    else if (kpl <= 16) RC_SK_LAUNCH(16);
    else {
        set_error("sinkhorn: K=%d > 512 is not supported", K);
        return RC_E_UNSUPPORTED;
    }
#undef RC_SK_LAUNCH
    RC_CHECK_LAUNCH("sinkhorn_pass_kernel");
    return RC_OK;
}

static int launch_reduce(const SkPart& p, const SkState& s, int M, int K, cudaStream_t st) {
    sinkhorn_reduce_kernel<<<M, 256, 0, st>>>(s.partial, p, K, s.P);
    RC_CHECK_LAUNCH("sinkhorn_reduce_kernel");
    return RC_OK;
}

static int launch_update(const SkState& s, int M, int K, int32_t* flags, cudaStream_t st) {
    const int64_t n = (int64_t)M * K;
    sinkhorn_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.lu, s.P, n, (double)K, flags);
    RC_CHECK_LAUNCH("sinkhorn_update_kernel");
    return RC_OK;
}

}  // namespace rc

using namespace rc;

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
    const int64_t n = (int64_t)M * K;
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.lu, n, 0.0);
    RC_CHECK_LAUNCH("fill_f64_kernel");
    rc = launch_pass<SK_BEGIN>(table, minmax, B, (double)B, M, K, eps, p, s, nullptr, nullptr, flags, st);
    if (rc) return rc;
    return launch_reduce(p, s, M, K, st);
}

RC_API int rc_sinkhorn_step(const float* table, int64_t B, int64_t B_global, int M, int K, double eps, void* state,
                            int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(B_global >= B, "rc_sinkhorn_step: B_global < B");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    rc = launch_update(s, M, K, flags, st);
    if (rc) return rc;
    rc = launch_pass<SK_STEP>(const_cast<float*>(table), nullptr, B, (double)B_global, M, K, eps, p, s, nullptr,
                              nullptr, flags, st);
    if (rc) return rc;
    return launch_reduce(p, s, M, K, st);
}

RC_API int rc_sinkhorn_finish(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                              int apply_rowsum, void* state, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                              void* stream) {
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
        rc = launch_update(s, M, K, flags, st);
        if (rc) return rc;
    }
    return launch_pass<SK_FINISH>(const_cast<float*>(table), nullptr, B, (double)B, M, K, eps, p, s, codes_mb,
                                  codes_u8, flags, st);
}
