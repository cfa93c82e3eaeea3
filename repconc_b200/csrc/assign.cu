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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
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
    int64_t total;  // M * tpm
    int G;          // CTAs
    int S;          // max distinct sub-vectors one CTA can touch (partial slots)
};

__host__ __device__ inline int64_t sk_lo(const SkPart& p, int g) { return (p.total * g) / p.G; }

constexpr int SK_MAX_CTAS_PER_SM = 4;  // partial buffer is sized for the largest grid any pass uses

static SkPart sk_partition(int64_t B, int M, int ctas_per_sm = SK_CTAS_PER_SM) {
    SkPart p;
    p.tpm = (B + SK_TILE - 1) / SK_TILE;
    p.total = p.tpm * M;
    p.G = num_sms() * ctas_per_sm;
    const int64_t tpc = (p.total + p.G - 1) / p.G;
    p.S = (int)((tpc + p.tpm - 2) / p.tpm) + 1;
    if (p.S < 2) p.S = 2;
    return p;
}

struct SkState {
    double* lu;       // (M,K)   log2 row scaling
    double* P;        // (M,K)   row sums (all-reduce operand)
    double* lv;       // (M,B)   log2 column scaling
    double* partial;  // (G,S,K) per-CTA row-sum partials
    double* drift;    // (M,2)   {max_k, max_k - min_k} of lu - lu_build      (sparse pass)
    double* lu_build; // (M,K)   lu at the last survivor selection             (sparse pass)
    unsigned long long* cursor;  // pool allocation cursor
    uint2* csr;          // (M, tiles, SK_WARPS) row-pair directory: {pool record of the pair's first row in 16-byte
                         // units, sk_dir_word(survivors of row b, of row b + SK_WARPS)} (0: absent / did not fit);
                         // the second record follows the first one directly
    unsigned char* pool; // survivor records (layout below)
    double* U;           // (M,K)  2^(lu - lu_build - max_k(lu - lu_build)): per-column factor since selection
    uint64_t pool_cap;   // pool size in 16-byte units
    double slack;        // selection depth beyond SK_MARGIN (log2 units)
};

// Survivor record of one table row (pool, 16-byte units).  Lane l of the warp that owns the row looks after the
// 8 columns k(l,j) = (j>>2)*128 + 4*l + (j&3), j < 8 (the two float4 it loads from the row), and the record is
//   [0,64)    32 x u16, one per lane: low byte = keep mask over j, high byte = survivors in the lanes below
//   [64,..)   E = 2^(w - rowmax) of the survivors in (lane, j) order, padded to an even count
// so a lane of the list pass finds its own survivors contiguous, with their columns implied by the mask:
// no column index is stored, no per-column table is gathered and the row sums stay in registers.
constexpr unsigned int SK_POOL_CHUNK = 1024;  // 16-byte units a warp grabs per atomic (16 KB; a full row is 132)
constexpr int SK_POOL_PER_ROW = 160;  // survivor pool sized for this many entries per table row on average
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

// test hook: usable survivor-pool capacity in entries per table row (0 = the allocation's SK_POOL_PER_ROW); lets a
// test exhaust the pool on one rank only (rc_sinkhorn_debug_pool_entries)
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
        pa = std::max(pa, (size_t)pc.G * pc.S * K * 8);
    }
    (void)p;
    const size_t o_pa = take(pa);
    const size_t o_drift = take((size_t)M * 16 + 16);   // + {int decision, int block counter}
    const size_t o_lub = take((size_t)M * K * 8);
    const size_t o_cur = take(256);
    // every warp of the selection pass may strand up to one chunk
    uint64_t pool = (uint64_t)M * (uint64_t)B * sk_record_units(SK_POOL_PER_ROW) +
                    (uint64_t)num_sms() * SK_MAX_CTAS_PER_SM * SK_WARPS * SK_POOL_CHUNK;
    if (pool > 0xFFFFFF00ull) pool = 0xFFFFFF00ull;      // record offsets are 32-bit
    const bool csr = (K == 256);
    const size_t o_csr = take(csr ? (size_t)M * ((B + SK_TILE - 1) / SK_TILE) * SK_WARPS * 8 : 0);
    const size_t o_U = take((size_t)M * K * 8);
    const size_t o_pool = take(csr ? (size_t)pool * 16 : 0);
    if (s) {
        s->drift = (double*)(b + o_drift);
        s->lu_build = (double*)(b + o_lub);
        s->cursor = (unsigned long long*)(b + o_cur);
        s->csr = (uint2*)(b + o_csr);
        s->U = (double*)(b + o_U);
        s->pool = (unsigned char*)(b + o_pool);
        s->slack = sk_slack();
        s->pool_cap = pool;
        if (g_pool_entries_override > 0) {
            const uint64_t cap = (uint64_t)M * (uint64_t)B * sk_record_units((uint32_t)g_pool_entries_override);
            if (cap < s->pool_cap) s->pool_cap = cap;
        }
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
// Sparse STEP pass (K == 256).  At the reference's eps (0.003) a column of Q spans hundreds of log2
// units: only ~10 % of a table row lies within 2^-72 of the row's largest element (probed: 28 of 256
// on synthetic data), and what lies below cannot change an fp64 column sum (dropped mass
// <= 256 * 2^-72 = 2^-64 of the sum) nor -- as long as every centroid keeps a sane share of the mass,
// which sinkhorn_update_kernel verifies (RC_FLAG_SPARSE_UNSAFE) -- an fp64 row sum.  Per row the warp
//   1. evaluates log2 Q up to the column constant in fp32 for all 256 elements (2 x LDS.128 from the TMA
//      ring, 8 FFMA) and takes the warp maximum,
//   2. compacts the k of the elements within SK_MARGIN + 0.5 of that maximum (the 0.5 covers the fp32
//      rounding, <= 2e-4) into a shared list by ballot / popc,
//   3. evaluates those elements in fp64, one per lane: w = a + lu[k] - max, 2^w by Estrin's scheme,
//   4. normalises by the column sum and adds Q / (B * sum) to the warp's private row-sum array.
// The column scaling lv is not needed at all here: each column is normalised by its own sum, and the
// row maximum keeps 2^w in range.  Accumulation order per (warp, k) is the row order -> deterministic.
//
// Survivor lists are REUSED across iterations.  The selection pass keeps everything within
// SK_MARGIN + SK_SLACK of the row maximum and emits the (k, d~) pairs of each row into a pool (CSR: per-row
// offset / count / fp32 maximum).  Between selections only lu changes: an element left out had
// w - rowmax < -(MARGIN + SLACK) at selection time, and afterwards w - rowmax can grow by at most
// spread = max_k(dlu) - min_k(dlu), dlu = lu - lu_build.  sinkhorn_update_kernel tracks that spread per
// sub-vector; while max_m spread <= SK_SLACK the cheap pass (sinkhorn_step_list_kernel) iterates on the
// lists alone -- no table read, no filter -- otherwise the selection pass runs again.  Both kernels are
// launched every iteration and decide on the device (same inputs, same decision) which one works.
// The dense kernel above remains the path for K != 256, for RC_SINKHORN_DENSE=1 and for the re-run
// after RC_FLAG_SPARSE_UNSAFE.
// ---------------------------------------------------------------------------------------------
constexpr double SK_MARGIN = 72.0;       // log2 units: dropped mass <= K * 2^-72 = 2^-64 of a column sum


// true -> this iteration re-selects (and re-emits) the survivors; false -> it iterates on the lists
// (the decision is made once per iteration by the last block of sinkhorn_update_kernel and published as
//  one int behind the drift array: drift[2*M] reinterpreted)
__device__ __forceinline__ bool csr_reselect(const double* __restrict__ drift, int M, int force) {
    if (force) return true;
    return reinterpret_cast<const int*>(drift + 2 * M)[0] != 0;
}

constexpr double SK_UNSAFE_LOG2 = -8.0;  // a row that keeps < 2^-8 / K of mass voids the row-sum bound
constexpr int SP_CTAS_PER_SM = 3;
constexpr int SP_K = 256;
constexpr int SP_OFF_ACC = SK_WARPS * SK_DEPTH * SP_K * 4;          // ring: 32 KB
constexpr int SP_OFF_Q = SP_OFF_ACC + SK_WARPS * SP_K * 8;          // acc : 16 KB
constexpr int SP_OFF_LU = SP_OFF_Q + SK_WARPS * SP_K * 8;           // q   : 16 KB
constexpr int SP_OFF_KL = SP_OFF_LU + SP_K * 8;                     // lu  :  2 KB
constexpr int SP_OFF_BAR = SP_OFF_KL + SK_WARPS * SP_K;             // k   :  2 KB
constexpr int SP_SMEM = SP_OFF_BAR + SK_WARPS * SK_DEPTH * 8;

__global__ void __launch_bounds__(SK_THREADS, SP_CTAS_PER_SM)
sinkhorn_step_sparse_kernel(const float* __restrict__ table, int64_t B, double rBg, int M, double scale2,
                            SkPart part, const double* __restrict__ lu_g, const double* __restrict__ drift,
                            int force, SkState st, double* __restrict__ partial, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char sp_smem[];
    pdl_wait();
    pdl_launch_dependents();
    if (!csr_reselect(drift, M, force)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* my_ring = reinterpret_cast<float*>(sp_smem) + warp * SK_DEPTH * SP_K;
    double* acc_all = reinterpret_cast<double*>(sp_smem + SP_OFF_ACC);
    double* my_acc = acc_all + warp * SP_K;
    double* my_q = reinterpret_cast<double*>(sp_smem + SP_OFF_Q) + warp * SP_K;
    double* lu_s = reinterpret_cast<double*>(sp_smem + SP_OFF_LU);
    uint8_t* my_kl = sp_smem + SP_OFF_KL + warp * SP_K;
    const uint32_t my_bars = smem_u32(sp_smem + SP_OFF_BAR) + warp * SK_DEPTH * 8;
    const int g = blockIdx.x;
    const int64_t t_lo = sk_lo(part, g), t_hi = sk_lo(part, g + 1);
    if (t_lo >= t_hi) return;
    const int m_first = (int)(t_lo / part.tpm);
    int bad = 0;
    uint32_t seq = 0;
    if (lane == 0)
        for (int d = 0; d < SK_DEPTH; ++d) mbar_init(my_bars + 8 * d, 1);
    fence_barrier_init();
    __syncwarp();
    const float scale32 = (float)scale2;
    const uint32_t row_bytes = SP_K * 4u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float sel_margin = (float)(SK_MARGIN + st.slack) + 0.5f;
    unsigned long long chunk_base = 0;   // this warp's current pool chunk
    unsigned int chunk_left = 0;

    int64_t t = t_lo;
    while (t < t_hi) {
        const int m = (int)(t / part.tpm);
        const int64_t t_end = min(t_hi, (int64_t)(m + 1) * part.tpm);
        const int64_t b_first = (t - (int64_t)m * part.tpm) * SK_TILE + warp;
        const int64_t b_stop = min(B, (t_end - (int64_t)m * part.tpm) * SK_TILE);
        const int nrows = b_first < b_stop ? (int)((b_stop - b_first + SK_WARPS - 1) / SK_WARPS) : 0;
        const float* tm = table + ((int64_t)m * B + b_first) * SP_K;   // this warp's first row
        uint2* dir = st.csr + t * SK_WARPS + warp;                      // this warp's first row pair (tile t)
        uint32_t pair_ptr = 0u, pair_cnt = 0u;

        lu_s[threadIdx.x] = lu_g[(int64_t)m * SP_K + threadIdx.x];
#pragma unroll
        for (int i = 0; i < SP_K / 32; ++i) my_acc[i * 32 + lane] = 0.0;
        __syncthreads();
        // this lane's 8 columns: k = 4*lane + j (j < 4) and 128 + 4*lane + (j - 4)
        float lu32[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lu32[j] = (float)lu_s[(j >> 2) * 128 + 4 * lane + (j & 3)];

        if (lane == 0) {
            const int pre = nrows < SK_DEPTH ? nrows : SK_DEPTH;
            fence_proxy_async();
            for (int d = 0; d < pre; ++d) {
                const uint32_t slot = (seq + (uint32_t)d) % SK_DEPTH;
                bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)d * SK_WARPS * SP_K, row_bytes, my_bars + 8 * slot);
            }
        }
        for (int r = 0; r < nrows; ++r) {
            const uint32_t slot = seq % SK_DEPTH;
            mbar_wait(my_bars + 8 * slot, (seq / SK_DEPTH) & 1u);
            const float* src = my_ring + slot * SP_K;
            // 1. fp32 log2 Q (up to the column constant) and its warp maximum
            const float4 d0 = reinterpret_cast<const float4*>(src)[lane];
            const float4 d1 = reinterpret_cast<const float4*>(src)[32 + lane];
            float wf[8];
            wf[0] = fmaf(-d0.x, scale32, lu32[0]); wf[1] = fmaf(-d0.y, scale32, lu32[1]);
            wf[2] = fmaf(-d0.z, scale32, lu32[2]); wf[3] = fmaf(-d0.w, scale32, lu32[3]);
            wf[4] = fmaf(-d1.x, scale32, lu32[4]); wf[5] = fmaf(-d1.y, scale32, lu32[5]);
            wf[6] = fmaf(-d1.z, scale32, lu32[6]); wf[7] = fmaf(-d1.w, scale32, lu32[7]);
            float mx = fmaxf(fmaxf(fmaxf(wf[0], wf[1]), fmaxf(wf[2], wf[3])),
                             fmaxf(fmaxf(wf[4], wf[5]), fmaxf(wf[6], wf[7])));
            mx = warp_max(mx);
            const float cutf = mx - sel_margin;
            // 2. compaction by ballot
            int base = 0;
            uint32_t keep_mask = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool keep = wf[j] >= cutf;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    my_kl[base + __popc(bal & lt_mask)] = (uint8_t)sp_col(lane, j);
                    keep_mask |= 1u << j;
                }
                base += __popc(bal);
            }
            const int total = base;
            // record header of this lane: survivors in the lanes below (exclusive scan) and the keep mask
            uint32_t meta16;
            {
                const int c = __popc(keep_mask);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                meta16 = ((uint32_t)(incl - c) << 8) | keep_mask;
            }
            __syncwarp();
            // 3. fp64 evaluation of the survivors, shifted by the row maximum
            const double shift = (double)mx;
            double s = 0.0, q0 = 0.0;
            int k0 = 0;
            {
                const bool valid = lane < total;
                k0 = valid ? my_kl[lane] : 0;
                const double w = fma(-(double)src[k0], scale2, lu_s[k0]) - shift;
                q0 = valid ? exp2_fast_estrin(w) : 0.0;
                s = q0;
            }
            for (int bs = 32; bs < total; bs += 32) {
                const int idx = bs + lane;
                const bool valid = idx < total;
                const int k = valid ? my_kl[idx] : 0;
                const double w = fma(-(double)src[k], scale2, lu_s[k]) - shift;
                const double q = valid ? exp2_fast_estrin(w) : 0.0;
                s += q;
                my_q[idx] = q;
            }
            // 3b. emit the survivor list of this row: (k, 2^(w - rowmax)) pairs.  Pool space comes in per-warp
            //     chunks (one atomic per ~50 rows); the order of rows in the pool is irrelevant, the order inside
            //     a row is the ballot order -> deterministic sums
            {
                const unsigned int units = sk_record_units((uint32_t)total);
                // the two records of a row pair (r even, r + 1) are contiguous in the pool: only the first row of
                // a pair may open a new chunk, and it does so unless the largest possible second record fits too
                const bool first = (r & 1) == 0;           // segments are tile-aligned: local parity = global parity
                if (first && units + sk_record_units(SP_K) > chunk_left) {
                    unsigned long long nb = 0;
                    if (lane == 0) nb = atomicAdd(st.cursor, (unsigned long long)SK_POOL_CHUNK);
                    chunk_base = __shfl_sync(0xffffffffu, nb, 0);
                    chunk_left = SK_POOL_CHUNK;
                }
                const unsigned long long off = chunk_base;
                const bool fits = off + (unsigned long long)units <= st.pool_cap;
                unsigned char* rec = st.pool + (fits ? off : 0ull) * 16ull;
                double* pe = reinterpret_cast<double*>(rec + 64);
                if (fits) {
                    reinterpret_cast<uint16_t*>(rec)[lane] = (uint16_t)meta16;
                    if (lane == 0 && (total & 1)) pe[total] = 0.0;     // padding entry (copied, never used)
                } else {
                    bad |= RC_FLAG_SPARSE_UNSAFE;   // pool exhausted: the host re-runs densely
                }
                // survivor idx of the ballot order -> its place in (lane, j) order: the owner lane's header
                // gives the lane's first entry, the mask bits below j the place inside the lane
                for (int bs = 0; bs < total; bs += 32) {
                    const int idx = bs + lane;
                    const bool valid = idx < total;
                    const int k = bs == 0 ? k0 : (valid ? (int)my_kl[idx] : 0);
                    const double q = bs == 0 ? q0 : (valid ? my_q[idx] : 0.0);
                    const int j = ((k >> 7) << 2) | (k & 3);
                    const uint32_t hm = __shfl_sync(0xffffffffu, meta16, (k >> 2) & 31);
                    if (valid && fits) pe[(hm >> 8) + __popc(hm & ((1u << j) - 1u))] = q;
                }
                // pair directory: written with the first row, completed with the second
                const uint32_t cnt = fits ? (uint32_t)total : 0u;
                uint32_t word;
                if (first) {
                    pair_ptr = (uint32_t)off;
                    pair_cnt = cnt;
                    word = sk_dir_word(cnt, 0u);
                } else {
                    if (pair_cnt == 0u) pair_ptr = (uint32_t)off;   // first record absent: the pair starts here
                    word = sk_dir_word(pair_cnt, cnt);
                }
                if (lane == 0) dir[(int64_t)(r >> 1) * SK_WARPS] = make_uint2(pair_ptr, word);
                chunk_base += (unsigned long long)units;
                chunk_left -= units;
            }
            s = warp_sum(s);
            if (!(s >= 0.5) || !isfinite(s)) bad |= RC_FLAG_NONFINITE;   // the maximum itself contributes ~1
            // 4. Q / (B_global * column sum) into the row sums (:155, :162-163)
            const double rz = __drcp_rn(s) * rBg;
            if (lane < total) my_acc[k0] = fma(q0, rz, my_acc[k0]);
            for (int bs = 32; bs < total; bs += 32) {
                const int idx = bs + lane;
                if (idx < total) {
                    const int k = my_kl[idx];
                    my_acc[k] = fma(my_q[idx], rz, my_acc[k]);
                }
            }
            __syncwarp();
            if (lane == 0 && r + SK_DEPTH < nrows) {
                fence_proxy_async();
                bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)(r + SK_DEPTH) * SK_WARPS * SP_K, row_bytes,
                              my_bars + 8 * slot);
            }
            ++seq;
        }
        __syncthreads();
        {
            double* dst = partial + ((int64_t)g * part.S + (m - m_first)) * SP_K;
            double sum = acc_all[threadIdx.x];
#pragma unroll
            for (int w = 1; w < SK_WARPS; ++w) sum += acc_all[w * SP_K + threadIdx.x];
            dst[threadIdx.x] = sum;
        }
        __syncthreads();
        t = t_end;
    }
    if (bad) atomicOr(flags, bad);
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
            __syncwarp();                                    // the slot is in registers: refill it
            if (lane == 0 && r + SK_DEPTH < nrows)
                bulk_load_row(my_ring + slot * SP_K, tm + (int64_t)(r + SK_DEPTH) * SK_WARPS * SP_K, row_bytes,
                              my_bars + 8 * slot);
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

// Iteration on the survivor lists alone (see the comment above sinkhorn_step_sparse_kernel).
// A record carries E = 2^(w - rowmax) as evaluated (fp64) by the selection pass; since then only lu moved, so
// the element's current value is E * 2^(lu[k] - lu_build[k]) -- up to a per-row constant that the column
// normalisation removes.  U[k] = 2^(dlu[k] - max_k dlu) is evaluated once per iteration by
// sinkhorn_update_kernel (256 values per sub-vector).
// A warp owns a table row; lane l owns the 8 columns k(l,j) of the record layout, so its factors U[k(l,j)]
// and its 8 row-sum accumulators live in REGISTERS for a whole sub-vector segment: per survivor the pass
// issues one shared-memory load (the staged E) and two DFMAs -- no column gather, no read-modify-write of a
// shared row-sum array (the kernel this replaces did three random 8-byte shared accesses per survivor and
// sat at 78 % of the LSU wavefront limit, half of the wavefronts bank conflicts; profiles/r01_sinkhorn_list_*).
// Records come in through a per-warp ring of cp.async.bulk copies (one per row, two rows per slot, LP_DEPTH
// slots in flight, an mbarrier per slot).  Two rows are processed together so that the cross-lane reduction
// (a transposing butterfly: 5 exchanges for both rows) and the reciprocal are shared.  The row sums are
// accumulated as sum_b E * rz_b and multiplied by U[k] once per segment.  Every sum has a fixed order.
constexpr int LP_CTAS_PER_SM = 2;
constexpr int LP_DEPTH = 3;                                         // slots (row pairs) in flight per warp
constexpr int LP_SLOT = 3072;                                       // bytes; a pair that does not fit is read from global
constexpr int LP_OFF_RED = SK_WARPS * LP_DEPTH * LP_SLOT;           // ring: 64 KB
constexpr int LP_OFF_META = LP_OFF_RED + SK_WARPS * SP_K * 8;       // red : 16 KB
constexpr int LP_OFF_BAR = LP_OFF_META + SK_WARPS * LP_DEPTH * 8;   // directory entry per slot
constexpr int LP_SMEM = LP_OFF_BAR + SK_WARPS * LP_DEPTH * 8;


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

__global__ void __launch_bounds__(SK_THREADS, LP_CTAS_PER_SM)
sinkhorn_step_list_kernel(int64_t B, double rBg, int M, SkPart part, const double* __restrict__ drift, int force,
                          SkState st, double* __restrict__ partial, int32_t* __restrict__ flags) {
    extern __shared__ __align__(128) unsigned char lp_smem[];
    pdl_wait();
    pdl_launch_dependents();
    if (csr_reselect(drift, M, force)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my_ring = lp_smem + warp * LP_DEPTH * LP_SLOT;
    double* red_all = reinterpret_cast<double*>(lp_smem + LP_OFF_RED);
    uint2* my_meta = reinterpret_cast<uint2*>(lp_smem + LP_OFF_META) + warp * LP_DEPTH;
    const uint32_t my_bars = smem_u32(lp_smem + LP_OFF_BAR) + warp * LP_DEPTH * 8;
    const int g = blockIdx.x;
    const int64_t t_lo = sk_lo(part, g), t_hi = sk_lo(part, g + 1);
    if (t_lo >= t_hi) return;
    const int m_first = (int)(t_lo / part.tpm);
    int bad = 0;
    // ring positions of this warp: next slot to consume (and its mbarrier phase), next slot to fill
    uint32_t c_slot = 0, c_phase = 0, i_slot = 0;
    if (lane == 0)
        for (int d = 0; d < LP_DEPTH; ++d) mbar_init(my_bars + 8 * d, 1);
    fence_barrier_init();
    __syncwarp();

    int64_t t = t_lo;
    while (t < t_hi) {
        const int m = (int)(t / part.tpm);
        const int64_t t_end = min(t_hi, (int64_t)(m + 1) * part.tpm);
        // rows of this warp inside the segment: b = b_first + SK_WARPS * i, i < nrows (the rows warp `warp` of
        // some selection CTA wrote back to back into its pool chunk)
        const int64_t b_first = (t - (int64_t)m * part.tpm) * SK_TILE + warp;
        const int64_t b_stop = min(B, (t_end - (int64_t)m * part.tpm) * SK_TILE);
        const int nrows = b_first < b_stop ? (int)((b_stop - b_first + SK_WARPS - 1) / SK_WARPS) : 0;
        const int npairs = (nrows + 1) >> 1;
        const uint2* dir = st.csr + t * SK_WARPS + warp;                 // this warp's first row pair (tile t)
        double U[8], A[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            U[j] = st.U[(int64_t)m * SP_K + sp_col(lane, j)];
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
            const uint32_t slot = i_slot;
            if (lane == 0) {
                const uint32_t tot = sk_dir_total(dn.y);
                const uint32_t bar = my_bars + 8 * slot;
                my_meta[slot] = dn;
                const bool staged = tot <= (uint32_t)LP_SLOT;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(staged ? tot : 0u)
                             : "memory");
                if (staged && tot)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            smem_u32(my_ring + slot * LP_SLOT)),
                        "l"(st.pool + (size_t)dn.x * 16), "r"(tot), "r"(bar)
                        : "memory");
            }
            fetch_meta(pair + 1);
            i_slot = i_slot + 1 == LP_DEPTH ? 0u : i_slot + 1;
        };
        fetch_meta(0);
        for (int q = 0; q < npairs && q < LP_DEPTH; ++q) issue(q);
        __syncwarp();

        for (int q = 0; q < npairs; ++q) {
            const uint32_t slot = c_slot;
            mbar_wait(my_bars + 8 * slot, c_phase);
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
                const unsigned char* rec0 = st.pool + (size_t)mt.x * 16;
                const unsigned char* rec1 = rec0 + len0;
                if (cnt0) hm0 = reinterpret_cast<const uint16_t*>(rec0)[lane];
                if (cnt1) hm1 = reinterpret_cast<const uint16_t*>(rec1)[lane];
                s0 = lp_gather_row(rec0, hm0, U, e0);
                s1 = lp_gather_row(rec1, hm1, U, e1);
            }
            // both records are in registers: refill the slot LP_DEPTH pairs ahead before the reduction
            __syncwarp();
            if (q + LP_DEPTH < npairs) issue(q + LP_DEPTH);
            // column sums of both rows: lanes 0-15 end up with row 0's, lanes 16-31 with row 1's
            const bool hi = lane & 16;
            double a = hi ? s1 : s0;
            a += __shfl_xor_sync(0xffffffffu, hi ? s0 : s1, 16);
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
            c_slot = c_slot + 1 == LP_DEPTH ? 0u : c_slot + 1;
            c_phase ^= c_slot == 0 ? 1u : 0u;
        }
        // row sums of the segment: warp 0..7 in order (deterministic)
        {
            double* my_red = red_all + warp * SP_K;
#pragma unroll
            for (int j = 0; j < 8; ++j) my_red[sp_col(lane, j)] = U[j] * A[j];
            __syncthreads();
            double* dst = partial + ((int64_t)g * part.S + (m - m_first)) * SP_K;
            double sum = red_all[threadIdx.x];
#pragma unroll
            for (int w = 1; w < SK_WARPS; ++w) sum += red_all[w * SP_K + threadIdx.x];
            dst[threadIdx.x] = sum;
            __syncthreads();
        }
        t = t_end;
    }
    if (bad) atomicOr(flags, bad);
}

// P[m,k] = sum over the CTAs that touched sub-vector m, in CTA order (deterministic).  One block per m.
// The rows of `partial` that hold sub-vector m are located once per block (they belong to a contiguous range of
// CTAs), then every thread sums its columns over that list with independent loads.
constexpr int SK_RED_LIST = 256;
__device__ __forceinline__ void sk_reduce_block(const double* __restrict__ partial, const SkPart& part_in,
                                                const SkPart& part_list, int K, double* __restrict__ P, int csr_mode,
                                                int M, int force, const SkState& st, int* s_off) {
    const int m = blockIdx.x;
    // the partials were written by the pass that did the work: the list pass has its own grid
    const bool resel = csr_mode && csr_reselect(st.drift, M, force);
    const SkPart part = (csr_mode && !resel) ? part_list : part_in;
    if (csr_mode) {
        // the pass that just ran re-selected the survivors iff csr_reselect() says so: remember its lu
        if (resel)
            for (int k = threadIdx.x; k < K; k += blockDim.x) st.lu_build[(int64_t)m * K + k] = st.lu[(int64_t)m * K + k];
        if (m == 0 && threadIdx.x == 0) *st.cursor = 0ull;   // only live during a selection pass
    }
    const int64_t m_lo = (int64_t)m * part.tpm, m_hi = m_lo + part.tpm;
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
            s_off[i] = hit ? g * part.S + (m - (int)(lo / part.tpm)) : -1;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int k = threadIdx.x + c * 256;
            if (k < K) {
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const int o = s_off[i];
                    if (o >= 0) sum[c] += partial[(int64_t)o * K + k];
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int k = threadIdx.x + c * 256;
        if (k < K) P[(int64_t)m * K + k] = sum[c];
    }
}

__global__ void __launch_bounds__(256)
sinkhorn_reduce_kernel(const double* __restrict__ partial, SkPart part_in, SkPart part_list, int K,
                       double* __restrict__ P, int csr_mode, int M, int force, SkState st) {
    __shared__ int s_off[SK_RED_LIST];
    pdl_wait();
    pdl_launch_dependents();
    sk_reduce_block(partial, part_in, part_list, K, P, csr_mode, M, force, st, s_off);
}

// row normalisation in log2 form: lu[m,k] -= log2(K * P[m,k])     (Q /= sum_of_rows; Q /= K, :158-159)
// Also tracks, per sub-vector, how far lu has moved since the last survivor selection
// (drift[m] = {max_k, max_k - min_k} of lu - lu_build), which is what decides between the list pass and a
// new selection, and checks that every centroid kept a sane share of the mass through the last column
// normalisation (the premise of the sparse pass's row-sum bound).  One block per m.
__device__ __forceinline__ void sk_update_block(double* __restrict__ lu, const double* __restrict__ P,
                                                const double* __restrict__ lu_build, int K, double Kd,
                                                int check_mass, double slack, double* __restrict__ drift,
                                                double* __restrict__ U, int32_t* __restrict__ flags, double* red_mx,
                                                double* red_mn, double* s_dmax) {
    const int m = blockIdx.x;
    double dmax = -INFINITY, dmin = INFINITY;
    int bad = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int64_t i = (int64_t)m * K + k;
        const double z = Kd * P[i];
        if (!(z > 0.0) || !isfinite(z)) bad |= RC_FLAG_NONFINITE;
        const double dl = -log2(z);
        if (check_mass && dl > -SK_UNSAFE_LOG2) bad |= RC_FLAG_SPARSE_UNSAFE;   // K*P[k] < 2^-8
        const double nl = lu[i] + dl;
        lu[i] = nl;
        const double dv = nl - lu_build[i];
        dmax = fmax(dmax, dv);
        dmin = fmin(dmin, dv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    }
    if ((threadIdx.x & 31) == 0) { red_mx[threadIdx.x >> 5] = dmax; red_mn[threadIdx.x >> 5] = dmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            dmax = fmax(dmax, red_mx[w]);
            dmin = fmin(dmin, red_mn[w]);
        }
        drift[2 * m] = dmax;
        drift[2 * m + 1] = dmax - dmin;
        *s_dmax = dmax;
        // last block to finish publishes the iteration's decision: re-select iff any spread exceeds the slack
        // (every block has read the previous decision before it gets here)
        int* ctl = reinterpret_cast<int*>(drift + 2 * gridDim.x);
        __threadfence();
        if (atomicAdd(ctl + 1, 1) == (int)gridDim.x - 1) {
            __threadfence();
            int resel = 0;
            for (int i = 0; i < (int)gridDim.x; ++i) {
                const double sp = reinterpret_cast<volatile double*>(drift)[2 * i + 1];
                if (!(sp <= slack)) resel = 1;   // also NaN
            }
            ctl[0] = resel;
            ctl[1] = 0;
        }
    }
    __syncthreads();
    // per-column factor of the list pass: 2^(dlu - max dlu) in (0, 1]
    const double dm = *s_dmax;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int64_t i = (int64_t)m * K + k;
        U[i] = exp2(lu[i] - lu_build[i] - dm);
    }
    if (bad) atomicOr(flags, bad);
}

__global__ void __launch_bounds__(256)
sinkhorn_update_kernel(double* __restrict__ lu, const double* __restrict__ P, const double* __restrict__ lu_build,
                       int K, double Kd, int check_mass, double slack, double* __restrict__ drift,
                       double* __restrict__ U, int32_t* __restrict__ flags) {
    __shared__ double red_mx[8], red_mn[8];
    __shared__ double s_dmax;
    pdl_wait();
    pdl_launch_dependents();
    sk_update_block(lu, P, lu_build, K, Kd, check_mass, slack, drift, U, flags, red_mx, red_mn, &s_dmax);
}

// reduce + update in one launch: the single-rank solve (rc_sinkhorn_solve) has no all-reduce between them.
// Same arithmetic, same order as the two kernels above.
__global__ void __launch_bounds__(256)
sinkhorn_reduce_update_kernel(const double* __restrict__ partial, SkPart part_in, SkPart part_list, int K, int csr_mode,
                              int M, int force, int check_mass, SkState st, int32_t* __restrict__ flags) {
    __shared__ int s_off[SK_RED_LIST];
    __shared__ double red_mx[8], red_mn[8];
    __shared__ double s_dmax;
    pdl_wait();
    pdl_launch_dependents();
    sk_reduce_block(partial, part_in, part_list, K, st.P, csr_mode, M, force, st, s_off);
    __syncthreads();   // P[m,:] and lu_build[m,:] of this block are final (each element is re-read by its writer)
    sk_update_block(st.lu, st.P, st.lu_build, K, (double)K, check_mass, st.slack, st.drift, st.U, flags, red_mx,
                    red_mn, &s_dmax);
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

__global__ void list_stats_kernel(const uint2* __restrict__ csr, int64_t pairs, unsigned long long* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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

static int launch_reduce(const SkPart& p, const SkState& s, int M, int K, cudaStream_t st, int csr_mode = 0,
                         int force = 0, const SkPart* p_list = nullptr) {
    RC_CUDA(launch_chain(sinkhorn_reduce_kernel, (unsigned)M, 256u, 0, st, s.partial, p, p_list ? *p_list : p, K, s.P,
                         csr_mode, M, force, s));
    RC_CHECK_LAUNCH("sinkhorn_reduce_kernel");
    return RC_OK;
}

static int launch_update(const SkState& s, int M, int K, double Bg, int check_mass, int32_t* flags,
                         cudaStream_t st) {
    (void)Bg;
    RC_CUDA(launch_chain(sinkhorn_update_kernel, (unsigned)M, 256u, 0, st, s.lu, s.P, s.lu_build, K, (double)K, check_mass,
                         s.slack, s.drift, s.U, flags));
    RC_CHECK_LAUNCH("sinkhorn_update_kernel");
    return RC_OK;
}

// One sparse iteration (K == 256): the selection pass and the list pass are both launched; exactly one of them
// works (device-side decision from the drift the update kernel just wrote; force = 1: select).
static int launch_sparse_step(const float* table, int64_t B, int64_t B_global, int M, double eps, int force,
                              const SkState& s, int32_t* flags, cudaStream_t st, SkPart* ps_out, SkPart* pl_out) {
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_step_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
        RC_CUDA(cudaFuncSetAttribute(sinkhorn_step_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LP_SMEM));
    }
    const SkPart ps = sk_partition(B, M, SP_CTAS_PER_SM), pl = sk_partition(B, M, LP_CTAS_PER_SM);
    const double rBg = 1.0 / (double)B_global, scale2 = RC_LOG2E / eps;
    RC_CUDA(launch_chain(sinkhorn_step_sparse_kernel, ps.G, SK_THREADS, SP_SMEM, st, table, B, rBg, M, scale2, ps, s.lu,
                         s.drift, force, s, s.partial, flags));
    RC_CHECK_LAUNCH("sinkhorn_step_sparse_kernel");
    RC_CUDA(launch_chain(sinkhorn_step_list_kernel, pl.G, SK_THREADS, LP_SMEM, st, B, rBg, M, pl, s.drift, force, s,
                         s.partial, flags));
    RC_CHECK_LAUNCH("sinkhorn_step_list_kernel");
    *ps_out = ps;
    *pl_out = pl;
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
    const int64_t n = (int64_t)M * K;
    if (K == SP_K)   // pair directory: entries of rows past B stay empty
        RC_CUDA(cudaMemsetAsync(s.csr, 0, (size_t)M * ((B + SK_TILE - 1) / SK_TILE) * SK_WARPS * sizeof(uint2), st));
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.lu, n, 0.0);
    RC_CHECK_LAUNCH("fill_f64_kernel");
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.lu_build, n, 0.0);
    RC_CHECK_LAUNCH("fill_f64_kernel");
    RC_CUDA(cudaMemsetAsync(s.drift + 2 * M, 0, 16, st));
    RC_CUDA(cudaMemsetAsync(s.cursor, 0, 8, st));
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
    return launch_reduce(p, s, M, K, st);
}

/* Single-rank solve: rc_sinkhorn_begin + (iters - 1) x rc_sinkhorn_step + rc_sinkhorn_finish in one call.  With no
 * all-reduce between the row-sum reduction and the row normalisation the two run as ONE kernel per iteration
 * (sinkhorn_reduce_update_kernel); everything else is the kernels of the step-wise entry points in the same
 * order, so the results are bit-identical to the step-wise sequence with B_global == B. */
RC_API int rc_sinkhorn_solve(float* table, const float* minmax, int64_t B, int M, int K, double eps, int iters,
                             int dense, void* state, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                             void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(minmax, "rc_sinkhorn_solve: null minmax");
    RC_REQUIRE(iters >= 0, "rc_sinkhorn_solve: iters < 0");
    RC_REQUIRE(codes_mb || codes_u8, "rc_sinkhorn_solve: no output");
    RC_REQUIRE(!codes_u8 || K <= 256, "rc_sinkhorn_solve: uint8 codes need K <= 256");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    rc = sk_reset(s, B, M, K, st);
    if (rc) return rc;
    rc = launch_pass<SK_BEGIN>(table, minmax, B, (double)B, M, K, eps, p, s, nullptr, nullptr, flags, st);
    if (rc) return rc;
    if (iters >= 1 && B == 1) {
        // one column: every entry is exactly 1/K after the row normalisation (see rc_sinkhorn_finish)
        if (codes_mb) RC_CUDA(cudaMemsetAsync(codes_mb, 0, (size_t)M * B * sizeof(int64_t), st));
        if (codes_u8) RC_CUDA(cudaMemsetAsync(codes_u8, 0, (size_t)M * B, st));
        return RC_OK;
    }
    const bool sparse = sk_use_sparse(table, K, dense);
    // which pass wrote the partials the next reduce+update consumes
    int csr_mode = 0, force = 0;
    SkPart pin = p, plist = p;
    for (int it = 0; it < iters; ++it) {
        const bool last = it == iters - 1;
        RC_CUDA(launch_chain(sinkhorn_reduce_update_kernel, (unsigned)M, 256u, 0, st, s.partial, pin, plist, K, csr_mode,
                             M, force, (sparse && it > 0) ? 1 : 0, s, flags));
        RC_CHECK_LAUNCH("sinkhorn_reduce_update_kernel");
        if (last) break;
        if (sparse) {
            force = it == 0 ? 1 : 0;
            rc = launch_sparse_step(table, B, B, M, eps, force, s, flags, st, &pin, &plist);
            if (rc) return rc;
            csr_mode = 1;
        } else {
            rc = launch_pass<SK_STEP>(table, nullptr, B, (double)B, M, K, eps, p, s, nullptr, nullptr, flags, st);
            if (rc) return rc;
        }
    }
    return launch_finish(table, B, M, K, eps, p, s, codes_mb, codes_u8, flags, dense, st);
}

RC_API int rc_sinkhorn_step(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                            int step_index, int dense, void* state, int32_t* flags, void* stream) {
    int rc = sk_args(table, B, M, K, eps, state, flags);
    if (rc) return rc;
    RC_REQUIRE(B_global >= B, "rc_sinkhorn_step: B_global < B");
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    // The sparse pass needs rows (centroids) that kept their share of the mass through the previous column
    // normalisation; the update kernel checks that from the second STEP on (the first update sees the row
    // sums of the unnormalised Q0, which say nothing about it).
    const bool sparse = sk_use_sparse(table, K, dense);
    rc = launch_update(s, M, K, (double)B_global, (sparse && step_index > 0) ? 1 : 0, flags, st);
    if (rc) return rc;
    if (sparse) {
        const int force = step_index == 0 ? 1 : 0;
        SkPart ps, pl;
        rc = launch_sparse_step(table, B, B_global, M, eps, force, s, flags, st, &ps, &pl);
        if (rc) return rc;
        return launch_reduce(ps, s, M, K, st, 1, force, &pl);
    }
    rc = launch_pass<SK_STEP>(const_cast<float*>(table), nullptr, B, (double)B_global, M, K, eps, p, s, nullptr,
                              nullptr, flags, st);
    if (rc) return rc;
    return launch_reduce(p, s, M, K, st);
}

RC_API int rc_sinkhorn_list_stats(void* state, int64_t B, int M, int K, int64_t* out, void* stream) {
    RC_REQUIRE(state && out && B >= 1 && M >= 1, "rc_sinkhorn_list_stats: bad argument");
    RC_REQUIRE(K == SP_K, "rc_sinkhorn_list_stats: survivor lists exist for K == 256 only (K=%d)", K);
    cudaStream_t st = (cudaStream_t)stream;
    const SkPart p = sk_partition(B, M);
    SkState s;
    sk_layout(B, M, K, p, state, &s);
    RC_CUDA(cudaMemsetAsync(out, 0, 36 * sizeof(int64_t), st));
    const int64_t pairs = (int64_t)M * p.tpm * SK_WARPS;
    list_stats_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(s.csr, pairs, (unsigned long long*)out);
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
        rc = launch_update(s, M, K, (double)B_global, 0, flags, st);
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
        const int check = (sk_use_sparse(table, K, dense) && steps_done >= 1) ? 1 : 0;
        rc = launch_update(s, M, K, (double)B_global, check, flags, st);
        if (rc) return rc;
    }
    return launch_finish(table, B, M, K, eps, p, s, codes_mb, codes_u8, flags, dense, st);
}
