// adc.cu -- PQ asymmetric-distance (inner-product) top-k search over the coded corpus.
// Reference: src/repconc/models/repconc/evaluate_repconc.py:78-98,180-206 (Faiss IndexPQ,
// METRIC_INNER_PRODUCT, nbits = 8) and src/repconc/models/jpq/finetune_jpq.py:176.
//
// Data layout in HBM
//   codes   (N, M) uint8, row-major -- exactly the bytes of Faiss' IndexPQ.codes / the reference's
//           `add_docs` input; a document is M consecutive bytes (48 B at M = 48)
//   lut     (nq, M, 256) fp32, <x_q[m,:], c[m,k,:]>
//   scores  never materialised for the full corpus: the scan keeps, per query, only documents whose
//           (quantised) score reaches a threshold taken from a strided sample of the corpus; survivors
//           are re-scored exactly
//
// Kernels (details at each definition)
//   adc_lut_tile_kernel / adc_pack_tile_kernel   fp32 inner-product tables + per-query integer tables written in the
//                             scan's shared-memory layout (a (query tile) x (sub-vector group) grid, two launches)
//   adc_scan_u8_kernel        THE HOT LOOP: bank-conflict-free filter scan over 8-bit fields, 16 queries per 16-byte
//                             shared-memory entry (8 per 8-byte entry for M > 48), 2 or 4 lanes per document;
//                             also scans the threshold sample (integer sums out)
//   adc_scan_cf_kernel        round 1's 16-bit-field version (RC_ADC_FIELDS=16, A/B measurements)
//   adc_scan_packed_kernel    same arithmetic, thread-per-document gather (M not a multiple of 8 / 16)
//   adc_lut_kernel / adc_quantise_lut_kernel   tables for the dense path and for the gather variant
//   adc_scan_kernel           un-quantised fp32 scan with dense score output (small corpora, exact fallback,
//                             fp32 sampling for the gather variant); accumulation fp32, m ascending =
//                             bit-identical to a sequential CPU scan
//   radix_select_*            r-th largest of a sample row (integer sums / fp32 keys)
//   adc_rescore_sort_kernel   exact fp32 re-score of the filter's survivors (table in shared memory) + radix select
//                             of the k largest + bitonic sort
//   gather_topk / sort_candidates / topk_merge   tie-exact dense top-k, per-shard list merge
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rc {

constexpr int ADC_K = 256;
constexpr int SCAN_THREADS = 512;
constexpr int SEL_THREADS = 1024;
constexpr int CAND_CAP = 8192;        // candidate list capacity per query == max sortable k
constexpr int CAND_CAP32 = 65536;    // capacity of the integer filter's list (document positions, re-scored exactly)
// list capacity actually used: the 8-bit fields of wide codes keep more candidates (M = 96: ~26 k per query measured)
static int cand_cap32(int M) { return M >= 80 ? CAND_CAP32 : CAND_CAP32 / 2; }
constexpr int Q_CHUNK = 2048;         // max queries processed per pass of the host loop
constexpr int FB_ROWS = 4;            // exact-fallback queries per dense scan
constexpr int64_t DENSE_N_MAX = 262144;  // corpora up to this size take the dense path
constexpr int SAMPLE_BLK = 1024;      // sample = evenly spaced blocks of this many documents

static thread_local int64_t g_stats[4] = {0, 0, 0, 0};

// optional CUDA-event timing of the filtered corpus scan (the dominant kernel) inside rc_adc_search
static int g_timing = 0;
static thread_local double g_scan_ms = 0.0;
static thread_local int g_scan_launches = 0;
static thread_local double g_scan_wavefronts = 0.0;     // analytic LSU wavefronts of the filtered-scan launches
static thread_local const char* g_scan_kernel = "none";

// ---------------------------------------------------------------------------------------------
// LUT
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adc_lut_kernel(const float* __restrict__ queries, int64_t ldq, const float* __restrict__ c, int M, int ds,
               float* __restrict__ lut) {
    extern __shared__ float qs[];  // ds floats
    const int64_t q = blockIdx.x;
    const int m = blockIdx.y;
    for (int j = threadIdx.x; j < ds; j += blockDim.x) qs[j] = queries[q * ldq + (int64_t)m * ds + j];
    __syncthreads();
    const int k = threadIdx.x;
    const float* ck = c + ((int64_t)m * ADC_K + k) * ds;
    float s = 0.0f;
    for (int j = 0; j < ds; ++j) s = __fadd_rn(s, __fmul_rn(qs[j], __ldg(ck + j)));
    lut[(q * M + m) * ADC_K + k] = s;
}

// ---------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------
template <int QT> struct LutVec;
template <> struct LutVec<1> { using type = float; };
template <> struct LutVec<2> { using type = float2; };
template <> struct LutVec<4> { using type = float4; };

template <int QT>
__device__ __forceinline__ void lut_add(float (&acc)[QT], const typename LutVec<QT>::type& v);
template <> __device__ __forceinline__ void lut_add<1>(float (&acc)[1], const float& v) { acc[0] += v; }
template <> __device__ __forceinline__ void lut_add<2>(float (&acc)[2], const float2& v) {
    acc[0] += v.x; acc[1] += v.y;
}
template <> __device__ __forceinline__ void lut_add<4>(float (&acc)[4], const float4& v) {
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
}

struct ScanArgs {
    const float* lut;        // (nq, M, 256)
    const uint8_t* codes;    // (N, M)
    int64_t nq;
    int64_t npos;            // positions to scan (documents, or sample positions)
    int64_t pos_per_split;   // positions per blockIdx.y
    int64_t blk, stride;     // blk == 0: doc = n0 + pos ; else doc = (pos / blk) * stride + pos % blk
    int64_t n0;
    int M;
    float* out; int64_t ld_out;   // dense scores: out[q * ld_out + pos]
};

__device__ __forceinline__ unsigned long long pack_cand(float score, uint32_t doc) {
    return ((unsigned long long)f32_to_key(score) << 32) | (unsigned long long)(0xFFFFFFFFu - doc);
}

template <int QT, int MT>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
adc_scan_kernel(ScanArgs a) {
    using V = typename LutVec<QT>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* lutS = reinterpret_cast<V*>(smem_raw);  // [M][256]
    const int M = MT > 0 ? MT : a.M;
    const int64_t q0 = (int64_t)blockIdx.x * QT;
    const int nqt = (int)min((int64_t)QT, a.nq - q0);

    // stage the QT look-up tables, interleaved per (m,k)
    {
        float* lf = reinterpret_cast<float*>(smem_raw);
        const int n = M * ADC_K;
        for (int qq = 0; qq < QT; ++qq) {
            const float* src = a.lut + (q0 + qq) * (int64_t)n;
            for (int i = threadIdx.x; i < n; i += SCAN_THREADS) lf[i * QT + qq] = qq < nqt ? __ldg(src + i) : 0.0f;
        }
    }
    __syncthreads();

    const int64_t p_lo = (int64_t)blockIdx.y * a.pos_per_split;
    const int64_t p_hi = min(a.npos, p_lo + a.pos_per_split);
    for (int64_t p = p_lo + threadIdx.x; p < p_hi; p += SCAN_THREADS) {
        const int64_t doc = a.blk == 0 ? a.n0 + p : (p / a.blk) * a.stride + (p % a.blk);
        const uint8_t* cp = a.codes + doc * M;
        float acc[QT];
#pragma unroll
        for (int qq = 0; qq < QT; ++qq) acc[qq] = 0.0f;
        if constexpr (MT > 0 && MT % 16 == 0) {
            uint4 w[MT / 16];
#pragma unroll
            for (int i = 0; i < MT / 16; ++i) w[i] = __ldg(reinterpret_cast<const uint4*>(cp) + i);
#pragma unroll
            for (int i = 0; i < MT / 16; ++i) {
                const uint32_t ww[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        const int m = i * 16 + j * 4 + bb;
                        const uint32_t code = (ww[j] >> (8 * bb)) & 0xffu;
                        lut_add<QT>(acc, lutS[m * ADC_K + code]);
                    }
            }
        } else if constexpr (MT > 0 && MT % 4 == 0) {
            uint32_t w[MT / 4];
#pragma unroll
            for (int i = 0; i < MT / 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(cp) + i);
#pragma unroll
            for (int i = 0; i < MT / 4; ++i)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const uint32_t code = (w[i] >> (8 * bb)) & 0xffu;
                    lut_add<QT>(acc, lutS[(i * 4 + bb) * ADC_K + code]);
                }
        } else {
            for (int m = 0; m < M; ++m) lut_add<QT>(acc, lutS[m * ADC_K + __ldg(cp + m)]);
        }
#pragma unroll
        for (int qq = 0; qq < QT; ++qq)
            if (qq < nqt) a.out[(q0 + qq) * a.ld_out + p] = acc[qq];

    }
}

// ---------------------------------------------------------------------------------------------
// packed integer filter scan.  The corpus scan is bound by shared-memory gather throughput (one
// 16-byte gather per (document, sub-vector) per CTA pass; random codes -> ~2.4-way bank-group
// conflicts), so the filter pass packs TWICE the queries into the same 16 bytes: each query's
// table is quantised to unsigned integers  q = round((v - lo_m) / step) <= QMAX = 65535 / M  and two
// queries share one 32-bit word; a plain 32-bit IADD then accumulates both (the 16-bit fields cannot
// carry: M * QMAX <= 65535).  With |v - (lo_m + step*q)| <= step/2 a document whose exact fp32 score
// reaches the threshold tau has  sum_m q  >=  (tau - sum_m lo_m) / step - M/2, so comparing the integer
// sum against that bound (minus slack for the fp32 roundings) never loses a true top-k document.
// Survivors are re-scored exactly (fp32, m ascending) by adc_rescore_sort_kernel, so the final scores
// and ids are bit-identical to the unquantised scan.
// ---------------------------------------------------------------------------------------------
template <int QP> struct PackVec;   // QP = queries per entry
template <> struct PackVec<2> { using type = uint32_t; };
template <> struct PackVec<4> { using type = uint2; };
template <> struct PackVec<8> { using type = uint4; };

template <int QP>
__device__ __forceinline__ void pack_add(uint32_t (&acc)[QP / 2], const typename PackVec<QP>::type& v);
template <> __device__ __forceinline__ void pack_add<2>(uint32_t (&acc)[1], const uint32_t& v) { acc[0] += v; }
template <> __device__ __forceinline__ void pack_add<4>(uint32_t (&acc)[2], const uint2& v) {
    acc[0] += v.x; acc[1] += v.y;
}
template <> __device__ __forceinline__ void pack_add<8>(uint32_t (&acc)[4], const uint4& v) {
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
}

// per query: lo[m] = min_k lut, step = max_m (max_k - min_k) / QMAX, qlut = round((v - lo) / step)
// grid nq, block 256 (thread = k)
__global__ void __launch_bounds__(256)
adc_quantise_lut_kernel(const float* __restrict__ lut, int M, int qmax, uint16_t* __restrict__ qlut,
                        float* __restrict__ step_out, double* __restrict__ sumlo_out,
                        double* __restrict__ sumabs_out) {
    extern __shared__ float q_lo[];   // M floats
    __shared__ float red_lo[8], red_hi[8];
    __shared__ float s_range;
    const int64_t q = blockIdx.x;
    const int k = threadIdx.x, lane = k & 31, warp = k >> 5;
    const float* src = lut + q * (int64_t)M * ADC_K;
    float range = 0.0f;
    double sumabs = 0.0;
    for (int m = 0; m < M; ++m) {
        const float v = src[m * ADC_K + k];
        const float lo = warp_min(v), hi = warp_max(v);
        if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
        __syncthreads();
        if (k == 0) {
            float l = red_lo[0], h = red_hi[0];
            for (int w = 1; w < 8; ++w) { l = fminf(l, red_lo[w]); h = fmaxf(h, red_hi[w]); }
            q_lo[m] = l;
            range = fmaxf(range, h - l);
            sumabs += (double)fmaxf(fabsf(l), fabsf(h));
        }
        __syncthreads();
    }
    if (k == 0) s_range = range;
    __syncthreads();
    // a degenerate (constant) table still needs a positive step
    const float step = fmaxf(s_range, 1e-30f) / (float)qmax;
    for (int m = 0; m < M; ++m) {
        const float v = src[m * ADC_K + k];
        int qi = __float2int_rn((v - q_lo[m]) / step);
        qi = qi < 0 ? 0 : (qi > qmax ? qmax : qi);
        qlut[(q * M + m) * ADC_K + k] = (uint16_t)qi;
    }
    if (k == 0) {
        double s = 0.0;
        for (int m = 0; m < M; ++m) s += (double)q_lo[m];
        step_out[q] = step;
        sumlo_out[q] = s;
        sumabs_out[q] = sumabs;
    }
}

// integer threshold of the packed scan from the exact fp32 threshold tau (see the bound above):
// clamping in adc_quantise_lut_kernel only happens at the ends of the range, where it moves q towards
// the exact value, so |v - (lo + step*q)| <= step * (0.5 + 1e-3) including the fp32 division rounding.
__global__ void adc_int_threshold_kernel(const float* __restrict__ thr, const float* __restrict__ step,
                                         const double* __restrict__ sumlo, const double* __restrict__ sumabs,
                                         int M, int64_t nq, int* __restrict__ thr_i) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double t = ((double)thr[q] - sumlo[q]) / (double)step[q] - (double)M * 0.501 - 1.0;
    // fp32 accumulation error of the exact score itself: <= M * 2^-24 * max|partial sum|, and every
    // partial sum is bounded by sumabs = sum_m max_k |lut|; in units of step
    const double slack = (double)M * 1.2e-7 * sumabs[q] / (double)step[q];
    double ti = floor(t - slack);
    if (!(ti > 0.0)) ti = 0.0;            // also catches NaN: everything passes, exact re-score decides
    if (ti > 65535.0) ti = 65536.0;       // nothing can pass (cannot happen for a threshold taken from a sample)
    thr_i[q] = (int)ti;
}

struct PackScanArgs {
    const uint16_t* qlut;    // (nq, M, 256)
    const uint8_t* codes;    // (N, M)
    const int* thr_i;        // (nq)
    int64_t nq, npos, pos_per_split;
    int M;
    unsigned int* cnt;       // (nq)
    uint32_t* cand;          // (nq, cap) document positions
    int cap;
    // sample mode of adc_scan_cf_kernel<MT, true>: position p -> document (p / blk) * stride + p % blk,
    // integer sums written densely: out16[q * ld16 + p]
    int64_t blk, stride;
    uint16_t* out16;
    int64_t ld16;
    // conflict-free kernel: work items = (split, tile), handed out through item_ctr
    int64_t tiles, splits;
    unsigned int* item_ctr;
};

template <int QP, int MT>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
adc_scan_packed_kernel(PackScanArgs a) {
    using V = typename PackVec<QP>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* lutS = reinterpret_cast<V*>(smem_raw);  // [M][256], QP uint16 per entry
    const int M = MT > 0 ? MT : a.M;
    const int64_t q0 = (int64_t)blockIdx.x * QP;
    const int nqt = (int)min((int64_t)QP, a.nq - q0);
    {
        uint16_t* lh = reinterpret_cast<uint16_t*>(smem_raw);
        const int n = M * ADC_K;
        for (int qq = 0; qq < QP; ++qq) {
            const uint16_t* src = a.qlut + (q0 + qq) * (int64_t)n;
            for (int i = threadIdx.x; i < n; i += SCAN_THREADS) lh[i * QP + qq] = qq < nqt ? src[i] : (uint16_t)0;
        }
    }
    int thr[QP];
#pragma unroll
    for (int qq = 0; qq < QP; ++qq) thr[qq] = qq < nqt ? a.thr_i[q0 + qq] : 0x7fffffff;
    __syncthreads();

    const int64_t p_lo = (int64_t)blockIdx.y * a.pos_per_split;
    const int64_t p_hi = min(a.npos, p_lo + a.pos_per_split);
    for (int64_t p = p_lo + threadIdx.x; p < p_hi; p += SCAN_THREADS) {
        const uint8_t* cp = a.codes + p * M;
        uint32_t acc[QP / 2];
#pragma unroll
        for (int i = 0; i < QP / 2; ++i) acc[i] = 0u;
        if constexpr (MT > 0 && MT % 16 == 0) {
            uint4 w[MT / 16];
#pragma unroll
            for (int i = 0; i < MT / 16; ++i) w[i] = __ldg(reinterpret_cast<const uint4*>(cp) + i);
#pragma unroll
            for (int i = 0; i < MT / 16; ++i) {
                const uint32_t ww[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        const uint32_t code = (ww[j] >> (8 * bb)) & 0xffu;
                        pack_add<QP>(acc, lutS[(i * 16 + j * 4 + bb) * ADC_K + code]);
                    }
            }
        } else if constexpr (MT > 0 && MT % 4 == 0) {
            uint32_t w[MT / 4];
#pragma unroll
            for (int i = 0; i < MT / 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(cp) + i);
#pragma unroll
            for (int i = 0; i < MT / 4; ++i)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const uint32_t code = (w[i] >> (8 * bb)) & 0xffu;
                    pack_add<QP>(acc, lutS[(i * 4 + bb) * ADC_K + code]);
                }
        } else {
            for (int m = 0; m < M; ++m) pack_add<QP>(acc, lutS[m * ADC_K + __ldg(cp + m)]);
        }
#pragma unroll
        for (int qq = 0; qq < QP; ++qq) {
            const int s = (int)((acc[qq >> 1] >> (16 * (qq & 1))) & 0xffffu);
            if (s >= thr[qq]) {
                const unsigned int pos = atomicAdd(a.cnt + q0 + qq, 1u);
                if (pos < (unsigned int)a.cap) a.cand[(q0 + qq) * (int64_t)a.cap + pos] = (uint32_t)p;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused table build for the integer filter scans.  Replaces adc_lut_kernel (0.44 ms per 1200 queries: one CTA per
// (query, sub-vector), every centroid re-read per query) + adc_quantise_lut_kernel (0.09 ms) + the scan's own
// transposing stage (2-byte strided shared stores per CTA) of round 1: 0.19 ms per 1200 queries.
// ---------------------------------------------------------------------------------------------
// (declared here, defined with the 8-bit scan) table position of sub-vector m in a tile of 8-bit fields
__host__ __device__ inline int u8_pos(int m, int M, int eb, int lpd);

// Two launches over a (tile of QP queries) x (group of LP_MG sub-vectors) grid, so that a small query batch (a rank's
// share of a split batch: 150 queries = 10 tiles) still fills the machine:
//   adc_lut_tile_kernel   thread k holds centroid k: the QP inner products of each sub-vector of its group (fp32, j
//                         ascending, multiply then add: the oracle's order), the fp32 table (the exact re-score reads
//                         it) and min / max per (query, sub-vector)
//   adc_pack_tile_kernel  step / sum of minima per query (every CTA recomputes them from the M extrema, in m order),
//                         then the arithmetic of adc_quantise_lut_kernel, value for value, written as entries in the
//                         scan's shared-memory layout
constexpr int LP_MG = 4;   // sub-vectors per CTA

template <int QP>
__global__ void __launch_bounds__(256)
adc_lut_tile_kernel(const float* __restrict__ queries, int64_t ldq, const float* __restrict__ c, int64_t nq, int M,
                    int ds, float* __restrict__ lut, float* __restrict__ q_lo, float* __restrict__ q_hi) {
    extern __shared__ __align__(16) float lp_sh[];
    float* qs = lp_sh;                             // [LP_MG * ds][QP]  (the QP queries of a dimension are contiguous)
    float* wred = qs + (size_t)LP_MG * ds * QP;    // [LP_MG][8 warps][QP][2]
    const int64_t tile = blockIdx.x;
    const int m0 = blockIdx.y * LP_MG, mg = min(LP_MG, M - m0);
    const int64_t q0 = tile * QP;
    const int nqt = (int)min((int64_t)QP, nq - q0);
    const int k = threadIdx.x, lane = k & 31, warp = k >> 5;
    for (int i = k; i < QP * mg * ds; i += 256) {
        const int qq = i / (mg * ds), d = i - qq * (mg * ds);
        qs[d * QP + qq] = qq < nqt ? queries[(q0 + qq) * ldq + (int64_t)m0 * ds + d] : 0.0f;
    }
    __syncthreads();
    for (int mi = 0; mi < mg; ++mi) {
        const int m = m0 + mi;
        const float* ck = c + ((int64_t)m * ADC_K + k) * ds;
        float sacc[QP];
#pragma unroll
        for (int qq = 0; qq < QP; ++qq) sacc[qq] = 0.0f;
        for (int j = 0; j < ds; ++j) {
            const float cj = __ldg(ck + j);
            const float4* qv = reinterpret_cast<const float4*>(qs + (size_t)(mi * ds + j) * QP);
#pragma unroll
            for (int v = 0; v < QP / 4; ++v) {
                const float4 x = qv[v];
                sacc[4 * v + 0] = __fadd_rn(sacc[4 * v + 0], __fmul_rn(x.x, cj));
                sacc[4 * v + 1] = __fadd_rn(sacc[4 * v + 1], __fmul_rn(x.y, cj));
                sacc[4 * v + 2] = __fadd_rn(sacc[4 * v + 2], __fmul_rn(x.z, cj));
                sacc[4 * v + 3] = __fadd_rn(sacc[4 * v + 3], __fmul_rn(x.w, cj));
            }
        }
#pragma unroll
        for (int qq = 0; qq < QP; ++qq) {
            if (qq < nqt) lut[((q0 + qq) * M + m) * ADC_K + k] = sacc[qq];
            const float lo = warp_min(sacc[qq]), hi = warp_max(sacc[qq]);
            if (lane == 0) {
                wred[((mi * 8 + warp) * QP + qq) * 2] = lo;
                wred[((mi * 8 + warp) * QP + qq) * 2 + 1] = hi;
            }
        }
    }
    __syncthreads();
    for (int i = k; i < QP * mg; i += 256) {
        const int qq = i / mg, mi = i - qq * mg;
        float l = wred[((mi * 8) * QP + qq) * 2], h = wred[((mi * 8) * QP + qq) * 2 + 1];
        for (int w = 1; w < 8; ++w) {
            l = fminf(l, wred[((mi * 8 + w) * QP + qq) * 2]);
            h = fmaxf(h, wred[((mi * 8 + w) * QP + qq) * 2 + 1]);
        }
        q_lo[(q0 + qq) * M + m0 + mi] = l;          // (rows of whole tiles: the arrays hold tiles * QP queries)
        q_hi[(q0 + qq) * M + m0 + mi] = h;
    }
}

// U8 = false: 16-bit fields, entry = QP * 2 bytes at position m;  U8 = true: 8-bit fields, entry = QP bytes at
// position u8_pos(m) (adc_scan_u8_kernel with `lpd` lanes per document)
template <int QP, bool U8>
__global__ void __launch_bounds__(256)
adc_pack_tile_kernel(const float* __restrict__ lut, const float* __restrict__ q_lo, const float* __restrict__ q_hi,
                     int64_t nq, int M, int qmax, int lpd, uint16_t* __restrict__ qpack, float* __restrict__ step_out,
                     double* __restrict__ sumlo_out, double* __restrict__ sumabs_out) {
    __shared__ float s_step[QP];
    __shared__ float s_lo[QP][LP_MG];
    const int64_t tile = blockIdx.x;
    const int m0 = blockIdx.y * LP_MG, mg = min(LP_MG, M - m0);
    const int64_t q0 = tile * QP;
    const int nqt = (int)min((int64_t)QP, nq - q0);
    const int k = threadIdx.x;
    if (k < QP) {
        const int qq = k;
        float range = 0.0f;
        double sumabs = 0.0, sumlo = 0.0;
        for (int m = 0; m < M; ++m) {
            const float l = q_lo[(q0 + qq) * M + m], h = q_hi[(q0 + qq) * M + m];
            range = fmaxf(range, h - l);
            sumabs += (double)fmaxf(fabsf(l), fabsf(h));
            sumlo += (double)l;
        }
        // a degenerate (constant) table still needs a positive step
        const float step = fmaxf(range, 1e-30f) / (float)qmax;
        s_step[qq] = step;
        for (int mi = 0; mi < mg; ++mi) s_lo[qq][mi] = q_lo[(q0 + qq) * M + m0 + mi];
        if (qq < nqt && blockIdx.y == 0) {
            step_out[q0 + qq] = step;
            sumlo_out[q0 + qq] = sumlo;
            sumabs_out[q0 + qq] = sumabs;
        }
    }
    __syncthreads();
    constexpr int FB = U8 ? 1 : 2;              // bytes per field
    constexpr int NW = QP * FB / 4;             // words per entry
    unsigned char* tp = reinterpret_cast<unsigned char*>(qpack) + (size_t)tile * M * ADC_K * QP * FB;
    for (int mi = 0; mi < mg; ++mi) {
        const int m = m0 + mi;
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = 0u;
#pragma unroll
        for (int qq = 0; qq < QP; ++qq) {
            int qi = 0;
            if (qq < nqt) {
                const float v = lut[((q0 + qq) * M + m) * ADC_K + k];
                qi = __float2int_rn((v - s_lo[qq][mi]) / s_step[qq]);
                qi = qi < 0 ? 0 : (qi > qmax ? qmax : qi);
            }
            if constexpr (U8) w[qq >> 2] |= (uint32_t)qi << (8 * (qq & 3));
            else w[qq >> 1] |= (uint32_t)qi << (16 * (qq & 1));
        }
        const int pos = U8 ? u8_pos(m, M, QP, lpd) : m;
        uint32_t* dst = reinterpret_cast<uint32_t*>(tp + ((size_t)k * M + pos) * (QP * FB));
        if constexpr (NW == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        else *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    }
}

// ---------------------------------------------------------------------------------------------
// bank-conflict-free packed scan.  The thread-per-document kernel above gathers at random 16-byte bank
// groups: ncu shows 2.5 wavefronts per ideal wavefront and the LSU wavefront pipe at 98.8 % -- it IS the
// shared-memory crossbar limit.  Here the table is laid out [k][m] (m fastest) and the lanes that share a
// document take CONSECUTIVE sub-vectors of it: entry index = code*M + m, so the bank group of an access equals
// the lane's position among the lanes of its document whatever the codes are -- every shared-memory phase is
// conflict-free.
//   QP = 8 (M % 8 == 0, M <= 48): 16-byte entries (8 queries x uint16), 8 lanes per document (a quarter-warp),
//       4 documents per warp at a time; each lane accumulates M/8 entries, then the 8 partial 4-word
//       accumulators of a document are reduced with a halving exchange (4 shuffles instead of 12) that leaves
//       lane j of the quarter with the complete 16-bit sums of query j.
//   QP = 4 (M % 16 == 0, M <= 96 -- the reference's M = 64 and BASELINE's M = 96): 8-byte entries (4 queries),
//       16 lanes per document (a half-warp: an LDS.64 phase covers 16 lanes), 2 documents per warp at a time;
//       one halving exchange and a 3-step butterfly leave every lane of an 8-lane group with the sums of two
//       queries.  The tables of wide codes do not fit at 8 queries (M * 256 * 16 B = 256 KB at M = 64).
// Work distribution: CTAs are persistent (one per SM -- the tables fill the shared memory) and take ITEMS
// = (split of the corpus, tile of QP queries) from an atomic counter, tile fastest: at any moment the CTAs work on
// the same few-MB split of the codes (L2-resident) for different query tiles, the load is balanced to one item
// (~32k documents) whatever the shard size, and a tile's tables (already in the shared-memory layout, see
// adc_lut_pack_kernel) arrive with bulk copies (TMA), ~2 us per item.
// ---------------------------------------------------------------------------------------------
constexpr int CF_THREADS = 1024;   // 32 warps: the gathers are latency-bound at 16

template <int MT, int QP, bool SAMPLE>
__global__ void __launch_bounds__(CF_THREADS, 1)
adc_scan_cf_kernel(PackScanArgs a) {
    static_assert(QP == 8 || QP == 4, "8 or 4 queries per entry");
    constexpr int LPD = QP == 8 ? 8 : 16;                 // lanes per document
    static_assert(MT % LPD == 0, "consecutive sub-vectors per lane group");
    constexpr int NE = MT / LPD;                          // entries per lane
    constexpr int DPW = 32 / LPD;                         // documents per warp at a time
    constexpr int DOCS_PER_IT = (CF_THREADS / 32) * DPW;  // documents per CTA iteration
    constexpr uint32_t TILE_BYTES = (uint32_t)MT * ADC_K * QP * 2;
    using Entry = typename PackVec<QP>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Entry* lutS = reinterpret_cast<const Entry*>(smem_raw);  // [256][MT] entries
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + TILE_BYTES);
    int* s_item = reinterpret_cast<int*>(smem_raw + TILE_BYTES + 8);
    const uint32_t bar_a = smem_u32(bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane & (LPD - 1), sub = lane / LPD;
    if (threadIdx.x == 0) mbar_init(bar_a, 1);
    fence_barrier_init();
    __syncthreads();
    const int64_t total = a.tiles * a.splits;
    int64_t cur_tile = -1;
    uint32_t phase = 0;
    int my_thr = 0x7fffffff;
    int64_t q0 = 0;
    int nqt = 0;

    for (;;) {
        if (threadIdx.x == 0) *s_item = (int)atomicAdd(a.item_ctr, 1u);
        __syncthreads();                                  // (also: every gather of the previous item has been consumed)
        const int64_t item = *s_item;
        __syncthreads();                                  // s_item is rewritten by the next fetch
        if (item >= total) break;
        const int64_t tile = item % a.tiles, split = item / a.tiles;
        if (tile != cur_tile) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar_a, TILE_BYTES);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.qlut) + (size_t)tile * TILE_BYTES;
                for (uint32_t o = 0; o < TILE_BYTES; o += 32768u)
                    bulk_load(smem_raw + o, src + o, min(32768u, TILE_BYTES - o), bar_a);
            }
            q0 = tile * QP;
            nqt = (int)min((int64_t)QP, a.nq - q0);
            // the lane that ends up with query qi's sum: QP = 8: lane j <-> query j; QP = 4: lanes (j & 7) < 2 of
            // each 8-lane group <-> query 2 * (j >> 3) + (j & 1)
            const int qi = QP == 8 ? j : ((j & 7) < 2 ? 2 * (j >> 3) + (j & 1) : QP);
            my_thr = (!SAMPLE && qi < nqt) ? a.thr_i[q0 + qi] : 0x7fffffff;
            mbar_wait(bar_a, phase);
            phase ^= 1u;
            cur_tile = tile;
        }
        const int qi = QP == 8 ? j : ((j & 7) < 2 ? 2 * (j >> 3) + (j & 1) : QP);
        const int64_t p_lo = split * a.pos_per_split;
        const int64_t p_hi = min(a.npos, p_lo + a.pos_per_split);

        // one group of DPW documents: gather, exchange over the document's lanes, threshold test
        auto scan_group = [&](const uint32_t (&cb)[NE], int64_t p, bool live) {
            int sum;
            if constexpr (QP == 8) {
                const bool b2 = (j & 4) != 0, b1 = (j & 2) != 0;
                uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    const uint4 v = lutS[cb[i] * MT + 8 * i + j];
                    acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
                }
                const uint32_t s0 = b2 ? acc0 : acc2, s1 = b2 ? acc1 : acc3;
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
                const uint32_t k0 = (b2 ? acc2 : acc0) + r0, k1 = (b2 ? acc3 : acc1) + r1;
                const uint32_t s2 = b1 ? k0 : k1;
                const uint32_t r2 = __shfl_xor_sync(0xffffffffu, s2, 2);
                uint32_t w = (b1 ? k1 : k0) + r2;
                w += __shfl_xor_sync(0xffffffffu, w, 1);
                sum = (int)((w >> (16 * (j & 1))) & 0xffffu);
            } else {
                const bool b3 = (j & 8) != 0;
                uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    const uint2 v = lutS[cb[i] * MT + 16 * i + j];
                    acc0 += v.x; acc1 += v.y;
                }
                const uint32_t r = __shfl_xor_sync(0xffffffffu, b3 ? acc0 : acc1, 8);
                uint32_t w = (b3 ? acc1 : acc0) + r;       // lanes 0-7: queries 0,1; lanes 8-15: queries 2,3
                w += __shfl_xor_sync(0xffffffffu, w, 4);
                w += __shfl_xor_sync(0xffffffffu, w, 2);
                w += __shfl_xor_sync(0xffffffffu, w, 1);
                sum = (int)((w >> (16 * (j & 1))) & 0xffffu);
            }
            if (SAMPLE) {
                if (live && qi < nqt) a.out16[(q0 + qi) * a.ld16 + p] = (uint16_t)sum;
            } else if (live && sum >= my_thr) {
                const unsigned int pos = atomicAdd(a.cnt + q0 + qi, 1u);
                if (pos < (unsigned int)a.cap) a.cand[(q0 + qi) * (int64_t)a.cap + pos] = (uint32_t)p;
            }
        };
        // position -> first code byte of this lane (sample mode scans evenly spaced blocks of the corpus)
        auto code_ptr = [&](int64_t pos) {
            const int64_t doc = SAMPLE ? (pos / a.blk) * a.stride + (pos % a.blk) : pos;
            return a.codes + doc * MT + j;
        };
        // full iterations (every document of every warp's group is inside the split): no predicates, code bytes
        // of the next group are loaded while this one is scanned
        const int64_t full = (p_hi - p_lo) / DOCS_PER_IT;
        int64_t p = p_lo + warp * DPW + sub;
        const uint8_t* cp = code_ptr(p);
        uint32_t cb[NE], cn[NE];
        if (full > 0) {
#pragma unroll
            for (int i = 0; i < NE; ++i) cb[i] = __ldg(cp + LPD * i);
        }
        for (int64_t t = 0; t < full; ++t) {
            if (t + 1 < full) {
                const uint8_t* cq = SAMPLE ? code_ptr(p + DOCS_PER_IT) : cp + (int64_t)DOCS_PER_IT * MT;
#pragma unroll
                for (int i = 0; i < NE; ++i) cn[i] = __ldg(cq + LPD * i);
            }
            scan_group(cb, p, true);
#pragma unroll
            for (int i = 0; i < NE; ++i) cb[i] = cn[i];
            p += DOCS_PER_IT;
            cp += (int64_t)DOCS_PER_IT * MT;
        }
        // tail: fewer than DOCS_PER_IT documents left in the split
        if (p_lo + full * DOCS_PER_IT < p_hi) {   // block-uniform
            const bool live = p < p_hi;
            const uint8_t* ct = live ? code_ptr(p) : a.codes + j;
#pragma unroll
            for (int i = 0; i < NE; ++i) cb[i] = live ? (uint32_t)__ldg(ct + LPD * i) : 0u;
            scan_group(cb, p, live);
        }
    }
}

// which conflict-free variant serves this M: 8 queries per entry, 4, or none (0)
static int cf_qp(int M) {
    if (M % 8 == 0 && M >= 8 && M <= 48) return 8;
    if (M % 16 == 0 && M <= 96) return 4;
    return 0;
}
static bool cf_capable(int M) { return cf_qp(M) != 0; }

template <int MT, int QP, bool SAMPLE>
static int launch_cf_inst(const PackScanArgs& a, cudaStream_t st) {
    auto kern = adc_scan_cf_kernel<MT, QP, SAMPLE>;
    const size_t smem = (size_t)MT * ADC_K * QP * 2 + 16;
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int64_t total = a.tiles * a.splits;
    const unsigned grid = (unsigned)std::min<int64_t>(total, num_sms());
    kern<<<grid, CF_THREADS, smem, st>>>(a);
    RC_CHECK_LAUNCH("adc_scan_cf_kernel");
    return RC_OK;
}

// items: splits of ~32k documents (at least ~32 items per CTA when the corpus allows, never more than 8 MB of codes
// so that a split stays L2-resident while the query tiles sweep over it)
static int launch_cf(PackScanArgs a, bool sample, cudaStream_t st) {
    if (a.nq <= 0 || a.npos <= 0) return RC_OK;
    const int qp = cf_qp(a.M);
    const int docs_per_it = qp == 8 ? 128 : 64;
    a.tiles = (a.nq + qp - 1) / qp;
    int64_t pps = (32 * (int64_t)num_sms() + a.tiles - 1) / a.tiles;        // splits wanted
    pps = (a.npos + pps - 1) / pps;                                          // -> positions per split
    pps = std::min<int64_t>(pps, (8 << 20) / a.M);
    pps = std::max<int64_t>(pps, 16384);
    pps = (pps + docs_per_it - 1) / docs_per_it * docs_per_it;
    a.pos_per_split = pps;
    a.splits = (a.npos + pps - 1) / pps;
    if (sample) RC_CUDA(cudaMemsetAsync(a.item_ctr, 0, 4, st));
    // LSU wavefronts of this launch (what ncu counts as l1tex__data_pipe_lsu_wavefronts)
    if (!sample) {
        const double per_group = qp == 8 ? a.M / 2.0 + 6.0 : a.M / 8.0 + 5.0;   // gathers + exchange + code bytes
        g_scan_wavefronts += (double)a.tiles * ((double)a.npos / (qp == 8 ? 4.0 : 2.0)) * per_group;
        g_scan_kernel = qp == 8 ? "adc_scan_cf_kernel<8 queries/entry>" : "adc_scan_cf_kernel<4 queries/entry>";
    }
#define RC_CF(MT, QP)                                                        \
    case MT: return sample ? launch_cf_inst<MT, QP, true>(a, st) : launch_cf_inst<MT, QP, false>(a, st)
    if (qp == 8) {
        switch (a.M) {
            RC_CF(8, 8); RC_CF(16, 8); RC_CF(24, 8); RC_CF(32, 8); RC_CF(40, 8); RC_CF(48, 8);
            default: break;
        }
    } else {
        switch (a.M) {
            RC_CF(64, 4); RC_CF(80, 4); RC_CF(96, 4);
            default: break;
        }
    }
#undef RC_CF
    set_error("launch_cf: unsupported M=%d", a.M);
    return RC_E_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// 8-bit-field filter scan: TWICE the queries per shared-memory byte.  adc_scan_cf_kernel sits on the LSU wavefront
// limit (ncu: 98 % of peak, issue 53 %), so the only way up is fewer shared-memory bytes per (document, query,
// sub-vector): here a table entry carries one BYTE per query (16 queries per 16-byte entry for M <= 48, 8 queries per
// 8-byte entry for M = 64 / 80 / 96 whose tables would not fit otherwise), four queries per 32-bit word.
//   * arithmetic: ACC8 consecutive entries of a lane are added as packed bytes (q <= QMAX = 255 / ACC8, no carry
//     between the byte fields; three entries are one IADD3), then widened to two words of 16-bit fields (PRMT) and
//     accumulated there; the cross-lane reduction is the halving exchange of the 16-bit kernel.  The coarser
//     quantisation only widens the candidate set: a document is kept when its integer sum reaches
//     T = S_r - ceil(0.501 M + slack) - 1 (S_r: r-th largest sum of the sample), the exact re-score then keeps what
//     is strictly above the largest exact score an excluded document can have -- same exactness argument as before.
//   * geometry: LPD lanes per document (8 or 4), each with NE = M / LPD CONSECUTIVE code bytes (one vector load
//     instead of NE byte loads).  Sub-vector m = NE*j + i (lane j, byte i) lives at table position
//     pos = BG*(i / DPP) + LPD*(i % DPP) + j   (BG = 128 / entry bytes = entries per 128-byte wavefront,
//     DPP = BG / LPD = documents that share a shared-memory phase), and document slot s of a phase walks its bytes in
//     the order i = t ^ s, so that at every step the BG lanes of a phase hit BG different bank groups whatever the
//     codes are: conflict-free with 4 lanes per document too (half the exchange steps per document).
// Work distribution, table staging (bulk copies of tiles already written in this layout by adc_lut_pack_kernel) and
// sample mode are those of adc_scan_cf_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// table position of sub-vector m (host + device)
__host__ __device__ inline int u8_pos(int m, int M, int eb, int lpd) {
    const int bg = 128 / eb, dpp = bg / lpd, ne = M / lpd;
    const int j = m / ne, i = m % ne;
    return bg * (i / dpp) + lpd * (i % dpp) + j;
}

template <int N>
__device__ __forceinline__ void u8_halve(uint32_t* w, bool hi, int d) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const uint32_t send = hi ? w[i] : w[N / 2 + i];
        const uint32_t keep = hi ? w[N / 2 + i] : w[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
    }
}

// reduction over the lanes of a document: exchange distance D, D/2, ... 1; halve the word set while more than NWF
// words are left, plain butterflies afterwards
template <int N, int D, int NWF>
__device__ __forceinline__ void u8_reduce(uint32_t* w, int j) {
    if constexpr (D >= 1) {
        if constexpr (N > NWF) {
            u8_halve<N>(w, (j & D) != 0, D);
            u8_reduce<N / 2, D / 2, NWF>(w, j);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] += __shfl_xor_sync(0xffffffffu, w[i], D);
            u8_reduce<N, D / 2, NWF>(w, j);
        }
    }
}

template <int MT, int EB, int LPD, int ACC8, bool SAMPLE>
__global__ void __launch_bounds__(CF_THREADS, 1)
adc_scan_u8_kernel(PackScanArgs a) {
    static_assert(EB == 16 || EB == 8, "16 or 8 queries per entry");
    static_assert(LPD == 8 || LPD == 4 || LPD == 2, "8, 4 or 2 lanes per document");
    static_assert(128 / EB / LPD <= 4, "the byte permutation of a document slot stays inside a 32-bit word");
    constexpr int BG = 128 / EB, DPP = BG / LPD, NE = MT / LPD, DPW = 32 / LPD;
    static_assert(MT % BG == 0, "whole phases");
    constexpr int DOCS_PER_IT = (CF_THREADS / 32) * DPW;
    constexpr int NW8 = EB / 4, NW16 = EB / 2;
    constexpr int LOG_LPD = LPD == 8 ? 3 : LPD == 4 ? 2 : 1, LOG_NW16 = NW16 == 8 ? 3 : 2;
    constexpr int HSTEPS = LOG_LPD < LOG_NW16 ? LOG_LPD : LOG_NW16;   // halving steps, then butterflies
    constexpr int NWF = NW16 >> HSTEPS;                               // words a lane ends up with
    constexpr int RSH = LPD >> HSTEPS;                                // lanes that end up with the same word (1 or 2)
    constexpr int AL = (NE % 8 == 0) ? 8 : (NE % 4 == 0) ? 4 : (NE % 2 == 0) ? 2 : 1;   // code load width
    constexpr int NU = NE / AL, NCW = (NE + 3) / 4;
    constexpr int NUR = AL == 8 ? NE / 4 : NU;            // registers holding the loaded code bytes
    constexpr uint32_t TILE_BYTES = (uint32_t)MT * ADC_K * EB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + TILE_BYTES);
    int* s_item = reinterpret_cast<int*>(smem_raw + TILE_BYTES + 8);
    const uint32_t bar_a = smem_u32(bar);
    const uint32_t lut_a = smem_u32(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane & (LPD - 1), sub = lane / LPD, slot = sub & (DPP - 1);
    // per-lane table bases: step t reads  base[t % DPP] + 128 * (t / DPP) + code * MT * EB
    uint32_t base[DPP];
#pragma unroll
    for (int r = 0; r < DPP; ++r) base[r] = lut_a + (uint32_t)EB * (uint32_t)(LPD * (r ^ slot) + j);
    const uint32_t perm_sel = 0x3210u ^ (0x1111u * (uint32_t)slot);   // byte t of the permuted word = byte t ^ slot
    if (threadIdx.x == 0) mbar_init(bar_a, 1);
    fence_barrier_init();
    __syncthreads();
    const int64_t total = a.tiles * a.splits;
    int64_t cur_tile = -1;
    uint32_t phase = 0;
    int64_t q0 = 0;
    int nqt = 0;
    // the queries this lane tests: final word x = widx * NWF + f holds queries 4 (x >> 1) + (x & 1) [low 16 bits] and
    // + 2 [high 16 bits]; with RSH == 2 the two lanes that share a word test one field each
    const int widx = j >> (LOG_LPD - HSTEPS);
    uint32_t bias[NWF];          // packed (0x8000 - threshold) per field: field >= threshold <=> bit 15 of field + bias
#pragma unroll
    for (int f = 0; f < NWF; ++f) bias[f] = 0u;

    for (;;) {
        if (threadIdx.x == 0) *s_item = (int)atomicAdd(a.item_ctr, 1u);
        __syncthreads();                                  // (also: every gather of the previous item has been consumed)
        const int64_t item = *s_item;
        __syncthreads();                                  // s_item is rewritten by the next fetch
        if (item >= total) break;
        const int64_t tile = item % a.tiles, split = item / a.tiles;
        if (tile != cur_tile) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar_a, TILE_BYTES);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.qlut) + (size_t)tile * TILE_BYTES;
                for (uint32_t o = 0; o < TILE_BYTES; o += 32768u)
                    bulk_load(smem_raw + o, src + o, min(32768u, TILE_BYTES - o), bar_a);
            }
            q0 = tile * EB;
            nqt = (int)min((int64_t)EB, a.nq - q0);
            if (!SAMPLE) {
#pragma unroll
                for (int f = 0; f < NWF; ++f) {
                    const int x = widx * NWF + f;
                    const int qa = 4 * (x >> 1) + (x & 1), qb = qa + 2;
                    // thresholds above 0x8000 cannot be reached (M * QMAX <= 24480): 0x8000 keeps the bias in range
                    const int ta = qa < nqt ? min(a.thr_i[q0 + qa], 0x8000) : 0x8000;
                    const int tb = qb < nqt ? min(a.thr_i[q0 + qb], 0x8000) : 0x8000;
                    bias[f] = (uint32_t)(0x8000 - ta) | ((uint32_t)(0x8000 - tb) << 16);
                }
            }
            mbar_wait(bar_a, phase);
            phase ^= 1u;
            cur_tile = tile;
        }
        const int64_t p_lo = split * a.pos_per_split;
        const int64_t p_hi = min(a.npos, p_lo + a.pos_per_split);

        // one group of DPW documents: gather, exchange over the document's lanes, threshold test
        auto scan_group = [&](const uint32_t (&cu)[NUR], int64_t p, bool live) {
            // code bytes as words, in the order this document slot walks them
            uint32_t cw[NCW];
            if constexpr (AL == 8) {
#pragma unroll
                for (int i = 0; i < NCW; ++i) cw[i] = cu[i];     // (cu already holds 32-bit halves, see the loads)
            } else if constexpr (AL == 4) {
#pragma unroll
                for (int i = 0; i < NCW; ++i) cw[i] = cu[i];
            } else if constexpr (AL == 2) {
#pragma unroll
                for (int i = 0; i < NCW; ++i) cw[i] = cu[2 * i] | ((2 * i + 1 < NU ? cu[2 * i + 1] : 0u) << 16);
            } else {
#pragma unroll
                for (int i = 0; i < NCW; ++i) {
                    uint32_t w = 0u;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (4 * i + b < NU) w |= cu[4 * i + b] << (8 * b);
                    cw[i] = w;
                }
            }
            if constexpr (DPP > 1) {
#pragma unroll
                for (int i = 0; i < NCW; ++i) cw[i] = prmt(cw[i], 0u, perm_sel);
            }
            uint32_t acc16[NW16];
#pragma unroll
            for (int i = 0; i < NW16; ++i) acc16[i] = 0u;
#pragma unroll
            for (int t0 = 0; t0 < NE; t0 += ACC8) {
                uint32_t acc8[NW8];
#pragma unroll
                for (int i = 0; i < NW8; ++i) acc8[i] = 0u;
#pragma unroll
                for (int t = t0; t < t0 + ACC8 && t < NE; ++t) {
                    const uint32_t code = (cw[t >> 2] >> (8 * (t & 3))) & 0xffu;
                    const uint32_t addr = base[t % DPP] + 128u * (uint32_t)(t / DPP) + code * (uint32_t)(MT * EB);
                    if constexpr (EB == 16) {
                        const uint4 v = lds_v4(addr);
                        acc8[0] += v.x; acc8[1] += v.y; acc8[2] += v.z; acc8[3] += v.w;
                    } else {
                        const uint2 v = lds_v2(addr);
                        acc8[0] += v.x; acc8[1] += v.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < NW8; ++i) {
                    acc16[2 * i] += prmt(acc8[i], 0u, 0x4240u);       // bytes 0, 2 -> 16-bit fields
                    acc16[2 * i + 1] += prmt(acc8[i], 0u, 0x4341u);   // bytes 1, 3
                }
            }
            // reduce over the LPD lanes of the document: halve the word set while more than NWF words are left
            u8_reduce<NW16, LPD / 2, NWF>(acc16, j);
#pragma unroll
            for (int f = 0; f < NWF; ++f) {
                const uint32_t w = acc16[f];
                const int x = widx * NWF + f;
                const int qa = 4 * (x >> 1) + (x & 1), qb = qa + 2;
                if (SAMPLE) {
                    if (live) {
                        if (RSH == 1 || (j & 1) == 0) {
                            if (qa < nqt) a.out16[(q0 + qa) * a.ld16 + p] = (uint16_t)(w & 0xffffu);
                        }
                        if (RSH == 1 || (j & 1) == 1) {
                            if (qb < nqt) a.out16[(q0 + qb) * a.ld16 + p] = (uint16_t)(w >> 16);
                        }
                    }
                } else {
                    uint32_t hit = (w + bias[f]) & 0x80008000u;
                    if (RSH == 2) hit &= (j & 1) ? 0x80000000u : 0x00008000u;
                    if (hit != 0u && live) {
                        if (hit & 0x8000u) {
                            const unsigned int pos = atomicAdd(a.cnt + q0 + qa, 1u);
                            if (pos < (unsigned int)a.cap) a.cand[(q0 + qa) * (int64_t)a.cap + pos] = (uint32_t)p;
                        }
                        if (hit & 0x80000000u) {
                            const unsigned int pos = atomicAdd(a.cnt + q0 + qb, 1u);
                            if (pos < (unsigned int)a.cap) a.cand[(q0 + qb) * (int64_t)a.cap + pos] = (uint32_t)p;
                        }
                    }
                }
            }
        };
        // position -> this lane's NE code bytes (sample mode scans evenly spaced blocks of the corpus)
        auto code_ptr = [&](int64_t pos) {
            const int64_t doc = SAMPLE ? (pos / a.blk) * a.stride + (pos % a.blk) : pos;
            return a.codes + doc * MT + NE * j;
        };
        auto load_codes = [&](const uint8_t* cp, uint32_t (&cu)[NUR]) {
            if constexpr (AL == 8) {
#pragma unroll
                for (int i = 0; i < NE / 8; ++i) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(cp) + i);
                    cu[2 * i] = v.x; cu[2 * i + 1] = v.y;
                }
            } else if constexpr (AL == 4) {
#pragma unroll
                for (int i = 0; i < NU; ++i) cu[i] = __ldg(reinterpret_cast<const uint32_t*>(cp) + i);
            } else if constexpr (AL == 2) {
#pragma unroll
                for (int i = 0; i < NU; ++i) cu[i] = __ldg(reinterpret_cast<const uint16_t*>(cp) + i);
            } else {
#pragma unroll
                for (int i = 0; i < NU; ++i) cu[i] = __ldg(cp + i);
            }
        };
        const int64_t full = (p_hi - p_lo) / DOCS_PER_IT;
        int64_t p = p_lo + warp * DPW + sub;
        const uint8_t* cp = code_ptr(p);
        // the code bytes of the next group are loaded while this one is scanned -- except where their registers
        // (24 bytes per lane at M = 48 with 2 lanes per document, twice) would spill: there the 32 warps hide the load
        constexpr bool PREFETCH = NUR <= 4;
        uint32_t cb[NUR], cn[PREFETCH ? NUR : 1];
        if (PREFETCH && full > 0) load_codes(cp, cb);
        for (int64_t t = 0; t < full; ++t) {
            if constexpr (PREFETCH) {
                if (t + 1 < full)
                    load_codes(SAMPLE ? code_ptr(p + DOCS_PER_IT) : cp + (int64_t)DOCS_PER_IT * MT,
                               reinterpret_cast<uint32_t (&)[NUR]>(cn));
                scan_group(cb, p, true);
#pragma unroll
                for (int i = 0; i < NUR; ++i) cb[i] = cn[PREFETCH ? i : 0];
            } else {
                load_codes(SAMPLE ? code_ptr(p) : cp, cb);
                scan_group(cb, p, true);
            }
            p += DOCS_PER_IT;
            cp += (int64_t)DOCS_PER_IT * MT;
        }
        // tail: fewer than DOCS_PER_IT documents left in the split
        if (p_lo + full * DOCS_PER_IT < p_hi) {   // block-uniform
            const bool live = p < p_hi;
            if (live) {
                load_codes(code_ptr(p), cb);
            } else {
#pragma unroll
                for (int i = 0; i < NUR; ++i) cb[i] = 0u;
            }
            scan_group(cb, p, live);
        }
    }
}

// which 8-bit-field variant serves this M: entry bytes = queries per tile (16, 8) or 0 (none)
static int u8_eb(int M) {
    if (M % 8 == 0 && M >= 8 && M <= 48) return 16;
    if (M % 16 == 0 && M <= 96) return 8;
    return 0;
}
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
// lanes per document and entries added in 8-bit fields before widening (QMAX = 255 / acc); RC_ADC_U8_LPD / RC_ADC_U8_ACC
// select another compiled variant for A/B measurements
struct U8Cfg { int eb, lpd, acc; };
static U8Cfg u8_cfg(int M) {
    // 16-byte entries: 2 lanes per document (measured at M = 48: 17.2 ms vs 18.2 with 4 lanes, 24.3 with 8);
    // 8-byte entries: 4 lanes (2 would put 8 documents in a phase: the byte permutation no longer fits a word)
    U8Cfg c{u8_eb(M), 2, 2};
    if (c.eb == 8) c.lpd = 4;
    static const int e_lpd = env_int("RC_ADC_U8_LPD", 0), e_acc = env_int("RC_ADC_U8_ACC", 0);
    if (M == 48 || M == 32 || M == 64 || M == 96) {
        if (e_lpd) c.lpd = e_lpd;
        if (e_acc) c.acc = e_acc;
    }
    return c;
}
// RC_ADC_FIELDS=16 keeps the 16-bit-field scan (adc_scan_cf_kernel) for A/B measurements
static bool u8_enabled(int M) {
    static const int fields = env_int("RC_ADC_FIELDS", 8);
    return fields != 16 && u8_eb(M) != 0;
}

template <int MT, int EB, int LPD, int ACC8, bool SAMPLE>
static int launch_u8_inst(const PackScanArgs& a, cudaStream_t st) {
    auto kern = adc_scan_u8_kernel<MT, EB, LPD, ACC8, SAMPLE>;
    const size_t smem = (size_t)MT * ADC_K * EB + 16;
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int64_t total = a.tiles * a.splits;
    const unsigned grid = (unsigned)std::min<int64_t>(total, num_sms());
    kern<<<grid, CF_THREADS, smem, st>>>(a);
    RC_CHECK_LAUNCH("adc_scan_u8_kernel");
    return RC_OK;
}

static int launch_u8(PackScanArgs a, bool sample, cudaStream_t st) {
    if (a.nq <= 0 || a.npos <= 0) return RC_OK;
    const U8Cfg c = u8_cfg(a.M);
    const int docs_per_it = (CF_THREADS / 32) * (32 / c.lpd);
    a.tiles = (a.nq + c.eb - 1) / c.eb;
    int64_t pps = (32 * (int64_t)num_sms() + a.tiles - 1) / a.tiles;        // splits wanted
    pps = (a.npos + pps - 1) / pps;                                          // -> positions per split
    pps = std::min<int64_t>(pps, (8 << 20) / a.M);
    pps = std::max<int64_t>(pps, 16384);
    pps = (pps + docs_per_it - 1) / docs_per_it * docs_per_it;
    a.pos_per_split = pps;
    a.splits = (a.npos + pps - 1) / pps;
    if (sample) RC_CUDA(cudaMemsetAsync(a.item_ctr, 0, 4, st));
    if (!sample) {
        // LSU wavefronts per group of (32 / lpd) documents x eb queries: gathers (eb / 4 per warp-wide load),
        // exchange shuffles, code loads
        const int ne = a.M / c.lpd;
        const int shuffles = c.eb == 16 ? (c.lpd == 8 ? 7 : c.lpd == 4 ? 6 : 4) : (c.lpd == 8 ? 4 : 3);   // exchange steps
        const int al = ne % 8 == 0 ? 8 : ne % 4 == 0 ? 4 : ne % 2 == 0 ? 2 : 1;
        const double per_group = (double)ne * (c.eb / 4) + shuffles + (double)(ne / al);
        g_scan_wavefronts += (double)a.tiles * ((double)a.npos / (32 / c.lpd)) * per_group;
        g_scan_kernel = c.eb == 16 ? "adc_scan_u8_kernel<16 queries/entry>" : "adc_scan_u8_kernel<8 queries/entry>";
    }
#define RC_U8(MT, EB, LPD, ACC)                                              \
    case (MT * 100 + LPD * 10 + ACC):                                        \
        return sample ? launch_u8_inst<MT, EB, LPD, ACC, true>(a, st) : launch_u8_inst<MT, EB, LPD, ACC, false>(a, st)
    switch (a.M * 100 + c.lpd * 10 + c.acc) {
        RC_U8(8, 16, 2, 2); RC_U8(16, 16, 2, 2); RC_U8(24, 16, 2, 2); RC_U8(40, 16, 2, 2); RC_U8(80, 8, 4, 2);
        RC_U8(32, 16, 2, 2); RC_U8(32, 16, 4, 2);
        RC_U8(48, 16, 2, 2); RC_U8(48, 16, 4, 2); RC_U8(48, 16, 4, 3); RC_U8(48, 16, 8, 2);
        RC_U8(64, 8, 4, 2); RC_U8(64, 8, 8, 2);
        RC_U8(96, 8, 4, 2); RC_U8(96, 8, 8, 2);
        default: break;
    }
#undef RC_U8
    set_error("launch_u8: unsupported M=%d lpd=%d acc=%d", a.M, c.lpd, c.acc);
    return RC_E_UNSUPPORTED;
}

// threshold of the 8-bit scan from the r-th largest sample sum S_r: T = S_r - ceil(0.501 M + slack) - 1, so that
// about `target` documents have an exact score above what an excluded document (sum <= T - 1) can reach (the
// bound adc_rescore_sort_kernel applies)
__global__ void adc_u8_threshold_kernel(const float* __restrict__ step, const double* __restrict__ sumabs, int M,
                                        int64_t nq, int* __restrict__ thr_i) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double slack = (double)M * 1.2e-7 * sumabs[q] / (double)step[q] + 1.0;
    double sh = ceil(0.501 * (double)M + slack) + 1.0;
    if (!(sh < 65536.0)) sh = 65536.0;        // NaN / degenerate tables: everything passes, the exact re-score decides
    const int t = thr_i[q] - (int)sh;
    thr_i[q] = t < 0 ? 0 : t;
}

// r-th largest 16-bit integer sum of a sample row (2 radix passes); the filter threshold of the packed scan
__global__ void __launch_bounds__(SEL_THREADS)
radix_select_u16_kernel(const uint16_t* __restrict__ dense, int64_t ld, int64_t n, int rank, int* __restrict__ thr_i) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_hi, s_rem;
    const int64_t q = blockIdx.x;
    const uint16_t* row = dense + q * ld;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += SEL_THREADS) atomicAdd(&hist[row[i] >> 8], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 255;
        for (; b > 0; --b) {
            if (cum + hist[b] >= (unsigned int)rank) break;
            cum += hist[b];
        }
        s_hi = (unsigned int)b;
        s_rem = (unsigned int)rank - cum;
    }
    __syncthreads();
    const unsigned int hi = s_hi;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += SEL_THREADS) {
        const unsigned int v = row[i];
        if ((v >> 8) == hi) atomicAdd(&hist[v & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int cum = 0;
        int b = 255;
        for (; b > 0; --b) {
            if (cum + hist[b] >= s_rem) break;
            cum += hist[b];
        }
        thr_i[q] = (int)((hi << 8) | (unsigned int)b);
    }
}

static int scan_qp(int M) {
    const size_t per_q = (size_t)M * ADC_K * 2;
    const size_t budget = 200 * 1024;
    if (8 * per_q <= budget) return 8;
    if (4 * per_q <= budget) return 4;
    if (2 * per_q <= budget) return 2;
    return 0;
}

template <int QP, int MT>
static int launch_packed_inst(const PackScanArgs& a, int splits, cudaStream_t st) {
    auto kern = adc_scan_packed_kernel<QP, MT>;
    const size_t smem = (size_t)a.M * ADC_K * 2 * QP;
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    dim3 grid((unsigned)((a.nq + QP - 1) / QP), (unsigned)splits);
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    RC_CHECK_LAUNCH("adc_scan_packed_kernel");
    return RC_OK;
}

static void scan_splits(int64_t nq, int qt, int64_t npos, int M, int64_t* pps_out, int64_t* splits_out) {
    // split the positions so that (a) the grid fills the machine and (b) one split's codes (<= ~8 MB) stay
    // L2-resident while every query tile sweeps over them
    const int64_t tiles = (nq + qt - 1) / qt;
    int64_t pps = (8 << 20) / M;
    const int64_t want = (4 * (int64_t)num_sms() + tiles - 1) / tiles;   // >= 4 CTAs per SM in total
    if (want > 1) pps = std::min(pps, (npos + want - 1) / want);
    pps = std::max<int64_t>(pps, SCAN_THREADS);
    pps = (pps + SCAN_THREADS - 1) / SCAN_THREADS * SCAN_THREADS;
    int64_t splits = (npos + pps - 1) / pps;
    if (splits > 65535) {
        pps = ((npos + 65534) / 65535 + SCAN_THREADS - 1) / SCAN_THREADS * SCAN_THREADS;
        splits = (npos + pps - 1) / pps;
    }
    *pps_out = pps;
    *splits_out = splits;
}

// RC_ADC_GATHER=1 selects the thread-per-document packed scan (for A/B measurements)
static bool adc_force_gather() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RC_ADC_GATHER");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

static int launch_packed(PackScanArgs a, cudaStream_t st) {
    if (a.nq <= 0 || a.npos <= 0) return RC_OK;
    const int qp = scan_qp(a.M);
    if (qp == 0) {
        set_error("adc packed scan: M=%d too large", a.M);
        return RC_E_UNSUPPORTED;
    }
    int64_t pps, splits;
    scan_splits(a.nq, qp, a.npos, a.M, &pps, &splits);
    a.pos_per_split = pps;
#define RC_PSCAN(QP, MT) return launch_packed_inst<QP, MT>(a, (int)splits, st)
    // LSU wavefronts: one 16 / 8 / 4-byte gather per (document, sub-vector) per tile at the measured 2.5 conflict
    // wavefronts per ideal wavefront
    g_scan_wavefronts += (double)((a.nq + qp - 1) / qp) * ((double)a.npos / 32.0) * a.M * (qp * 2 / 4.0) * 2.5;
    g_scan_kernel = "adc_scan_packed_kernel";
    if (qp == 8) {
        switch (a.M) {
            case 8: RC_PSCAN(8, 8);
            case 16: RC_PSCAN(8, 16);
            case 24: RC_PSCAN(8, 24);
            case 32: RC_PSCAN(8, 32);
            case 48: RC_PSCAN(8, 48);
            default: RC_PSCAN(8, 0);
        }
    } else if (qp == 4) {
        switch (a.M) {
            case 64: RC_PSCAN(4, 64);
            case 96: RC_PSCAN(4, 96);
            default: RC_PSCAN(4, 0);
        }
    }
    RC_PSCAN(2, 0);
#undef RC_PSCAN
}

static int scan_qt(int M) {
    const size_t per_q = (size_t)M * ADC_K * 4;
    const size_t budget = 200 * 1024;
    if (4 * per_q <= budget) return 4;
    if (2 * per_q <= budget) return 2;
    if (per_q <= 220 * 1024) return 1;
    return 0;
}

template <int QT, int MT>
static int launch_scan_inst(const ScanArgs& a, int splits, cudaStream_t st) {
    auto kern = adc_scan_kernel<QT, MT>;
    const size_t smem = (size_t)a.M * ADC_K * 4 * QT;
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    dim3 grid((unsigned)((a.nq + QT - 1) / QT), (unsigned)splits);
    kern<<<grid, SCAN_THREADS, smem, st>>>(a);
    RC_CHECK_LAUNCH("adc_scan_kernel");
    return RC_OK;
}

static int launch_scan(ScanArgs a, cudaStream_t st) {
    if (a.nq <= 0 || a.npos <= 0) return RC_OK;
    const int qt = scan_qt(a.M);
    if (qt == 0) {
        set_error("adc scan: M=%d too large for the shared-memory look-up table", a.M);
        return RC_E_UNSUPPORTED;
    }
    // split the positions so that (a) the grid fills the machine and (b) one split's codes
    // (<= ~8 MB) stay L2-resident while every query tile sweeps over them
    const int64_t tiles = (a.nq + qt - 1) / qt;
    int64_t pps = (8 << 20) / a.M;
    const int64_t want = (4 * (int64_t)num_sms() + tiles - 1) / tiles;   // >= 4 CTAs per SM in total
    if (want > 1) pps = std::min(pps, (a.npos + want - 1) / want);
    pps = std::max<int64_t>(pps, SCAN_THREADS);
    pps = (pps + SCAN_THREADS - 1) / SCAN_THREADS * SCAN_THREADS;
    int64_t splits = (a.npos + pps - 1) / pps;
    if (splits > 65535) {
        pps = ((a.npos + 65534) / 65535 + SCAN_THREADS - 1) / SCAN_THREADS * SCAN_THREADS;
        splits = (a.npos + pps - 1) / pps;
    }
    a.pos_per_split = pps;
#define RC_SCAN(QT, MT) return launch_scan_inst<QT, MT>(a, (int)splits, st)
    if (qt == 4) {
        switch (a.M) {
            case 8: RC_SCAN(4, 8);
            case 16: RC_SCAN(4, 16);
            case 24: RC_SCAN(4, 24);
            case 32: RC_SCAN(4, 32);
            case 48: RC_SCAN(4, 48);
            default: RC_SCAN(4, 0);
        }
    } else if (qt == 2) {
        switch (a.M) {
            case 64: RC_SCAN(2, 64);
            case 96: RC_SCAN(2, 96);
            default: RC_SCAN(2, 0);
        }
    }
    RC_SCAN(1, 0);
#undef RC_SCAN
}

// ---------------------------------------------------------------------------------------------
// radix select on dense score rows: key of the r-th largest score and #scores strictly above it
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS)
radix_select_kernel(const float* __restrict__ dense, int64_t ld, int64_t n, const int* __restrict__ rank_of_q,
                    int rank_all, uint32_t* __restrict__ thr_key, float* __restrict__ thr_f,
                    unsigned int* __restrict__ count_gt) {
    __shared__ unsigned int hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ unsigned int s_remaining;
    const int64_t q = blockIdx.x;
    const float* row = dense + q * ld;
    const unsigned int r = (unsigned int)(rank_of_q ? rank_of_q[q] : rank_all);  // 1-based from the top
    if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_remaining = r; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int64_t i = threadIdx.x; i < n; i += SEL_THREADS) {
            const uint32_t key = f32_to_key(row[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int rem = s_remaining, cum = 0;
            int b = 255;
            for (; b > 0; --b) {
                if (cum + hist[b] >= rem) break;
                cum += hist[b];
            }
            s_remaining = rem - cum;
            s_prefix = prefix | ((uint32_t)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (thr_key) thr_key[q] = s_prefix;
        if (thr_f) thr_f[q] = key_to_f32(s_prefix);
        if (count_gt) count_gt[q] = r - s_remaining;
    }
}

// exact gather of the top-k of a dense row into the candidate list: everything above the k-th key,
// plus the (k - count_gt) smallest positions among the ties at the k-th key.
__global__ void __launch_bounds__(SEL_THREADS)
gather_topk_kernel(const float* __restrict__ dense, int64_t ld, int64_t n, int k, const uint32_t* __restrict__ thr_key,
                   const unsigned int* __restrict__ count_gt, unsigned long long* __restrict__ cand, int cap,
                   unsigned int* __restrict__ cnt) {
    __shared__ unsigned int s_cnt;
    __shared__ unsigned int warp_eq[SEL_THREADS / 32];
    const int64_t q = blockIdx.x;
    const float* row = dense + q * ld;
    const uint32_t tk = thr_key[q];
    const unsigned int need_eq = (unsigned int)k - count_gt[q];
    unsigned long long* out = cand + q * (int64_t)cap;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned int eq_base = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += SEL_THREADS) {
        const int64_t i = base + threadIdx.x;
        const bool in = i < n;
        const float s = in ? row[i] : 0.0f;
        const uint32_t key = f32_to_key(s);
        if (in && key > tk) {
            const unsigned int pos = atomicAdd(&s_cnt, 1u);
            if (pos < (unsigned int)cap) out[pos] = pack_cand(s, (uint32_t)i);
        }
        if (eq_base < need_eq) {  // block-uniform
            const bool is_eq = in && key == tk;
            const unsigned int bal = __ballot_sync(0xffffffffu, is_eq);
            if (lane == 0) warp_eq[warp] = __popc(bal);
            __syncthreads();
            unsigned int before = 0, total = 0;
            for (int w = 0; w < SEL_THREADS / 32; ++w) {
                const unsigned int c = warp_eq[w];
                if (w < warp) before += c;
                total += c;
            }
            if (is_eq) {
                const unsigned int rank = eq_base + before + __popc(bal & ((1u << lane) - 1u));
                if (rank < need_eq) {
                    const unsigned int pos = atomicAdd(&s_cnt, 1u);
                    if (pos < (unsigned int)cap) out[pos] = pack_cand(s, (uint32_t)i);
                }
            }
            eq_base += total;
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cnt[q] = s_cnt;
}

// ---------------------------------------------------------------------------------------------
// per-query bitonic sort of the candidate list, descending composite key; writes the top k
// status[q] = 0 ok, 1 = too few candidates (< k_eff), 2 = candidate list overflowed
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SEL_THREADS)
sort_candidates_kernel(const unsigned long long* __restrict__ cand, int cap, const unsigned int* __restrict__ cnt,
                       int k, int k_eff, int64_t id_offset, float* __restrict__ scores, int64_t ld_s,
                       int64_t* __restrict__ ids, int64_t ld_i, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    const int64_t q = blockIdx.x;
    const unsigned int c = cnt[q];
    if (c < (unsigned int)k_eff || c > (unsigned int)cap) {
        if (threadIdx.x == 0) status[q] = c > (unsigned int)cap ? 2 : 1;
        return;
    }
    int n = 2;
    while (n < (int)c) n <<= 1;
    const unsigned long long* src = cand + q * (int64_t)cap;
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) skeys[i] = i < (int)c ? src[i] : 0ull;
    bitonic_sort_desc(skeys, n);
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
        if (i < k_eff) {
            const unsigned long long key = skeys[i];
            scores[q * ld_s + i] = key_to_f32((uint32_t)(key >> 32));
            ids[q * ld_i + i] = id_offset + (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        } else {
            scores[q * ld_s + i] = -FLT_MAX;   // Faiss pads missing results with (lowest, -1)
            ids[q * ld_i + i] = -1;
        }
    }
    if (threadIdx.x == 0) status[q] = 0;
}

constexpr int RS_THREADS = 1024;    // re-score + select + sort: one CTA per query
constexpr int RS_SEL_MAX = 2048;    // largest k the selection buffer holds (larger k: the filtered path is not used)
// auxiliary shared memory behind the keys: the fp32 table while re-scoring, then selection buffer + histograms
constexpr size_t RS_AUX_BYTES = (size_t)RS_SEL_MAX * 8 + (RS_THREADS / 32) * 256 * 4 + 256 * 4;
// exact fp32 scores (m ascending) of the candidates of one query, NC candidates per thread in flight; rows are read
// as NV 8-byte words (M a multiple of 8, M <= 8 NV)
template <int NV, int NC, typename Keep>
__device__ __forceinline__ void rescore_rows(const uint32_t* __restrict__ src, unsigned int n,
                                             const uint8_t* __restrict__ codes, int M, const float* t, Keep keep) {
    const int nv = M >> 3;
    for (unsigned int i = threadIdx.x; i < n; i += NC * RS_THREADS) {
        uint32_t doc[NC];
        uint2 a[NC][NV];
#pragma unroll
        for (int c = 0; c < NC; ++c) doc[c] = i + c * RS_THREADS < n ? src[i + c * RS_THREADS] : src[i];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const uint2* r = reinterpret_cast<const uint2*>(codes + (int64_t)doc[c] * M);
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (v < nv) a[c][v] = __ldg(r + v);
        }
        float sc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) sc[c] = 0.0f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            if (v < nv) {
                const float* tv = t + v * 8 * ADC_K;
#pragma unroll
                for (int bb = 0; bb < 8; ++bb)
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint32_t w = bb < 4 ? a[c][v].x : a[c][v].y;
                        sc[c] += tv[bb * ADC_K + ((w >> (8 * (bb & 3))) & 0xffu)];
                    }
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (i + c * RS_THREADS < n) keep(sc[c], doc[c]);
    }
}

// The k largest of n distinct 64-bit keys in shared memory -> aux[0 .. k) (unordered); returns that buffer.
// Radix select from the highest bit in which the keys differ, 8 bits per pass, per-warp histograms (keys that share
// their leading bits would serialise on one counter), the bucket search by warp 0.  Ends as soon as a bucket holds
// exactly the number of keys still needed.  All threads of the block call it; ends with a block barrier.
__device__ __forceinline__ unsigned long long* rs_select_topk(const unsigned long long* keys, int n, int k,
                                                              unsigned char* aux) {
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(aux);
    unsigned int* hist = reinterpret_cast<unsigned int*>(aux + (size_t)RS_SEL_MAX * 8);      // [warps][256]
    unsigned int* tot = hist + (RS_THREADS / 32) * 256;                                       // [256]
    __shared__ unsigned long long s_min, s_max, s_prefix, s_mask;
    __shared__ unsigned int s_need, s_done, s_out;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_min = ~0ull; s_max = 0ull; s_out = 0u; }
    __syncthreads();
    {
        unsigned long long lo = ~0ull, hi = 0ull;
        for (int i = threadIdx.x; i < n; i += RS_THREADS) {
            const unsigned long long v = keys[i];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = a < lo ? a : lo;
            hi = b > hi ? b : hi;
        }
        if (lane == 0) { atomicMin(&s_min, lo); atomicMax(&s_max, hi); }
    }
    __syncthreads();
    const unsigned long long diff = s_min ^ s_max;          // != 0: the keys are distinct and n > k >= 1
    int low = 64 - __clzll((long long)diff);                // bits [low, 64) are common to all keys
    if (threadIdx.x == 0) {
        s_mask = low >= 64 ? 0ull : ~((1ull << low) - 1ull);
        s_prefix = s_max & s_mask;
        s_need = (unsigned int)k;
        s_done = 0u;
    }
    __syncthreads();
    while (low > 0) {
        const int w = low < 8 ? low : 8, shift = low - w;
        for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix, mask = s_mask;
        for (int i = threadIdx.x; i < n; i += RS_THREADS) {
            const unsigned long long v = keys[i];
            if ((v & mask) == prefix) atomicAdd(&hist[warp * 256 + (int)((v >> shift) & ((1u << w) - 1u))], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            unsigned int t = 0u;
#pragma unroll 8
            for (int ww = 0; ww < RS_THREADS / 32; ++ww) t += hist[ww * 256 + threadIdx.x];
            tot[threadIdx.x] = t;
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns buckets [8 l, 8 l + 8); above[l] = keys in the buckets of the lanes above
            unsigned int mine = 0u;
#pragma unroll
            for (int b = 0; b < 8; ++b) mine += tot[8 * lane + b];
            unsigned int above = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_down_sync(0xffffffffu, above, o);
                if (lane + o < 32) above += v;
            }
            above -= mine;                                   // exclusive suffix sum
            const unsigned int need = s_need;
            __syncwarp();                                    // every lane has read s_need before one of them rewrites it
            if (above < need && above + mine >= need) {      // exactly one lane
                unsigned int cum = above;
                int b = 7;
                for (; b > 0; --b) {
                    if (cum + tot[8 * lane + b] >= need) break;
                    cum += tot[8 * lane + b];
                }
                const unsigned int in_bucket = tot[8 * lane + b];
                s_prefix = prefix | ((unsigned long long)(8 * lane + b) << shift);
                s_mask = mask | ((unsigned long long)((1u << w) - 1u) << shift);
                s_need = need - cum;
                s_done = (in_bucket == need - cum) ? 1u : 0u;
            }
        }
        __syncthreads();
        low = shift;
        if (s_done) break;
    }
    // everything >= the k-th key: s_prefix with the undecided low bits zero (when the loop ended early the whole
    // bucket belongs to the result; when it ran to bit 0 the prefix IS the k-th key)
    const unsigned long long kth = s_prefix;
    for (int i = threadIdx.x; i < n; i += RS_THREADS) {
        const unsigned long long v = keys[i];
        if (v >= kth) {
            const unsigned int pos = atomicAdd(&s_out, 1u);
            if (pos < (unsigned int)RS_SEL_MAX) sel[pos] = v;
        }
    }
    __syncthreads();
    return sel;
}

// exact re-scoring of the integer filter's survivors + sort.  One CTA per query: the query's fp32 table is staged
// in shared memory (when it fits), every candidate document is scored in fp32, m ascending (= a sequential CPU
// scan, bit for bit), kept if it can belong to the result, and the kept ones are sorted like
// sort_candidates_kernel does.
//   thr != NULL : exact fp32 threshold tau known (fp32 sampling) -> keep score >= tau
//   thr == NULL : only the integer threshold T is known (integer-domain sampling).  A document the scan left out
//                 has integer sum <= T - 1, hence (quantisation bound, see adc_int_threshold_kernel) exact score
//                 <= ub = sumlo + step * (T - 1 + 0.501 M + slack): keep what is STRICTLY above ub; the result is
//                 exact iff at least k documents are kept.
// status: 0 ok, 1 fewer than k_eff kept, 2 the approximate list overflowed, 3 too many kept for the sort
__global__ void __launch_bounds__(RS_THREADS, 1)
adc_rescore_sort_kernel(const float* __restrict__ lut, const uint8_t* __restrict__ codes, int M, int lut_in_smem,
                        const float* __restrict__ thr, const int* __restrict__ thr_i,
                        const float* __restrict__ qstep, const double* __restrict__ qsumlo,
                        const double* __restrict__ qsumabs, const uint32_t* __restrict__ cand, int cap,
                        const unsigned int* __restrict__ cnt, int k, int k_eff, int64_t id_offset,
                        float* __restrict__ scores, int64_t ld_s, int64_t* __restrict__ ids, int64_t ld_i,
                        int* __restrict__ status, unsigned int* __restrict__ exact_cnt) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    __shared__ unsigned int s_cnt;
    const int64_t q = blockIdx.x;
    const unsigned int c_approx = cnt[q];
    if (c_approx > (unsigned int)cap) {
        if (threadIdx.x == 0) { status[q] = 2; exact_cnt[q] = c_approx; }
        return;
    }
    if (threadIdx.x == 0) s_cnt = 0;
    const float* t = lut + q * (int64_t)M * ADC_K;
    if (lut_in_smem) {
        float* ts = reinterpret_cast<float*>(skeys + CAND_CAP);
        const float4* src4 = reinterpret_cast<const float4*>(t);
        for (int i = threadIdx.x; i < M * (ADC_K / 4); i += RS_THREADS) reinterpret_cast<float4*>(ts)[i] = __ldg(src4 + i);
        t = ts;
    }
    __syncthreads();
    const float tau = thr ? thr[q] : -FLT_MAX;
    double ub = -DBL_MAX;
    if (!thr) {
        const double st = (double)qstep[q];
        const double slack = (double)M * 1.2e-7 * qsumabs[q] / st + 1.0;
        ub = qsumlo[q] + st * ((double)thr_i[q] - 1.0 + (double)M * 0.501 + slack);
        if (!(ub == ub)) {                   // NaN bound (non-finite table): exact fallback
            if (threadIdx.x == 0) { status[q] = 1; exact_cnt[q] = 0; }
            return;
        }
    }
    const uint32_t* src = cand + q * (int64_t)cap;
    auto keep = [&](float sc, uint32_t doc) {
        if (thr ? (sc >= tau) : ((double)sc > ub)) {
            const unsigned int pos = atomicAdd(&s_cnt, 1u);
            if (pos < (unsigned int)CAND_CAP) skeys[pos] = pack_cand(sc, doc);
        }
    };
    if ((M & 7) == 0 && M <= 96 && ((uintptr_t)codes & 7) == 0) {
        // the usual widths: the whole code row of a candidate (two candidates for M <= 48) is requested before
        // anything is used -- the loop is bound by the latency of these scattered reads -- then scored m ascending
        if (M <= 48) rescore_rows<6, 2>(src, c_approx, codes, M, t, keep);
        else rescore_rows<12, 1>(src, c_approx, codes, M, t, keep);
    } else {
        for (unsigned int i = threadIdx.x; i < c_approx; i += RS_THREADS) {
            const uint32_t doc = src[i];
            const uint8_t* cp = codes + (int64_t)doc * M;
            float sc = 0.0f;
            if ((M & 3) == 0) {
                for (int w = 0; w < M / 4; ++w) {
                    const uint32_t ww = __ldg(reinterpret_cast<const uint32_t*>(cp) + w);
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) sc += t[(w * 4 + bb) * ADC_K + ((ww >> (8 * bb)) & 0xffu)];
                }
            } else {
                for (int m = 0; m < M; ++m) sc += t[m * ADC_K + __ldg(cp + m)];
            }
            keep(sc, doc);
        }
    }
    __syncthreads();
    const unsigned int c = s_cnt;
    if (threadIdx.x == 0) exact_cnt[q] = c;
    if (c < (unsigned int)k_eff || c > (unsigned int)CAND_CAP) {
        if (threadIdx.x == 0) status[q] = c > (unsigned int)CAND_CAP ? 3 : 1;
        return;
    }
    // the k_eff largest of the c kept keys (all different: the document is part of the key) are selected with a radix
    // select and only those are sorted: sorting all ~3 k_eff kept keys cost 6-13x the compare-exchanges
    unsigned long long* sorted = skeys;
    int n = 2;
    if (c > (unsigned int)k_eff && k_eff <= RS_SEL_MAX) {
        sorted = rs_select_topk(skeys, (int)c, k_eff, reinterpret_cast<unsigned char*>(skeys + CAND_CAP));
        while (n < k_eff) n <<= 1;
        for (int i = k_eff + threadIdx.x; i < n; i += RS_THREADS) sorted[i] = 0ull;
    } else {
        while (n < (int)c) n <<= 1;
        for (int i = (int)c + threadIdx.x; i < n; i += RS_THREADS) skeys[i] = 0ull;
    }
    bitonic_sort_desc(sorted, n);
    for (int i = threadIdx.x; i < k; i += RS_THREADS) {
        if (i < k_eff) {
            const unsigned long long key = sorted[i];
            scores[q * ld_s + i] = key_to_f32((uint32_t)(key >> 32));
            ids[q * ld_i + i] = id_offset + (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        } else {
            scores[q * ld_s + i] = -FLT_MAX;
            ids[q * ld_i + i] = -1;
        }
    }
    if (threadIdx.x == 0) status[q] = 0;
}

// merge W sorted lists per query: (W, nq, k) -> (nq, k).  ids must be in [0, 2^32) or -1 (padding).
__global__ void __launch_bounds__(SEL_THREADS)
topk_merge_kernel(const float* __restrict__ s_in, const int64_t* __restrict__ i_in, int W, int64_t nq, int k,
                  int n_sort, float* __restrict__ scores, int64_t* __restrict__ ids) {
    extern __shared__ __align__(16) unsigned long long skeys[];
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < n_sort; i += SEL_THREADS) {
        unsigned long long key = 0ull;
        if (i < W * k) {
            const int w = i / k, j = i - w * k;
            const int64_t id = i_in[((int64_t)w * nq + q) * k + j];
            if (id >= 0) key = pack_cand(s_in[((int64_t)w * nq + q) * k + j], (uint32_t)id);
        }
        skeys[i] = key;
    }
    bitonic_sort_desc(skeys, n_sort);
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
        const unsigned long long key = skeys[i];
        if (key != 0ull) {
            scores[q * k + i] = key_to_f32((uint32_t)(key >> 32));
            ids[q * k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        } else {
            scores[q * k + i] = -FLT_MAX;
            ids[q * k + i] = -1;
        }
    }
}

__global__ void fill_pad_kernel(float* scores, int64_t* ids, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { scores[i] = -FLT_MAX; ids[i] = -1; }
}

// ---------------------------------------------------------------------------------------------
// host plan + orchestration
// ---------------------------------------------------------------------------------------------
struct AdcPlan {
    bool dense_all;
    int k_eff;
    int64_t n_sample;     // sample positions (multiple of SAMPLE_BLK)
    int64_t stride;       // distance between sample block starts
    int rank_sample;      // rank of the threshold inside the sample
    int64_t dense_floats; // size of the dense score buffer
    int dense_rows;       // query rows per dense-all pass
};

static AdcPlan adc_plan(int64_t nq, int64_t N, int64_t k) {
    AdcPlan p{};
    p.k_eff = (int)std::min<int64_t>(k, N);
    p.dense_all = N <= DENSE_N_MAX || p.k_eff > 2048;
    const int64_t qc = std::min<int64_t>(nq, Q_CHUNK);
    if (p.dense_all) {
        // rows per pass bounded by ~1 GiB of scores
        int64_t rows = std::max<int64_t>(FB_ROWS, std::min<int64_t>(qc, (1ll << 28) / std::max<int64_t>(N, 1)));
        rows = (rows + 3) / 4 * 4;
        p.dense_rows = (int)rows;
        p.dense_floats = rows * N;
        return p;
    }
    const int64_t target = std::min<int64_t>(CAND_CAP / 2, std::max<int64_t>(3 * (int64_t)p.k_eff, 2048));
    int64_t ns = std::max<int64_t>(32768, (32 * N + target - 1) / target);   // threshold = ~32nd largest sample sum
    ns = (ns + SAMPLE_BLK - 1) / SAMPLE_BLK * SAMPLE_BLK;
    const int64_t nblk = ns / SAMPLE_BLK;
    p.n_sample = ns;
    p.stride = N / nblk;                      // >= SAMPLE_BLK because N > DENSE_N_MAX >= 4 * ns is not
    if (p.stride < SAMPLE_BLK) p.stride = SAMPLE_BLK;  // guaranteed; clamp keeps blocks disjoint
    while (nblk * p.stride - (p.stride - SAMPLE_BLK) > N) --p.stride;  // last block must end inside N
    p.rank_sample = (int)std::max<int64_t>(1, (target * ns + N - 1) / N);
    p.dense_rows = FB_ROWS;
    p.dense_floats = std::max<int64_t>(qc * ns, (int64_t)FB_ROWS * N);
    return p;
}

struct AdcWs {
    float* lut; float* dense; unsigned long long* cand; unsigned int* cnt; float* thr; uint32_t* thr_key;
    unsigned int* count_gt; int* status; float* fb_lut; float* fb_scores; int64_t* fb_ids;
    uint16_t* qlut; float* qstep; double* qsumlo; double* qsumabs; int* thr_i; uint32_t* cand32;
    unsigned int* exact_cnt;
    unsigned int* item_ctr;      // 2 work-item counters (sample scan, corpus scan)
    float* q_lo; float* q_hi;    // (whole tiles of queries, M) extrema of the fp32 tables
};

static size_t adc_ws_layout(int64_t nq, int64_t N, int M, int64_t k, const AdcPlan& p, void* base, AdcWs* w) {
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t n) {
        size_t o = off;
        off = align_up(off + n, 256);
        return o;
    };
    const int64_t qc = std::max<int64_t>(std::min<int64_t>(nq, Q_CHUNK), FB_ROWS);
    const size_t o_lut = take((size_t)qc * M * ADC_K * 4);
    const size_t o_dense = take((size_t)p.dense_floats * 4);
    const size_t o_cand = take((size_t)qc * CAND_CAP * 8);
    const size_t o_cnt = take((size_t)qc * 4);
    const size_t o_thr = take((size_t)qc * 4);
    const size_t o_tk = take((size_t)qc * 4);
    const size_t o_gt = take((size_t)qc * 4);
    const size_t o_st = take((size_t)qc * 4);
    const size_t o_fl = take((size_t)FB_ROWS * M * ADC_K * 4);
    const size_t o_fs = take((size_t)FB_ROWS * std::max<int64_t>(k, 1) * 4);
    const size_t o_fi = take((size_t)FB_ROWS * std::max<int64_t>(k, 1) * 8);
    const size_t o_ql = take((size_t)((qc + 15) / 16 * 16) * M * ADC_K * 2);   // whole tiles of 16 / 8 / 4 queries
    const size_t o_qs = take((size_t)qc * 4);
    const size_t o_qo = take((size_t)qc * 8);
    const size_t o_qa = take((size_t)qc * 8);
    const size_t o_ti = take((size_t)qc * 4);
    const size_t o_c32 = take((size_t)qc * (u8_eb(M) ? cand_cap32(M) : CAND_CAP) * 4);
    const size_t o_ec = take((size_t)qc * 4);
    const size_t o_ic = take(16);
    const size_t o_lo = take((size_t)((qc + 15) / 16 * 16) * M * 4);
    const size_t o_hi = take((size_t)((qc + 15) / 16 * 16) * M * 4);
    if (w) {
        w->item_ctr = (unsigned int*)(b + o_ic);
        w->q_lo = (float*)(b + o_lo);
        w->q_hi = (float*)(b + o_hi);
        w->qlut = (uint16_t*)(b + o_ql);
        w->qstep = (float*)(b + o_qs);
        w->qsumlo = (double*)(b + o_qo);
        w->qsumabs = (double*)(b + o_qa);
        w->thr_i = (int*)(b + o_ti);
        w->cand32 = (uint32_t*)(b + o_c32);
        w->exact_cnt = (unsigned int*)(b + o_ec);
        w->lut = (float*)(b + o_lut);
        w->dense = (float*)(b + o_dense);
        w->cand = (unsigned long long*)(b + o_cand);
        w->cnt = (unsigned int*)(b + o_cnt);
        w->thr = (float*)(b + o_thr);
        w->thr_key = (uint32_t*)(b + o_tk);
        w->count_gt = (unsigned int*)(b + o_gt);
        w->status = (int*)(b + o_st);
        w->fb_lut = (float*)(b + o_fl);
        w->fb_scores = (float*)(b + o_fs);
        w->fb_ids = (int64_t*)(b + o_fi);
    }
    return off;
}

static int launch_lut(const float* queries, int64_t ldq, const float* c, int64_t nq, int M, int ds, float* lut,
                      cudaStream_t st) {
    if (nq <= 0) return RC_OK;
    for (int64_t q0 = 0; q0 < nq; q0 += 32768) {  // grid.x is fine up to 2^31, keep launches modest anyway
        const int64_t n = std::min<int64_t>(32768, nq - q0);
        dim3 grid((unsigned)n, (unsigned)M);
        adc_lut_kernel<<<grid, ADC_K, (size_t)ds * 4, st>>>(queries + q0 * ldq, ldq, c, M, ds,
                                                           lut + q0 * (int64_t)M * ADC_K);
        RC_CHECK_LAUNCH("adc_lut_kernel");
    }
    return RC_OK;
}

static int sort_smem_attr() {
    static unsigned long long attr_seen = 0ull;
    if (first_use_on_device(attr_seen)) {
        RC_CUDA(cudaFuncSetAttribute(sort_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CAND_CAP * 8));
        RC_CUDA(cudaFuncSetAttribute(adc_rescore_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RC_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CAND_CAP * 8));
    }
    return RC_OK;
}

// exact top-k of `rows` dense rows (each covering all N documents) into (scores, ids)
static int dense_topk(const AdcWs& w, int rows, int64_t N, int k, int k_eff, int64_t id_offset, float* scores,
                      int64_t ld_s, int64_t* ids, int64_t ld_i, cudaStream_t st) {
    radix_select_kernel<<<rows, SEL_THREADS, 0, st>>>(w.dense, N, N, nullptr, k_eff, w.thr_key, nullptr, w.count_gt);
    RC_CHECK_LAUNCH("radix_select_kernel");
    gather_topk_kernel<<<rows, SEL_THREADS, 0, st>>>(w.dense, N, N, k_eff, w.thr_key, w.count_gt, w.cand, CAND_CAP,
                                                     w.cnt);
    RC_CHECK_LAUNCH("gather_topk_kernel");
    sort_candidates_kernel<<<rows, SEL_THREADS, CAND_CAP * 8, st>>>(w.cand, CAND_CAP, w.cnt, k, k_eff, id_offset,
                                                                   scores, ld_s, ids, ld_i, w.status);
    RC_CHECK_LAUNCH("sort_candidates_kernel");
    return RC_OK;
}


}  // namespace rc

using namespace rc;

RC_API void rc_adc_last_stats(int64_t out4[4]) {
    for (int i = 0; i < 4; ++i) out4[i] = g_stats[i];
}

RC_API void rc_adc_enable_timing(int enable) { g_timing = enable ? 1 : 0; }
RC_API double rc_adc_last_scan_ms(void) { return g_scan_ms; }
RC_API int rc_adc_last_scan_launches(void) { return g_scan_launches; }
RC_API double rc_adc_last_scan_wavefronts(void) { return g_scan_wavefronts; }
RC_API const char* rc_adc_last_scan_kernel(void) { return g_scan_kernel; }

RC_API int rc_adc_lut(const float* queries, int64_t ldq, const float* centroids, int64_t nq, int M, int K, int ds,
                      float* lut, void* stream) {
    RC_REQUIRE(queries && centroids && lut, "rc_adc_lut: null pointer");
    RC_REQUIRE(K == ADC_K, "rc_adc_lut: K must be 256 (8-bit codes), got %d", K);
    RC_REQUIRE(nq >= 0 && M >= 1 && M <= 65535 && ds >= 1 && ldq >= (int64_t)M * ds, "rc_adc_lut: bad shape");
    return launch_lut(queries, ldq, centroids, nq, M, ds, lut, (cudaStream_t)stream);
}

RC_API int rc_adc_scores(const float* lut, const uint8_t* codes, int64_t nq, int64_t n0, int64_t n, int M, float* out,
                         void* stream) {
    RC_REQUIRE(lut && codes && out, "rc_adc_scores: null pointer");
    RC_REQUIRE(nq >= 0 && n0 >= 0 && n >= 0 && M >= 1, "rc_adc_scores: bad shape");
    ScanArgs a{};
    a.lut = lut; a.codes = codes; a.nq = nq; a.npos = n; a.blk = 0; a.stride = 0; a.n0 = n0; a.M = M;
    a.out = out; a.ld_out = n;
    return launch_scan(a, (cudaStream_t)stream);
}

RC_API size_t rc_adc_search_workspace_bytes(int64_t nq, int64_t N, int M, int K, int64_t k) {
    if (nq < 1 || N < 1 || M < 1 || k < 1 || K != ADC_K) return 256;
    const AdcPlan p = adc_plan(nq, N, k);
    return adc_ws_layout(nq, N, M, k, p, nullptr, nullptr) + 256;
}

RC_API int rc_adc_search(const float* queries, int64_t ldq, const float* centroids, const uint8_t* codes, int64_t nq,
                         int64_t N, int M, int K, int ds, int64_t k, int64_t id_offset, float* scores, int64_t* ids,
                         void* workspace, size_t workspace_bytes, void* stream) {
    RC_REQUIRE(K == ADC_K, "rc_adc_search: K must be 256 (8-bit codes), got %d", K);
    RC_REQUIRE(nq >= 0 && N >= 0 && M >= 1 && ds >= 1 && k >= 1, "rc_adc_search: bad shape");
    RC_REQUIRE(N < (1ll << 32), "rc_adc_search: N must be < 2^32 per shard");
    if (nq == 0) return RC_OK;
    RC_REQUIRE(queries && centroids && scores && ids, "rc_adc_search: null pointer");
    RC_REQUIRE(ldq >= (int64_t)M * ds, "rc_adc_search: bad query stride");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < 4; ++i) g_stats[i] = 0;
    if (N == 0) {
        const int64_t n = nq * k;
        fill_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, ids, n);
        RC_CHECK_LAUNCH("fill_pad_kernel");
        return RC_OK;
    }
    RC_REQUIRE(codes && workspace, "rc_adc_search: null pointer");
    RC_REQUIRE(((uintptr_t)workspace & 255) == 0, "rc_adc_search: workspace must be 256-byte aligned");
    const AdcPlan p = adc_plan(nq, N, k);
    if (p.k_eff > CAND_CAP) {
        set_error("rc_adc_search: k=%lld > %d is not supported", (long long)k, CAND_CAP);
        return RC_E_UNSUPPORTED;
    }
    if (scan_qt(M) == 0 || scan_qp(M) == 0) {
        set_error("rc_adc_search: M=%d too large", M);
        return RC_E_UNSUPPORTED;
    }
    AdcWs w;
    const size_t need = adc_ws_layout(nq, N, M, k, p, workspace, &w);
    if (need > workspace_bytes) {
        set_error("rc_adc_search: workspace %zu < %zu bytes", workspace_bytes, need);
        return RC_E_WORKSPACE;
    }
    int rc = sort_smem_attr();
    if (rc) return rc;
    const int ik = (int)k;
    g_stats[3] = p.dense_all ? 0 : p.n_sample;
    std::vector<int> status_h;
    std::vector<unsigned int> cnt_h;
    g_scan_ms = 0.0;
    g_scan_launches = 0;
    g_scan_wavefronts = 0.0;
    g_scan_kernel = "none";
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_timing) {
        RC_CUDA(cudaEventCreate(&ev0));
        RC_CUDA(cudaEventCreate(&ev1));
    }

    // balanced chunks (a multiple of 16 queries: whole CTA tiles of the packed scans)
    const int64_t n_chunks = (nq + Q_CHUNK - 1) / Q_CHUNK;
    const int64_t chunk = std::min<int64_t>(Q_CHUNK, ((nq + n_chunks - 1) / n_chunks + 15) / 16 * 16);
    const bool codes_aligned = ((uintptr_t)codes & 7) == 0;     // the 8-bit scan loads a lane's code bytes as one vector
    for (int64_t c0 = 0; c0 < nq; c0 += chunk) {
        const int64_t qc = std::min<int64_t>(chunk, nq - c0);
        const bool int_sampling = !p.dense_all && cf_capable(M) && !adc_force_gather();
        const bool fields8 = int_sampling && u8_enabled(M) && codes_aligned;
        const U8Cfg u8 = u8_cfg(M);
        const int qmax = fields8 ? 255 / u8.acc : 65535 / M;
        const int cand_cap = fields8 ? cand_cap32(M) : CAND_CAP;
        if (int_sampling) {
            // fp32 tables + integer tables in the scan's shared-memory layout
            const int qp = fields8 ? u8.eb : cf_qp(M);
            const unsigned tiles = (unsigned)((qc + qp - 1) / qp);
            const dim3 grid(tiles, (unsigned)((M + LP_MG - 1) / LP_MG));
            const size_t sm = ((size_t)LP_MG * ds * qp + (size_t)LP_MG * 8 * qp * 2) * 4;
            if (sm > 48 * 1024) {
                set_error("rc_adc_search: sub-vector dimension %d too large for the table builder", ds);
                return RC_E_UNSUPPORTED;
            }
            if (qp == 16) adc_lut_tile_kernel<16><<<grid, 256, sm, st>>>(queries + c0 * ldq, ldq, centroids, qc, M, ds, w.lut, w.q_lo, w.q_hi);
            else if (qp == 8) adc_lut_tile_kernel<8><<<grid, 256, sm, st>>>(queries + c0 * ldq, ldq, centroids, qc, M, ds, w.lut, w.q_lo, w.q_hi);
            else adc_lut_tile_kernel<4><<<grid, 256, sm, st>>>(queries + c0 * ldq, ldq, centroids, qc, M, ds, w.lut, w.q_lo, w.q_hi);
            RC_CHECK_LAUNCH("adc_lut_tile_kernel");
#define RC_PACK(QP, U8)                                                                                            \
    adc_pack_tile_kernel<QP, U8><<<grid, 256, 0, st>>>(w.lut, w.q_lo, w.q_hi, qc, M, qmax, u8.lpd, w.qlut, w.qstep, \
                                                       w.qsumlo, w.qsumabs)
            if (fields8 && qp == 16) RC_PACK(16, true);
            else if (fields8) RC_PACK(8, true);
            else if (qp == 8) RC_PACK(8, false);
            else RC_PACK(4, false);
#undef RC_PACK
            RC_CHECK_LAUNCH("adc_pack_tile_kernel");
        } else {
            rc = launch_lut(queries + c0 * ldq, ldq, centroids, qc, M, ds, w.lut, st);
            if (rc) return rc;
        }
        if (p.dense_all) {
            for (int64_t r0 = 0; r0 < qc; r0 += p.dense_rows) {
                const int rows = (int)std::min<int64_t>(p.dense_rows, qc - r0);
                ScanArgs a{};
                a.lut = w.lut + r0 * (int64_t)M * ADC_K; a.codes = codes; a.nq = rows; a.npos = N; a.n0 = 0; a.M = M;
                a.out = w.dense; a.ld_out = N;
                rc = launch_scan(a, st);
                if (rc) return rc;
                rc = dense_topk(w, rows, N, ik, p.k_eff, id_offset, scores + (c0 + r0) * k, k, ids + (c0 + r0) * k, k, st);
                if (rc) return rc;
            }
            g_stats[1] += qc;
            continue;
        }
        if (!int_sampling) {
            adc_quantise_lut_kernel<<<(unsigned)qc, ADC_K, (size_t)M * 4, st>>>(w.lut, M, qmax, w.qlut, w.qstep, w.qsumlo,
                                                                              w.qsumabs);
            RC_CHECK_LAUNCH("adc_quantise_lut_kernel");
        }
        // 1. thresholds from a strided sample of the corpus
        if (int_sampling) {
            // integer domain: the sample is scanned by the packed kernel itself and the threshold is the
            // r-th largest 16-bit sum (exactness is verified after the re-score, adc_rescore_sort_kernel)
            PackScanArgs a{};
            a.qlut = w.qlut; a.codes = codes; a.nq = qc; a.npos = p.n_sample; a.M = M;
            a.blk = SAMPLE_BLK; a.stride = p.stride; a.out16 = reinterpret_cast<uint16_t*>(w.dense);
            a.ld16 = p.n_sample;
            a.item_ctr = w.item_ctr;
            rc = fields8 ? launch_u8(a, true, st) : launch_cf(a, true, st);
            if (rc) return rc;
            radix_select_u16_kernel<<<(unsigned)qc, SEL_THREADS, 0, st>>>(reinterpret_cast<uint16_t*>(w.dense),
                                                                           p.n_sample, p.n_sample, p.rank_sample,
                                                                           w.thr_i);
            RC_CHECK_LAUNCH("radix_select_u16_kernel");
            if (fields8) {
                adc_u8_threshold_kernel<<<(unsigned)((qc + 255) / 256), 256, 0, st>>>(w.qstep, w.qsumabs, M, qc, w.thr_i);
                RC_CHECK_LAUNCH("adc_u8_threshold_kernel");
            }
        } else {
            ScanArgs a{};
            a.lut = w.lut; a.codes = codes; a.nq = qc; a.npos = p.n_sample; a.blk = SAMPLE_BLK; a.stride = p.stride;
            a.M = M; a.out = w.dense; a.ld_out = p.n_sample;
            rc = launch_scan(a, st);
            if (rc) return rc;
            radix_select_kernel<<<(unsigned)qc, SEL_THREADS, 0, st>>>(w.dense, p.n_sample, p.n_sample, nullptr,
                                                                       p.rank_sample, nullptr, w.thr, nullptr);
            RC_CHECK_LAUNCH("radix_select_kernel");
            adc_int_threshold_kernel<<<(unsigned)((qc + 255) / 256), 256, 0, st>>>(w.thr, w.qstep, w.qsumlo,
                                                                                   w.qsumabs, M, qc, w.thr_i);
            RC_CHECK_LAUNCH("adc_int_threshold_kernel");
        }
        // 2. packed integer filter scan of the whole corpus
        RC_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)qc * 4, st));
        {
            PackScanArgs a{};
            a.qlut = w.qlut; a.codes = codes; a.thr_i = w.thr_i; a.nq = qc; a.npos = N; a.M = M;
            a.cnt = w.cnt; a.cand = w.cand32; a.cap = cand_cap;
            a.item_ctr = w.item_ctr + 1;
            if (int_sampling) RC_CUDA(cudaMemsetAsync(a.item_ctr, 0, 4, st));   // (outside the timed region)
            if (ev0) RC_CUDA(cudaEventRecord(ev0, st));
            rc = fields8 ? launch_u8(a, false, st) : int_sampling ? launch_cf(a, false, st) : launch_packed(a, st);
            if (rc) return rc;
            if (ev1) RC_CUDA(cudaEventRecord(ev1, st));
        }
        // 3. exact re-score of the survivors + per-query sort
        const size_t lut_bytes = (size_t)M * ADC_K * 4;
        const int lut_in_smem = (size_t)CAND_CAP * 8 + lut_bytes <= 200 * 1024 ? 1 : 0;
        const size_t rs_smem = (size_t)CAND_CAP * 8 + std::max(lut_in_smem ? lut_bytes : (size_t)0, RS_AUX_BYTES);
        adc_rescore_sort_kernel<<<(unsigned)qc, RS_THREADS, rs_smem, st>>>(
            w.lut, codes, M, lut_in_smem, int_sampling ? nullptr : w.thr, w.thr_i, w.qstep, w.qsumlo, w.qsumabs, w.cand32,
            cand_cap, w.cnt, ik, p.k_eff, id_offset, scores + c0 * k, k, ids + c0 * k, k, w.status, w.exact_cnt);
        RC_CHECK_LAUNCH("adc_rescore_sort_kernel");
        // 4. queries whose list under/overflowed take the exact dense path
        status_h.resize(qc);
        cnt_h.resize(qc);
        RC_CUDA(cudaMemcpyAsync(status_h.data(), w.status, (size_t)qc * 4, cudaMemcpyDeviceToHost, st));
        RC_CUDA(cudaMemcpyAsync(cnt_h.data(), w.cnt, (size_t)qc * 4, cudaMemcpyDeviceToHost, st));
        RC_CUDA(cudaStreamSynchronize(st));
        if (ev0) {
            float ms = 0.0f;
            RC_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            g_scan_ms += ms;
            g_scan_launches += 1;
        }
        std::vector<int64_t> failed;
        for (int64_t i = 0; i < qc; ++i) {
            if (status_h[i] != 0) failed.push_back(i);
            g_stats[2] = std::max<int64_t>(g_stats[2], cnt_h[i]);
        }
        g_stats[0] += qc - (int64_t)failed.size();
        g_stats[1] += (int64_t)failed.size();
        for (size_t f0 = 0; f0 < failed.size(); f0 += FB_ROWS) {
            const int rows = (int)std::min<size_t>(FB_ROWS, failed.size() - f0);
            for (int r = 0; r < rows; ++r)
                RC_CUDA(cudaMemcpyAsync(w.fb_lut + (int64_t)r * M * ADC_K, w.lut + failed[f0 + r] * (int64_t)M * ADC_K,
                                        (size_t)M * ADC_K * 4, cudaMemcpyDeviceToDevice, st));
            ScanArgs a{};
            a.lut = w.fb_lut; a.codes = codes; a.nq = rows; a.npos = N; a.n0 = 0; a.M = M;
            a.out = w.dense; a.ld_out = N;
            rc = launch_scan(a, st);
            if (rc) return rc;
            rc = dense_topk(w, rows, N, ik, p.k_eff, id_offset, w.fb_scores, k, w.fb_ids, k, st);
            if (rc) return rc;
            for (int r = 0; r < rows; ++r) {
                const int64_t q = c0 + failed[f0 + r];
                RC_CUDA(cudaMemcpyAsync(scores + q * k, w.fb_scores + (int64_t)r * k, (size_t)k * 4,
                                        cudaMemcpyDeviceToDevice, st));
                RC_CUDA(cudaMemcpyAsync(ids + q * k, w.fb_ids + (int64_t)r * k, (size_t)k * 8,
                                        cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    RC_CUDA(cudaStreamSynchronize(st));
    if (ev0) {
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    return RC_OK;
}

// external ids: out[i] = corpus_ids[idx[i]] (negative idx wraps like numpy's corpus_ids[-1], which is what
// the reference's `corpus_ids[x]` does with Faiss' -1 padding, evaluate_repconc.py:183)
__global__ void map_ids_kernel(const int64_t* __restrict__ idx, const int64_t* __restrict__ table, int64_t n_table,
                               int64_t n, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t j = idx[i];
    if (j < 0) j += n_table;
    out[i] = (j >= 0 && j < n_table) ? table[j] : -1;
}

RC_API int rc_map_ids(const int64_t* idx, const int64_t* corpus_ids, int64_t n_corpus, int64_t n, int64_t* out,
                      void* stream) {
    RC_REQUIRE(idx && corpus_ids && out && n >= 0 && n_corpus >= 1, "rc_map_ids: bad argument");
    if (n == 0) return RC_OK;
    map_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(idx, corpus_ids, n_corpus, n, out);
    RC_CHECK_LAUNCH("map_ids_kernel");
    return RC_OK;
}

RC_API int rc_topk_merge(const float* scores_in, const int64_t* ids_in, int W, int64_t nq, int64_t k, float* scores,
                         int64_t* ids, void* stream) {
    RC_REQUIRE(scores_in && ids_in && scores && ids, "rc_topk_merge: null pointer");
    RC_REQUIRE(W >= 1 && nq >= 0 && k >= 1, "rc_topk_merge: bad shape");
    if ((int64_t)W * k > CAND_CAP) {
        set_error("rc_topk_merge: W*k = %lld > %d is not supported", (long long)W * k, CAND_CAP);
        return RC_E_UNSUPPORTED;
    }
    if (nq == 0) return RC_OK;
    int rc = sort_smem_attr();
    if (rc) return rc;
    int n_sort = 2;
    while (n_sort < W * (int)k) n_sort <<= 1;
    topk_merge_kernel<<<(unsigned)nq, SEL_THREADS, (size_t)n_sort * 8, (cudaStream_t)stream>>>(
        scores_in, ids_in, W, nq, (int)k, n_sort, scores, ids);
    RC_CHECK_LAUNCH("topk_merge_kernel");
    return RC_OK;
}
