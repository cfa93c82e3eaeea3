/*
 * repconc_oracle.c -- CPU restatement of the RepCONC constrained-PQ hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (repconc_b200/) never links, imports or calls it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * RepCONC tree, src/repconc/...).  Pinning status:
 *   - table / NN assign / centring / Sinkhorn / decode / MSE: PINNED against
 *     vectors produced by importing the reference module in the build
 *     container (tools/gen_golden.py -> tests/golden/).
 *   - ADC search: the arithmetic lives in Faiss 1.7.1 (un-vendored, absent
 *     here, setup.py:21) -> "parity unpinned" w.r.t. Faiss itself; anchored on
 *     the reference's own decode(): score == <q, decode(codes)> (golden vectors
 *     from modeling_repconc.decode + matmul).
 *
 * fp32 summation order: the reference builds the table with
 *   ((x - c)**2).sum(-1)                       (modeling_repconc.py:50)
 * executed by ATen's CPU sum kernel (aten/src/ATen/native/cpu/SumKernel.cpp,
 * 256-bit float vectors, 4-way ILP, cascade levels).  orc_sum_torch_order()
 * restates that published order so the table is BIT-IDENTICAL to the
 * reference's (probed for dsub in {1..768}, tools/probe_sum_order.py).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 *        (contraction must stay off: fp32 (x-c)*(x-c) then add, no FMA).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* ATen CPU sum order (SumKernel.cpp: multi_row_sum / row_sum /             */
/* vectorized_inner_sum / scalar_inner_sum), restated for a single row.     */
/* `W` is the lane count of one item: 8 for the vector path, 1 for scalars. */
/* ------------------------------------------------------------------------ */
#define ORC_V 8
#define ORC_ILP 4
#define ORC_LEVELS 4

static int ceil_log2_i64(int64_t x) {
    int r = 0;
    if (x <= 1) return 0;
    --x;
    while (x > 0) { x >>= 1; ++r; }
    return r;
}

/* items: n items of W floats laid out contiguously; out: W floats. */
static void row_sum_items(const float *items, int64_t n, int W, float *out) {
    float acc[ORC_LEVELS][ORC_ILP][ORC_V];
    const int64_t size_ilp = n / ORC_ILP;
    memset(acc, 0, sizeof(acc));
    {
        /* multi_row_sum over size_ilp rows of ORC_ILP items */
        const int64_t size = size_ilp;
        int lp = ceil_log2_i64(size) / ORC_LEVELS;
        const int level_power = lp > 4 ? lp : 4;
        const int64_t level_step = (int64_t)1 << level_power;
        const int64_t level_mask = level_step - 1;
        int64_t i = 0;
        while (i + level_step <= size) {
            for (int64_t j = 0; j < level_step; ++j, ++i)
                for (int k = 0; k < ORC_ILP; ++k)
                    for (int l = 0; l < W; ++l)
                        acc[0][k][l] += items[(i * ORC_ILP + k) * W + l];
            for (int j = 1; j < ORC_LEVELS; ++j) {
                for (int k = 0; k < ORC_ILP; ++k)
                    for (int l = 0; l < W; ++l) {
                        acc[j][k][l] += acc[j - 1][k][l];
                        acc[j - 1][k][l] = 0.0f;
                    }
                if ((i & (level_mask << (j * level_power))) != 0) break;
            }
        }
        for (; i < size; ++i)
            for (int k = 0; k < ORC_ILP; ++k)
                for (int l = 0; l < W; ++l)
                    acc[0][k][l] += items[(i * ORC_ILP + k) * W + l];
        for (int j = 1; j < ORC_LEVELS; ++j)
            for (int k = 0; k < ORC_ILP; ++k)
                for (int l = 0; l < W; ++l) acc[0][k][l] += acc[j][k][l];
    }
    for (int64_t i = size_ilp * ORC_ILP; i < n; ++i)
        for (int l = 0; l < W; ++l) acc[0][0][l] += items[i * W + l];
    for (int k = 1; k < ORC_ILP; ++k)
        for (int l = 0; l < W; ++l) acc[0][0][l] += acc[0][k][l];
    for (int l = 0; l < W; ++l) out[l] = acc[0][0][l];
}

/* sum of `n` floats in ATen-CPU order (inner contiguous reduction). */
ORC_API float orc_sum_torch_order(const float *v, int64_t n) {
    if (n >= ORC_V) {
        float part[ORC_V];
        const int64_t nvec = n / ORC_V;
        float fin = 0.0f;
        row_sum_items(v, nvec, ORC_V, part);
        for (int64_t k = nvec * ORC_V; k < n; ++k) fin += v[k];
        for (int k = 0; k < ORC_V; ++k) fin += part[k];
        return fin;
    } else {
        float r;
        row_sum_items(v, n, 1, &r);
        return r;
    }
}

/* squared L2 distance of two dsub-vectors, reference arithmetic:           */
/* fp32 subtract, fp32 square (pow(2) == x*x), ATen-order sum.              */
/* modeling_repconc.py:50                                                   */
static float sqdist_ref(const float *x, const float *c, int ds, float *tmp) {
    for (int j = 0; j < ds; ++j) {
        const float d = x[j] - c[j];
        tmp[j] = d * d;
    }
    return orc_sum_torch_order(tmp, ds);
}

/* ------------------------------------------------------------------------ */
/* a1: distance table d[m][b][k]  (modeling_repconc.py:47-50)               */
/* x (B, M*ds) row-major; c (M, K, ds); out (M, B, K).                      */
/* ------------------------------------------------------------------------ */
ORC_API void orc_dist_table(const float *x, const float *c, int64_t B, int M, int K, int ds,
                            float *out) {
    const int64_t D = (int64_t)M * ds;
#pragma omp parallel
    {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)(ds > 0 ? ds : 1));
#pragma omp for collapse(2) schedule(static)
        for (int m = 0; m < M; ++m)
            for (int64_t b = 0; b < B; ++b) {
                const float *xr = x + b * D + (int64_t)m * ds;
                float *o = out + ((int64_t)m * B + b) * K;
                for (int k = 0; k < K; ++k)
                    o[k] = sqdist_ref(xr, c + ((int64_t)m * K + k) * ds, ds, tmp);
            }
        free(tmp);
    }
}

/* ------------------------------------------------------------------------ */
/* a2: NN assign, codes[b][m] = argmin_k d (first minimum)  (:51-52,66)     */
/* ------------------------------------------------------------------------ */
ORC_API void orc_nn_assign(const float *x, const float *c, int64_t B, int M, int K, int ds,
                           int64_t *codes) {
    const int64_t D = (int64_t)M * ds;
#pragma omp parallel
    {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)(ds > 0 ? ds : 1));
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b)
            for (int m = 0; m < M; ++m) {
                const float *xr = x + b * D + (int64_t)m * ds;
                float best = 0.0f;
                int bi = 0;
                for (int k = 0; k < K; ++k) {
                    const float d = sqdist_ref(xr, c + ((int64_t)m * K + k) * ds, ds, tmp);
                    /* torch.argmin: first minimum; NaN wins */
                    if (k == 0 || d < best || (d != d && best == best)) { best = d; bi = k; }
                }
                codes[b * M + m] = bi;
            }
        free(tmp);
    }
}

/* ------------------------------------------------------------------------ */
/* a3: per-m max / min over (B,K)   (modeling_repconc.py:76-77)             */
/* ------------------------------------------------------------------------ */
ORC_API void orc_table_minmax(const float *table, int M, int64_t B, int K, float *maxd,
                              float *mind) {
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
        const float *t = table + (int64_t)m * B * K;
        float mx = -INFINITY, mn = INFINITY;
        for (int64_t i = 0; i < B * K; ++i) {
            if (t[i] > mx) mx = t[i];
            if (t[i] < mn) mn = t[i];
        }
        maxd[m] = mx;
        mind[m] = mn;
    }
}

/* a3: centring with (possibly all-reduced) max/min   (:78-85)              */
/* returns 0, or -1 if the reference's `assert amplitude > 0` would fire.   */
ORC_API int orc_center_table(float *table, int M, int64_t B, int K, const float *maxd,
                             const float *mind) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int m = 0; m < M; ++m) {
        float *t = table + (int64_t)m * B * K;
        const float middle = (maxd[m] + mind[m]) / 2.0f;
        const float amplitude = (maxd[m] - middle) + 1e-5f;
        if (!(amplitude > 0.0f)) bad |= 1;
        for (int64_t i = 0; i < B * K; ++i) t[i] = (t[i] - middle) / amplitude;
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------ */
/* a4: sinkhorn_algorithm(out, eps, iters, distrib)  (:137-165)             */
/* Q: in = `out` (M,K,B) fp64, overwritten with the returned Q.             */
/* `world` emulates dist.get_world_size(): when the caller passes the       */
/* CONCATENATED global batch, B already is B_local*world and world must be  */
/* 1 (sums over B are then global, identical to the all-reduced reference). */
/* Per element the arithmetic is the reference's (exp, /sum, /r, /K, /c,    */
/* /B, *B in that order); passes over memory are fused where that does not  */
/* change any rounding.                                                     */
/* ------------------------------------------------------------------------ */
ORC_API void orc_sinkhorn(double *Q, int M, int K, int64_t B, double eps, int iters) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m < M; ++m) {
        double *q = Q + (int64_t)m * K * B;
        double *colsum = (double *)malloc(sizeof(double) * (size_t)B);
        double *rowsum = (double *)malloc(sizeof(double) * (size_t)K);
        double total = 0.0;
        /* Q = exp(out / eps); sum_Q = Q.sum(-1).sum(-2)   (:141,148) */
        for (int k = 0; k < K; ++k) {
            double rs = 0.0;
            double *row = q + (int64_t)k * B;
            for (int64_t b = 0; b < B; ++b) {
                row[b] = exp(row[b] / eps);
                rs += row[b];
            }
            total += rs;
        }
        /* Q /= sum_Q   (:152) */
        for (int64_t i = 0; i < (int64_t)K * B; ++i) q[i] /= total;
        for (int it = 0; it < iters; ++it) {
            /* sum_of_rows = Q.sum(dim=2)   (:155) */
            for (int k = 0; k < K; ++k) {
                double rs = 0.0;
                const double *row = q + (int64_t)k * B;
                for (int64_t b = 0; b < B; ++b) rs += row[b];
                rowsum[k] = rs;
            }
            /* Q /= sum_of_rows; Q /= K; then column sums over k  (:158-159,162) */
            memset(colsum, 0, sizeof(double) * (size_t)B);
            for (int k = 0; k < K; ++k) {
                double *row = q + (int64_t)k * B;
                const double rs = rowsum[k];
                for (int64_t b = 0; b < B; ++b) {
                    double v = row[b] / rs;
                    v = v / (double)K;
                    row[b] = v;
                    colsum[b] += v;
                }
            }
            /* Q /= colsum; Q /= B   (:162-163) */
            for (int k = 0; k < K; ++k) {
                double *row = q + (int64_t)k * B;
                for (int64_t b = 0; b < B; ++b) {
                    double v = row[b] / colsum[b];
                    row[b] = v / (double)B;
                }
            }
        }
        /* Q *= B   (:164) */
        for (int64_t i = 0; i < (int64_t)K * B; ++i) q[i] *= (double)B;
        free(colsum);
        free(rowsum);
    }
}

/* ------------------------------------------------------------------------ */
/* a1+a3+a4+a5: RepCONC.quantize with use_constraint=True   (:47-66)        */
/* codes (B,M) int64; *nonfinite = 1 iff Q holds NaN/Inf (the warning :64). */
/* Optional outputs (may be NULL): maxd/mind (M), Qout (M,K,B) fp64.        */
/* If ext_max/ext_min != NULL they replace the locally computed extrema     */
/* (emulates the all_reduce MAX/MIN of :78-80).                             */
/* returns 0 ok, -1 amplitude assert, -2 out of memory.                     */
/* ------------------------------------------------------------------------ */
ORC_API int orc_constrained_assign(const float *x, const float *c, int64_t B, int M, int K,
                                   int ds, double eps, int iters, const float *ext_max,
                                   const float *ext_min, int64_t *codes, int *nonfinite,
                                   float *maxd_out, float *mind_out, double *Qout) {
    const int64_t n = (int64_t)M * B * K;
    float *table = (float *)malloc(sizeof(float) * (size_t)n);
    double *Q = Qout ? Qout : (double *)malloc(sizeof(double) * (size_t)n);
    float *mx = (float *)malloc(sizeof(float) * (size_t)M);
    float *mn = (float *)malloc(sizeof(float) * (size_t)M);
    int rc = 0, bad = 0;
    if (!table || !Q || !mx || !mn) { rc = -2; goto done; }
    orc_dist_table(x, c, B, M, K, ds, table);
    orc_table_minmax(table, M, B, K, mx, mn);
    if (ext_max && ext_min) {
        memcpy(mx, ext_max, sizeof(float) * (size_t)M);
        memcpy(mn, ext_min, sizeof(float) * (size_t)M);
    }
    if (maxd_out) memcpy(maxd_out, mx, sizeof(float) * (size_t)M);
    if (mind_out) memcpy(mind_out, mn, sizeof(float) * (size_t)M);
    if (orc_center_table(table, M, B, K, mx, mn) != 0) { rc = -1; goto done; }
    /* distances.double(); out = -distances.transpose(1,2)   (:56-58) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k)
            for (int64_t b = 0; b < B; ++b)
                Q[((int64_t)m * K + k) * B + b] = -(double)table[((int64_t)m * B + b) * K + k];
    orc_sinkhorn(Q, M, K, B, eps, iters);
    /* codes = argmax_k Q.transpose(1,2); NaN/Inf check; .t()   (:62-66) */
#pragma omp parallel for collapse(2) schedule(static) reduction(| : bad)
    for (int m = 0; m < M; ++m)
        for (int64_t b = 0; b < B; ++b) {
            double best = 0.0;
            int bi = 0;
            for (int k = 0; k < K; ++k) {
                const double v = Q[((int64_t)m * K + k) * B + b];
                if (isnan(v) || isinf(v)) bad |= 1;
                /* torch.argmax: first maximum; NaN wins */
                if (k == 0 || v > best || (v != v && best == best)) { best = v; bi = k; }
            }
            codes[b * M + m] = bi;
        }
    if (nonfinite) *nonfinite = bad;
done:
    free(table);
    if (!Qout) free(Q);
    free(mx);
    free(mn);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* a6: decode(codes, centroids)  (modeling_repconc.py:168-184)              */
/* ------------------------------------------------------------------------ */
ORC_API void orc_decode(const int64_t *codes, const float *c, int64_t B, int M, int K, int ds,
                        float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int m = 0; m < M; ++m)
            memcpy(out + (b * M + m) * ds, c + ((int64_t)m * K + codes[b * M + m]) * ds,
                   sizeof(float) * (size_t)ds);
}

/* ------------------------------------------------------------------------ */
/* a8: surrogate + MSE loss and closed-form gradients                       */
/* (finetune_repconc.py:367-374, 389-396).  n rows of D.                    */
/*   surrogate = <g,x> + <g,q>;  mse = mean_b sum_d (q-x)^2 * w             */
/*   total = scale*mse + surrogate                                          */
/*   d total/dx = g + (2 w scale / n)(x - q);  d total/dq = g - (...)(x-q)  */
/*   d total/dc[m,k,:] = sum_{b: codes[b,m]=k} dq[b,m,:]  (b ascending)     */
/* Accumulations in fp64 then rounded: this is a checker, tolerance 1e-4.   */
/* ------------------------------------------------------------------------ */
ORC_API void orc_mse_surrogate(const float *x, const float *q, const float *g,
                               const int64_t *codes, int64_t n, int M, int K, int ds, float w,
                               float scale, float *mse_out, float *surrogate_out, float *grad_x,
                               float *grad_q, float *grad_c) {
    const int64_t D = (int64_t)M * ds;
    double mse = 0.0, sur = 0.0;
    const double coef = 2.0 * (double)w * (double)scale / (double)n;
    for (int64_t b = 0; b < n; ++b)
        for (int64_t j = 0; j < D; ++j) {
            const double xv = x[b * D + j], qv = q[b * D + j], gv = g[b * D + j];
            mse += (qv - xv) * (qv - xv);
            sur += gv * xv + gv * qv;
            if (grad_x) grad_x[b * D + j] = (float)(gv + coef * (xv - qv));
            if (grad_q) grad_q[b * D + j] = (float)(gv - coef * (xv - qv));
        }
    if (mse_out) *mse_out = (float)(mse / (double)n * (double)w);
    if (surrogate_out) *surrogate_out = (float)sur;
    if (grad_c) {
        double *acc = (double *)calloc((size_t)M * K * ds, sizeof(double));
        for (int64_t b = 0; b < n; ++b)
            for (int m = 0; m < M; ++m) {
                const int64_t k = codes[b * M + m];
                for (int j = 0; j < ds; ++j) {
                    const double xv = x[b * D + m * ds + j], qv = q[b * D + m * ds + j];
                    acc[((int64_t)m * K + k) * ds + j] +=
                        (double)g[b * D + m * ds + j] - coef * (xv - qv);
                }
            }
        for (int64_t i = 0; i < (int64_t)M * K * ds; ++i) grad_c[i] = (float)acc[i];
        free(acc);
    }
}

/* ------------------------------------------------------------------------ */
/* a12: ADC search with Faiss IndexPQ(d, M, 8, METRIC_INNER_PRODUCT)        */
/* semantics as used at evaluate_repconc.py:81-85,94-97,182.                */
/* LUT[q][m][k] = <q_m, c_{m,k}> fp32, j ascending, mul then add (no FMA).  */
/* ------------------------------------------------------------------------ */
ORC_API void orc_adc_lut(const float *queries, const float *c, int64_t nq, int M, int K, int ds,
                         float *lut) {
    const int64_t D = (int64_t)M * ds;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t q = 0; q < nq; ++q)
        for (int m = 0; m < M; ++m) {
            const float *qv = queries + q * D + (int64_t)m * ds;
            for (int k = 0; k < K; ++k) {
                const float *cv = c + ((int64_t)m * K + k) * ds;
                float s = 0.0f;
                for (int j = 0; j < ds; ++j) {
                    const float p = qv[j] * cv[j];
                    s = s + p;
                }
                lut[(q * M + m) * K + k] = s;
            }
        }
}

/* "a ranks before b": larger score first, ties -> smaller id first */
static inline int ranks_before(float sa, int64_t ia, float sb, int64_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

static void heap_sift_down(float *hs, int64_t *hi, int64_t n, int64_t i) {
    /* heap root = the WORST kept element (ranks after every other) */
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && ranks_before(hs[w], hi[w], hs[l], hi[l])) w = l;
        if (r < n && ranks_before(hs[w], hi[w], hs[r], hi[r])) w = r;
        if (w == i) return;
        { float ts = hs[i]; hs[i] = hs[w]; hs[w] = ts; }
        { int64_t ti = hi[i]; hi[i] = hi[w]; hi[w] = ti; }
        i = w;
    }
}

/* score[q][n] = sum_{m ascending} LUT[q][m][code[n][m]] (fp32);            */
/* keep the k largest, sorted descending (ties: smaller id first); if k > N */
/* pad with (lowest float, -1) as Faiss' heap does.  ids are positions      */
/* id_offset + n (id_offset emulates a corpus shard).                       */
ORC_API void orc_adc_search_lut(const float *lut, const uint8_t *codes, int64_t nq, int64_t N,
                                int M, int K, int64_t k, int64_t id_offset, float *scores,
                                int64_t *ids) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t q = 0; q < nq; ++q) {
        const float *t = lut + q * M * K;
        float *hs = scores + q * k;
        int64_t *hi = ids + q * k;
        int64_t cnt = 0;
        for (int64_t n = 0; n < N; ++n) {
            const uint8_t *cd = codes + n * M;
            float s = 0.0f;
            for (int m = 0; m < M; ++m) s = s + t[m * K + cd[m]];
            if (cnt < k) {
                /* append, sift up */
                int64_t i = cnt++;
                hs[i] = s; hi[i] = id_offset + n;
                while (i > 0) {
                    int64_t p = (i - 1) / 2;
                    if (ranks_before(hs[p], hi[p], hs[i], hi[i])) {
                        float ts = hs[i]; hs[i] = hs[p]; hs[p] = ts;
                        int64_t ti = hi[i]; hi[i] = hi[p]; hi[p] = ti;
                        i = p;
                    } else break;
                }
            } else if (k > 0 && ranks_before(s, id_offset + n, hs[0], hi[0])) {
                hs[0] = s; hi[0] = id_offset + n;
                heap_sift_down(hs, hi, k, 0);
            }
        }
        /* heap sort: repeatedly move the worst to the end -> descending order */
        for (int64_t end = cnt; end > 1; --end) {
            float ts = hs[0]; hs[0] = hs[end - 1]; hs[end - 1] = ts;
            int64_t ti = hi[0]; hi[0] = hi[end - 1]; hi[end - 1] = ti;
            heap_sift_down(hs, hi, end - 1, 0);
        }
        for (int64_t i = cnt; i < k; ++i) { hs[i] = -FLT_MAX; hi[i] = -1; }
    }
}

ORC_API int orc_adc_search(const float *queries, const float *c, const uint8_t *codes, int64_t nq,
                           int64_t N, int M, int K, int ds, int64_t k, int64_t id_offset,
                           float *scores, int64_t *ids) {
    float *lut = (float *)malloc(sizeof(float) * (size_t)(nq * M * K));
    if (!lut) return -2;
    orc_adc_lut(queries, c, nq, M, K, ds, lut);
    orc_adc_search_lut(lut, codes, nq, N, M, K, k, id_offset, scores, ids);
    free(lut);
    return 0;
}

/* merge W sorted (descending) top-k lists per query into one (shard merge) */
ORC_API void orc_topk_merge(const float *scores_in, const int64_t *ids_in, int W, int64_t nq,
                            int64_t k, float *scores, int64_t *ids) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nq; ++q) {
        int64_t *pos = (int64_t *)calloc((size_t)W, sizeof(int64_t));
        for (int64_t o = 0; o < k; ++o) {
            int best = -1;
            for (int w = 0; w < W; ++w) {
                if (pos[w] >= k) continue;
                const float s = scores_in[((int64_t)w * nq + q) * k + pos[w]];
                const int64_t id = ids_in[((int64_t)w * nq + q) * k + pos[w]];
                if (id < 0) continue;
                if (best < 0) { best = w; continue; }
                {
                    const float sb = scores_in[((int64_t)best * nq + q) * k + pos[best]];
                    const int64_t ib = ids_in[((int64_t)best * nq + q) * k + pos[best]];
                    if (ranks_before(s, id, sb, ib)) best = w;
                }
            }
            if (best < 0) { scores[q * k + o] = -FLT_MAX; ids[q * k + o] = -1; continue; }
            scores[q * k + o] = scores_in[((int64_t)best * nq + q) * k + pos[best]];
            ids[q * k + o] = ids_in[((int64_t)best * nq + q) * k + pos[best]];
            pos[best]++;
        }
        free(pos);
    }
}
