/*
 * repconc_b200.h -- C ABI of librepconc_b200.so (hand-written sm_100a CUDA kernels for the
 * RepCONC constrained-clustering product-quantization hot path).
 *
 * Conventions
 *   - every pointer is a CALLER-OWNED DEVICE pointer unless the name ends in `_host`;
 *     nothing is allocated, freed or retained by the library (workspaces are sized by the
 *     matching *_workspace_bytes function and passed in);
 *   - `stream` is a cudaStream_t passed as void* (torch: torch.cuda.current_stream().cuda_stream);
 *   - every function returns 0 on success, <0 on error (RC_E_*); the message of the last error
 *     on the calling thread is returned by rc_last_error();
 *   - functions are asynchronous on `stream` unless documented "synchronises";
 *   - layouts are row-major; M sub-vectors, K centroids per sub-vector, ds = D / M floats each.
 *
 * Each entry point names the reference interface (jingtaozhan/RepCONC, paths under
 * src/repconc/) it replaces.  INTEGRATION.md shows the reference-side bindings.
 */
#ifndef REPCONC_B200_H
#define REPCONC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_OK 0
#define RC_E_INVALID (-1)   /* bad argument */
#define RC_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define RC_E_UNSUPPORTED (-3)
#define RC_E_WORKSPACE (-4) /* workspace too small */

/* flag bits written by the assign kernels into `flags` (one int32 on the device) */
#define RC_FLAG_NONFINITE 1 /* Q would hold NaN/Inf: models/repconc/modeling_repconc.py:64-65 */
#define RC_FLAG_AMPLITUDE 2 /* `assert torch.all(amplitude > 0)` would fire: modeling_repconc.py:83 */
#define RC_FLAG_BADCODE 4   /* rc_decode saw a code outside [0, K) */
#define RC_FLAG_SPARSE_UNSAFE 8 /* a row kept < 2^-8/K of mass or the survivor pool ran out: re-run with dense = 1 */
#define RC_FLAG_PEER_TIMEOUT 16 /* rc_peer_allreduce_f64: a peer never signalled (bounded spin expired) */

const char* rc_last_error(void);
/* library / build identification: "repconc_b200 <version> sm_100a" */
const char* rc_version(void);
/* number of kernel launches issued through this library by the calling process so far */
int64_t rc_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * a2  NN assign (use_constraint == False):  codes = argmin_k ||x[b,m,:] - c[m,k,:]||^2
 *     replaces modeling_repconc.py:47-52,66 (RepCONC.quantize, argmin branch); the uint8 output
 *     additionally fuses evaluate_repconc.py:69 (`codes.astype(np.uint8)`).
 *   x          (B, M*ds) fp32, row stride `ldx` floats
 *   centroids  (M, K, ds) fp32
 *   codes_mb   (M, B) int64 or NULL   -- the reference's pre-`.t()` layout
 *   codes_u8   (B, M) uint8 or NULL   -- requires K <= 256
 * The fp32 arithmetic (subtract, square, ATen-CPU summation order, first-minimum ties) is
 * bit-identical to the reference's CPU path.
 * ------------------------------------------------------------------------------------------- */
int rc_nn_assign(const float* x, int64_t ldx, const float* centroids, int64_t B, int M, int K, int ds,
                 int64_t* codes_mb, uint8_t* codes_u8, void* stream);

/* f3 -- fused corpus-encode epilogue: what RepCONC.forward does after the encoder when
 * use_constraint is False -- `dense_embed @ rotation.T` (modeling_repconc.py:98), the per-sub-vector L2
 * normalisation of METRIC_CENTROID_COS (:99-100), `quantize` = argmin of the squared distances (:101, :51-52)
 * -- and evaluate_repconc.py:69's `.cpu().numpy().astype(np.uint8)`, in one kernel: the rotated embeddings
 * stay in registers, the codes come out as (B, M) uint8 rows ready for GpuIndexPQ.add.
 *   pooled      (B, D = M*ds) fp32 encoder output, row stride ld
 *   rotation    (D, D) fp32 row-major, y = pooled @ rotation^T (the module's `rotation` buffer)
 *   normalize   1 for METRIC_CENTROID_COS
 *   rotated_out optional (B, D) fp32, row stride ld_out: the module's `continuous_embeds`; NULL to skip
 *   codes_mb / codes_u8  as for rc_nn_assign (either may be NULL, not both)
 * An identity rotation gives rc_nn_assign's codes bit for bit. */
int rc_encode_assign(const float* pooled, int64_t ld, const float* rotation, const float* centroids, int64_t B,
                     int M, int K, int ds, int normalize, float* rotated_out, int64_t ld_out,
                     int64_t* codes_mb, uint8_t* codes_u8, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a1+a3  distance table + per-sub-vector extrema
 *     replaces modeling_repconc.py:50 (table) and :76-77 (max / min over (B,K)).
 *   table   (M, B, K) fp32 out
 *   minmax  (2, M) fp32 out: minmax[0][m] = max, minmax[1][m] = min  (all-reduce MAX / MIN these
 *           across ranks exactly as modeling_repconc.py:78-80 does, then call rc_sinkhorn_begin)
 *   flags   int32*, RC_FLAG_NONFINITE is OR-ed in if the table holds NaN
 * minmax must be initialised by rc_minmax_init before the first (or only) call; several calls
 * with different row blocks accumulate into the same extrema.
 * ------------------------------------------------------------------------------------------- */
int rc_minmax_init(float* minmax, int M, void* stream);
int rc_dist_table(const float* x, int64_t ldx, const float* centroids, int64_t B, int M, int K, int ds,
                  float* table, float* minmax, int32_t* flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3+a4+a5  Sinkhorn uniform assignment in scaling-vector (log-domain) form
 *     replaces RepCONC.center_distance_for_constraint (modeling_repconc.py:73-85), `.double()`
 *     (:56) and sinkhorn_algorithm (:137-165) + argmax / NaN check (:62-66).
 *
 * State (all caller-owned, sized by rc_sinkhorn_state_bytes, laid out by the library):
 *   lu (M,K) fp64  log row scaling,   lv (M,B) fp64  log column scaling,
 *   P  (M,K) fp64  row sums of Q (the tensor the reference all-reduces at :157),
 *   partial sums scratch.
 * Q_t[m,k,b] = exp(-d~[m,b,k]/eps + lu[m,k] + lv[m,b]) is never materialised.
 *
 * Call sequence for one quantize() (W = world size, B = rows on this rank):
 *   rc_minmax_init; rc_dist_table; [all_reduce MAX/MIN minmax]
 *   rc_sinkhorn_begin      centre the table in place (fp32, bit-identical to :81-84), zero lu/lv,
 *                          and compute P = sum_b exp(-d~/eps)          (first row sums, :155)
 *   repeat iters-1 times:  [all_reduce SUM P]; rc_sinkhorn_step        (row+column normalisation
 *                          of iteration t and the row sums of iteration t+1 in ONE table pass)
 *   [all_reduce SUM P]; rc_sinkhorn_finish                            (last row normalisation,
 *                          argmax_k, NaN/Inf flag)
 * The all-reduces are the caller's (torch.distributed / NCCL on the same stream) -- the library
 * has no communicator.  iters == 0 is allowed (begin, finish).
 * ------------------------------------------------------------------------------------------- */
/* Sinkhorn pass selection is PER CALL: the `dense` argument of rc_sinkhorn_step / finish / solve.
 * dense = 0: for K == 256 the iteration passes evaluate only the table elements within 2^-72 of their
 * column sum (the rest cannot change an fp64 sum; bound in csrc/assign.cu); dense = 1: always the dense
 * pass (what the host re-runs with after RC_FLAG_SPARSE_UNSAFE).  rc_sinkhorn_set_dense only sets the
 * process-wide default that is OR-ed with the argument (debugging / A-B runs; env RC_SINKHORN_DENSE=1
 * sets its initial value); it returns the previous default. */
int rc_sinkhorn_set_dense(int dense);
/* test hook: usable survivor-pool capacity in entries per table row (0 = full allocation); returns the
 * previous value.  Lets a test exhaust the pool -> RC_FLAG_SPARSE_UNSAFE on one rank only. */
int64_t rc_sinkhorn_debug_pool_entries(int64_t entries_per_row);
size_t rc_sinkhorn_state_bytes(int64_t B, int M, int K);
/* device pointer to the (M,K) fp64 row-sum buffer inside `state` (the all-reduce operand) */
double* rc_sinkhorn_rowsum_ptr(void* state, int64_t B, int M, int K);
int rc_sinkhorn_begin(float* table, const float* minmax, int64_t B, int M, int K, double eps,
                      void* state, int32_t* flags, void* stream);
/* step_index: 0 for the first step after rc_sinkhorn_begin, then 1, 2, ... (the first step always
 * runs the dense pass: its input columns are not normalised yet) */
int rc_sinkhorn_step(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                     int step_index, int dense, void* state, int32_t* flags, void* stream);
/* the transport plan itself, Q (M,K,B) fp64 with columns summing to 1 -- what the reference's
 * sinkhorn_algorithm returns (modeling_repconc.py:164-165); call instead of rc_sinkhorn_finish */
int rc_sinkhorn_expand(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                       int apply_rowsum, void* state, double* Q, int32_t* flags, void* stream);
/* Single-rank solve = rc_sinkhorn_begin + (iters-1) x rc_sinkhorn_step + rc_sinkhorn_finish with
 * B_global == B, in one call and bit-identical to that sequence.  For the sparse passes (K == 256,
 * dense == 0) the whole iteration loop is ONE persistent cooperative kernel: per sub-vector, the last
 * CTA that finishes a pass reduces the row sums, updates the row scaling and releases the next pass --
 * three launches per iteration become none.  This is what RepCONC.quantize runs when torch.distributed
 * is not initialised (modeling_repconc.py:61). */
int rc_sinkhorn_solve(float* table, const float* minmax, int64_t B, int M, int K, double eps, int iters,
                      int dense, void* state, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                      void* stream);
/* W-rank solve with the per-iteration exchange of the row sums FUSED into the row-sum kernel: the
 * `dist.all_reduce(sum_of_rows)` of modeling_repconc.py:156-157 (M x K fp64 every iteration) happens
 * inside the kernel that reduces the sums -- block m PUSHES its 256 sums into every peer's symmetric
 * buffer as 16-byte words {low half, seq, high half, seq} (every 8-byte half carries the sequence number
 * of the exchange, so no fence and no separate flag are needed), polls its own buffer for the peers'
 * words, sums the W vectors in rank order (bitwise identical on every rank) and updates the row scaling
 * of m.  One NVLink store latency per exchange (measured 1-3 us on 2 B200s).
 *   peer_buffers_host  HOST array of W device pointers: rank p's symmetric buffer as mapped into THIS
 *                      process, rc_sinkhorn_peer_buffer_bytes(M, K) bytes each, zero-initialised ONCE
 *                      (e.g. torch.distributed._symmetric_memory); reused by later calls
 *   seq_base           exchange sequence numbers consumed so far (0 for the first call, then the sum of
 *                      max(iters, 1) over the previous calls on this buffer) -- the same on every rank
 *   minmax             already all-reduced (MAX / MIN) by the caller, as for rc_sinkhorn_begin
 * Sparse passes only (K == 256); a dense re-run goes through the step-wise entry points + the caller's
 * all-reduce.  Waits are bounded (env RC_PEER_TIMEOUT_MS, default 30000): RC_FLAG_PEER_TIMEOUT. */
size_t rc_sinkhorn_peer_buffer_bytes(int M, int K);
int rc_sinkhorn_solve_peer(float* table, const float* minmax, int64_t B, int64_t B_global, int M, int K,
                           double eps, int iters, void* state, const uint64_t* peer_buffers_host, int rank,
                           int W, uint32_t seq_base, int64_t* codes_mb, uint8_t* codes_u8, int32_t* flags,
                           void* stream);
/* diagnostics of the survivor lists the sparse pass currently holds (K == 256 only): out[0] = entries,
 * out[1] = longest row, out[2] = rows, out[3 + i] = rows with i*8 <= count < i*8+8 (i < 33); counters of
 * the last assignment on this state: out[36] = segments that ran the selection pass, out[37] = segments
 * that ran the list pass, out[38] = centroids that failed the mass check, out[39] = rows that found the
 * survivor pool full.  `out` is a DEVICE array of 40 int64. */
/* diagnostics: per-CTA nanoseconds of the last persistent-kernel run (wait, selection, list, arrive+update);
 * out_host is a HOST array of max_ctas x 4 int64; returns the number of CTAs written; synchronises the device */
int rc_sinkhorn_debug_cta_times(void* state, int64_t B, int M, int K, int64_t* out_host, int max_ctas);
/* diagnostics: per sub-vector {max_k, max_k - min_k} of lu - lu_build after the last update (host array of 2*M
 * doubles): how far the row scaling has moved since the survivor lists were selected.  Synchronises the device. */
int rc_sinkhorn_debug_drift(void* state, int64_t B, int M, int K, double* out_host);
int rc_sinkhorn_list_stats(void* state, int64_t B, int M, int K, int64_t* out, void* stream);
/* apply_rowsum: 1 = apply the pending row normalisation from P first (iters >= 1); 0 = iters == 0
 * steps_done : number of rc_sinkhorn_step calls since rc_sinkhorn_begin (the last row sums get the sparse
 *              pass's mass check only if they came from a step) */
int rc_sinkhorn_finish(const float* table, int64_t B, int64_t B_global, int M, int K, double eps,
                       int apply_rowsum, int steps_done, int dense, void* state, int64_t* codes_mb,
                       uint8_t* codes_u8, int32_t* flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * One-shot all-reduce (SUM, fp64, in place) of the row sums over NVLink peer memory: the
 * `dist.all_reduce(sum_of_rows)` of modeling_repconc.py:156-157 without NCCL.
 *   peer_buffers_host  HOST array of W device pointers: rank p's symmetric buffer as mapped into THIS
 *                      process (rc_peer_allreduce_buffer_bytes(n) bytes each, zero-initialised once,
 *                      e.g. torch.distributed._symmetric_memory buffers)
 *   seq                1, 2, 3, ... -- the same value on every rank for the same exchange
 *   flags              RC_FLAG_PEER_TIMEOUT is OR-ed in if a peer never arrives (the result is then undefined)
 * Every rank sums the W vectors in rank order, so all ranks hold bitwise identical results.
 * ------------------------------------------------------------------------------------------- */
size_t rc_peer_allreduce_buffer_bytes(int64_t n);
int rc_peer_allreduce_f64(const uint64_t* peer_buffers_host, int rank, int W, int64_t n, uint32_t seq,
                          double* inout, int32_t* flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a6  decode: q[b, m*ds:(m+1)*ds] = c[m, codes[b,m], :]        (modeling_repconc.py:168-184)
 *   codes: int64, element strides (stride_b, stride_m) -- accepts the non-contiguous `.t()` view
 *   returned by quantize; codes_u8 (B,M) contiguous alternative (exactly one of the two non-NULL)
 *   out (B, M*ds) fp32.  RC_FLAG_NONFINITE is NOT used; out-of-range codes set *flags |= 4.
 * a6' backward of decode w.r.t. centroids (index_put_(accumulate=True)):
 *   grad_c[m,k,:] = sum_{b: codes[b,m]==k} grad_q[b,m,:], deterministic, b ascending per chunk.
 *   workspace: rc_decode_bwd_workspace_bytes.
 * ------------------------------------------------------------------------------------------- */
int rc_decode(const int64_t* codes, int64_t stride_b, int64_t stride_m, const uint8_t* codes_u8,
              const float* centroids, int64_t B, int M, int K, int ds, float* out, int32_t* flags,
              void* stream);
/* counts[m,k] = #{b : codes[b,m] == k} (int32, (M,K)): the code histogram of eval_balance
 * (finetune_repconc.py:604-611) and the cluster sizes of the k-means warm-up (run_warmup.py:85-132).
 * codes as for rc_decode; out-of-range codes set *flags |= 4 and are not counted. */
int rc_code_histogram(const int64_t* codes, int64_t stride_b, int64_t stride_m, const uint8_t* codes_u8,
                      int64_t B, int M, int K, int32_t* counts, int32_t* flags, void* stream);
size_t rc_decode_bwd_workspace_bytes(int64_t B, int M, int K, int ds);
int rc_decode_bwd(const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* grad_q,
                  int64_t ldg, int64_t B, int M, int K, int ds, float* grad_c, void* workspace,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  quantisation (MSE) loss + surrogate and its closed-form backward
 *     replaces finetune_repconc.py:367-374 (forward) and the autograd of :389-396.
 *   forward : out2[0] = mse = mean_b sum_d (q-x)^2 * w ; out2[1] = <g,x> + <g,q>   (fp64 accum)
 *             q may be NULL -> decoded on the fly from codes (fused decode)
 *             g may be NULL -> surrogate = 0
 *   backward: grad_x = gs*g + coef*(x-q), grad_q = gs*g - coef*(x-q), coef = 2*w*gm/n
 *             (gm, gs = upstream gradients of mse and surrogate; AMP loss scale folds into gm)
 *             grad_q may be NULL; grad_c (M,K,ds) gets the scatter-add of grad_q (may be NULL)
 * ------------------------------------------------------------------------------------------- */
int rc_mse_fwd(const float* x, int64_t ldx, const float* q, int64_t ldq, const float* g, int64_t ldg,
               const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* centroids,
               int64_t n, int M, int K, int ds, float w, float* out2, void* workspace, void* stream);
size_t rc_mse_workspace_bytes(int64_t n, int M, int K, int ds);
int rc_mse_bwd(const float* x, int64_t ldx, const float* q, int64_t ldq, const float* g, int64_t ldg,
               const int64_t* codes, int64_t stride_b, int64_t stride_m, const float* centroids,
               int64_t n, int M, int K, int ds, float w, float gm, float gs, float* grad_x,
               float* grad_q, float* grad_c, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a12  PQ asymmetric-distance (inner product) search, Faiss IndexPQ(d, M, 8, METRIC_INNER_PRODUCT)
 *      semantics as the reference uses them (evaluate_repconc.py:78-98,180-185;
 *      finetune_jpq.py:176):  LUT[q,m,k] = <x_q[m,:], c[m,k,:]>,
 *      score[q,n] = sum_{m ascending} LUT[q,m,code[n,m]] (fp32), the k largest sorted descending,
 *      ties -> smaller id first, k > N padded with (-FLT_MAX, -1).
 *   queries (nq, M*ds) fp32 row stride ldq;  codes (N, M) uint8;  K == 256
 *   scores (nq,k) fp32 out, ids (nq,k) int64 out (= id_offset + row position)
 * rc_adc_search runs LUT build, threshold sampling, the filtered corpus scan and the final
 * per-query sort; it SYNCHRONISES the stream (it reads per-query candidate counts back to pick
 * the exact fallback for queries whose candidate buffer under/overflowed).
 * The corpus scan is an integer FILTER (8-bit-quantised tables, conservative threshold); every
 * surviving document is re-scored in fp32 exactly as above and a query whose result cannot be
 * proven complete takes an exact dense pass, so scores and ids are bit-identical to a sequential
 * fp32 scan whatever the data.  `codes` 8-byte aligned selects the fastest scan (any alignment
 * works); N < 2^32 per call (shard larger corpora and merge with rc_topk_merge); k <= 8192.
 * ------------------------------------------------------------------------------------------- */
size_t rc_adc_search_workspace_bytes(int64_t nq, int64_t N, int M, int K, int64_t k);
int rc_adc_search(const float* queries, int64_t ldq, const float* centroids, const uint8_t* codes,
                  int64_t nq, int64_t N, int M, int K, int ds, int64_t k, int64_t id_offset,
                  float* scores, int64_t* ids, void* workspace, size_t workspace_bytes, void* stream);

/* building blocks of rc_adc_search, exported for tests / profiling */
int rc_adc_lut(const float* queries, int64_t ldq, const float* centroids, int64_t nq, int M, int K,
               int ds, float* lut, void* stream);
/* dense scores of docs [n0, n0+n) for all queries: out (nq, n) fp32 */
int rc_adc_scores(const float* lut, const uint8_t* codes, int64_t nq, int64_t n0, int64_t n, int M,
                  float* out, void* stream);
/* out[i] = corpus_ids[idx[i]]: the position -> external id mapping of evaluate_repconc.py:183 on the device
 * (negative positions wrap like numpy indexing, as the reference's `corpus_ids[x]` does) */
int rc_map_ids(const int64_t* idx, const int64_t* corpus_ids, int64_t n_corpus, int64_t n, int64_t* out,
               void* stream);
/* merge W per-shard sorted top-k lists (W, nq, k) into (nq, k); workspace-free */
int rc_topk_merge(const float* scores_in, const int64_t* ids_in, int W, int64_t nq, int64_t k,
                  float* scores, int64_t* ids, void* stream);

/* statistics of the last rc_adc_search on this thread: [0] queries served by the filtered scan,
 * [1] queries that took the exact dense fallback, [2] max candidates of any query, [3] sample size */
void rc_adc_last_stats(int64_t out4[4]);
/* measurement hooks: with timing enabled rc_adc_search brackets every filtered-scan launch (the
 * dominant kernel) with CUDA events on `stream`; the sum of their durations (ms) and the number of
 * launches of the last search on this thread are returned by the two getters */
void rc_adc_enable_timing(int enable);
double rc_adc_last_scan_ms(void);
int rc_adc_last_scan_launches(void);
/* LSU (shared-memory data pipe) wavefronts the filtered-scan launches of the last search must issue, counted
 * analytically from the kernel's instruction mix (the quantity ncu reports as
 * l1tex__data_pipe_lsu_wavefronts), and the name of the scan kernel that ran */
double rc_adc_last_scan_wavefronts(void);
const char* rc_adc_last_scan_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* REPCONC_B200_H */
