"""GPU parity of the PQ asymmetric-distance search against the oracle (Faiss IndexPQ semantics as the
reference uses them) and the golden fixtures (reference decode + matmul).  The scan accumulates in
fp32, m ascending, exactly like the oracle, so scores and ids are compared bit-exactly; against the
fp64 decode+matmul anchor the tolerance is the north_star's 1e-4 relative."""
import numpy as np
import pytest
import torch

from tests import golden_cases as GC
from tests.conftest import golden

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _index(c, codes):
    from repconc_b200.faiss_compat import GpuIndexPQ
    return GpuIndexPQ(_dev(codes), _dev(c))


@pytest.mark.parametrize("name", list(GC.ADC_CASES))
def test_search_matches_oracle_and_golden(name, oracle):
    case = GC.ADC_CASES[name]
    g = golden(name)
    q, c, codes = GC.adc_inputs(case)
    idx = _index(c, codes)
    for k in case["ks"]:
        s, i = idx.search(q, k)
        os_, oi = oracle.adc_search(q, c, codes, k)
        assert np.array_equal(s, os_) and np.array_equal(i, oi)
        gs, gi = g[f"scores_k{k}"], g[f"ids_k{k}"].astype(np.int64)
        np.testing.assert_allclose(s, gs, rtol=1e-4, atol=1e-4)
        diff = i != gi
        assert np.all(np.abs(s[diff] - gs[diff]) <= 1e-4 * np.abs(gs[diff]) + 1e-4) and diff.mean() < 0.01
    # CUDA tensors in -> CUDA tensors out (finetune_jpq.py:176)
    st, it = idx.search(_dev(q), 10)
    assert st.is_cuda and it.is_cuda and it.dtype == torch.int64
    assert np.array_equal(it.cpu().numpy(), oracle.adc_search(q, c, codes, 10)[1])


def test_lut_and_dense_scores_bit_exact(oracle):
    from repconc_b200 import _lib, ops
    from oracle import oracle_np as ON
    lib = _lib.load()
    case = GC.ADC_CASES["adc_m48"]
    q, c, codes = GC.adc_inputs(case)
    qd, cd, kd = _dev(q), _dev(c), _dev(codes)
    M, K, ds = c.shape
    lut = torch.empty((len(q), M, K), device="cuda")
    _lib.check(lib.rc_adc_lut(qd.data_ptr(), qd.stride(0), cd.data_ptr(), len(q), M, K, ds, lut.data_ptr(),
                              ops._stream()), "lut")
    assert np.array_equal(lut.cpu().numpy(), oracle.adc_lut(q, c))
    n0, n = 100, 3001
    out = torch.empty((len(q), n), device="cuda")
    _lib.check(lib.rc_adc_scores(lut.data_ptr(), kd.data_ptr(), len(q), n0, n, M, out.data_ptr(), ops._stream()),
               "scores")
    assert np.array_equal(out.cpu().numpy()[:4], ON.adc_scores(q[:4], c, codes)[:, n0:n0 + n])


@pytest.mark.parametrize("M,ds", [(8, 4), (12, 4), (32, 8), (64, 4), (96, 8), (7, 3)])
def test_all_code_widths(oracle, M, ds):
    r = np.random.default_rng(100 + M)
    N, nq = 3000, 9
    c = r.standard_normal((M, 256, ds), dtype=np.float32)
    codes = r.integers(0, 256, size=(N, M), dtype=np.uint8)
    q = r.standard_normal((nq, M * ds), dtype=np.float32)
    s, i = _index(c, codes).search(q, 17)
    os_, oi = oracle.adc_search(q, c, codes, 17)
    assert np.array_equal(s, os_) and np.array_equal(i, oi)


def test_edge_cases(oracle):
    case = dict(GC.ADC_CASES["adc_m8"], N=37, nq=3)
    q, c, codes = GC.adc_inputs(case)
    idx = _index(c, codes)
    s, i = idx.search(q, 50)                                   # k > N pads like Faiss
    os_, oi = oracle.adc_search(q, c, codes, 50)
    assert np.array_equal(s, os_) and np.array_equal(i, oi)
    assert np.all(i[:, 37:] == -1) and np.all(s[:, 37:] == np.finfo(np.float32).min)
    s0, i0 = idx.search(np.zeros((0, 128), np.float32), 5)     # no queries
    assert s0.shape == (0, 5) and i0.shape == (0, 5)
    # all documents identical: N-way exact tie -> ids ascending
    codes2 = np.repeat(codes[:1], 5000, axis=0)
    s2, i2 = _index(c, codes2).search(q, 100)
    assert np.array_equal(i2, np.tile(np.arange(100), (3, 1)))
    # id_offset (a corpus shard)
    from repconc_b200.faiss_compat import GpuIndexPQ
    sh = GpuIndexPQ(_dev(codes[10:]), _dev(c), id_offset=10)
    s3, i3 = sh.search(q, 5)
    o3 = oracle.adc_search(q, c, codes[10:], 5, id_offset=10)
    assert np.array_equal(s3, o3[0]) and np.array_equal(i3, o3[1])


def test_filtered_scan_path_large_corpus(oracle):
    """N above the dense-path limit: sampled thresholds + filtered scan + sort, exact vs the oracle;
    includes duplicate documents (ties) and a skewed block of near-identical high scorers."""
    r = np.random.default_rng(77)
    M, ds, N, nq, k = 16, 4, 600_000, 24, 1000
    c = r.standard_normal((M, 256, ds), dtype=np.float32)
    codes = r.integers(0, 256, size=(N, M), dtype=np.uint8)
    codes[N // 2: N // 2 + 300] = codes[7]                     # 300 exact duplicates of doc 7
    q = r.standard_normal((nq, M * ds), dtype=np.float32)
    idx = _index(c, codes)
    s, i = idx.search(q, k)
    assert idx.last_stats["sample"] > 0 and idx.last_stats["filtered"] > 0
    os_, oi = oracle.adc_search(q, c, codes, k)
    assert np.array_equal(s, os_) and np.array_equal(i, oi)
    s10, i10 = idx.search(q, 10)
    assert np.array_equal(i10, oi[:, :10]) and np.array_equal(s10, os_[:, :10])


@pytest.mark.parametrize("M,ds,N", [(32, 24, 300_000), (64, 12, 280_000), (96, 8, 270_000), (48, 16, 330_000),
                                    (40, 4, 300_000), (12, 8, 300_000), (56, 4, 290_001), (8, 16, 300_000),
                                    (16, 8, 270_017), (24, 4, 265_000), (80, 4, 263_000)])
def test_filtered_scan_every_kernel_family_vs_oracle(oracle, M, ds, N):
    """Above the dense-path limit (262,144 docs) every M takes sampled thresholds + a packed integer filter scan +
    exact re-score: the conflict-free kernel (M % 8 == 0), the 4-query / 2-query variants for wide codes
    (M = 64, 96: the reference's other canonical settings) and the thread-per-document gather (other M).
    Scores and ids must equal the oracle's bit for bit, for k = 10 and k = 1000."""
    r = np.random.default_rng(1000 + M)
    nq = 19
    c = r.standard_normal((M, 256, ds), dtype=np.float32)
    codes = r.integers(0, 256, size=(N, M), dtype=np.uint8)
    codes[N // 3: N // 3 + 40] = codes[11]                     # exact ties across the corpus
    q = r.standard_normal((nq, M * ds), dtype=np.float32)
    q[3] = 0.0                                                 # constant table: every document ties
    idx = _index(c, codes)
    for k in (1000, 10):
        s, i = idx.search(q, k)
        assert idx.last_stats["sample"] > 0
        os_, oi = oracle.adc_search(q, c, codes, k)
        assert np.array_equal(s, os_) and np.array_equal(i, oi), (M, k)
    assert idx.last_stats["filtered"] >= nq - 1


def test_full_size_queries_vs_oracle(oracle):
    """BASELINE configs[1] corpus size (8,841,823 docs, M=48, k=1000) generated by the section-8(d) recipe in
    miniature (codes = NN assign of synthetic documents by the path itself, queries = document + 0.5 noise):
    16 queries against the oracle's exhaustive scan, bit for bit, and identical MRR@10."""
    from repconc_b200 import ops
    from repconc_b200.faiss_compat import GpuIndexPQ
    N, M, ds, nq, k = 8_841_823, 48, 16, 16, 1000
    D = M * ds
    gen = torch.Generator(device="cuda").manual_seed(4242)
    seed_docs = torch.randn((256, D), generator=gen, device="cuda")
    c = seed_docs.reshape(256, M, ds).transpose(0, 1).contiguous()          # trained-like centroids
    codes = torch.empty((N, M), dtype=torch.uint8, device="cuda")
    rel = torch.randint(0, N, (nq,), generator=gen, device="cuda")
    qs = torch.empty((nq, D), device="cuda")
    blk = 1 << 18
    for lo in range(0, N, blk):
        hi = min(N, lo + blk)
        docs = torch.randn((hi - lo, D), generator=gen, device="cuda")
        codes[lo:hi] = ops.nn_assign(docs, c, uint8=True)
        sel = ((rel >= lo) & (rel < hi)).nonzero().flatten()
        if len(sel):
            qs[sel] = docs[rel[sel] - lo]
    qs += 0.5 * torch.randn((nq, D), generator=gen, device="cuda")
    idx = GpuIndexPQ(codes, c)
    s, i = idx.search_tensor(qs, k)
    os_, oi = oracle.adc_search(qs.cpu().numpy(), c.cpu().numpy(), codes.cpu().numpy(), k)
    assert np.array_equal(s.cpu().numpy(), os_) and np.array_equal(i.cpu().numpy(), oi)
    relh = rel.cpu().numpy()
    assert oracle.mrr_at_k(i.cpu().numpy(), relh, 10) == oracle.mrr_at_k(oi, relh, 10)


def test_fallback_when_thresholds_fail(oracle):
    """Adversarial corpus for the sampler: every high scorer sits outside the sampled blocks, or a
    huge tie group overflows the candidate list -> those queries must take the exact fallback."""
    r = np.random.default_rng(78)
    M, ds, N, nq, k = 8, 4, 400_000, 6, 50
    c = r.standard_normal((M, 256, ds), dtype=np.float32)
    codes = r.integers(0, 256, size=(N, M), dtype=np.uint8)
    codes[1000:21000] = codes[0]                               # 20k-way exact tie (> candidate capacity)
    q = r.standard_normal((nq, M * ds), dtype=np.float32)
    # make the tie group the top scorer for query 0
    from oracle import oracle as O
    lut = O.adc_lut(q[:1], c)[0]
    codes[0] = lut.argmax(1).astype(np.uint8)
    codes[1000:21000] = codes[0]
    idx = _index(c, codes)
    s, i = idx.search(q, k)
    os_, oi = oracle.adc_search(q, c, codes, k)
    assert np.array_equal(s, os_) and np.array_equal(i, oi)
    assert idx.last_stats["dense"] >= 1


def test_reference_wrappers_and_mrr(oracle):
    """initialize_index / add_docs / from_pq_to_ivfpq / load_index_to_gpu / batch_search end to end on a
    synthetic eval set; MRR@10 identical to the oracle's run."""
    from repconc_b200 import evaluate_repconc as E, ops
    r = np.random.default_rng(5)
    D, M, K, N, nq = 128, 8, 256, 20000, 200
    ds = D // M
    docs = r.standard_normal((N, D), dtype=np.float32)
    c = np.ascontiguousarray(docs[:K].reshape(K, M, ds).transpose(1, 0, 2))
    rel = r.integers(0, N, size=nq)
    queries = docs[rel] + 0.5 * r.standard_normal((nq, D), dtype=np.float32)

    class Cfg:
        hidden_size, MCQ_M, MCQ_K = D, M, K

    class Model:
        config = Cfg()
        centroids = torch.nn.Parameter(torch.from_numpy(c))

    codes = ops.nn_assign(_dev(docs), _dev(c), uint8=True).cpu().numpy()
    assert np.array_equal(codes, oracle.nn_assign(docs, c).astype(np.uint8))
    index = E.initialize_index(Model())
    E.add_docs(index, codes[:12000])
    E.add_docs(index, codes[12000:])
    assert index.ntotal == N
    gpu = E.load_index_to_gpu(E.from_pq_to_ivfpq(index))
    corpus_ids = np.arange(N)[::-1].copy()                     # arbitrary external ids
    qids = np.arange(nq)
    s, ids = E.batch_search(qids, queries, corpus_ids, gpu, topk=10, batch_size=64)
    os_, oi = oracle.adc_search(queries, c, codes, 10)
    assert np.array_equal(s, os_) and np.array_equal(ids, corpus_ids[oi])
    mrr = oracle.mrr_at_k(ids, corpus_ids[rel], 10)
    assert mrr == oracle.mrr_at_k(corpus_ids[oi], corpus_ids[rel], 10) and mrr > 0.3
    # the pipelined multi-batch path (copy-back of batch i overlapping the scan of batch i+1) returns what the
    # single-batch path and the generic per-batch loop (non-int64 ids) return
    s1, ids1 = E.batch_search(qids, queries, corpus_ids, gpu, topk=10, batch_size=nq)
    assert np.array_equal(s1, s) and np.array_equal(ids1, ids)
    s3, ids3 = E.batch_search(qids, queries, corpus_ids.astype(np.int32), gpu, topk=10, batch_size=37)
    assert np.array_equal(s3, s) and np.array_equal(ids3, ids)
    s4, ids4 = E.batch_search(qids, queries, corpus_ids, gpu, topk=10, batch_size=37)
    assert np.array_equal(s4, s) and np.array_equal(ids4, ids)
    # host-side IndexPQ.search (the --cpu_search call site) lands on the same kernels
    s2, i2 = index.search(queries[:5], 10)
    assert np.array_equal(i2, oi[:5])


def test_shard_merge(oracle):
    from repconc_b200 import evaluate_repconc as E
    from repconc_b200.faiss_compat import GpuIndexPQ
    case = GC.ADC_CASES["adc_m8"]
    q, c, codes = GC.adc_inputs(case)
    W, k = 4, 100
    ss, ii = [], []
    for w in range(W):
        lo, hi = E.shard_bounds(len(codes), w, W)
        s, i = GpuIndexPQ(_dev(codes[lo:hi]), _dev(c), id_offset=lo).search_tensor(_dev(q), k)
        ss.append(s)
        ii.append(i)
    ms, mi = E.merge_shard_results(torch.stack(ss), torch.stack(ii))
    os_, oi = oracle.adc_search(q, c, codes, k)
    assert np.array_equal(ms.cpu().numpy(), os_) and np.array_equal(mi.cpu().numpy(), oi)


def test_map_ids_on_device_matches_numpy_indexing():
    """search(..., corpus_ids) maps positions to external ids on the device exactly like the reference's
    `corpus_ids[x]` (evaluate_repconc.py:183), including the -1 padding wrapping to corpus_ids[-1]."""
    from repconc_b200 import evaluate_repconc as E
    case = dict(GC.ADC_CASES["adc_m8"], N=37, nq=3)
    q, c, codes = GC.adc_inputs(case)
    idx = _index(c, codes)
    corpus_ids = (np.arange(37, dtype=np.int64) * 7 + 1000)
    s_pos, i_pos = idx.search(q, 50)                       # positions, padded with -1 beyond N
    s_map, i_map = E.search(np.arange(3), q, corpus_ids, idx, 50)
    assert np.array_equal(s_map, s_pos) and np.array_equal(i_map, corpus_ids[i_pos])
    corpus_ids[5] = -42                                     # in-place edit must not hit a stale device copy
    _, i_map2 = E.search(np.arange(3), q, corpus_ids, idx, 50)
    assert np.array_equal(i_map2, corpus_ids[i_pos])


def test_batch_search_sees_in_place_id_edits_anywhere():
    """batch_search maps ids with the cached device copy of `corpus_ids` while a helper validates it against the
    caller's array (full comparison); an in-place edit at ANY position -- also one a sampled fingerprint would miss --
    must show up in the results of the next call (the call is repeated with the refreshed copy)."""
    from repconc_b200 import evaluate_repconc as E
    case = dict(GC.ADC_CASES["adc_m8"], N=5000, nq=40)
    q, c, codes = GC.adc_inputs(case)
    idx = _index(c, codes)
    corpus_ids = np.arange(5000, dtype=np.int64) * 3 + 7
    s0, i0 = E.batch_search(np.arange(40), q, corpus_ids, idx, 20, 16)       # 3 batches -> the pipelined path
    _, pos = idx.search(q, 20)
    assert np.array_equal(i0, corpus_ids[pos])
    hit = int(pos[7, 3])
    corpus_ids[hit] = -99                                                     # a position that IS in the results
    corpus_ids[4999 if hit != 4999 else 4998] = -7                            # and one far from any sample stride
    s1, i1 = E.batch_search(np.arange(40), q, corpus_ids, idx, 20, 16)
    assert np.array_equal(i1, corpus_ids[pos]) and i1[7, 3] == -99 and np.array_equal(s1, s0)
    s2, i2 = E.batch_search(np.arange(40), q, corpus_ids, idx, 20, 16)       # steady state again
    assert np.array_equal(i2, i1)


def test_full_size_properties():
    """BASELINE configs[1] size (8,841,823 docs, M=48, k=1000): size-independent properties --
    sorted descending, ids unique and in range, every returned score equals an independent dense
    re-computation, the k-th score is the true k-th largest of the whole corpus, sharded == unsharded,
    and the search is idempotent."""
    from repconc_b200 import _lib, ops
    from repconc_b200 import evaluate_repconc as E
    from repconc_b200.faiss_compat import GpuIndexPQ
    lib = _lib.load()
    N, M, ds, nq, k = 8_841_823, 48, 16, 64, 1000
    gen = torch.Generator(device="cuda").manual_seed(42)
    codes = torch.randint(0, 256, (N, M), generator=gen, device="cuda", dtype=torch.uint8)
    c = torch.randn((M, 256, ds), generator=gen, device="cuda")
    q = torch.randn((nq, M * ds), generator=gen, device="cuda")
    idx = GpuIndexPQ(codes, c)
    s, i = idx.search_tensor(q, k)
    assert idx.last_stats["filtered"] == nq and idx.last_stats["dense"] == 0
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    assert int(i.min()) >= 0 and int(i.max()) < N
    assert all(len(torch.unique(i[r])) == k for r in range(nq))
    # dense fp32 scores of 4 queries over the whole corpus through the un-quantised scan kernel
    lut = torch.empty((4, M, 256), device="cuda")
    _lib.check(lib.rc_adc_lut(q.data_ptr(), q.stride(0), c.data_ptr(), 4, M, 256, ds, lut.data_ptr(), ops._stream()),
               "lut")
    dense = torch.empty((4, N), device="cuda")
    _lib.check(lib.rc_adc_scores(lut.data_ptr(), codes.data_ptr(), 4, 0, N, M, dense.data_ptr(), ops._stream()),
               "scores")
    assert torch.equal(torch.gather(dense, 1, i[:4]), s[:4])            # bit-exact scores
    top = torch.topk(dense, k, dim=1).values
    assert torch.equal(top, s[:4])                                      # the true top-k score multiset
    # 3 shards merged == unsharded
    parts = []
    for w in range(3):
        lo, hi = E.shard_bounds(N, w, 3)
        parts.append(GpuIndexPQ(codes[lo:hi], c, id_offset=lo).search_tensor(q, k))
    ms, mi = E.merge_shard_results(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(ms, s) and torch.equal(mi, i)
    s2, i2 = idx.search_tensor(q, k)
    assert torch.equal(s2, s) and torch.equal(i2, i)


def test_gpu_side_add_and_centroid_refresh(oracle):
    """Device-side append of codes (corpus encoding never leaves HBM) and the in-place centroid refresh
    that replaces JPQ's per-step index re-clone (finetune_jpq.py:208-214)."""
    from repconc_b200 import ops
    from repconc_b200.faiss_compat import GpuIndexPQ
    case = GC.ADC_CASES["adc_m8"]
    q, c, codes = GC.adc_inputs(case)
    idx = GpuIndexPQ(_dev(codes[:100]), _dev(c))
    for lo in range(100, len(codes), 3300):
        idx.add(_dev(codes[lo:lo + 3300]))
    assert idx.ntotal == len(codes)
    s, i = idx.search(q, 10)
    os_, oi = oracle.adc_search(q, c, codes, 10)
    assert np.array_equal(s, os_) and np.array_equal(i, oi)
    c2 = (c * 1.5 + 0.25).astype(np.float32)
    idx.set_centroids(_dev(c2))
    s2, i2 = idx.search(q, 10)
    o2 = oracle.adc_search(q, c2, codes, 10)
    assert np.array_equal(s2, o2[0]) and np.array_equal(i2, o2[1])


@pytest.mark.parametrize("M,ds,N,nq,k", [(48, 4, 262_145, 1, 1), (48, 4, 300_001, 17, 2048), (32, 4, 270_000, 33, 7),
                                          (96, 2, 265_000, 5, 2048), (64, 2, 290_000, 16, 1), (8, 8, 400_000, 49, 1000)])
def test_filtered_scan_edge_shapes_vs_oracle(oracle, M, ds, N, nq, k):
    """edge shapes of the filtered path: a single query (one tile with 15 empty slots), k = 1, k = 2048 (the largest k
    the re-score's selection buffer holds), ragged tiles / splits, one document above the dense-path limit"""
    r = np.random.default_rng(7000 + M + nq)
    c = r.standard_normal((M, 256, ds), dtype=np.float32)
    codes = r.integers(0, 256, size=(N, M), dtype=np.uint8)
    q = r.standard_normal((nq, M * ds), dtype=np.float32)
    idx = _index(c, codes)
    s, i = idx.search(q, k)
    os_, oi = oracle.adc_search(q, c, codes, k)
    assert np.array_equal(s, os_) and np.array_equal(i, oi), (M, nq, k)
    assert idx.last_stats["sample"] > 0


def test_single_process_multi_device_index_vs_oracle(oracle):
    """`load_index_to_gpu(index)` from ONE process (the unchanged evaluator, run_repconc_eval.py:93-100,155) returns a
    MultiGpuIndexPQ when several GPUs are visible: corpus shards, one host thread per device, per-shard lists merged
    on the first device.  On a single-GPU box the shards share the device (the threading, the global ids, the
    merge and the pipelined batch_search are the same code)."""
    from repconc_b200 import evaluate_repconc as E
    from repconc_b200 import faiss_compat as F
    case = dict(GC.ADC_CASES["adc_m8"], N=7001, nq=50)
    q, c, codes = GC.adc_inputs(case)
    host = F.IndexPQ(case["D"], case["M"], 8, F.METRIC_INNER_PRODUCT)
    F.copy_array_to_vector(np.ascontiguousarray(c).ravel(), host.pq.centroids)
    host.is_trained = True
    E.add_docs(host, codes[:4000])
    E.add_docs(host, codes[4000:])
    ndev = torch.cuda.device_count()
    multi = F.MultiGpuIndexPQ.from_host(host, list(range(ndev)) if ndev >= 2 else [0, 0, 0])
    assert multi.ntotal == 7001 and len(multi.shards) >= 2
    corpus_ids = np.arange(7001, dtype=np.int64) * 5 + 3
    for k in (10, 300):
        os_, oi = oracle.adc_search(q, c, codes, k)
        s, i = multi.search(q, k)
        assert np.array_equal(s, os_) and np.array_equal(i, oi)
        st, it = multi.search(_dev(q), k)                                       # CUDA tensors in -> CUDA tensors out
        assert np.array_equal(st.cpu().numpy(), os_) and np.array_equal(it.cpu().numpy(), oi)
        sb, ib = E.batch_search(np.arange(50), q, corpus_ids, multi, k, 16)     # 4 batches, ids mapped on the device
        assert np.array_equal(sb, os_) and np.array_equal(ib, corpus_ids[oi])
    if ndev >= 2:
        assert isinstance(E.load_index_to_gpu(host), F.MultiGpuIndexPQ)
    assert isinstance(E.load_index_to_gpu(host, 0), F.GpuIndexPQ)
