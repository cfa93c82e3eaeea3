"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side mirror of
the reference interface behaves (module surface, Faiss-free index objects, index file round trip,
shard bounds), and the product refuses to run without CUDA instead of falling back."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from repconc_b200.build import build_library
    return build_library()


def test_header_symbols_are_exported(built):
    import ctypes
    from repconc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "repconc_b200.h")).read()
    declared = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/repconc_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = _lib.load()
    assert b"sm_100a" in l.rc_version()
    assert l.rc_launch_count() >= 0


def test_argument_validation_without_gpu(built):
    """Validation happens before any CUDA call, so it is testable here."""
    from repconc_b200 import _lib
    l = _lib.load()
    assert l.rc_nn_assign(None, 0, None, 4, 8, 256, 16, None, None, None) == -1
    assert b"null pointer" in l.rc_last_error()
    assert l.rc_adc_lut(1, 128, 1, 4, 8, 64, 16, 1, None) == -1 and b"256" in l.rc_last_error()
    assert l.rc_sinkhorn_state_bytes(0, 8, 256) == 0
    n = l.rc_sinkhorn_state_bytes(8192, 48, 256)
    assert n >= 2 * 48 * 256 * 8 + 48 * 8192 * 8
    assert l.rc_adc_search_workspace_bytes(1024, 8841823, 48, 256, 1000) > 1024 * 8192 * 8
    assert l.rc_topk_merge(1, 1, 16, 4, 1000, 1, 1, None) == -3          # W*k above the sortable cap


def test_no_cpu_fallback():
    from repconc_b200 import RepCONC, _lib, decode, ops
    x = torch.zeros((4, 128))
    c = torch.zeros((8, 256, 16))
    with pytest.raises(_lib.RepconcLibraryError):
        ops.nn_assign(x, c)
    with pytest.raises(_lib.RepconcLibraryError):
        ops.constrained_assign(x, c, 0.003, 5, distributed=False)
    with pytest.raises(_lib.RepconcLibraryError):
        decode(torch.zeros((4, 8), dtype=torch.long), c)
    with pytest.raises(NotImplementedError):
        decode([[0] * 8], c)                                   # modeling_repconc.py:183


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "repconc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "librepconc_oracle" not in src, f


def _cfg(D=128, M=8, K=256, metric="METRIC_IP"):
    from transformers import PretrainedConfig
    cfg = PretrainedConfig(hidden_size=D)
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = M, K, metric
    return cfg


class _Enc(torch.nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.lin = torch.nn.Linear(4, cfg.hidden_size)

    def forward(self, input_ids, attention_mask):
        return self.lin(input_ids.float())

    def save_pretrained(self, d):
        os.makedirs(d, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(d, "enc.bin"))


def test_module_surface(tmp_path):
    """Constructor, attributes, state-dict keys, save/load: modeling_repconc.py:30-45,118-134."""
    from repconc_b200 import QuantizeOutput, RepCONC
    cfg = _cfg()
    torch.manual_seed(0)
    m = RepCONC(cfg, _Enc(cfg), True, 0.003, 100)
    assert tuple(m.centroids.shape) == (8, 256, 16) and m.centroids.requires_grad
    assert torch.equal(m.rotation, torch.eye(128))
    assert (m.use_constraint, m.sk_epsilon, m.sk_iters) == (True, 0.003, 100)
    keys = set(m.state_dict())
    assert {"centroids", "rotation"} <= keys and all(
        k in ("centroids", "rotation") or k.startswith("dense_encoder.") for k in keys)
    # forward without codes requested never touches the CUDA path
    out = m(torch.ones((3, 4), dtype=torch.long), torch.ones((3, 4), dtype=torch.long))
    assert isinstance(out, QuantizeOutput) and out.discrete_codes is None and out.quantized_embeds is None
    assert tuple(out.continuous_embeds.shape) == (3, 128)
    # METRIC_CENTROID_COS normalises centroids at construction and sub-vectors in forward (:42-43,:99-100)
    cfg2 = _cfg(metric="METRIC_CENTROID_COS")
    m2 = RepCONC(cfg2, _Enc(cfg2), False, None, None)
    assert torch.allclose(m2.centroids.norm(dim=-1), torch.ones(8, 256), atol=1e-5)
    o2 = m2(torch.ones((3, 4), dtype=torch.long), None)
    assert torch.allclose(o2.continuous_embeds.reshape(3, 8, 16).norm(dim=-1), torch.ones(3, 8), atol=1e-5)
    # save / load round trip through the same file names
    m.save_pretrained(str(tmp_path))
    assert os.path.exists(tmp_path / "pytorch_model.bin") and os.path.exists(tmp_path / "config.json")

    def loader(path):
        e = _Enc(cfg)
        e.load_state_dict(torch.load(os.path.join(path, "enc.bin")))
        return e
    m3 = RepCONC.from_pretrained(str(tmp_path), False, None, None, encoder_loader=loader)
    assert torch.equal(m3.centroids, m.centroids) and m3.use_constraint is False


def test_index_objects_and_file_round_trip(tmp_path):
    from repconc_b200 import evaluate_repconc as E, faiss_compat as faiss
    cfg = _cfg()

    class Model:
        config = cfg
        centroids = torch.nn.Parameter(torch.randn(8, 256, 16))

    index = E.initialize_index(Model())
    assert (index.pq.M, index.pq.d, index.pq.nbits, index.pq.code_size, index.ntotal) == (8, 128, 8, 8, 0)
    assert np.array_equal(faiss.vector_to_array(index.pq.centroids), Model.centroids.detach().numpy().ravel())
    r = np.random.default_rng(0)
    a, b = r.integers(0, 256, (10, 8), dtype=np.uint8), r.integers(0, 256, (7, 8), dtype=np.uint8)
    E.add_docs(index, a)
    E.add_docs(index, b)
    assert index.ntotal == 17 and np.array_equal(index.code_array(), np.vstack([a, b]))
    with pytest.raises(AssertionError):
        E.add_docs(index, np.zeros((3, 9), np.uint8))           # evaluate_repconc.py:93
    # the reference's own add_docs body (Faiss vector API, evaluate_repconc.py:94-98) works on the shim
    M, ntotal = index.pq.code_size, index.ntotal
    index.codes.resize((ntotal + 2) * M)
    codes = faiss.vector_to_array(index.codes)
    codes.reshape(-1, M)[-2:] = 9
    faiss.copy_array_to_vector(codes, index.codes)
    index.ntotal += 2
    assert index.ntotal == 19 and np.all(index.code_array()[-2:] == 9)
    # replace_pq_centroids (run_repconc_eval.py:123-127)
    newc = r.standard_normal(8 * 256 * 16).astype(np.float32)
    faiss.copy_array_to_vector(newc, index.pq.centroids)
    assert np.array_equal(index.pq.centroid_array().ravel(), newc)
    ivf = E.from_pq_to_ivfpq(index)
    assert ivf.ntotal == 19 and ivf.pq.M == 8 and ivf.nlist == 1
    # JPQ.__init__ (finetune_jpq.py:160-161)
    jc = faiss.vector_to_array(index.codes).astype(np.int64).reshape(-1, 8)
    assert jc.shape == (19, 8)
    # index file
    p = str(tmp_path / "index")
    faiss.write_index(index, p)
    back = faiss.read_index(p)
    assert back.ntotal == 19 and np.array_equal(back.code_array(), index.code_array())
    assert np.array_equal(back.pq.centroid_array(), index.pq.centroid_array()) and back.is_trained
    assert os.path.getsize(p) == 4 + 28 + 5 + 24 + 8 + 4 * newc.size + 8 + 19 * 8 + 9
    with pytest.raises(NotImplementedError):
        faiss.IndexPQ(128, 8, 4)
    with pytest.raises(NotImplementedError):
        faiss.IndexPQ(128, 8, 8, faiss.METRIC_L2)


def test_shard_bounds_cover_exactly():
    from repconc_b200.evaluate_repconc import shard_bounds
    for n, w in [(64_000_000, 8), (8841823, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_warmup_host_helpers_match_the_oracle():
    """The pure-torch pieces of the warm-up (empty-cluster split, Procrustes, seeded rotation, initial
    centroids) run on CPU tensors too: same values as oracle/warmup_np.py."""
    import numpy as np
    import torch
    from oracle import warmup_np as W
    from repconc_b200 import warmup
    r = np.random.default_rng(0)
    c = r.standard_normal((3, 16, 6)).astype(np.float32)
    counts = r.integers(1, 50, (3, 16))
    counts[0, 4] = counts[0, 9] = counts[2, 0] = 0
    counts[1, :] = 1                                   # nothing to split from: left alone
    counts[1, 7] = 0
    got = warmup.split_empty_clusters(torch.from_numpy(c), torch.from_numpy(counts.astype(np.int32))).numpy()
    want = W.split_empty_clusters(c, counts)
    assert np.array_equal(got, want)
    assert np.array_equal(got[1], c[1])
    x = r.standard_normal((500, 12)).astype(np.float32)
    rec = (x @ np.linalg.qr(r.standard_normal((12, 12)))[0]).astype(np.float32)
    A = warmup.procrustes(torch.from_numpy(x), torch.from_numpy(rec)).numpy()
    np.testing.assert_allclose(A, W.procrustes(x, rec), atol=1e-5)
    np.testing.assert_allclose(x @ A.T, rec, atol=1e-4)           # recovers the rotation that generated `rec`
    R = warmup.random_rotation(24, 5, "cpu")
    np.testing.assert_allclose((R @ R.t()).numpy(), np.eye(24), atol=1e-5)
    assert torch.equal(R, warmup.random_rotation(24, 5, "cpu"))
    c0 = warmup.initial_centroids(torch.from_numpy(x), 3, 16, seed=2)
    assert c0.shape == (3, 16, 4)
    rows = {tuple(np.round(v, 6)) for v in x.reshape(500, 3, 4)[:, 0, :]}
    assert all(tuple(np.round(v, 6)) in rows for v in c0[0].numpy())     # centroids are training points


def _u8_pos(m, M, eb, lpd):
    """mirror of csrc/adc.cu: u8_pos -- table position of sub-vector m in a tile of 8-bit fields"""
    bg = 128 // eb
    dpp, ne = bg // lpd, M // lpd
    j, i = divmod(m, ne)
    return bg * (i // dpp) + lpd * (i % dpp) + j


@pytest.mark.parametrize("M,eb,lpd", [(48, 16, 2), (48, 16, 4), (48, 16, 8), (32, 16, 2), (8, 16, 2), (40, 16, 2),
                                      (64, 8, 4), (96, 8, 4), (80, 8, 4), (64, 8, 8)])
def test_u8_scan_layout_is_conflict_free_for_any_codes(M, eb, lpd):
    """The invariant the 8-bit-field ADC scan rests on: with the table laid out [code][position] and
    document slot s of a shared-memory phase walking its bytes in the order i = t ^ s, the lanes of a phase (128 bytes
    worth of entries) hit pairwise different bank groups at EVERY step, whatever the codes are."""
    bg = 128 // eb                      # entries per 128-byte line = lanes per phase
    dpp, ne = bg // lpd, M // lpd       # documents per phase, code bytes per lane
    assert sorted(_u8_pos(m, M, eb, lpd) for m in range(M)) == list(range(M))        # a permutation of the sub-vectors
    rng = np.random.default_rng(M * 100 + eb + lpd)
    for _ in range(50):
        codes = rng.integers(0, 256, size=(dpp, M))
        seen = np.zeros((dpp, M), dtype=int)
        for t in range(ne):
            groups = set()
            for s in range(dpp):
                for j in range(lpd):
                    m = ne * j + (t ^ s)                                              # lane j's byte t ^ s
                    seen[s, m] += 1
                    groups.add((int(codes[s, m]) * M + _u8_pos(m, M, eb, lpd)) % bg)
            assert len(groups) == bg, (t, groups)
        assert (seen == 1).all()                                                      # every sub-vector exactly once


@pytest.mark.parametrize("M,acc", [(48, 2), (32, 2), (96, 2), (8, 2), (48, 3)])
def test_u8_filter_bound_never_excludes_a_better_document(M, acc):
    """The exactness argument of the 8-bit-field ADC filter, replayed in numpy with the kernels'
    arithmetic: tables quantised to q = clamp(round((v - lo_m) / step), 0, QMAX), step = max_m range / QMAX,
    QMAX = 255 / acc.  For EVERY document  |exact - (sum lo + step * sum q)| <= step * (0.501 M + slack), hence a
    document with integer sum <= T - 1 cannot beat ub = sum lo + step * (T - 1 + 0.501 M + slack): keeping what is
    strictly above ub loses nothing."""
    rng = np.random.default_rng(M * 10 + acc)
    qmax = 255 // acc
    N = 20000
    for trial in range(3):
        scale = [1.0, 37.5, 1e-3][trial]
        lut = (rng.standard_normal((M, 256)) * scale).astype(np.float32)
        if trial == 2:
            lut[0] *= 50                                        # one sub-vector dominates the range
        codes = rng.integers(0, 256, size=(N, M))
        lo = lut.min(1)
        rng_m = (lut.max(1) - lo).max()
        step = np.float32(max(rng_m, 1e-30)) / np.float32(qmax)
        q8 = np.clip(np.rint((lut - lo[:, None]) / step), 0, qmax).astype(np.int64)
        assert q8.max() <= qmax and acc * qmax <= 255          # `acc` entries add up inside a byte
        exact = np.zeros(N, dtype=np.float32)
        for m in range(M):                                     # fp32, m ascending, like the scan / the oracle
            exact = (exact + lut[m, codes[:, m]]).astype(np.float32)
        s8 = q8[np.arange(M)[None, :], codes].sum(1)
        sumlo = lo.astype(np.float64).sum()
        sumabs = np.maximum(np.abs(lo), np.abs(lut.max(1))).astype(np.float64).sum()
        slack = M * 1.2e-7 * sumabs / float(step) + 1.0
        err = np.abs(exact.astype(np.float64) - (sumlo + float(step) * s8))
        assert (err <= float(step) * (0.501 * M + slack)).all()
        # threshold from the r-th largest integer sum, shifted as adc_u8_threshold_kernel does
        k = 100
        T = int(np.sort(s8)[-3 * k]) - int(np.ceil(0.501 * M + slack)) - 1
        ub = sumlo + float(step) * (T - 1 + 0.501 * M + slack)
        excluded = s8 <= T - 1
        assert (exact[excluded].astype(np.float64) <= ub).all()
        kept = (~excluded) & (exact.astype(np.float64) > ub)
        topk = np.argsort(-exact, kind="stable")[:k]
        assert kept.sum() >= k and kept[topk].all()
