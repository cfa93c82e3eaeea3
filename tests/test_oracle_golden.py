"""Pins the CPU oracle (oracle/) against fixtures produced by the REFERENCE itself
(tools/gen_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest

from tests import golden_cases as GC
from tests.conftest import golden


@pytest.mark.parametrize("name", list(GC.ASSIGN_CASES))
def test_table_and_centring_bit_exact(oracle, name):
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    assert GC.digest(x, c) == str(g["input_sha"]), "seeded inputs drifted (numpy RNG?)"
    table = oracle.dist_table(x, c)
    # fp32 table must be bit-identical to ((x-c)**2).sum(-1) of the reference
    assert GC.digest(table) == str(g["table_sha"])
    assert np.array_equal(table[:, :2, :], g["table_head"])
    mx, mn = oracle.table_minmax(table)
    assert np.array_equal(mx, g["max"]) and np.array_equal(mn, g["min"])
    centred = oracle.center_table(table)
    assert GC.digest(centred) == str(g["centred_sha"])


@pytest.mark.parametrize("name", list(GC.ASSIGN_CASES))
def test_assign_codes_bit_exact(oracle, name):
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    assert np.array_equal(oracle.nn_assign(x, c), g["codes_nn"].astype(np.int64))
    res = oracle.constrained_assign(x, c, case["eps"], case["iters"], return_q=True)
    assert not res["nonfinite"]
    assert np.array_equal(res["codes"], g["codes_conc"].astype(np.int64))
    # Q itself: fp64, tolerance 1e-9 relative (libm vs sleef exp, summation order)
    Q = res["Q"]
    np.testing.assert_allclose(Q.sum(2), g["q_rowsum"], rtol=1e-9)
    np.testing.assert_allclose(Q.transpose(0, 2, 1)[:, :2, :], g["q_head"], rtol=1e-9, atol=1e-300)
    assert abs(Q.sum(1) - 1).max() < 1e-12


@pytest.mark.parametrize("name", ["ds16_b512", "ds12_b300", "ds5_k64"])
def test_numpy_restatement_agrees(oracle, name):
    from oracle import oracle_np as ON
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    assert GC.digest(ON.dist_table(x, c)) == str(g["table_sha"])
    assert np.array_equal(ON.constrained_assign(x, c, case["eps"], case["iters"]),
                          g["codes_conc"].astype(np.int64))
    assert np.array_equal(ON.nn_assign(x, c), g["codes_nn"].astype(np.int64))


def test_distributed_reference_equals_global_batch(oracle):
    """The reference run on 2 gloo ranks (all_reduce MAX/MIN/SUM, B *= world) gave the same
    codes as the oracle on the concatenated batch."""
    case = GC.DIST_CASES["dist2_ds16"]
    g = golden("assign_dist2_ds16")
    x, c = GC.assign_inputs(case)
    assert GC.digest(x, c) == str(g["input_sha"])
    res = oracle.constrained_assign(x, c, case["eps"], case["iters"])
    assert np.array_equal(res["codes"], g["codes_conc"].astype(np.int64))
    assert np.array_equal(res["codes"], g["codes_single"].astype(np.int64))
    # externally supplied extrema (the all-reduced ones) reproduce a rank's slab
    half = case["B"] // 2
    full = oracle.constrained_assign(x, c, case["eps"], case["iters"])
    assert np.array_equal(full["max"], np.maximum(
        oracle.table_minmax(oracle.dist_table(x[:half], c))[0],
        oracle.table_minmax(oracle.dist_table(x[half:], c))[0]))


@pytest.mark.parametrize("name", list(GC.MSE_CASES))
def test_decode_and_mse(oracle, name):
    case = GC.MSE_CASES[name]
    g = golden(name)
    x, c, gr, codes = GC.mse_inputs(case)
    assert GC.digest(x, c, gr, codes) == str(g["input_sha"])
    q = oracle.decode(codes, c)
    assert np.array_equal(q, g["quantized"])          # gather is exact
    r = oracle.mse_surrogate(x, q, gr, codes, case["K"], case["w"], case["scale"])
    np.testing.assert_allclose(r["mse"], g["mse"], rtol=1e-5)
    np.testing.assert_allclose(r["surrogate"], g["surrogate"], rtol=1e-4)
    np.testing.assert_allclose(r["grad_x"], g["grad_x"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(r["grad_c"], g["grad_c"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name", list(GC.ADC_CASES))
def test_adc_search_against_decode_matmul(oracle, name):
    from oracle import oracle_np as ON
    case = GC.ADC_CASES[name]
    g = golden(name)
    q, c, codes = GC.adc_inputs(case)
    assert GC.digest(q, c, codes) == str(g["input_sha"])
    for k in case["ks"]:
        s, ids = oracle.adc_search(q, c, codes, k)
        gs, gi = g[f"scores_k{k}"], g[f"ids_k{k}"].astype(np.int64)
        # ADC scores within 1e-4 relative of <q, decode(codes)> (north_star tolerance)
        np.testing.assert_allclose(s, gs, rtol=1e-4, atol=1e-4)
        assert np.all(np.diff(s, axis=1) <= 0)
        # ids agree except inside near-ties (|score gap| below fp32 noise)
        diff = ids != gi
        assert np.all(np.abs(s[diff] - gs[diff]) <= 1e-4 * np.abs(gs[diff]) + 1e-4)
        assert diff.mean() < 0.01
    # bit-level: C and numpy restatements of the fp32 accumulation agree exactly
    full = ON.adc_scores(q[:4], c, codes)
    s, ids = oracle.adc_search(q[:4], c, codes, 5)
    assert np.array_equal(s, np.take_along_axis(full, ids, 1))
    # exact duplicate documents tie -> smaller id first
    N = case["N"]
    s_all, i_all = oracle.adc_search(q[:2], c, codes, N)
    for r in range(2):
        pos = {int(d): p for p, d in enumerate(i_all[r])}
        assert pos[N // 3] + 1 == pos[N // 2]


def test_adc_edge_cases(oracle):
    case = dict(GC.ADC_CASES["adc_m8"], N=37, nq=3)
    q, c, codes = GC.adc_inputs(case)
    s, ids = oracle.adc_search(q, c, codes, 50)      # k > N pads like Faiss: (-FLT_MAX, -1)
    assert np.all(ids[:, 37:] == -1) and np.all(s[:, 37:] == np.finfo(np.float32).min)
    assert sorted(ids[0, :37]) == list(range(37))
    s2, ids2 = oracle.adc_search(q, c, codes, 5, id_offset=1000)
    assert np.array_equal(ids2, ids[:, :5] + 1000) and np.array_equal(s2, s[:, :5])
    # shard merge == unsharded search
    sa, ia = oracle.adc_search(q, c, codes[:20], 7)
    sb, ib = oracle.adc_search(q, c, codes[20:], 7, id_offset=20)
    sm, im = oracle.topk_merge(np.stack([sa, sb]), np.stack([ia, ib]))
    assert np.array_equal(sm, s[:, :7]) and np.array_equal(im, ids[:, :7])


def test_mrr_definition(oracle):
    run = np.array([[5, 3, 9], [1, 2, 3], [7, 8, 9]])
    assert oracle.mrr_at_k(run, np.array([3, 1, 4]), k=10) == round((0.5 + 1.0 + 0.0) / 3, 5)
    assert oracle.mrr_at_k(run, np.array([9, 3, 9]), k=2) == 0.0


def test_sum_order_model(oracle):
    from oracle import oracle_np as ON
    r = np.random.default_rng(3)
    for n in (1, 2, 5, 7, 8, 12, 16, 24, 48, 96, 192, 384, 768):
        v = (r.standard_normal((9, n), dtype=np.float32) ** 2).astype(np.float32)
        want = ON.sum_last_torch_order(v)
        got = np.array([oracle.sum_torch_order(row) for row in v], np.float32)
        assert np.array_equal(want, got), n


def test_warmup_oracle_properties(oracle):
    """oracle/warmup_np.py (k-means PQ + OPQ restatement, parity unpinned w.r.t. Faiss): Lloyd's objective
    never increases, empty clusters are re-seeded, the Procrustes rotation is orthogonal and optimal."""
    from oracle import warmup_np as W
    r = np.random.default_rng(1)
    centres = r.standard_normal((40, 32)).astype(np.float32)
    x = (centres[r.integers(0, 40, 4096)] + 0.35 * r.standard_normal((4096, 32))).astype(np.float32)
    c0 = x[r.permutation(4096)[:16]].reshape(16, 4, 8).transpose(1, 0, 2).copy()
    c, objs = W.train_pq(x, c0, 8)
    assert all(b <= a * (1 + 1e-9) for a, b in zip(objs, objs[1:]))
    far = c0.copy()
    far[:, 5, :] += 1000.0
    c1, _, _, counts = W.lloyd_step(x, far)
    assert (counts[:, 5] == 0).all() and np.abs(c1[:, 5, :]).max() < 100.0
    A = np.linalg.qr(r.standard_normal((32, 32)))[0].astype(np.float32)
    A2, c2, err = W.opq_alternation(x, A, c0, 4)
    np.testing.assert_allclose(A2 @ A2.T, np.eye(32), atol=1e-5)
    # A2 is the best rotation for the reconstruction it was fitted to: no worse than the one it replaces
    xp = x @ A.T
    codes = oracle.nn_assign(np.ascontiguousarray(xp), c2)
    rec = W.decode(codes, c2)
    assert ((x @ A2.T - rec) ** 2).sum() <= ((x @ A.T - rec) ** 2).sum() * (1 + 1e-6)


@pytest.mark.parametrize("name", list(GC.ENCODE_CASES))
def test_encode_epilogue_restatement_against_reference_forward(name):
    """oracle_np.encode_assign (rotation + COS normalisation + NN assign) vs the reference's RepCONC.forward fixtures:
    codes equal wherever the reference's two smallest distances are not a rounding apart (numpy's and torch's fp32
    GEMMs may sum in different orders)."""
    from oracle import oracle as O
    from oracle import oracle_np as ONP
    case = GC.ENCODE_CASES[name]
    g = golden("encode_" + name)
    x, rot, c = GC.encode_inputs(case)
    assert GC.digest(x, rot, c) == str(g["input_sha"])
    y, codes = ONP.encode_assign(x, rot, c, case["metric"] == "METRIC_CENTROID_COS", nn=O.nn_assign)
    bad = codes != g["codes"].astype(np.int64)
    assert not (bad & (g["gap"] >= 1e-4)).any()
    assert bad.mean() < 1e-3
    np.testing.assert_allclose(y[:4], g["rotated_head"], rtol=1e-4, atol=2e-6)
