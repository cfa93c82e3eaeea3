"""GPU parity of the assignment path (NN assign, table, centring, Sinkhorn) against the oracle and
the reference-generated golden fixtures.  Integer results are compared bit-exactly."""
import numpy as np
import pytest
import torch

from tests import golden_cases as GC
from tests.conftest import golden

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", list(GC.ASSIGN_CASES))
def test_nn_assign_matches_reference(name):
    from repconc_b200 import ops
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    codes = ops.nn_assign(_dev(x), _dev(c))
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (case["B"], case["M"])
    assert not codes.is_contiguous() or case["B"] == 1 or case["M"] == 1   # the reference's `.t()` view
    assert np.array_equal(codes.cpu().numpy(), g["codes_nn"].astype(np.int64))
    if case["K"] <= 256:
        u8 = ops.nn_assign(_dev(x), _dev(c), uint8=True)
        assert u8.dtype == torch.uint8 and u8.is_contiguous()
        assert np.array_equal(u8.cpu().numpy(), g["codes_nn"].astype(np.uint8))


@pytest.mark.parametrize("name", list(GC.ASSIGN_CASES))
def test_table_extrema_and_centring_bit_exact(name):
    from repconc_b200 import ops
    from repconc_b200.modeling_repconc import RepCONC
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    table, minmax, flags = ops.dist_table(_dev(x), _dev(c))
    t = table.cpu().numpy()
    assert GC.digest(t) == str(g["table_sha"])                     # fp32 table, bit for bit
    assert np.array_equal(minmax[0].cpu().numpy(), g["max"])
    assert np.array_equal(minmax[1].cpu().numpy(), g["min"])
    assert int(flags.item()) == 0
    centred = RepCONC.center_distance_for_constraint(table)
    assert GC.digest(centred.cpu().numpy()) == str(g["centred_sha"])


@pytest.mark.parametrize("name", list(GC.ASSIGN_CASES))
def test_constrained_assign_codes_bit_exact(name, oracle):
    from repconc_b200 import ops
    case = GC.ASSIGN_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    codes = ops.constrained_assign(_dev(x), _dev(c), case["eps"], case["iters"], distributed=False)
    got = codes.cpu().numpy()
    want = g["codes_conc"].astype(np.int64)
    bad = np.argwhere(got != want)
    # any mismatch must be reported together with the reference's own top-1/top-2 gap
    msg = "; ".join(f"(b={b},m={m}) gap={g['top2_gap'][b, m]:.3e}" for b, m in bad[:8])
    assert len(bad) == 0, f"{len(bad)} code mismatches vs reference: {msg}"
    # and the oracle agrees (same inputs, C restatement)
    assert np.array_equal(got, oracle.constrained_assign(x, c, case["eps"], case["iters"])["codes"])


@pytest.mark.parametrize("name", list(GC.ASSIGN_BIG_CASES))
def test_baseline_batch_bit_exact_vs_reference(name):
    """BASELINE-size batches (8192 x 768; M=48 = configs[2], M=96 = the per-rank slab of configs[4], T=50) against
    fixtures produced by the imported reference: fp32 table and centred table by sha256, extrema, NN codes and the
    constrained codes, bit for bit.  The sparse passes' behaviour (re-selections, pool use, drift) depends on B, so
    this is the size that has to be pinned."""
    from repconc_b200 import ops
    from repconc_b200.modeling_repconc import RepCONC
    case = GC.ASSIGN_BIG_CASES[name]
    g = golden("assign_" + name)
    x, c = GC.assign_inputs(case)
    assert GC.digest(x, c) == str(g["input_sha"])
    xd, cd = _dev(x), _dev(c)
    table, minmax, flags = ops.dist_table(xd, cd)
    assert GC.digest(table.cpu().numpy()) == str(g["table_sha"])
    assert np.array_equal(minmax[0].cpu().numpy(), g["max"]) and np.array_equal(minmax[1].cpu().numpy(), g["min"])
    centred = RepCONC.center_distance_for_constraint(table)
    assert GC.digest(centred.cpu().numpy()) == str(g["centred_sha"])
    del table, centred
    assert np.array_equal(ops.nn_assign(xd, cd, uint8=True).cpu().numpy(), g["codes_nn"])
    for stepwise in (False, True):          # the persistent solve and the begin/step/finish sequence
        if stepwise:
            kern = ops.CudaAssignKernels(xd, cd)
            kern.table()
            kern.begin(case["eps"])
            for _ in range(case["iters"] - 1):
                kern.step(case["eps"], case["B"])
            got = kern.finish(case["eps"], True).cpu().numpy()
            assert kern.read_flags() == 0
        else:
            got = ops.constrained_assign(xd, cd, case["eps"], case["iters"], distributed=False).cpu().numpy()
        want = g["codes_conc"].astype(np.int64)
        bad = np.argwhere(got != want)
        msg = "; ".join(f"(b={b},m={m}) gap={float(g['top2_gap'][b, m]):.3e}" for b, m in bad[:8])
        assert len(bad) == 0, f"{len(bad)} code mismatches vs reference (stepwise={stepwise}): {msg}"


def test_pool_exhaustion_flags_and_reruns_densely():
    """Survivor-pool exhaustion (rank-local data decides it) raises RC_FLAG_SPARSE_UNSAFE and the driver redoes the
    assignment with the dense pass: same codes as the reference."""
    from repconc_b200 import ops, _lib
    lib = _lib.load()
    case = GC.ASSIGN_CASES["m48_b1024"]
    g = golden("assign_m48_b1024")
    x, c = GC.assign_inputs(case)
    prev = lib.rc_sinkhorn_debug_pool_entries(1)
    try:
        kern = ops.CudaAssignKernels(_dev(x), _dev(c))
        kern.table()
        kern.solve(case["eps"], case["iters"])
        assert kern.read_flags() & ops.FLAG_SPARSE_UNSAFE
        got = ops.constrained_assign(_dev(x), _dev(c), case["eps"], case["iters"], distributed=False)
    finally:
        lib.rc_sinkhorn_debug_pool_entries(prev)
    assert np.array_equal(got.cpu().numpy(), g["codes_conc"].astype(np.int64))


def test_constrained_assign_uint8_and_strided_input():
    from repconc_b200 import ops
    case = GC.ASSIGN_CASES["ds16_b512"]
    g = golden("assign_ds16_b512")
    x, c = GC.assign_inputs(case)
    wide = torch.zeros((case["B"], case["D"] + 32), device="cuda")
    wide[:, : case["D"]] = _dev(x)
    view = wide[:, : case["D"]]                      # row stride > width
    u8 = ops.constrained_assign(view, _dev(c), case["eps"], case["iters"], distributed=False, uint8=True)
    assert np.array_equal(u8.cpu().numpy(), g["codes_conc"].astype(np.uint8))
    half = ops.nn_assign(_dev(x).half(), _dev(c))     # fp16 in -> promoted like the reference
    ref = ops.nn_assign(_dev(x).half().float(), _dev(c))
    assert torch.equal(half, ref)


def test_iters_zero_and_one(oracle):
    from repconc_b200 import ops
    case = GC.ASSIGN_CASES["ds8_b256"]
    x, c = GC.assign_inputs(case)
    for it in (0, 1, 2):
        got = ops.constrained_assign(_dev(x), _dev(c), 0.05, it, distributed=False).cpu().numpy()
        assert np.array_equal(got, oracle.constrained_assign(x, c, 0.05, it)["codes"]), it


def test_distributed_emulation_with_external_extrema(oracle):
    """A rank's slab with the all-reduced extrema and globally summed row sums gives the codes the
    reference computes on 2 ranks (golden `assign_dist2_ds16`)."""
    from repconc_b200 import ops
    case = GC.DIST_CASES["dist2_ds16"]
    g = golden("assign_dist2_ds16")
    x, c = GC.assign_inputs(case)
    half = case["B"] // 2
    kerns = [ops.CudaAssignKernels(_dev(x[:half]), _dev(c)), ops.CudaAssignKernels(_dev(x[half:]), _dev(c))]
    mms = [k.table() for k in kerns]
    mx = torch.maximum(mms[0][0], mms[1][0])
    mn = torch.minimum(mms[0][1], mms[1][1])
    for k in kerns:
        k.minmax[0].copy_(mx)
        k.minmax[1].copy_(mn)
    Ps = [k.begin(case["eps"]) for k in kerns]
    for _ in range(case["iters"] - 1):
        tot = Ps[0] + Ps[1]
        for P in Ps:
            P.copy_(tot)
        Ps = [k.step(case["eps"], case["B"]) for k in kerns]
    tot = Ps[0] + Ps[1]
    for P in Ps:
        P.copy_(tot)
    codes = torch.cat([k.finish(case["eps"], True) for k in kerns], 0).cpu().numpy()
    assert all(k.read_flags() == 0 for k in kerns)
    assert np.array_equal(codes, g["codes_conc"].astype(np.int64))


def test_balance_property_full_size():
    """Size-independent property at the BASELINE batch (8192 x 768, M=48): Sinkhorn's assignment is
    far better balanced than NN assign, every code is in range, and the run is deterministic."""
    from repconc_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((8192, 768), generator=gen, device="cuda")
    c = torch.randn((48, 256, 16), generator=gen, device="cuda")
    conc = ops.constrained_assign(x, c, 0.003, 50, distributed=False)
    again = ops.constrained_assign(x, c, 0.003, 50, distributed=False)
    assert torch.equal(conc, again)
    nn = ops.nn_assign(x, c)
    assert conc.min() >= 0 and conc.max() < 256

    def imbalance(codes):
        cnt = torch.stack([torch.bincount(codes[:, m], minlength=256) for m in range(48)]).double()
        return (cnt / (8192 / 256) - 1).abs().mean().item()
    assert imbalance(conc) < 0.5 * imbalance(nn)


def test_flags_and_errors():
    from repconc_b200 import ops, _lib
    x = torch.zeros((64, 128), device="cuda")
    c = torch.zeros((8, 256, 16), device="cuda")
    # all distances equal -> amplitude = 1e-5 > 0: fine; NaN input -> reference asserts on amplitude
    ops.constrained_assign(x, c, 0.003, 3, distributed=False)
    x[3, 5] = float("nan")
    with pytest.raises(AssertionError):
        ops.constrained_assign(x, c, 0.003, 3, distributed=False)
    with pytest.raises(_lib.RepconcLibraryError):
        ops.nn_assign(torch.zeros((4, 128)), c)            # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.nn_assign(torch.zeros((4, 100), device="cuda"), c)


def test_sinkhorn_algorithm_function_matches_reference_q():
    """Module-level sinkhorn_algorithm(out, eps, iters, False) -> Q (M,K,B): row sums and the first
    rows of Q against the reference's own Q (golden), 1e-9 relative."""
    from repconc_b200 import sinkhorn_algorithm
    from repconc_b200.modeling_repconc import RepCONC
    from repconc_b200 import ops
    case = GC.ASSIGN_CASES["ds16_b512"]
    g = golden("assign_ds16_b512")
    x, c = GC.assign_inputs(case)
    table, _, _ = ops.dist_table(_dev(x), _dev(c))
    centred = RepCONC.center_distance_for_constraint(table)
    Q = sinkhorn_algorithm(-centred.double().transpose(1, 2), case["eps"], case["iters"], False)
    Q = Q.cpu().numpy()
    np.testing.assert_allclose(Q.sum(2), g["q_rowsum"], rtol=1e-9)
    np.testing.assert_allclose(Q.transpose(0, 2, 1)[:, :2, :], g["q_head"], rtol=1e-9, atol=1e-300)
    assert abs(Q.sum(1) - 1).max() < 1e-12


def test_sparse_and_dense_passes_agree():
    """The default (sparse) Sinkhorn pass and the dense pass give identical codes and row sums that
    agree to 1e-13 relative (the sparse pass drops only terms below 2^-64 of a column sum)."""
    from repconc_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((2048, 768), generator=gen, device="cuda")
    c = torch.randn((48, 256, 16), generator=gen, device="cuda")
    out = {}
    for dense in (False, True):
        kern = ops.CudaAssignKernels(x, c)
        prev = kern.set_dense(dense)
        try:
            kern.table()
            P = kern.begin(0.003)
            for _ in range(19):
                P = kern.step(0.003, 2048)
            out[dense] = (P.clone(), kern.finish(0.003, True).clone(), kern.read_flags())
        finally:
            kern.set_dense(prev)
    assert out[False][2] == 0 and out[True][2] == 0
    assert torch.equal(out[False][1], out[True][1])
    rel = ((out[False][0] - out[True][0]).abs() / out[True][0]).max().item()
    assert rel < 1e-13, rel


def test_sparse_unsafe_flag_triggers_dense_rerun(oracle):
    """A centroid far away from every point keeps almost no mass after the first column
    normalisation: the sparse pass must flag it and the driver must re-run densely (same codes as
    the oracle either way)."""
    from repconc_b200 import ops
    r = np.random.default_rng(3)
    B, M, K, ds = 512, 2, 256, 8
    x = r.standard_normal((B, M * ds), dtype=np.float32)
    c = r.standard_normal((M, K, ds), dtype=np.float32)
    c[:, 5, :] += 40.0                              # one outlier centroid per sub-vector
    got = ops.constrained_assign(_dev(x), _dev(c), 0.003, 10, distributed=False).cpu().numpy()
    assert np.array_equal(got, oracle.constrained_assign(x, c, 0.003, 10)["codes"])


@pytest.mark.parametrize("shape", [(2048, 48, 16, 20), (300, 96, 8, 7), (1000, 32, 24, 3), (64, 4, 16, 1),
                                   (64, 4, 16, 0), (1, 4, 16, 5)])
@pytest.mark.parametrize("dense", [False, True])
def test_solve_matches_stepwise(shape, dense):
    """rc_sinkhorn_solve (one call, fused reduce+update) is bit-identical to begin/step/finish:
    same codes, same row sums P."""
    from repconc_b200 import ops
    B, M, ds, T = shape
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((B, M * ds), generator=gen, device="cuda")
    c = torch.randn((M, 256, ds), generator=gen, device="cuda")
    ks = ops.CudaAssignKernels(x, c)
    prev = ks.set_dense(dense)
    try:
        ks.table()
        P = ks.begin(0.003)
        for _ in range(max(T - 1, 0)):
            P = ks.step(0.003, B)
        codes_step = ks.finish(0.003, T >= 1).clone()
        kf = ops.CudaAssignKernels(x, c)
        kf.table()
        codes_solve = kf.solve(0.003, T)
        assert torch.equal(codes_step, codes_solve)
        assert ks.read_flags() == kf.read_flags()
        if T >= 2 and B > 1:
            assert torch.equal(ks.P, kf.P)
        u8 = ops.CudaAssignKernels(x, c)
        u8.table()
        assert torch.equal(u8.solve(0.003, T, uint8=True).long(), codes_solve)
    finally:
        ks.set_dense(prev)


def test_list_pass_oversized_rows():
    """Rows with more survivors than a ring slot holds (near-duplicate centroids make whole rows survive)
    take the list pass's global-memory path: codes must still match the dense pass and the step-wise API."""
    from repconc_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(9)
    B, M, ds = 777, 6, 16
    x = torch.randn((B, M * ds), generator=gen, device="cuda")
    c = torch.randn((M, 1, ds), generator=gen, device="cuda") + 1e-3 * torch.randn((M, 256, ds), generator=gen,
                                                                                    device="cuda")
    out = {}
    for dense in (False, True):
        k = ops.CudaAssignKernels(x, c)
        prev = k.set_dense(dense)
        try:
            k.table()
            out[dense] = (k.solve(0.003, 12).clone(), k.P.clone(), k.read_flags())
            if not dense:
                st = torch.zeros(44, dtype=torch.int64, device="cuda")
                from repconc_b200 import _lib
                _lib.check(k.lib.rc_sinkhorn_list_stats(k.state.data_ptr(), B, M, 256, st.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream), "stats")
                assert int(st[1]) > 128, st[:3]          # the case really has oversized rows
        finally:
            k.set_dense(prev)
    assert out[False][2] == 0 and out[True][2] == 0
    assert torch.equal(out[False][0], out[True][0])
    rel = ((out[False][1] - out[True][1]).abs() / out[True][1]).max().item()
    assert rel < 1e-12, rel


def test_sparse_finish_ties_and_odd_rows_match_dense():
    """The fp32-filtered FINISH (K = 256) and the dense FINISH agree on exact ties (duplicate centroids ->
    smallest k), on rows with several near-maximal candidates and on NaN rows (NaN counts as largest)."""
    from repconc_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(21)
    B, M, ds = 1500, 6, 16
    x = torch.randn((B, M * ds), generator=gen, device="cuda")
    c = torch.randn((M, 256, ds), generator=gen, device="cuda")
    c[:, 200:256] = c[:, 100:156]                       # 56 exact duplicates per sub-vector
    c[:, 7] = c[:, 3] + 1e-6                            # and one near-duplicate
    out = {}
    for T in (0, 6):
        for dense in (False, True):
            k = ops.CudaAssignKernels(x, c)
            prev = k.set_dense(dense)
            try:
                k.table()
                out[(T, dense)] = k.solve(0.003, T).clone()
            finally:
                k.set_dense(prev)
        assert torch.equal(out[(T, False)], out[(T, True)]), T
    assert (out[(6, False)] < 200).all()                # ties resolve to the smaller index
    # NaN rows: poison a few table entries after the table is built (begin propagates them)
    res = {}
    for dense in (False, True):
        k = ops.CudaAssignKernels(x, c)
        prev = k.set_dense(dense)
        try:
            k.table()
            k.begin(0.003)
            k.tab[2, 17, 33] = float("nan")
            k.tab[4, 900, 250] = float("nan")
            k.tab[4, 900, 12] = float("nan")
            res[dense] = (k.finish(0.003, False).clone(), k.read_flags())
        finally:
            k.set_dense(prev)
    assert torch.equal(res[False][0], res[True][0])
    assert res[False][0][17, 2] == 33 and res[False][0][900, 4] == 12
    assert res[False][1] == res[True][1] and res[False][1] & 1
