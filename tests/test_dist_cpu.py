"""world_size-2 `gloo` tests of the host-side multi-rank logic (no GPU):
  * `constrained_assign_driver` -- where the three all-reduces of the reference go -- driven with an
    emulated kernel set (numpy restatement of the kernels' contract, test-only) must reproduce the
    codes the REFERENCE produced on 2 gloo ranks (golden `assign_dist2_ds16`);
  * corpus shard bounds + top-k merge semantics (oracle merge) reproduce the unsharded search.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import golden_cases as GC  # noqa: E402


class EmulatedAssignKernels:
    """numpy stand-in for ops.CudaAssignKernels: same methods, same state semantics
    (table/minmax, lu, lv, P) as csrc/assign.cu, built on the oracle's bit-exact fp32 table."""

    def __init__(self, x, c):
        from oracle import oracle as O
        self.O = O
        self.x, self.c = x, c
        self.M, self.K, self.ds = c.shape
        self.B = x.shape[0]
        self.calls = []
        self._flags = 0
        self.unsafe_on_sparse = False             # test hook: this rank's sparse pass reports pool exhaustion

    def table(self):
        self.tab = self.O.dist_table(self.x, self.c)                     # (M,B,K) fp32
        mx, mn = self.O.table_minmax(self.tab)
        self.minmax = torch.from_numpy(np.stack([mx, mn]))
        return self.minmax

    def begin(self, eps):
        mm = self.minmax.numpy()
        self.tab = self.O.center_table(self.tab, mm[0], mm[1])           # centred in place
        self.a = -self.tab.astype(np.float64) / eps                      # (M,B,K)
        self.lu = np.zeros((self.M, self.K))
        self.lv = np.zeros((self.M, self.B))
        self.P = torch.from_numpy(np.exp(self.a).sum(1))                 # (M,K) row sums
        return self.P

    def _update(self):
        self.lu = self.lu - np.log(self.K * self.P.numpy())

    def step(self, eps, B_global, dense=False):
        self.calls.append(("step", bool(dense)))
        self._update()
        w = self.a + self.lu[:, None, :] + self.lv[:, :, None]
        q = np.exp(w)
        z = B_global * q.sum(2)                                          # (M,B)
        self.lv = self.lv - np.log(z)
        self.P = torch.from_numpy((q / z[:, :, None]).sum(1))
        return self.P

    def finish(self, eps, apply_rowsum, uint8=False, B_global=None, dense=False):
        self.calls.append(("finish", bool(dense)))
        if not dense and self.unsafe_on_sparse:
            self._flags |= 8                      # RC_FLAG_SPARSE_UNSAFE raised from this rank's own data
        if apply_rowsum and B_global == 1:          # single column: exact tie in the reference -> code 0
            return torch.zeros((self.B, self.M), dtype=torch.int64)
        if apply_rowsum:
            self._update()
        w = self.a + self.lu[:, None, :]
        return torch.from_numpy(np.argmax(w, axis=2).T.copy())

    def read_flags(self):
        return self._flags

    def clear_flags(self):
        self._flags = 0


def _worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from repconc_b200.ops import constrained_assign_driver
    x, c = GC.assign_inputs(case)
    per = case["B"] // world
    kern = EmulatedAssignKernels(x[rank * per:(rank + 1) * per], c)
    codes = constrained_assign_driver(kern, case["eps"], case["iters"], distributed=True)
    ret[rank] = codes.numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_driver_two_ranks_gloo_matches_reference_two_ranks():
    case = GC.DIST_CASES["dist2_ds16"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "assign_dist2_ds16.npz"))
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, 29733, case, ret), nprocs=2, join=True)
        codes = np.concatenate([ret[0], ret[1]], 0)
    assert np.array_equal(codes, g["codes_conc"].astype(np.int64))


def _asym_worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from repconc_b200.ops import constrained_assign_driver
    x, c = GC.assign_inputs(case)
    per = case["B"] // world
    kern = EmulatedAssignKernels(x[rank * per:(rank + 1) * per], c)
    kern.unsafe_on_sparse = rank == 1            # ONLY rank 1 sees RC_FLAG_SPARSE_UNSAFE
    codes = constrained_assign_driver(kern, case["eps"], case["iters"], distributed=True)
    ret[rank] = (codes.numpy(), kern.calls)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_local_sparse_unsafe_flag_reruns_densely_on_every_rank():
    """RC_FLAG_SPARSE_UNSAFE raised on ONE rank only (survivor-pool exhaustion depends on the rank's own rows):
    the flag word is OR-ed over the ranks, so both ranks redo the assignment with the dense pass in lock step
    (a rank re-entering the collective sequence alone would hang) and the codes are the reference's."""
    case = GC.DIST_CASES["dist2_ds16"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "assign_dist2_ds16.npz"))
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_asym_worker, args=(2, 29735, case, ret), nprocs=2, join=True)
        (c0, calls0), (c1, calls1) = ret[0], ret[1]
    assert np.array_equal(np.concatenate([c0, c1], 0), g["codes_conc"].astype(np.int64))
    assert calls0 == calls1                                     # identical kernel / collective sequence
    n = case["iters"]
    assert calls0 == [("step", False)] * (n - 1) + [("finish", False)] + [("step", True)] * (n - 1) + [("finish", True)]


def test_driver_single_rank_matches_golden():
    from repconc_b200.ops import constrained_assign_driver
    for name in ("ds16_b512", "ds12_b300", "ds5_k64", "b1"):
        case = GC.ASSIGN_CASES[name]
        g = np.load(os.path.join(ROOT, "tests", "golden", f"assign_{name}.npz"))
        x, c = GC.assign_inputs(case)
        codes = constrained_assign_driver(EmulatedAssignKernels(x, c), case["eps"], case["iters"], False)
        assert np.array_equal(codes.numpy(), g["codes_conc"].astype(np.int64)), name


def _shard_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from repconc_b200.evaluate_repconc import shard_bounds
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    lo, hi = shard_bounds(len(codes), rank, world)
    s, i = O.adc_search(q, c, codes[lo:hi], 20, id_offset=lo)          # stands in for the shard's GPU scan
    gs = [torch.empty(s.shape, dtype=torch.float32) for _ in range(world)]
    gi = [torch.empty(i.shape, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gs, torch.from_numpy(s))
    dist.all_gather(gi, torch.from_numpy(i))
    ms, mi = O.topk_merge(torch.stack(gs).numpy(), torch.stack(gi).numpy())
    ret[rank] = (ms, mi)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_protocol_two_ranks_gloo():
    from oracle import oracle as O
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_shard_worker, args=(2, 29734, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    s, i = O.adc_search(q, c, codes, 20)
    for ms, mi in (r0, r1):
        assert np.array_equal(ms, s) and np.array_equal(mi, i)


class _OracleIndex:
    """stands in for a GpuIndexPQ replica: the oracle's scan on CPU tensors"""

    def __init__(self, c, codes):
        self.c, self.codes = c, codes
        self.device = torch.device("cpu")
        self.ntotal = len(codes)
        self.last_stats = None

    def search_tensor(self, x, k):
        from oracle import oracle as O
        if x.shape[0] == 0:
            return torch.empty((0, k), dtype=torch.float32), torch.empty((0, k), dtype=torch.int64)
        s, i = O.adc_search(x.numpy(), self.c, self.codes, k)
        return torch.from_numpy(s), torch.from_numpy(i)


def _oracle_search_batches(self, batches, k, corpus_ids=None, search_fn=None, copy_back=True, result_sizes=None):
    """CPU stand-in for GpuIndexPQ.search_batches (no pinned memory, no streams): same contract -- `batches` are the
    rows this rank uploads, `result_sizes` the rows every batch returns, empty batches are skipped, only
    `copy_back` ranks receive arrays"""
    sizes = [int(v) for v in result_sizes] if result_sizes is not None else [len(b) for b in batches]
    outs, outi = [], []
    for b, n in zip(batches, sizes):
        if n == 0:
            continue
        s, i = search_fn(torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)), k)
        assert tuple(s.shape) == (n, k)
        if copy_back:
            ids = i.numpy()
            outs.append(s.numpy())
            outi.append(corpus_ids[ids] if corpus_ids is not None else ids)
    if not copy_back:
        return np.empty((0, k), np.float32), np.empty((0, k), np.int64)
    return np.concatenate(outs), np.concatenate(outi)


_OracleIndex.search_batches = _oracle_search_batches


def _replica_batches_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from repconc_b200.evaluate_repconc import ReplicatedSearcher
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    rep = ReplicatedSearcher(_OracleIndex(c, codes))
    corpus_ids = np.arange(len(codes), dtype=np.int64) * 3 + 11
    # ragged batches, an empty one, and a batch with fewer queries than ranks
    batches = [q[:25], q[25:25], q[25:26], q[26:64]]
    ret[rank] = rep.search_batches(batches, 20, corpus_ids=corpus_ids)
    dist.barrier()
    dist.destroy_process_group()


def test_replicated_batch_search_uploads_blocks_and_returns_whole_batches_gloo():
    """ReplicatedSearcher.search_batches: every rank passes only ITS block of each batch down, receives the gathered
    results of the whole batch, and only rank 0 gets arrays (ids mapped through corpus_ids)"""
    from oracle import oracle as O
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_replica_batches_worker, args=(2, 29738, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    s, i = O.adc_search(q[:64], c, codes, 20)
    assert np.array_equal(r0[0], s) and np.array_equal(r0[1], i * 3 + 11)
    assert len(r1[0]) == 0 and len(r1[1]) == 0


def _replica_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from repconc_b200.evaluate_repconc import ReplicatedSearcher
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    rep = ReplicatedSearcher(_OracleIndex(c, codes))
    out = []
    for nq in (len(q), 3, 1):                       # even split, ragged split, fewer queries than ranks
        s, i = rep.search(torch.from_numpy(q[:nq]), 20)
        out.append((s.numpy(), i.numpy()))
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_replicated_search_protocol_two_ranks_gloo():
    """query-split over index replicas (reference evaluate_repconc.py:130-134): every rank ends with the full result"""
    from oracle import oracle as O
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_replica_worker, args=(2, 29736, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    q, c, codes = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    for j, nq in enumerate((len(q), 3, 1)):
        s, i = O.adc_search(q[:nq], c, codes, 20)
        for r in (r0, r1):
            assert np.array_equal(r[j][0], s) and np.array_equal(r[j][1], i)
