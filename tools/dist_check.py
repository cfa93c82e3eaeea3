#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs of one box):
  * constrained assign, batch sharded over ranks, NCCL all-reduces -> codes identical to the codes the
    REFERENCE produced on 2 gloo ranks (golden assign_dist2_ds16) and to the oracle on the global batch;
  * corpus-sharded ADC search + all_gather + merge -> identical to the oracle's unsharded search.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_cases as GC  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import oracle as O
    from repconc_b200 import ops
    from repconc_b200 import evaluate_repconc as E
    from repconc_b200.faiss_compat import GpuIndexPQ
    ok = True
    # --- assign
    case = GC.DIST_CASES["dist2_ds16"]
    x, c = GC.assign_inputs(case)
    per = case["B"] // world
    xs = torch.from_numpy(x[rank * per:(rank + 1) * per]).to(dev)
    codes = ops.constrained_assign(xs, torch.from_numpy(c).to(dev), case["eps"], case["iters"])   # dist on
    allc = [torch.empty_like(codes.contiguous()) for _ in range(world)]
    dist.all_gather(allc, codes.contiguous())
    got = torch.cat(allc, 0).cpu().numpy()
    want = O.constrained_assign(x[: per * world], c, case["eps"], case["iters"])["codes"]
    a_ok = bool(np.array_equal(got, want))
    if world == 2:
        g = np.load(os.path.join(ROOT, "tests", "golden", "assign_dist2_ds16.npz"))
        a_ok = a_ok and bool(np.array_equal(got, g["codes_conc"].astype(np.int64)))
    ok &= a_ok
    # --- BASELINE-size assignments (global batch 8192, M = 48 and M = 96, T = 50) sharded over the ranks, against
    #     the codes the REFERENCE produced for the global batch (goldens assign_m48_b8192 / assign_m96_b8192;
    #     the reference's distributed run is the same computation: sums over the global batch, B *= world)
    b_ok = True
    for name in ("m48_b8192", "m96_b8192"):
        bc = GC.ASSIGN_BIG_CASES[name]
        gb = np.load(os.path.join(ROOT, "tests", "golden", f"assign_{name}.npz"))
        xb, cb = GC.assign_inputs(bc)
        perb = bc["B"] // world
        cd = ops.constrained_assign(torch.from_numpy(xb[rank * perb:(rank + 1) * perb]).to(dev),
                                    torch.from_numpy(cb).to(dev), bc["eps"], bc["iters"])
        wantb = gb["codes_conc"][rank * perb:(rank + 1) * perb].astype(np.int64)
        nbad = int((cd.cpu().numpy() != wantb).sum())
        if nbad:
            print(f"rank {rank}: {name}: {nbad} code mismatches vs the reference", flush=True)
        b_ok = b_ok and nbad == 0
    ok &= b_ok
    # --- RC_FLAG_SPARSE_UNSAFE raised on ONE rank only (rank 1 pretends its survivor records hold 8 entries): the
    #     flag word is OR-ed over the ranks, every rank re-runs densely in lock step, codes stay the reference's
    from repconc_b200 import _lib
    lib = _lib.load()
    prev = lib.rc_sinkhorn_debug_pool_entries(8 if rank == 1 else 0)
    try:
        codes2 = ops.constrained_assign(xs, torch.from_numpy(c).to(dev), case["eps"], case["iters"])
    finally:
        lib.rc_sinkhorn_debug_pool_entries(prev)
    allc2 = [torch.empty_like(codes2.contiguous()) for _ in range(world)]
    dist.all_gather(allc2, codes2.contiguous())
    u_ok = bool(np.array_equal(torch.cat(allc2, 0).cpu().numpy(), want))
    ok &= u_ok
    # --- sharded ADC
    q, cc, codes_h = GC.adc_inputs(GC.ADC_CASES["adc_m8"])
    lo, hi = E.shard_bounds(len(codes_h), rank, world)
    shard = GpuIndexPQ(torch.from_numpy(codes_h[lo:hi]).to(dev), torch.from_numpy(cc).to(dev), id_offset=lo)
    s, i = E.sharded_search(shard, torch.from_numpy(q).to(dev), 100)      # CUDA tensors in -> result on every rank
    # host arrays in (the evaluator's batch_search): rank 0 receives the merged result, the others empty arrays
    hs, hi_ = E.batch_search(np.arange(len(q)), q, np.arange(len(codes_h), dtype=np.int64), E.ShardedSearcher(shard), 100, 24)
    so, io = O.adc_search(q, cc, codes_h, 100)
    s_ok = bool(np.array_equal(s.cpu().numpy(), so) and np.array_equal(i.cpu().numpy(), io))
    s_ok = s_ok and (bool(np.array_equal(hs, so) and np.array_equal(hi_, io)) if rank == 0 else len(hs) == 0)
    ok &= s_ok
    # --- index replicas + query split (the reference's multi-GPU mode)
    full = GpuIndexPQ(torch.from_numpy(codes_h).to(dev), torch.from_numpy(cc).to(dev))
    rep = E.ReplicatedSearcher(full)
    rs, ri = rep.search(torch.from_numpy(q).to(dev), 100)
    rhs, rhi = E.batch_search(np.arange(len(q)), q, np.arange(len(codes_h), dtype=np.int64), rep, 100, 24)
    r_ok = bool(np.array_equal(rs.cpu().numpy(), so) and np.array_equal(ri.cpu().numpy(), io))
    r_ok = r_ok and (bool(np.array_equal(rhs, so) and np.array_equal(rhi, io)) if rank == 0 else len(rhs) == 0)
    ok &= r_ok
    # --- peer-memory all-reduce (rc_peer_allreduce_f64) vs a rank-ordered sum of all-gathered vectors
    from repconc_b200.peer import PeerAllReduce
    n = 48 * 256
    red = PeerAllReduce.get(n, dev)
    p_ok, p_used = True, red is not None
    if red is not None:
        fl = torch.zeros(1, dtype=torch.int32, device=dev)
        for it in range(5):
            g = torch.Generator(device=dev).manual_seed(100 * it + rank)
            v = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (10.0 ** (rank - it))
            allv = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(allv, v)
            want = allv[0].clone()
            for r in range(1, world):
                want += allv[r]
            red.all_reduce(v, fl)
            p_ok = p_ok and bool(torch.equal(v, want))
        p_ok = p_ok and int(fl.item()) == 0
    ok &= p_ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_check world={world}: assign_golden={a_ok} assign_baseline_size_vs_reference={b_ok} "
              f"one_rank_unsafe_flag_dense_rerun={u_ok} sharded_adc={s_ok} replicated_adc={r_ok} "
              f"peer_allreduce={'exact' if (p_used and p_ok) else ('MISMATCH' if p_used else 'unavailable (NCCL used)')} "
              f"-> {'PASS' if flag.item() == 1 else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
