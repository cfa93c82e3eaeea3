#!/usr/bin/env python
"""Run the REFERENCE (imported read-only from /root/reference/src) on the seeded cases of
tests/golden_cases.py and write small fixtures to tests/golden/*.npz.

Only runs in the build container (the GPU box has no /root/reference).  The fixtures pin
the oracle (oracle/) and, through it, the CUDA path.  Regenerate with:
    python tools/gen_golden.py
What is taken from the reference:
  * assign:  RepCONC.quantize (constraint on and off), center_distance_for_constraint,
             sinkhorn_algorithm          (modeling_repconc.py:47-85,137-165)
  * decode:  modeling_repconc.decode     (:168-184)
  * encode:  RepCONC.forward(return_code=True) behind a dummy encoder: rotation, METRIC_CENTROID_COS
             normalisation, NN assign    (:87-110)
  * MSE:     the expression of finetune_repconc.py:367-374 + autograd (the trainer module
             itself does not import under transformers 5.5, so the three lines are quoted)
  * ADC:     Faiss is absent -> anchor = q @ decode(codes).T with the reference's decode
             ("parity unpinned" w.r.t. Faiss)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from tests import golden_cases as GC  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_ref_model(D, M, K, c, use_constraint, eps, iters):
    from transformers import PretrainedConfig
    from repconc.models.repconc.modeling_repconc import RepCONC
    cfg = PretrainedConfig(hidden_size=D)
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = M, K, "METRIC_IP"

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = cfg
    model = RepCONC(cfg, Dummy(), use_constraint, eps, iters)
    with torch.no_grad():
        model.centroids.copy_(torch.from_numpy(c))
    return model


def sha_table(t):
    return GC.digest(np.ascontiguousarray(t, dtype=np.float32))


def run_assign(name, case):
    from repconc.models.repconc.modeling_repconc import RepCONC, sinkhorn_algorithm
    x, c = GC.assign_inputs(case)
    D, M, K, B = case["D"], case["M"], case["K"], case["B"]
    model = make_ref_model(D, M, K, c, True, case["eps"], case["iters"])
    xt = torch.from_numpy(x)
    codes_conc = model.quantize(xt).contiguous().numpy()
    model.use_constraint = False
    codes_nn = model.quantize(xt).contiguous().numpy()
    # the pieces, for finer-grained pins
    with torch.no_grad():
        table = ((xt.reshape(B, M, 1, -1).transpose(0, 1) - model.centroids.unsqueeze(1)) ** 2).sum(-1)
        mx = table.max(-1).values.max(-1).values
        mn = table.min(-1).values.min(-1).values
        centred = RepCONC.center_distance_for_constraint(table)
        Q = sinkhorn_algorithm(-centred.double().transpose(1, 2), case["eps"], case["iters"], False)
        Qt = Q.transpose(1, 2)  # M,B,K
        top2 = torch.topk(Qt, 2, dim=-1).values
        gap = ((top2[..., 0] - top2[..., 1]) / top2[..., 0]).t().contiguous().numpy().astype(np.float32)
        assert np.array_equal(torch.argmax(Qt, -1).t().numpy(), codes_conc)
    dt = np.uint8 if K <= 256 else np.int16
    np.savez_compressed(
        os.path.join(OUT, f"assign_{name}.npz"),
        input_sha=GC.digest(x, c),
        codes_conc=codes_conc.astype(dt), codes_nn=codes_nn.astype(dt),
        table_sha=sha_table(table.numpy()), table_head=table[:, :2, :].numpy(),
        centred_sha=sha_table(centred.numpy()), centred_head=centred[:, :2, :].numpy(),
        max=mx.numpy(), min=mn.numpy(),
        q_rowsum=Q.sum(2).numpy(), q_colsum_err=np.abs(Q.sum(1).numpy() - 1).max(),
        q_head=Qt[:, :2, :].numpy(), top2_gap=gap,
    )
    print(f"assign_{name}: B={B} M={M} K={K} min top2 gap {gap.min():.3e} "
          f"nn!=conc {np.mean(codes_nn != codes_conc):.3f}")


def run_assign_big(name, case):
    """BASELINE-size batch: codes (constraint on / off), extrema, table hashes, top-1/top-2 gap of Q."""
    import gc
    import time
    from repconc.models.repconc.modeling_repconc import RepCONC, sinkhorn_algorithm
    x, c = GC.assign_inputs(case)
    D, M, K, B = case["D"], case["M"], case["K"], case["B"]
    model = make_ref_model(D, M, K, c, True, case["eps"], case["iters"])
    xt = torch.from_numpy(x)
    t0 = time.perf_counter()
    codes_conc = model.quantize(xt).contiguous().numpy()
    t_ref = time.perf_counter() - t0
    model.use_constraint = False
    codes_nn = model.quantize(xt).contiguous().numpy()
    gc.collect()
    with torch.no_grad():
        # the table sub-vector by sub-vector (same broadcast-sub / pow / sum ATen kernels, without the 13 GB temp)
        table = torch.empty((M, B, K))
        xm = xt.reshape(B, M, 1, -1).transpose(0, 1)
        for m in range(M):
            table[m] = ((xm[m:m + 1] - model.centroids[m:m + 1].unsqueeze(1)) ** 2).sum(-1)[0]
        mx = table.max(-1).values.max(-1).values
        mn = table.min(-1).values.min(-1).values
        tsha = sha_table(table.numpy())
        centred = RepCONC.center_distance_for_constraint(table)
        del table
        csha = sha_table(centred.numpy())
        Q = sinkhorn_algorithm(-centred.double().transpose(1, 2), case["eps"], case["iters"], False)
        del centred
        Qt = Q.transpose(1, 2)  # M,B,K
        top2 = torch.topk(Qt, 2, dim=-1).values
        gap = ((top2[..., 0] - top2[..., 1]) / top2[..., 0]).t().contiguous().numpy().astype(np.float16)
        assert np.array_equal(torch.argmax(Qt, -1).t().numpy(), codes_conc)
        q_rowsum = Q.sum(2).numpy()
    np.savez_compressed(
        os.path.join(OUT, f"assign_{name}.npz"),
        input_sha=GC.digest(x, c), codes_conc=codes_conc.astype(np.uint8), codes_nn=codes_nn.astype(np.uint8),
        table_sha=tsha, centred_sha=csha, max=mx.numpy(), min=mn.numpy(), q_rowsum=q_rowsum, top2_gap=gap,
        ref_quantize_seconds=t_ref, ref_threads=torch.get_num_threads(),
    )
    print(f"assign_{name}: B={B} M={M} reference quantize {t_ref:.1f} s on {torch.get_num_threads()} threads, "
          f"min top2 gap {float(gap.astype(np.float32).min()):.3e} nn!=conc {np.mean(codes_nn != codes_conc):.3f}",
          flush=True)


def _dist_worker(rank, world, case, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29611"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    x, c = GC.assign_inputs(case)
    per = case["B"] // world
    model = make_ref_model(case["D"], case["M"], case["K"], c, True, case["eps"], case["iters"])
    codes = model.quantize(torch.from_numpy(x[rank * per:(rank + 1) * per])).contiguous().numpy()
    ret[rank] = codes
    dist.destroy_process_group()


def run_dist(name, case):
    world = case["world"]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_dist_worker, args=(world, case, ret), nprocs=world, join=True)
        codes = np.concatenate([ret[r] for r in range(world)], 0)
    x, c = GC.assign_inputs(case)
    # the same global batch through the single-process reference (documents equivalence)
    model = make_ref_model(case["D"], case["M"], case["K"], c, True, case["eps"], case["iters"])
    single = model.quantize(torch.from_numpy(x)).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, f"assign_{name}.npz"), input_sha=GC.digest(x, c),
                        codes_conc=codes.astype(np.uint8), codes_single=single.astype(np.uint8))
    print(f"assign_{name}: world={world} dist-vs-single mismatches {(codes != single).sum()}")


def run_encode(name, case):
    """RepCONC.forward(return_code=True) of the reference with a dummy encoder that returns the pooled embeddings:
    codes, the rotated (normalised) embeddings' digest / head, and the relative gap between the two smallest
    distances of every (row, sub-vector) (to adjudicate a code that differs by summation order of the rotation)."""
    from transformers import PretrainedConfig
    from repconc.models.repconc.modeling_repconc import RepCONC
    x, rot, c = GC.encode_inputs(case)
    D, M, K, B = case["D"], case["M"], case["K"], case["B"]
    cfg = PretrainedConfig(hidden_size=D)
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = M, K, case["metric"]

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = cfg

        def forward(self, input_ids=None, attention_mask=None):
            return torch.from_numpy(x)
    model = RepCONC(cfg, Dummy(), False, None, None)
    with torch.no_grad():
        model.centroids.copy_(torch.from_numpy(c))
        model.rotation.copy_(torch.from_numpy(rot))
        out = model(None, None, return_code=True)
        y = out.continuous_embeds
        codes = out.discrete_codes.contiguous().numpy()
        table = ((y.reshape(B, M, 1, -1).transpose(0, 1) - model.centroids.unsqueeze(1)) ** 2).sum(-1)   # M,B,K
        low2 = torch.topk(table, 2, dim=-1, largest=False).values
        gap = ((low2[..., 1] - low2[..., 0]) / low2[..., 1].clamp_min(1e-30)).t().contiguous().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"encode_{name}.npz"), input_sha=GC.digest(x, rot, c),
                        codes=codes.astype(np.uint8), gap=gap, rotated_sha=sha_table(y.numpy()),
                        rotated_head=y[:4].numpy())
    print(f"{name}: codes {codes.shape}, min gap {gap.min():.3e}")


def run_mse(name, case):
    from repconc.models.repconc.modeling_repconc import decode
    x, c, g, codes = GC.mse_inputs(case)
    xt = torch.from_numpy(x).requires_grad_(True)
    ct = torch.from_numpy(c).requires_grad_(True)
    gt = torch.from_numpy(g)
    q = decode(torch.from_numpy(codes), ct)
    # finetune_repconc.py:367-374 (is-doc branch) and :396 / :390 (scaler.scale(mse) + surrogate)
    surrogate = torch.dot(gt.flatten(), xt.flatten())
    surrogate = surrogate + torch.dot(gt.flatten(), q.flatten())
    mse_loss = ((q - xt) ** 2).sum(-1).mean() * case["w"]
    (case["scale"] * mse_loss + surrogate).backward()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), input_sha=GC.digest(x, c, g, codes),
                        quantized=q.detach().numpy(), mse=mse_loss.item(), surrogate=surrogate.item(),
                        grad_x=xt.grad.numpy(), grad_c=ct.grad.numpy())
    print(f"{name}: mse={mse_loss.item():.6e} surrogate={surrogate.item():.6e}")


def run_adc(name, case):
    from repconc.models.repconc.modeling_repconc import decode
    q, c, codes = GC.adc_inputs(case)
    docs = decode(codes.astype(np.int64), c)           # numpy branch of the reference decode
    scores = (torch.from_numpy(q).double() @ torch.from_numpy(docs).double().T)  # exact-ish anchor
    out = dict(input_sha=GC.digest(q, c, codes))
    for k in case["ks"]:
        top = torch.topk(scores, k, dim=1)
        out[f"scores_k{k}"] = top.values.numpy().astype(np.float32)
        out[f"ids_k{k}"] = top.indices.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(f"{name}: N={case['N']} nq={case['nq']} ks={case['ks']}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if "--big" in sys.argv:          # the two BASELINE-size cases only (minutes, tens of GB)
        for name, case in GC.ASSIGN_BIG_CASES.items():
            run_assign_big(name, case)
        return
    if "--encode" in sys.argv:       # only the forward() epilogue cases
        for name, case in GC.ENCODE_CASES.items():
            run_encode(name, case)
        return
    for name, case in GC.ASSIGN_CASES.items():
        run_assign(name, case)
    for name, case in GC.DIST_CASES.items():
        run_dist(name, case)
    for name, case in GC.ENCODE_CASES.items():
        run_encode(name, case)
    for name, case in GC.MSE_CASES.items():
        run_mse(name, case)
    for name, case in GC.ADC_CASES.items():
        run_adc(name, case)


if __name__ == "__main__":
    main()
