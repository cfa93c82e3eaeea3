#!/usr/bin/env python
"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, sinkhorn_algorithm
from repconc_b200.faiss_compat import GpuIndexPQ
gen = torch.Generator(device="cuda").manual_seed(0)
for (B, M, K, ds, T) in [(300, 4, 256, 16, 12), (70, 3, 64, 5, 4), (513, 2, 256, 8, 12)]:
    x = torch.randn((B, M * ds), generator=gen, device="cuda")
    c = torch.randn((M, K, ds), generator=gen, device="cuda", requires_grad=True)
    for dense in (False, True):
        from repconc_b200 import _lib
        prev = _lib.load().rc_sinkhorn_set_dense(1 if dense else 0)
        codes = ops.constrained_assign(x, c, 0.003, T, distributed=False)
        _lib.load().rc_sinkhorn_set_dense(prev)
    nn = ops.nn_assign(x, c, uint8=(K <= 256))
    xg = x.clone().requires_grad_(True)
    g = torch.randn_like(x) / B
    mse, sur = ops.quantization_loss(xg, c, codes, g, 1e-4)
    (mse + sur).backward()
    q = ops.decode(codes, c)
    q.sum().backward()
# step-wise entry points, oversized survivor rows (list pass reading straight from the pool), warm-up kernels
x = torch.randn((333, 4 * 16), generator=gen, device="cuda")
c = torch.randn((4, 1, 16), generator=gen, device="cuda") + 1e-3 * torch.randn((4, 256, 16), generator=gen, device="cuda")
k = ops.CudaAssignKernels(x, c)
k.table(); k.begin(0.003)
for _ in range(5):
    k.step(0.003, 333)
k.finish(0.003, True)
k.table(); k.solve(0.003, 8, uint8=True)
from repconc_b200 import warmup
xw = torch.randn((2048, 32), generator=gen, device="cuda")
cw, objs = warmup.train_pq(xw, 4, 256, niter=2, seed=1)
warmup.code_histogram(ops.nn_assign(xw, cw), 256)
M, ds = 16, 4
c = torch.randn((M, 256, ds), generator=gen, device="cuda")
for N, nq, k in [(3000, 9, 17), (300_000, 20, 100)]:
    codes = torch.randint(0, 256, (N, M), generator=gen, device="cuda", dtype=torch.uint8)
    qv = torch.randn((nq, M * ds), generator=gen, device="cuda")
    s, i = GpuIndexPQ(codes, c).search_tensor(qv, k)
# the 8-bit-field scans: 16 queries per entry with 4 and 2 lanes per document (M = 48, 32), 8 queries per entry
# (M = 64), a ragged query tile and a ragged last split; the re-score's radix select
for M, nq in [(48, 20), (32, 17), (64, 9)]:
    ds = 4
    c = torch.randn((M, 256, ds), generator=gen, device="cuda")
    codes = torch.randint(0, 256, (270_001, M), generator=gen, device="cuda", dtype=torch.uint8)
    qv = torch.randn((nq, M * ds), generator=gen, device="cuda")
    s, i = GpuIndexPQ(codes, c).search_tensor(qv, 50)
# fused encode epilogue (cp.async pipeline, packed FMA, 4 / 2 / 1 sub-vectors per CTA), appended in place
for (B, M, ds) in [(300, 8, 16), (70, 6, 24), (129, 3, 12)]:
    D = M * ds
    xe = torch.randn((B, D), generator=gen, device="cuda")
    ce = torch.randn((M, 256, ds), generator=gen, device="cuda")
    rot = torch.linalg.qr(torch.randn((D, D), generator=gen, device="cuda"))[0].contiguous()
    ops.encode_assign(xe, rot, ce, normalize=True)
    idx = GpuIndexPQ(torch.empty((0, M), dtype=torch.uint8, device="cuda"), ce)
    idx.add_encoded(xe, rot)
torch.cuda.synchronize()
print("sanitize_smoke done", float(s.sum()))
