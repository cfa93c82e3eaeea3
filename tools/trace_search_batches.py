#!/usr/bin/env python
"""Developer probe: where the host time of GpuIndexPQ.search_batches goes (per batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200.faiss_compat import GpuIndexPQ
from repconc_b200 import evaluate_repconc as E
N, M, D, K = 8841823, 48, 768, 1000
g = torch.Generator(device="cuda").manual_seed(0)
codes = torch.randint(0, 256, (N, M), generator=g, device="cuda", dtype=torch.uint8)
c = torch.randn((M, 256, D // M), generator=g, device="cuda")
idx = GpuIndexPQ(codes, c)
q = torch.randn((7200, D), generator=g, device="cuda").cpu().numpy()
ids = np.arange(N, dtype=np.int64)
E.batch_search(np.arange(2400), q[:2400], ids, idx, K, 1200)
for name, fn in [("pipelined", lambda: E.batch_search(np.arange(7200), q, ids, idx, K, 1200)),
                 ("per-batch", lambda: [E.batch_search(np.arange(1200), q[i * 1200:(i + 1) * 1200], ids, idx, K, 1200) for i in range(6)]),
                 ("device-only", lambda: [idx.search_tensor(torch.from_numpy(q[i * 1200:(i + 1) * 1200]).cuda(), K) for i in range(6)])]:
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{name}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
idx._trace = []
t0 = time.perf_counter(); E.batch_search(np.arange(7200), q, ids, idx, K, 1200); tt = 1e3 * (time.perf_counter() - t0)
tr = idx._trace
print("pipelined total", round(tt, 1))
prev = tr[0][1]
agg = {}
for lab, t in tr[1:]:
    agg.setdefault(lab, []).append(1e3 * (t - prev)); prev = t
for k, v in agg.items():
    print(f"  -> {k:10s} " + " ".join(f"{x:6.2f}" for x in v))
