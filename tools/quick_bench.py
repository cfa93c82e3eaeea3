#!/usr/bin/env python
"""Developer timing probe (NOT the official bench): constrained assign and ADC search on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, _lib
from repconc_b200.faiss_compat import GpuIndexPQ

def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

what = sys.argv[1:] or ["assign", "adc"]
gen = torch.Generator(device="cuda").manual_seed(0)
if "assign" in what:
    for (B, M, T) in [(8192, 48, 50), (8192, 96, 50)]:
        ds = 768 // M
        x = torch.randn((B, 768), generator=gen, device="cuda")
        c = torch.randn((M, 256, ds), generator=gen, device="cuda")
        t_nn = timeit(lambda: ops.nn_assign(x, c))
        t_tab = timeit(lambda: ops.dist_table(x, c))
        t = timeit(lambda: ops.constrained_assign(x, c, 0.003, T, distributed=False))
        per_pass = (t[0] - t_tab[0]) / T
        print(f"assign B={B} M={M} T={T}: nn {t_nn[0]:.3f} ms, table {t_tab[0]:.3f} ms, constrained {t[0]:.3f} ms "
              f"({B / t[0] * 1e3:.0f} emb/s), ~{per_pass * 1e3:.1f} us/iter, table pass at "
              f"{M * B * 256 * 4 / per_pass / 1e6:.0f} GB/s", flush=True)
if "adc" in what:
    N = int(os.environ.get("QB_N", 8841823))
    cases = [(48, 1200, 1000), (32, 1200, 1000), (64, 1200, 1000), (96, 1200, 1000), (48, 128, 200)]
    if os.environ.get("QB_M"):
        cases = [(int(m), int(os.environ.get("QB_NQ", 1200)), 1000) for m in os.environ["QB_M"].split(",")]
    _lib.load().rc_adc_enable_timing(1)
    for (M, nq, k) in cases:
        ds = 768 // M
        codes = torch.randint(0, 256, (N, M), generator=gen, device="cuda", dtype=torch.uint8)
        c = torch.randn((M, 256, ds), generator=gen, device="cuda")
        q = torch.randn((nq, 768), generator=gen, device="cuda")
        idx = GpuIndexPQ(codes, c)
        t = timeit(lambda: idx.search_tensor(q, k), n=2)
        print(f"adc N={N} M={M} nq={nq} k={k}: {t[0]:.1f} ms -> {nq / t[0] * 1e3:.0f} QPS, "
              f"{nq * N * M / t[0] / 1e6:.0f} G lookups/s, scan {_lib.load().rc_adc_last_scan_ms():.2f} ms, "
              f"stats {idx.last_stats}", flush=True)
        del codes, idx
print("launches", _lib.launch_count())
