#!/usr/bin/env python
"""Developer probe: one constrained assignment (B=8192, M=48 by default) for ncu, plus the survivor-list
statistics of the sparse Sinkhorn pass.   python tools/prof_assign.py [T] [B] [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repconc_b200 import ops, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
M = int(sys.argv[3]) if len(sys.argv) > 3 else 48
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 768), generator=gen, device="cuda")
c = torch.randn((M, 256, 768 // M), generator=gen, device="cuda")
kern = ops.CudaAssignKernels(x, c)
codes = ops.constrained_assign_driver(kern, 0.003, T, False)
out = torch.zeros(44, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.rc_sinkhorn_list_stats(kern.state.data_ptr(), B, M, 256, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "stats")
o = out.cpu().tolist()
print(f"lists: entries {o[0]}, rows {o[2]}, mean {o[0] / max(o[2], 1):.1f}, max {o[1]}")
print("hist (bucket of 8):", o[3:36])
print(f"segments: selection {o[36]}, list {o[37]}; mass-check failures {o[38]}, pool-full rows {o[39]}; flags {kern.read_flags()}")
import numpy as np
buf = np.zeros((1024, 4), dtype=np.int64)
n = lib.rc_sinkhorn_debug_cta_times(kern.state.data_ptr(), B, M, 256, buf.ctypes.data, 1024)
t = buf[:n] / 1e3
for i, nm in enumerate(["wait", "selection", "list", "arrive+update"]):
    print(f"  per-CTA {nm:14s} us: min {t[:, i].min():8.1f} mean {t[:, i].mean():8.1f} max {t[:, i].max():8.1f}")
busy = t[:, 1] + t[:, 2]
print(f"  busy (sel+list): min {busy.min():.1f} max {busy.max():.1f}; even/odd CTA mean {busy[0::2].mean():.1f} / {busy[1::2].mean():.1f}")
print("  busy of CTAs 0..15:", np.round(busy[:16]).tolist())
if os.environ.get("PROF_ALL"):
    np.set_printoptions(linewidth=200)
    print("  list us, all CTAs (rows of 37):")
    print(np.round(t[:, 2]).astype(int).reshape(-1, 37))
    print("  selection us:")
    print(np.round(t[:, 1]).astype(int).reshape(-1, 37))
print("codes checksum", int(codes.sum().item()))
