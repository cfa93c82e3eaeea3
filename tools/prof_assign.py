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
out = torch.zeros(36, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.rc_sinkhorn_list_stats(kern.state.data_ptr(), B, M, 256, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "stats")
o = out.cpu().tolist()
print(f"lists: entries {o[0]}, rows {o[2]}, mean {o[0] / max(o[2], 1):.1f}, max {o[1]}")
print("hist (bucket of 8):", o[3:])
print("codes checksum", int(codes.sum().item()))
