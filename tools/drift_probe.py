#!/usr/bin/env python
"""Developer probe: spread of lu - lu_build (max over sub-vectors) after every Sinkhorn update, step-wise entry points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, _lib
B = 8192; M = int(sys.argv[1]) if len(sys.argv) > 1 else 48; T = 50
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 768), generator=gen, device="cuda")
c = torch.randn((M, 256, 768 // M), generator=gen, device="cuda")
lib = _lib.load()
k = ops.CudaAssignKernels(x, c)
k.table(); k.begin(0.003)
buf = np.zeros((M, 2), dtype=np.float64)
out = []
for it in range(T - 1):
    k.step(0.003, B)
    lib.rc_sinkhorn_debug_drift(k.state.data_ptr(), B, M, 256, buf.ctypes.data)
    out.append((it, float(buf[:, 1].max()), float(np.median(buf[:, 1]))))
print("iteration: max spread / median spread over sub-vectors (after the update that follows pass `it`)")
print(" ".join(f"{it}:{a:.1f}/{b:.1f}" for it, a, b in out))
