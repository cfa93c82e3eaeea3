#!/usr/bin/env python
"""Developer probe: after ONE pass (T=2), compare the per-CTA partial rows / lu / P of the persistent kernel with the
step-wise path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops
from tests import golden_cases as GC
case = GC.ASSIGN_CASES["m48_b1024"]
x, c = GC.assign_inputs(case)
xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
M, K, B = 48, 256, 1024
al = lambda n: (n + 255) // 256 * 256
o_lu = 0; o_P = al(M * K * 8); o_lv = o_P + al(M * K * 8); o_pa = o_lv + al(M * B * 8)
rows = 296 * 4
def grab(k):
    st = k.state
    lu = st[o_lu:o_lu + M * K * 8].view(torch.float64).view(M, K).clone()
    pa = st[o_pa:o_pa + rows * K * 8].view(torch.float64).view(296, 4, K).clone()
    return lu, pa, k.P.clone()
for trial in range(6):
    k = ops.CudaAssignKernels(xd, cd); k.table(); k.solve(case["eps"], 2); a = grab(k)
    k2 = ops.CudaAssignKernels(xd, cd); k2.table(); k2.begin(case["eps"]); k2.step(case["eps"], B); b = grab(k2)
    # lu here is AFTER update(1) in the persistent run and after update(0) in the step-wise run: apply the finish update
    dpa = (a[1] != b[1])
    idx = dpa.nonzero()
    print(f"trial {trial}: P differs {int((a[2] != b[2]).sum())}; partial rows differing: "
          f"{sorted(set((int(g), int(s)) for g, s, _ in idx.tolist()))[:10]} columns: {sorted(set(int(kk) for _, _, kk in idx.tolist()))[:40]}")
    if len(idx):
        g, s, kk = idx[0].tolist()
        print("   first diff: cta", g, "slot", s, "col", kk, "persistent", a[1][g, s, kk].item(), "stepwise", b[1][g, s, kk].item())
