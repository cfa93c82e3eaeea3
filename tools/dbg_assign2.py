#!/usr/bin/env python
"""Developer probe: first iteration count at which the persistent kernel's row sums differ from the step-wise path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, _lib
from tests import golden_cases as GC
name = sys.argv[1] if len(sys.argv) > 1 else "m48_b1024"
case = (GC.ASSIGN_BIG_CASES | GC.ASSIGN_CASES)[name]
x, c = GC.assign_inputs(case)
xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
for T in [1, 2, 3, 4, 5, 6, 8, 12, 20, 50]:
    k = ops.CudaAssignKernels(xd, cd); k.table()
    c1 = k.solve(case["eps"], T).clone(); P1 = k.P.clone(); lu1 = None
    k2 = ops.CudaAssignKernels(xd, cd); k2.table(); k2.begin(case["eps"])
    for _ in range(T - 1):
        k2.step(case["eps"], case["B"])
    P2 = k2.P.clone()
    c2 = k2.finish(case["eps"], True)
    dP = (P1 != P2)
    bad_m = dP.any(1).nonzero().flatten().tolist()
    rel = ((P1 - P2).abs() / P2.abs()).max().item()
    print(f"T={T}: P differs in {int(dP.sum())} entries, sub-vectors {bad_m[:12]}{'...' if len(bad_m) > 12 else ''}, max rel {rel:.3e}; "
          f"codes differ {int((c1 != c2).sum())}; flags {k.read_flags()} {k2.read_flags()}")
