import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from repconc_b200 import ops
from tests import golden_cases as GC
case = GC.ASSIGN_CASES["m48_b1024"]
x, c = GC.assign_inputs(case)
xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
for T in [2, 3]:
    k = ops.CudaAssignKernels(xd, cd); k.table()
    c1 = k.solve(case["eps"], T).clone()
torch.cuda.synchronize()
print("done")
