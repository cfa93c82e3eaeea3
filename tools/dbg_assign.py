#!/usr/bin/env python
"""Developer probe: a golden assignment case through the persistent and the step-wise path, with the pass counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, _lib
from tests import golden_cases as GC
name = sys.argv[1] if len(sys.argv) > 1 else "m96_b8192"
case = (GC.ASSIGN_BIG_CASES | GC.ASSIGN_CASES)[name]
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"assign_{name}.npz"))
x, c = GC.assign_inputs(case)
xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
lib = _lib.load()
def stats(k):
    out = torch.zeros(44, dtype=torch.int64, device="cuda")
    _lib.check(lib.rc_sinkhorn_list_stats(k.state.data_ptr(), k.B, k.M, 256, out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "stats")
    o = out.cpu().tolist()
    return f"mean {o[0] / max(o[2], 1):.1f} max {o[1]} sel {o[36]} list {o[37]} massfail {o[38]} poolfull {o[39]}"
want = g["codes_conc"].astype(np.int64)
k = ops.CudaAssignKernels(xd, cd); k.table()
got = k.solve(case["eps"], case["iters"]).cpu().numpy()
print("persistent: flags", k.read_flags(), "mismatch", int((got != want).sum()), stats(k))
k = ops.CudaAssignKernels(xd, cd); k.table(); k.begin(case["eps"])
for _ in range(case["iters"] - 1):
    k.step(case["eps"], case["B"])
got = k.finish(case["eps"], True).cpu().numpy()
print("stepwise  : flags", k.read_flags(), "mismatch", int((got != want).sum()), stats(k))
