#!/usr/bin/env python
"""Developer probe: after ONE selection pass (T=2), compare the survivor records of the persistent kernel with the
step-wise path and decode the first differing one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repconc_b200 import ops, _lib
from tests import golden_cases as GC
case = GC.ASSIGN_CASES["m48_b1024"]
x, c = GC.assign_inputs(case)
xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
M, K, B = 48, 256, 1024
lib = _lib.load()
nbytes = lib.rc_sinkhorn_state_bytes(B, M, K)
pairs = M * (B // 16) * 8
pool_off = None
def grab(k):
    st = k.state
    # the pool is the last region of the state: pairs * 4224 bytes, 256-aligned end
    end = (st.numel() // 256) * 256
    size = (pairs * 4224 + 255) // 256 * 256
    pool = st[end - size:end - size + pairs * 4224].view(pairs, 4224).clone()
    return pool
for trial in range(8):
    k = ops.CudaAssignKernels(xd, cd); k.table(); k.solve(case["eps"], 2); a = grab(k); ta = k.tab.clone()
    k2 = ops.CudaAssignKernels(xd, cd); k2.table(); k2.begin(case["eps"]); k2.step(case["eps"], B); b = grab(k2)
    tb = k2.tab.clone()
    # compare only the bytes the directory says are live: simpler -- compare headers (first 64 bytes of each record 0)
    d = (a != b).any(1).nonzero().flatten().tolist()
    print(f"trial {trial}: tables equal {bool(torch.equal(ta, tb))}; pairs with differing bytes: {len(d)} {d[:6]}")
    for p in d[:2]:
        tile_g, warp = divmod(p, 8)
        m, tile = divmod(tile_g, B // 16)
        ra, rb = a[p].cpu().numpy(), b[p].cpu().numpy()
        ha, hb = ra[:64].view(np.uint16), rb[:64].view(np.uint16)
        ca, cb = int((ha[31] >> 8) + bin(int(ha[31]) & 255).count("1")), int((hb[31] >> 8) + bin(int(hb[31]) & 255).count("1"))
        ea, eb = ra[64:64 + 8 * ca].view(np.float64), rb[64:64 + 8 * cb].view(np.float64)
        print(f"   pair {p}: m {m} tile {tile} warp {warp} rows {tile * 16 + warp},{tile * 16 + warp + 8}; "
              f"first-record counts {ca} vs {cb}; header equal {bool((ha == hb).all())}")
        if ca == cb:
            dd = np.nonzero(ea != eb)[0]
            print(f"   E differs at {dd[:10].tolist()} of {ca}: {ea[dd[:4]].tolist()} vs {eb[dd[:4]].tolist()}")
        nb = np.nonzero(ra != rb)[0]
        print(f"   differing byte range {nb.min()}..{nb.max()} ({len(nb)} bytes)")
