#!/usr/bin/env python
"""Developer probe (W ranks, library built with RC_NVCC_EXTRA=-DRC_XCHG_PROFILE): where the time of the fused
reduce + peer exchange + update kernel goes (block 0, mean per exchange)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from repconc_b200 import ops, _lib
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
M = int(os.environ.get("PM", 48)); B = 8192; T = 50
gen = torch.Generator(device="cuda").manual_seed(rank)
x = torch.randn((B, 768), generator=gen, device="cuda")
c = torch.randn((M, 256, 768 // M), generator=torch.Generator(device="cuda").manual_seed(7), device="cuda")
kern = ops.CudaAssignKernels(x, c)
for _ in range(3):
    ops.constrained_assign_driver(kern, 0.003, T, True)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    ops.constrained_assign_driver(kern, 0.003, T, True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
lib = _lib.load()
buf = np.zeros((8, 4), dtype=np.int64)
lib.rc_sinkhorn_debug_cta_times(kern.state.data_ptr(), B, M, 256, buf.ctypes.data, 8)
v = buf.reshape(-1)[:6]
n = max(int(v[5]), 1)
names = ["reduce", "publish+fence", "signal+wait", "peer sum", "update"]
print(f"rank {rank}: assignment {ms:.3f} ms; exchanges {n}: " +
      ", ".join(f"{nm} {v[i] / n / 1e3:.2f} us" for i, nm in enumerate(names)), flush=True)
dist.destroy_process_group()
