#!/usr/bin/env python
"""ncu target: one constrained assign at the BASELINE shape with few iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repconc_b200 import ops
M = int(os.environ.get("PM", 48)); T = int(os.environ.get("PT", 4)); B = int(os.environ.get("PB", 8192))
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 768), generator=gen, device="cuda")
c = torch.randn((M, 256, 768 // M), generator=gen, device="cuda")
codes = ops.constrained_assign(x, c, 0.003, T, distributed=False)
torch.cuda.synchronize()
print(codes.sum().item())
