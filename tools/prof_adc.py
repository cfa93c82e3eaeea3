#!/usr/bin/env python
"""ncu target: one ADC search at the BASELINE config-2 shape (1200 queries x 8,841,823 docs, M=48, k=1000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repconc_b200.faiss_compat import GpuIndexPQ
N = int(os.environ.get("PN", 8841823)); M = int(os.environ.get("PM", 48)); nq = int(os.environ.get("PQ", 1200))
gen = torch.Generator(device="cuda").manual_seed(0)
codes = torch.randint(0, 256, (N, M), generator=gen, device="cuda", dtype=torch.uint8)
c = torch.randn((M, 256, 768 // M), generator=gen, device="cuda")
q = torch.randn((nq, 768), generator=gen, device="cuda")
idx = GpuIndexPQ(codes, c)
s, i = idx.search_tensor(q, 1000)
torch.cuda.synchronize()
print(idx.last_stats, float(s.sum()))
