#!/usr/bin/env python
"""Does programmatic dependent launch change anything?  200 dependent tiny kernels, eager and in a CUDA graph.
Run twice: AFLDM_PDL=0 and AFLDM_PDL=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afldm_b200 import ops

x = torch.randn(1 << 14, device="cuda")
e = torch.randn(1 << 14, device="cuda")
def chain(n=200):
    for _ in range(n):
        ops.axpby(x, e, 0.999, 0.001, out=x)
chain(10)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); chain(); ev1.record(); ev1.synchronize()
print("PDL", os.environ.get("AFLDM_PDL", "unset"), "eager  us/kernel", ev0.elapsed_time(ev1) * 1e3 / 200)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
g.replay(); torch.cuda.synchronize()
ev0.record(); g.replay(); ev1.record(); ev1.synchronize()
print("PDL", os.environ.get("AFLDM_PDL", "unset"), "graph  us/kernel", ev0.elapsed_time(ev1) * 1e3 / 200)
