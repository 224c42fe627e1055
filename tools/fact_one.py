"""One filtered-activation shape, a few launches (ncu target).  usage: fact_one.py B C n [half]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afldm_b200 import ops  # noqa: E402

b, c, n = (int(v) for v in sys.argv[1:4])
half = len(sys.argv) > 4
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
x = torch.randn(b, n, n, c, generator=g).to(dev)
sc = (torch.rand(b, c, generator=g) + 0.5).to(dev)
sh = (torch.randn(b, c, generator=g) * 0.3).to(dev)
for _ in range(4):
    y = ops.filtered_act(x, sc, sh, out_half=half)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
