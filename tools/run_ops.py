#!/usr/bin/env python
"""Launch a few representative hot-path ops once each (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afldm_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for _ in range(3):
    if which in ("all", "fact"):
        x = r(16, 32, 32, 192)
        sc, sh = r(16, 192) * 0.1 + 1, r(16, 192) * 0.1
        ops.filtered_act(x, sc, sh)
        ops.filtered_act(r(16, 16, 16, 384))
    if which in ("all", "conv"):
        x = r(16, 32, 32, 192)
        w = ops.pack_conv_weight(r(192, 192, 3, 3) * 0.02)
        ops.conv2d(x, w, r(192), 3, residual=r(16, 32, 32, 192), algo="tf32", gn_stats=True)
        x2 = r(16, 16, 16, 384)
        w2 = ops.pack_conv_weight(r(384, 384, 3, 3) * 0.02)
        ops.conv2d(x2, w2, r(384), 3, algo="tf32", gn_stats=True)
        x3 = r(16, 32, 32, 384)
        w3 = ops.pack_conv_weight(r(384, 384, 3, 3) * 0.02)
        ops.conv2d(x3, w3, r(384), 3, algo="tf32")
    if which in ("conv16",):
        # the same three layers with fp16 operands (tcgen05.mma.kind::f16), as the resnets / the up-sampler run them
        x = r(16, 32, 32, 192).half()
        w = ops.pack_conv_weight(r(192, 192, 3, 3) * 0.02).half()
        ops.set_default_conv_algo("tf32")
        ops.conv2d(x, w, r(192), 3, residual=r(16, 32, 32, 192), gn_stats=True)
        x2 = r(16, 16, 16, 384).half()
        w2 = ops.pack_conv_weight(r(384, 384, 3, 3) * 0.02).half()
        ops.conv2d(x2, w2, r(384), 3, gn_stats=True)
        x3 = r(16, 32, 32, 384).half()
        w3 = ops.pack_conv_weight(r(384, 384, 3, 3) * 0.02).half()
        ops.conv2d(x3, w3, r(384), 3)
    if which in ("all", "attn"):
        qkv = r(16, 1024, 576)
        ops.attention(qkv[:, :, :192], qkv[:, :, 192:384], qkv[:, :, 384:], 8, algo="tf32")
        qh = qkv.half()
        ops.attention_f16(qh[:, :, :192], qh[:, :, 192:384], qh[:, :, 384:], 8)
    if which in ("large",):
        ops.filtered_act(r(4, 128, 128, 256))
    if which in ("all", "lin"):
        ops.linear_rows(r(16, 768), r(14016, 768) * 0.03, r(14016), act_in="silu")
torch.cuda.synchronize()
print("done")
