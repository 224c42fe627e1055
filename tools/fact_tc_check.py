"""GPU check + timing of the filtered activation (run with AFLDM_FACT_TC=0 / 1 to compare the mma.sync and tcgen05 kernels).
Compares with the oracle's FFT form (oracle/ideal_lpf.py, pinned to the reference goldens) on the same inputs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afldm_b200 import ops  # noqa: E402
from oracle import ideal_lpf as OL  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
print("AFLDM_FACT_TC =", os.environ.get("AFLDM_FACT_TC", "(default)"))
shapes = [(2, 8, 32), (2, 192, 32), (16, 192, 32), (16, 576, 32), (16, 384, 32), (2, 16, 16), (16, 384, 16), (16, 768, 16), (3, 48, 16)]
for (b, c, n) in shapes:
    g = torch.Generator().manual_seed(b * 1000 + c + n)
    x = torch.randn(b, c, n, n, generator=g).to(dev) * 1.5
    sc = (torch.rand(b, c, generator=g) + 0.5).to(dev).contiguous()
    sh = (torch.randn(b, c, generator=g) * 0.3).to(dev).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    want = OL.filtered_act_fft(x * sc[:, :, None, None] + sh[:, :, None, None])
    got = ops.filtered_act(xn, sc, sh).permute(0, 3, 1, 2)
    goth = ops.filtered_act(xn, sc, sh, out_half=True)
    err = (got - want).abs().max().item()
    errh = (goth.float().permute(0, 3, 1, 2) - want).abs().max().item()
    ident = ops.filtered_act(xn, act="identity").permute(0, 3, 1, 2)
    erri = (ident - OL.filtered_act_fft(x, act=lambda t: t)).abs().max().item()
    out = torch.empty_like(xn)
    for _ in range(3):
        ops.filtered_act(xn, sc, sh, out=out)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(20):
            ops.filtered_act(xn, sc, sh, out=out)
    gph.replay()
    ev0.record()
    gph.replay()
    ev1.record()
    ev1.synchronize()
    us = ev0.elapsed_time(ev1) * 1000 / 20
    gbs = 8.0 * x.numel() / (us * 1e-6) / 1e9
    print(f"B={b:3d} C={c:4d} n={n:2d}  max|err| fp32 {err:.2e}  fp16-out {errh:.2e}  identity {erri:.2e}   {us:8.2f} us  {gbs:7.1f} GB/s  (max|y| {want.abs().max().item():.2f})", flush=True)
