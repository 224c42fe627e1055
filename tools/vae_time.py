#!/usr/bin/env python
"""Time the alias-free VAE decode (BASELINE config #3 architecture) on cuda:0: this repo vs the oracle in eager PyTorch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afldm_b200 import ops, _lib
from afldm_b200.models import AliasFreeAutoencoderKL
from oracle import nn as ON, af_blocks as OA

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
torch.manual_seed(0)
ref = ON.AutoencoderKL().to(dev).eval()
mine = AliasFreeAutoencoderKL.from_config().to(dev).eval()
mine.load_state_dict(ref.state_dict())
OA.make_af_vae_from_config(ref)
z = torch.randn(B, 4, 32, 32, device=dev)
ops.set_default_conv_algo("tf32")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for name, f in (("afldm_b200", lambda: mine.decode(z / 0.6).sample), ("oracle eager", lambda: ref.decode(z / 0.6).sample)):
        f(); torch.cuda.synchronize()
        n0 = _lib.launch_count()
        ev0.record(); out = f(); ev1.record(); ev1.synchronize()
        print(f"{name}: B={B} decode {ev0.elapsed_time(ev1):.1f} ms  ({B / ev0.elapsed_time(ev1) * 1e3:.1f} img/s)  launches {_lib.launch_count() - n0}")
    a, b = mine.decode(z / 0.6).sample, ref.decode(z / 0.6).sample
    print("max |diff| tf32 class:", (a - b).abs().max().item(), "ref max", b.abs().max().item())

# per-kernel breakdown (device time, CUDA-graph timed) of one decode
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rec = []
ops.record_to(rec)
with torch.no_grad():
    mine.decode(z / 0.6)
ops.record_to(None)
table = bench.time_records(rec, torch, reps=3)
agg = {}
for (name, key), (count, ms, meta, _n) in table.items():
    k = name + (f" N={meta['N']}" if "N" in meta else "") + (f" {meta['H']}x{meta['Cin']}->{meta['Cout']} k{meta['k']}" if "Cin" in meta else "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += count; a[1] += count * ms
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:45s} x{v[0]:3d} {v[1]:8.2f} ms")
print("sum", sum(v[1] for v in agg.values()))
