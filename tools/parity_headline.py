"""Measure (print) the error of the benchmarked numeric class against the fp32 oracle at the headline configuration."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afldm_b200 import ops  # noqa: E402
from afldm_b200.af_modules import af_api  # noqa: E402
from afldm_b200.models import AliasFreeAutoencoderKL, UNet2DModel  # noqa: E402
from afldm_b200.pipelines import MyLDMPipeline  # noqa: E402
from afldm_b200.schedulers import DDIMScheduler  # noqa: E402
from oracle import af_blocks as OA  # noqa: E402
from oracle import nn as ON  # noqa: E402
from oracle import pipeline as OP  # noqa: E402

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def stats(a, b, name):
    d = (a - b).abs()
    print(f"{name}: max|d| {d.max().item():.3e} mean|d| {d.mean().item():.3e} max|ref| {b.abs().max().item():.3f} "
          f"rms ref {b.pow(2).mean().sqrt().item():.3f} rel-rms {(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item():.3e}", flush=True)


torch.manual_seed(0)
ref = ON.UNet2DModel().to(dev).eval()
mine = UNet2DModel.from_config().to(dev).eval()
mine.load_state_dict(ref.state_dict())
OA.make_af_unet(ref)
af_api.make_af_unet(mine)
g = torch.Generator().manual_seed(0)
lat16 = torch.randn(16, 4, 32, 32, generator=g).to(dev)
for algo in ("tf32", "simt"):
    ops.set_default_conv_algo(algo)
    with torch.no_grad():
        for t in (981, 501, 1):
            want = ref(lat16, torch.tensor(t, device=dev)).sample
            got = mine(lat16, t).sample
            stats(got, want, f"[{algo}] eps one step B=16 t={t}")
# 50-step drift, B=2
for algo in ("tf32", "simt"):
    ops.set_default_conv_algo(algo)
    pipe = MyLDMPipeline(None, mine, DDIMScheduler.from_config())
    x = lat16[:2]
    t0 = time.time()
    got = pipe.denoise(x, 50, use_cuda_graph=(algo == "tf32"))
    sr = ON.DDIMScheduler()
    traj = {}
    want = OP.denoise(ref, sr, x, 50, on_step=lambda i, l: traj.__setitem__(i, l.clone()))
    stats(got, want, f"[{algo}] latents after 50 DDIM steps B=2")
    # per-step drift for the tf32 class (eager, to see intermediate steps)
    if algo == "tf32":
        sm = DDIMScheduler.from_config()
        sm.set_timesteps(50)
        a = x
        with torch.no_grad():
            for i, t in enumerate(sm.timesteps):
                a = sm.step(mine(a, int(t)).sample, int(t), a).prev_sample
                if i in (0, 9, 24, 49):
                    stats(a.contiguous(), traj[i], f"   step {i + 1}")
    # inversion
    sm = DDIMScheduler.from_config()
    sm.set_timesteps(10)
    pipe.scheduler = sm
    sr.set_timesteps(10)
    inv = pipe.ddim_inversion(x, bar=False)
    stats(inv, OP.ddim_inversion(ref, sr, x), f"[{algo}] ddim_inversion 10 steps B=2")
del ref, mine
torch.cuda.empty_cache()
# VAE decode B=4
torch.manual_seed(0)
rv = ON.AutoencoderKL().to(dev).eval()
mv = AliasFreeAutoencoderKL.from_config().to(dev).eval()
mv.load_state_dict(rv.state_dict())
OA.make_af_vae_from_config(rv)
z = torch.randn(4, 4, 32, 32, generator=g).to(dev)
with torch.no_grad():
    want = rv.decode(z / 0.6).sample
    for algo in ("tf32", "simt"):
        ops.set_default_conv_algo(algo)
        stats(mv.decode(z / 0.6).sample, want, f"[{algo}] AF-VAE decode B=4")
