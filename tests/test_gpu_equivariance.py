"""GPU: the fractional-shift equivariance measurement of scripts/shift_ldm_ffhq.py:50-159, run on the
CUDA path and on the oracle with identical weights / latents: the masked PSNR between
"shift the latent, then denoise" and "denoise, then shift" (afldm/shift_utils/metrics.py:11-15,
shifters.py:163-191) must agree.  Small UNet, 16x16 latents, 4 DDIM steps, cross-frame attention in
STORE (reference trajectory) / LOAD (shifted trajectories) mode, exact-fp32 kernels."""
import pytest
import torch

from afldm_b200.shift_utils import ImageShifter
from afldm_b200 import ops
from afldm_b200.af_modules import af_api
from afldm_b200.models import UNet2DModel
from afldm_b200.pipelines import (AttnState, CrossFrameAttnProcessor, get_unet_attn_processors,
                                  set_unet_attn_processor)
from afldm_b200.schedulers import DDIMScheduler
from oracle import af_blocks as OA
from oracle import cross_frame as OC
from oracle import nn as ON
from oracle import shift as OS

pytestmark = pytest.mark.gpu
DEV = "cuda"
CFG = dict(block_out_channels=[64, 128, 128], down_block_types=["AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
           up_block_types=["UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D"], attention_head_dim=8, sample_size=16)
STEPS = 4


def _denoise(unet, sched, state, x, tensor_t):
    sched.set_timesteps(STEPS)
    for t in sched.timesteps:
        state.set_timestep(t)
        eps = unet(x, t.to(DEV) if tensor_t else int(t), return_dict=False)[0]
        x = sched.step(eps, int(t), x, return_dict=False)[0]
    return x.contiguous()


@pytest.mark.parametrize("algo,tol_db,tol_traj", [("simt", 0.05, 1e-3), ("tf32", 0.05, 3e-2)])
def test_shift_equivariance_metric_matches_oracle(algo, tol_db, tol_traj):
    """``simt``: exact-fp32 kernels, the masked shift-PSNR equals the oracle's to 0.05 dB.  ``tf32``: the benchmarked
    class (TF32 convolutions, fp16-operand attention where the projections are un-split, tensor-core filtered
    activation) against the SAME fp32 oracle: the equivariance metric moves by less than 0.05 dB (measured 0.008 dB:
    41.132 / 34.096 vs 41.124 / 34.088, fp16 operand storage on)."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    prev_algo = ops.default_conv_algo()
    ops.set_default_conv_algo(algo)
    try:
        torch.manual_seed(0)
        ref = ON.UNet2DModel(**CFG).to(DEV).eval()
        mine = UNet2DModel.from_config(CFG).to(DEV).eval()
        mine.load_state_dict(ref.state_dict())
        OA.make_af_unet(ref)
        af_api.make_af_unet(mine)
        st_r, st_m = OC.AttnState(), AttnState()
        OC.set_attn_processor(ref, lambda: OC.CrossFrameAttnProcessor(st_r))
        set_unet_attn_processor(mine, {k: CrossFrameAttnProcessor(st_m) for k in get_unet_attn_processors(mine)})
        g = torch.Generator().manual_seed(7)
        init = torch.randn(1, 4, 16, 16, generator=g).to(DEV)
        with torch.no_grad():
            st_r.reset()
            st_m.reset()
            base_r = _denoise(ref, ON.DDIMScheduler(), st_r, init, True)
            base_m = _denoise(mine, DDIMScheduler.from_config(), st_m, init, False)
            torch.testing.assert_close(base_m, base_r, rtol=0, atol=tol_traj)
            st_r.to_load()
            st_m.to_load()
            psnr_r, psnr_m, psnr_d = [], [], []
            dev_shifter = ImageShifter("ideal_crop", 8)
            for k in (1, 4):                       # shifts of k/8 latent pixels (scripts: offsets i/8)
                tj = k / 8.0
                shifted, mask = OS.ideal_shift(init, 0.0, tj, 8, crop=True)
                want_r, _ = OS.ideal_shift(base_r, 0.0, tj, 8, crop=True)
                want_m, _ = OS.ideal_shift(base_m, 0.0, tj, 8, crop=True)
                out_r = _denoise(ref, ON.DDIMScheduler(), st_r, shifted, True)
                out_m = _denoise(mine, DDIMScheduler.from_config(), st_m, shifted, False)
                psnr_r.append(float(OS.mask_psnr(out_r, want_r, mask)))
                psnr_m.append(float(OS.mask_psnr(out_m, want_m, mask)))
                # the same measurement with the warps on device too (afldm_b200.shift_utils: one fused operator per axis)
                shifted_d, mask_d = dev_shifter.shift(init, 0.0, tj)
                want_d, _ = dev_shifter.shift(base_m, 0.0, tj)
                out_d = _denoise(mine, DDIMScheduler.from_config(), st_m, shifted_d, False)
                psnr_d.append(float(OS.mask_psnr(out_d, want_d, mask_d)))
        print(f"masked shift-PSNR (dB) [{algo}]  oracle:", [round(p, 3) for p in psnr_r], " cuda:", [round(p, 3) for p in psnr_m])
        print("   with on-device shifters:", [round(p, 3) for p in psnr_d])
        for pr, pm, pd in zip(psnr_r, psnr_m, psnr_d):
            assert abs(pr - pm) < tol_db, (psnr_r, psnr_m)
            assert abs(pr - pd) < tol_db, (psnr_r, psnr_d)
    finally:
        ops.set_default_conv_algo(prev_algo)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
