"""CPU: host-side mirror of the reference interface (no kernels are launched here)."""
import os
import sys

import numpy as np
import pytest
import torch

from afldm_b200 import _lib, ops
from afldm_b200.af_modules import af_api
from afldm_b200.af_modules.af_blocks import AliasFreeDownsample2D, AliasFreeUpsample2D, WarpedNonlinearity
from afldm_b200.configs import FFHQ_AFVAE, FFHQ_DDIM, FFHQ_UNET
from afldm_b200.models import AliasFreeAutoencoderKL, UNet2DModel
from afldm_b200.pipelines.cross_frame_attn import (AttnState, CrossFrameAttnProcessor, get_unet_attn_processors,
                                                   set_unet_attn_processor)
from afldm_b200.schedulers import DDIMScheduler
from oracle import af_blocks as OA
from oracle import ideal_lpf as OL
from oracle import nn as ON

SMALL = dict(block_out_channels=[32, 64, 64], down_block_types=["AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
             up_block_types=["UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D"], attention_head_dim=8, sample_size=8)


def test_taps_match_oracle_operators():
    sys.path.insert(0, os.path.join(os.path.dirname(_lib.CSRC_DIR), "csrc"))
    import gen_taps
    for n in gen_taps.SIZES:
        u, d = OL.up2_matrix(n), OL.down2_matrix(n)
        dt, gt = gen_taps.up_taps(n), gen_taps.down_taps(n)
        for i in range(n):
            assert np.allclose(u[2 * i + 1], [dt[(i - j) % n] for j in range(n)], atol=1e-14)
            assert np.allclose(d[i], [gt[(2 * i - m) % (2 * n)] for m in range(2 * n)], atol=1e-14)
    # the committed include is what the generator renders
    with open(os.path.join(_lib.CSRC_DIR, "taps.inc")) as f:
        assert f.read() == gen_taps.render()


def test_unet_state_dict_matches_diffusers_layout():
    mine, ref = UNet2DModel.from_config(SMALL), ON.UNet2DModel(**SMALL)
    af_api.make_af_unet(mine)
    OA.make_af_unet(ref)
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    mine.load_state_dict(b)         # weights are interchangeable
    assert "down_blocks.0.downsamplers.0.conv.weight" in a and "up_blocks.0.upsamplers.0.conv.bias" in a
    assert "mid_block.attentions.0.to_out.0.weight" in a


def test_full_size_unet_parameter_count():
    with torch.device("meta"):
        u = UNet2DModel.from_config(FFHQ_UNET)
    assert sum(p.numel() for p in u.parameters()) == 256_401_796         # 256.4 M


def test_make_af_unet_swaps_exactly_what_the_reference_swaps():
    u = UNet2DModel.from_config(SMALL)
    convs = [b.downsamplers[0].conv for b in u.down_blocks if b.downsamplers is not None]
    assert af_api.make_af_unet(u) is None
    for blk in list(u.down_blocks) + [u.mid_block] + list(u.up_blocks):
        assert all(isinstance(r.nonlinearity, WarpedNonlinearity) for r in blk.resnets)
    downs = [b.downsamplers[0] for b in u.down_blocks if b.downsamplers is not None]
    assert all(isinstance(d, AliasFreeDownsample2D) for d in downs) and len(downs) == 2
    assert all(d.conv is c and d.conv.stride == (1, 1) for d, c in zip(downs, convs))
    assert all(isinstance(b.upsamplers[0], AliasFreeUpsample2D) for b in u.up_blocks if b.upsamplers is not None)
    assert u.down_blocks[-1].downsamplers is None and u.up_blocks[-1].upsamplers is None
    assert isinstance(u.conv_act, torch.nn.SiLU)        # the tail activation is NOT wrapped (af_api.py:70-83)


def test_make_af_vae_flags_follow_config():
    with torch.device("meta"):
        v = AliasFreeAutoencoderKL.from_config(FFHQ_AFVAE)
        o = ON.AutoencoderKL()
    OA.make_af_vae_from_config(o)
    wrapped = lambda blk: [isinstance(r.nonlinearity, (WarpedNonlinearity, OA.WarpedNonlinearity)) for r in blk.resnets]
    for mine, ref in zip(list(v.encoder.down_blocks) + list(v.decoder.up_blocks),
                         list(o.encoder.down_blocks) + list(o.decoder.up_blocks)):
        assert wrapped(mine) == wrapped(ref)
    assert wrapped(v.decoder.up_blocks[3]) == [False] * 3 and wrapped(v.encoder.down_blocks[0]) == [False] * 2
    assert all(wrapped(v.decoder.mid_block)) and all(wrapped(v.encoder.mid_block))
    assert [isinstance(b.downsamplers[0], AliasFreeDownsample2D) for b in v.encoder.down_blocks[:3]] == [True] * 3
    assert v.encoder.down_blocks[0].downsamplers[0].padding == 0
    assert v.downsample_ratio == 8 and v.config.scaling_factor == 0.6
    assert list(v.state_dict().keys()) == list(o.state_dict().keys())


def test_make_af_vae_reversed_resampling_flags():
    with torch.device("meta"):
        v = AliasFreeAutoencoderKL.from_config(FFHQ_AFVAE, up_rescale=[True, False, False], mid_act=False,
                                               down_filtered_act=[False] * 4, up_filtered_act=[False] * 4)
    ups = [isinstance(b.upsamplers[0], AliasFreeUpsample2D) for b in v.decoder.up_blocks[:3]]
    downs = [isinstance(b.downsamplers[0], AliasFreeDownsample2D) for b in v.encoder.down_blocks[:3]]
    assert ups == [True, False, False] and downs == [False, False, True]       # af_api.py:42


def test_ddim_schedule_matches_oracle():
    mine, ref = DDIMScheduler.from_config(FFHQ_DDIM), ON.DDIMScheduler()
    mine.set_timesteps(50)
    ref.set_timesteps(50)
    assert mine.timesteps.tolist() == ref.timesteps.tolist()
    assert mine.timesteps[0] == 981 and mine.timesteps[-1] == 1
    assert torch.equal(mine.alphas_cumprod, ref.alphas_cumprod)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for t in (981, 501, 21, 1):
        cx, ce = mine.coefficients(t)
        want = ref.step(e, t, x, return_dict=False)[0]
        torch.testing.assert_close(cx * x + ce * e, want, rtol=0, atol=2e-6)


def test_attn_state_and_processor_plumbing():
    st = AttnState()
    assert st.state == AttnState.STORE and st.timestep == 0 and st.store_id == 0
    st.set_timestep(torch.tensor(981))
    assert st.timestep == 981 and isinstance(st.timestep, int)
    st.to_load()
    assert st.state == AttnState.LOAD
    st.to_idle()
    assert st.state == AttnState.IDLE
    st.reset()
    assert st.state == AttnState.STORE
    u = UNet2DModel.from_config(SMALL)
    procs = get_unet_attn_processors(u)
    assert len(procs) == 2 + 2 + 1 + 3 + 3 and all(k.endswith(".processor") for k in procs)
    assert "mid_block.attentions.0.processor" in procs
    new = {k: CrossFrameAttnProcessor(st) for k in procs}
    set_unet_attn_processor(u, dict(new))
    assert all(get_unet_attn_processors(u)[k] is new[k] for k in new)
    set_unet_attn_processor(u, procs)
    with pytest.raises(ValueError):
        set_unet_attn_processor(u, {"x": None})


def test_ops_fail_loudly_without_cuda_tensors():
    x = torch.zeros(1, 8, 8, 32)
    with pytest.raises(_lib.AfldmError):
        ops.filtered_act(x)
    with pytest.raises(_lib.AfldmError):
        ops.conv2d(x, torch.zeros(32, 9, 32), None, 3)


def test_pack_conv_weight_layout():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_conv_weight(w)
    assert p.shape == (2, 9, 3)
    assert p[1, 5, 2] == w[1, 2, 1, 2]          # tap (kh=1, kw=2) -> index 5, channel last


def test_i2sb_schedule_matches_oracle():
    """I2SB scheduler scalars (i2sb_scheduler.py:188-197, 382-459): the collapsed one-kernel update
    x_prev = x + c_eps * eps equals the reference's mu_x0 * x0 + mu_xt * x_t formulation."""
    from afldm_b200.schedulers import I2SBScheduler
    from oracle.i2sb import I2SBScheduler as Ref
    mine, ref = I2SBScheduler.from_config(), Ref()
    mine.set_timesteps(100)
    ref.set_timesteps(100)
    assert mine.timesteps.tolist() == ref.timesteps.tolist() and mine.timesteps[0] == 991
    assert torch.equal(mine.std_fwd, ref.std_fwd) and torch.equal(mine.mu_x0, ref.mu_x0)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for t in (991, 501, 21, 11):
        c_eps, sqrt_var = mine.coefficients(t)
        want = ref.step(e, t, x, is_ode=True, return_dict=False)[0]
        torch.testing.assert_close(x + c_eps * e, want, rtol=0, atol=2e-6)
        g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
        want_sde = ref.step(e, t, x, is_ode=False, generator=g1, return_dict=False)[0]
        noise = torch.randn(e.shape, generator=g2)
        torch.testing.assert_close(x + c_eps * e + sqrt_var * noise, want_sde, rtol=0, atol=2e-6)
    x1 = torch.randn(2, 4, 8, 8, generator=g)
    ts = torch.tensor([10, 700])
    torch.testing.assert_close(mine.add_noise(x, x1, ts, is_ode=True), ref.add_noise(x, x1, ts, is_ode=True))
    torch.testing.assert_close(mine.compute_label(ts, x, x1), ref.compute_label(ts, x, x1))


# ------------------------------------------------------------------------- fp16 operand storage: host-side decisions
def test_conv_plan_query_for_fp16_operands():
    """afldm_conv2d_supported is a host-side plan query (no GPU): fp16 operands need 64-channel stages."""
    from afldm_b200 import _lib
    L = _lib.lib()
    assert L.afldm_conv2d_supported(16, 32, 32, 192, 192, 3, 2) == 1        # resnet conv at the 32x32 level
    assert L.afldm_conv2d_supported(16, 2, 2, 1536, 768, 3, 2) == 1         # split-K level
    assert L.afldm_conv2d_supported(16, 32, 32, 32, 192, 3, 2) == 0         # padded conv_in: 32 channels = half a stage
    assert L.afldm_conv2d_supported(16, 32, 32, 32, 192, 3, 1) == 1         # ... which the TF32 instantiation takes
    assert L.afldm_conv2d_supported(16, 32, 32, 4, 192, 3, 1) == 0          # raw latents: SIMT only
    assert L.afldm_conv2d_supported(16, 32, 32, 4, 192, 3, 0) == 1
    assert L.afldm_conv2d_supported(16, 24, 24, 192, 192, 3, 2) == 0        # non-power-of-two plane
    assert L.afldm_conv2d_supported(0, 32, 32, 192, 192, 3, 2) == 0
    # the workspace / slot queries accept the fp16 algo id
    assert L.afldm_conv2d_workspace_floats(16, 2, 2, 1536, 768, 3, 2) > 0
    assert L.afldm_conv2d_gn_slots(16, 32, 32, 192, 192, 3, 2) == 8


def test_conv_f16_supported_follows_class_and_switch():
    from afldm_b200 import ops
    prev = ops.default_conv_algo()
    try:
        ops.set_default_conv_algo("simt")
        assert not ops.conv_f16_supported(16, 32, 32, 192, 192, 3)           # exact-fp32 class never stores fp16
        ops.set_default_conv_algo("tf32")
        assert ops.conv_f16_supported(16, 32, 32, 192, 192, 3) == ops.F16_CONV
        assert not ops.conv_f16_supported(16, 32, 32, 96, 192, 3)
    finally:
        ops.set_default_conv_algo(prev)


def test_fp16_weight_pack_is_cached_and_invalidated():
    import torch
    from afldm_b200.packing import conv_params, conv_params_f16
    conv = torch.nn.Conv2d(64, 32, 3, padding=1)
    w16, b16, k = conv_params_f16(conv)
    w32, _, _ = conv_params(conv)
    assert w16.dtype == torch.float16 and w16.shape == (32, 9, 64) and k == 3
    assert torch.equal(w16, w32.half())
    assert conv_params_f16(conv)[0] is w16                                   # cached
    with torch.no_grad():
        conv.weight.mul_(2.0)                                                # in-place update bumps the version
    w16b = conv_params_f16(conv)[0]
    assert w16b is not w16 and torch.equal(w16b, conv_params(conv)[0].half())


def test_conv_plans_never_oversubscribe_tmem():
    """Host-side invariant behind the hang fix of DESIGN.md section 3: every tcgen05 convolution plan requests shared
    memory in proportion to its TMEM columns (>= 454 B per column, the 1 KB per-CTA reserve included), so whatever mix
    of conv CTAs fits on one SM (228 KB of shared memory) holds at most 512 columns and `tcgen05.alloc` never blocks.
    Checked over the layer shapes of all five BASELINE configurations (and more), both operand classes; no GPU needed."""
    import ctypes as C
    from afldm_b200 import _lib
    L = _lib.lib()
    plan = (C.c_int * 8)()
    SM_SMEM, RESERVE, MAX_DYN = 233472, 1024, 232448
    chans = (32, 64, 128, 192, 256, 320, 384, 512, 576, 640, 768, 960, 1152, 1280, 1536, 1920, 2560, 4096)
    seen, pairs, widest = 0, 0, 0
    for algo in (1, 2):
        for ks in (1, 3):
            for b in (1, 2, 4, 16, 32, 128):
                for n in (2, 4, 8, 16, 32, 64, 128, 256, 512):
                    if b * n * n > (1 << 25):
                        continue
                    for cin in chans:
                        for cout in (4, 32) + chans:
                            if L.afldm_conv2d_plan(b, n, n, cin, cout, ks, algo, plan) != 0:
                                continue
                            cols, smem, ctas, two = plan[0], plan[1], plan[2], plan[3]
                            seen += 1
                            pairs += two
                            widest = max(widest, cols)
                            assert cols in (32, 64, 128, 256, 512), (b, n, cin, cout, ks, algo, cols)
                            assert smem <= MAX_DYN and ctas >= 1
                            assert smem + RESERVE >= 454 * cols, (b, n, cin, cout, ks, algo, cols, smem)
                            if two:
                                assert ctas % 2 == 0
    assert seen > 5000 and pairs > 500 and widest >= 256
    # the arithmetic of the rule: CTAs that fit together have sum(smem + reserve) <= 228 KB, hence sum(cols) <= 514,
    # and column counts are multiples of 32
    assert SM_SMEM // 454 < 512 + 32
    assert L.afldm_conv2d_plan(16, 32, 32, 192, 192, 3, 0, plan) == -3          # the exact-FMA class has no such plan
    assert L.afldm_conv2d_plan(16, 32, 32, 100, 192, 3, 1, plan) == -3          # outside the tcgen05 family
