"""GPU parity of the host-side module mirror (afldm_b200.models / af_modules / pipelines) against
the oracle's restated diffusers blocks with IDENTICAL weights (state_dicts are interchangeable).
fp32 everywhere; the torch reference runs with TF32 disabled."""
import pytest
import torch

from afldm_b200 import _lib, ops
from afldm_b200.af_modules import af_api
from afldm_b200.models import UNet2DModel
from afldm_b200.models import blocks as B
from afldm_b200.pipelines import AttnState, CrossFrameAttnProcessor, MyLDMPipeline, set_unet_attn_processor
from afldm_b200.schedulers import DDIMScheduler
from oracle import af_blocks as OA
from oracle import cross_frame as OC
from oracle import nn as ON

pytestmark = pytest.mark.gpu
DEV = "cuda"

SMALL = dict(block_out_channels=[64, 128, 128], down_block_types=["AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
             up_block_types=["UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D"], attention_head_dim=8, sample_size=16)


@pytest.fixture(autouse=True, scope="module")
def _exact_torch_reference():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ops.set_default_conv_algo("simt")
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def jitter(module, seed=0):
    """Non-trivial norm affines / biases so that every parameter matters."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * 0.1)
            elif n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * 0.05)


def pair(mine, ref, seed=0):
    torch.manual_seed(seed)
    ref = ref.to(DEV)
    jitter(ref, seed)
    mine = mine.to(DEV)
    mine.load_state_dict(ref.state_dict())
    return mine.eval(), ref.eval()


@pytest.mark.parametrize("cin,cout,n,af", [(192, 192, 32, True), (576, 192, 32, True), (1536, 768, 2, True),
                                            (384, 384, 16, False), (768, 384, 8, True)])
def test_resnet_block(cin, cout, n, af):
    mine, ref = pair(B.ResnetBlock2D(cin, cout, 768), ON.ResnetBlock2D(cin, cout, 768), seed=cin)
    if af:
        mine.nonlinearity = af_api.wrap_nonlinearity(mine.nonlinearity)
        ref.nonlinearity = OA.WarpedNonlinearity(ref.nonlinearity)
    x, temb = randn(2, cin, n, n, seed=1), randn(2, 768, seed=2)
    with torch.no_grad():
        want = ref(x, temb)
        got = mine(x, temb)
    torch.testing.assert_close(got.contiguous(), want, rtol=0, atol=1e-4)


@pytest.mark.parametrize("c,n,heads", [(192, 32, 8), (384, 16, 16), (768, 2, 32)])
def test_attention_block(c, n, heads):
    mine, ref = pair(B.Attention(c, heads, c // heads), ON.Attention(c, heads, c // heads), seed=c)
    x = randn(2, c, n, n, seed=3)
    with torch.no_grad():
        torch.testing.assert_close(mine(x).contiguous(), ref(x), rtol=0, atol=1e-4)


def test_alias_free_resamplers():
    d_mine, d_ref = pair(B.Downsample2D(192, True, 192, padding=1), ON.Downsample2D(192, True, 192, padding=1))
    d_mine, d_ref = af_api.replace_downsampler(d_mine), OA.AliasFreeDownsample2D(192, True, 192, 1, d_ref.conv)
    x = randn(2, 192, 32, 32, seed=4)
    with torch.no_grad():
        torch.testing.assert_close(d_mine(x).contiguous(), d_ref(x), rtol=0, atol=5e-5)
    u_mine, u_ref = pair(B.Upsample2D(384, True, 384), ON.Upsample2D(384, True, 384))
    u_mine, u_ref = af_api.replace_upsampler(u_mine), OA.AliasFreeUpsample2D(384, True, 384, u_ref.conv)
    x = randn(2, 384, 16, 16, seed=5)
    with torch.no_grad():
        torch.testing.assert_close(u_mine(x).contiguous(), u_ref(x), rtol=0, atol=5e-5)
    # padding == 0 variant (VAE encoder, af_blocks.py:142-144)
    p_mine, p_ref = pair(B.Downsample2D(64, True, 64, padding=0), ON.Downsample2D(64, True, 64, padding=0))
    p_mine, p_ref = af_api.replace_downsampler(p_mine), OA.AliasFreeDownsample2D(64, True, 64, 0, p_ref.conv)
    x = randn(1, 64, 16, 16, seed=6)
    with torch.no_grad():
        torch.testing.assert_close(p_mine(x).contiguous(), p_ref(x), rtol=0, atol=5e-5)


def test_warped_nonlinearity_module_dropin():
    m = af_api.wrap_nonlinearity(torch.nn.SiLU())
    r = OA.WarpedNonlinearity(torch.nn.SiLU())
    x = randn(2, 64, 8, 8, seed=7)                                   # plain NCHW-contiguous input
    torch.testing.assert_close(m(x).contiguous(), r(x), rtol=0, atol=1e-5)
    t = randn(4, 768, seed=8)                                        # ndim < 4: plain activation
    torch.testing.assert_close(m(t), r(t), rtol=0, atol=2e-6)


def _small_unets(seed=0):
    mine, ref = pair(UNet2DModel.from_config(SMALL), ON.UNet2DModel(**SMALL), seed=seed)
    af_api.make_af_unet(mine)
    OA.make_af_unet(ref)
    return mine, ref


def test_small_unet_forward_and_ddim_trajectory():
    mine, ref = _small_unets()
    x = randn(3, 4, 16, 16, seed=9)
    with torch.no_grad():
        want = ref(x, torch.tensor(501, device=DEV)).sample
        got = mine(x, 501).sample
        torch.testing.assert_close(got.contiguous(), want, rtol=0, atol=2e-4)
        got_t = mine(x, torch.tensor([501.0] * 3, device=DEV), return_dict=False)[0]
        assert torch.equal(got_t, got)
        # 10 DDIM steps: trajectory drift stays small
        sm, sr = DDIMScheduler.from_config(), ON.DDIMScheduler()
        sm.set_timesteps(10)
        sr.set_timesteps(10)
        a, b = x, x
        for t in sm.timesteps:
            a = sm.step(mine(a, int(t)).sample, int(t), a).prev_sample
            b = sr.step(ref(b, t.to(DEV)).sample, int(t), b, return_dict=False)[0]
        torch.testing.assert_close(a.contiguous(), b, rtol=0, atol=1e-3)


def test_graphed_denoise_equals_eager():
    mine, _ = _small_unets(seed=1)
    pipe = MyLDMPipeline(None, mine, DDIMScheduler.from_config())
    x = randn(2, 4, 16, 16, seed=10)
    eager = pipe.denoise(x, 6, use_cuda_graph=False)
    n0 = _lib.launch_count()
    graphed = pipe.denoise(x, 6, use_cuda_graph=True)
    assert torch.equal(eager, graphed)                              # same kernels, same order: bitwise
    assert pipe.graphed(2).launches_per_step > 50
    again = pipe.denoise(x, 6, use_cuda_graph=True)                 # replay of the cached graph
    assert torch.equal(again, graphed)
    assert _lib.launch_count() >= n0


def test_cross_frame_attention_store_load():
    mine, ref = _small_unets(seed=2)
    st_m, st_r = AttnState(), OC.AttnState()
    set_unet_attn_processor(mine, {k: CrossFrameAttnProcessor(st_m) for k in
                                   __import__("afldm_b200.pipelines.cross_frame_attn", fromlist=["x"]).get_unet_attn_processors(mine)})
    OC.set_attn_processor(ref, lambda: OC.CrossFrameAttnProcessor(st_r))
    x0, x1 = randn(1, 4, 16, 16, seed=11), randn(4, 4, 16, 16, seed=12)
    with torch.no_grad():
        for st in (st_m, st_r):
            st.reset()
            st.set_timestep(torch.tensor(301))
        a0, b0 = mine(x0, 301).sample, ref(x0, torch.tensor(301, device=DEV)).sample
        torch.testing.assert_close(a0.contiguous(), b0, rtol=0, atol=2e-4)
        st_m.to_load()
        st_r.to_load()
        a1, b1 = mine(x1, 301).sample, ref(x1, torch.tensor(301, device=DEV)).sample    # K/V of frame 0, batch 4
        torch.testing.assert_close(a1.contiguous(), b1, rtol=0, atol=2e-4)
        assert (a1 - mine(x1, 301).sample).abs().max() == 0


def test_afvae_decode_and_encode_vs_oracle():
    """BASELINE config #3 architecture (configs/vae/model_afvae.json), B = 1: 4x32x32 -> 3x256x256,
    including the 64 / 128 planes, the d = 512 mid-block attention and the plain-SiLU last block."""
    from afldm_b200.models import AliasFreeAutoencoderKL
    torch.manual_seed(0)
    ref = ON.AutoencoderKL().to(DEV).eval()
    jitter(ref, 3)
    mine = AliasFreeAutoencoderKL.from_config().to(DEV).eval()
    mine.load_state_dict(ref.state_dict())
    OA.make_af_vae_from_config(ref)
    z = randn(1, 4, 32, 32, seed=13)
    with torch.no_grad():
        want = ref.decode(z / 0.6).sample
        got = mine.decode(z / 0.6).sample
        assert got.shape == (1, 3, 256, 256) and got.is_contiguous()
        err = (got - want).abs().max().item()
        assert err < 2e-4 * max(1.0, want.abs().max().item()), err
        assert (mine.decode_scale(z) - got).abs().max() == 0
        img = randn(1, 3, 64, 64, seed=14)
        m_want = ref.encode(img).latent_dist.mean
        m_got = mine.encode(img).latent_dist.mean
        assert m_got.shape == (1, 4, 8, 8)
        torch.testing.assert_close(m_got, m_want, rtol=0, atol=2e-4)


def test_afvae_decode_512_planes_vs_oracle():
    """The config #4 image size: 4x64x64 latents -> 3x512x512.  The 256 x 256 filtered activations and the 128 -> 256
    ideal up-sampler are above the fused kernels' plane range and run the operator-matrix form (ops._general_resample);
    the encoder's 512 -> 256 down-sampler likewise."""
    from afldm_b200.models import AliasFreeAutoencoderKL
    torch.manual_seed(0)
    ref = ON.AutoencoderKL().to(DEV).eval()
    jitter(ref, 5)
    mine = AliasFreeAutoencoderKL.from_config().to(DEV).eval()
    mine.load_state_dict(ref.state_dict())
    OA.make_af_vae_from_config(ref)
    z = randn(1, 4, 64, 64, seed=21)
    with torch.no_grad():
        want = ref.decode(z / 0.6).sample
        got = mine.decode(z / 0.6).sample
        assert got.shape == (1, 3, 512, 512)
        err = (got - want).abs().max().item()
        assert err < 3e-4 * max(1.0, want.abs().max().item()), err
        img = randn(1, 3, 512, 512, seed=22)
        m_want = ref.encode(img).latent_dist.mean
        m_got = mine.encode(img).latent_dist.mean
        assert m_got.shape == (1, 4, 64, 64)
        torch.testing.assert_close(m_got, m_want, rtol=0, atol=3e-4)


def test_i2sb_bridge_matches_oracle():
    """Config #5 path on a small UNet: the I2SB sampling loop (i2sb_pipeline.py:45-56), ODE mode."""
    from afldm_b200.pipelines import I2SBLDMPipeline
    from afldm_b200.schedulers import I2SBScheduler
    from oracle.i2sb import I2SBScheduler as RefSched
    mine, ref = _small_unets(seed=4)
    pipe = I2SBLDMPipeline(None, mine, I2SBScheduler.from_config())
    x = randn(2, 4, 16, 16, seed=21)
    got = pipe.bridge(x, num_inference_steps=6, is_ode=True)
    rs = RefSched()
    rs.set_timesteps(6)
    lat = x
    with torch.no_grad():
        for i, t in enumerate(rs.timesteps):
            if i == 5:
                break
            lat = rs.step(ref(lat, t.to(DEV)).sample, int(t), lat, is_ode=True, return_dict=False)[0]
    torch.testing.assert_close(got, lat, rtol=0, atol=1e-3)


def test_ldm_pipeline_call_end_to_end():
    """MyLDMPipeline.__call__ (ldm_pipeline.py:32-131): latents -> DDIM (CUDA graph) -> alias-free VAE decode."""
    from afldm_b200.models import AliasFreeAutoencoderKL
    mine, _ = _small_unets(seed=5)
    torch.manual_seed(1)
    vae = AliasFreeAutoencoderKL.from_config(block_out_channels=[32, 64], down_block_types=["DownEncoderBlock2D"] * 2,
                                             up_block_types=["UpDecoderBlock2D"] * 2, layers_per_block=1,
                                             down_filtered_act=[False, True], up_filtered_act=[True, False],
                                             up_rescale=[True]).to(DEV).eval()
    pipe = MyLDMPipeline(vae, mine, DDIMScheduler.from_config())
    g = torch.Generator().manual_seed(3)
    img = pipe(batch_size=2, generator=g, num_inference_steps=3, output_type="pt")
    assert img.shape == (2, 3, 32, 32) and torch.isfinite(img).all()
    g = torch.Generator().manual_seed(3)
    out = pipe(batch_size=2, generator=g, num_inference_steps=3, output_type="np")
    assert out.images.shape == (2, 32, 32, 3) and out.images.min() >= 0 and out.images.max() <= 1
    lat = pipe(batch_size=2, generator=torch.Generator().manual_seed(3), num_inference_steps=3, output_type="latent")
    eager = pipe(batch_size=2, generator=torch.Generator().manual_seed(3), num_inference_steps=3, output_type="latent",
                 use_cuda_graph=False)
    assert torch.equal(lat, eager)


@pytest.mark.slow
def test_full_ffhq_unet_step_vs_oracle():
    """BASELINE config #2 architecture (256.4 M parameters), B = 2, one forward."""
    mine, ref = pair(UNet2DModel.from_config(), ON.UNet2DModel(), seed=0)
    af_api.make_af_unet(mine)
    OA.make_af_unet(ref)
    x = randn(2, 4, 32, 32, seed=0)
    with torch.no_grad():
        want = ref(x, torch.tensor(981, device=DEV)).sample
        got = mine(x, 981).sample
    err = (got - want).abs().max().item()
    assert err < 5e-4 * max(1.0, want.abs().max().item()), err


def test_shift_sweep_batched_graph_matches_sequential_reference_loop():
    """MyLDMPipeline.shift_sweep (captured STORE pass + ONE batched, captured LOAD pass over all shifts, device-indexed
    map tables) against the reference's own flow (scripts/shift_ldm_ffhq.py:85-151): dictionary-keyed processors, one
    eager trajectory per shift."""
    from afldm_b200.pipelines import get_unet_attn_processors
    from afldm_b200.shift_utils import ImageShifter
    mine, _ = _small_unets(seed=6)
    pipe = MyLDMPipeline(None, mine, DDIMScheduler.from_config())
    init = randn(1, 4, 16, 16, seed=30)
    shifts = [(0.0, k / 8.0) for k in (1, 3, 8)] + [(0.5, -0.25)]
    steps = 5
    prev = get_unet_attn_processors(mine)
    base, outs, shifted, masks = pipe.shift_sweep(init, shifts, steps)
    assert outs.shape == (4, 4, 16, 16) and masks.shape == (4, 4, 16, 16)
    assert get_unet_attn_processors(mine) == prev                        # processors restored (:153-157)
    eager = pipe.shift_sweep(init, shifts, steps, use_cuda_graph=False)
    for a, b in zip((base, outs), eager[:2]):
        assert torch.equal(a, b)                                          # captured == eager, same kernels
    again = pipe.shift_sweep(init, shifts, steps)                         # cached graphs / tables replay
    assert torch.equal(again[1], outs)
    # the reference's loop: dict-keyed maps, sequential B = 1 trajectories
    st = AttnState()
    set_unet_attn_processor(mine, {k: CrossFrameAttnProcessor(st) for k in prev})
    shifter = ImageShifter("ideal_crop", 8)
    sch = DDIMScheduler.from_config()

    def denoise(x):
        sch.set_timesteps(steps)
        for t in sch.timesteps:
            st.set_timestep(t)
            x = sch.step(mine(x, int(t)).sample, int(t), x).prev_sample
        return x.contiguous()

    with torch.no_grad():
        st.reset()
        ref_base = denoise(init)
        st.to_load()
        torch.testing.assert_close(base, ref_base, rtol=0, atol=2e-5)
        for k, (ti, tj) in enumerate(shifts):
            x0, m = shifter.shift(init, ti, tj)
            torch.testing.assert_close(shifted[k:k + 1], x0, rtol=0, atol=1e-6)
            torch.testing.assert_close(outs[k:k + 1], denoise(x0), rtol=0, atol=5e-5)
    set_unet_attn_processor(mine, dict(prev))
