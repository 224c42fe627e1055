"""GPU: the SD-1.5 ``UNet2DConditionModel`` mirror and ``VideoEquivariantEditingPipeline`` (BASELINE config #4,
/root/reference/afldm/pipelines/video_equiv_editing_pipeline.py:330-748) against the oracle's plain-PyTorch restatement
with identical weights.  fp32 oracle (TF32 off); exact-fp32 kernel class unless a test says otherwise."""
import pytest
import torch

from afldm_b200 import ops
from afldm_b200.af_modules import af_api
from afldm_b200.models import AliasFreeAutoencoderKL, UNet2DConditionModel
from afldm_b200.models import unet_2d_condition as MC
from afldm_b200.pipelines import VideoEquivariantEditingPipeline
from afldm_b200.schedulers import DDIMScheduler
from oracle import af_blocks as OA
from oracle import nn as ON
from oracle import nn_cond as OCN
from oracle import video as OV

pytestmark = pytest.mark.gpu
DEV = "cuda"
SMALL = dict(block_out_channels=[64, 128, 128, 128], attention_head_dim=8, cross_attention_dim=96, sample_size=16)
SD_DDIM = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1, set_alpha_to_one=False)


@pytest.fixture(autouse=True, scope="module")
def _exact_torch_reference():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    prev = ops.default_conv_algo()
    ops.set_default_conv_algo("simt")
    yield
    ops.set_default_conv_algo(prev)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def jitter(module, seed=0):
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


@pytest.mark.parametrize("dim,heads,tokens", [(320, 8, 256), (640, 8, 64), (1280, 8, 16)])
def test_basic_transformer_block(dim, heads, tokens):
    """LayerNorm -> attn1 -> LayerNorm -> attn2 (77 text tokens) -> LayerNorm -> GEGLU feed-forward; head dims 40 / 80 / 160."""
    mine, ref = pair(MC.BasicTransformerBlock(dim, heads, dim // heads, 768), OCN.BasicTransformerBlock(dim, heads, dim // heads, 768), seed=dim)
    h, ehs = randn(2, tokens, dim, seed=1), randn(2, 77, 768, seed=2)
    with torch.no_grad():
        torch.testing.assert_close(mine(h, ehs), ref(h, ehs), rtol=0, atol=2e-4)


def test_transformer_2d_model():
    mine, ref = pair(MC.Transformer2DModel(8, 40, 320, 768), OCN.Transformer2DModel(8, 40, 320, 768), seed=3)
    x, ehs = randn(2, 320, 16, 16, seed=4), randn(2, 77, 768, seed=5)
    with torch.no_grad():
        torch.testing.assert_close(mine(x, ehs).contiguous(), ref(x, ehs), rtol=0, atol=2e-4)


def test_geglu_and_layer_norm_ops():
    p = randn(3, 50, 256, seed=6)
    a, g = p.chunk(2, dim=-1)
    torch.testing.assert_close(ops.geglu(p), a * torch.nn.functional.gelu(g), rtol=0, atol=2e-6)
    x, w, b = randn(4, 33, 320, seed=7), randn(320, seed=8), randn(320, seed=9)
    torch.testing.assert_close(ops.layer_norm(x, w, b, 1e-5), torch.nn.functional.layer_norm(x, (320,), w, b, 1e-5), rtol=0, atol=1e-5)


def _small_pair(seed=0):
    mine, ref = pair(UNet2DConditionModel.from_config(SMALL), OCN.UNet2DConditionModel(**SMALL), seed=seed)
    af_api.make_af_unet(mine)
    OA.make_af_unet(ref)
    return mine, ref


def test_small_cond_unet_forward_and_processor_surface():
    mine, ref = _small_pair()
    x, ehs = randn(2, 4, 16, 16, seed=10), randn(2, 77, 96, seed=11)
    with torch.no_grad():
        want = ref(x, torch.tensor(401, device=DEV), encoder_hidden_states=ehs).sample
        got = mine(x, 401, encoder_hidden_states=ehs).sample
    torch.testing.assert_close(got.contiguous(), want, rtol=0, atol=3e-4)
    procs = mine.attn_processors                                  # diffusers surface (video_equiv_editing_pipeline.py:505-512)
    assert len(procs) == 2 * (2 * 3 + 1 + 3 * 3) and all(k.endswith(".processor") for k in procs)
    mine.set_attn_processor(dict(procs))
    with pytest.raises(ValueError):
        mine.set_attn_processor({k: v for k, v in list(procs.items())[:3]})


def test_video_pipeline_matches_oracle_loop():
    """3 frames, 6 DDIM steps at strength 0.7 (4 used), guidance 7.5: per-frame inversion with cross-frame attention,
    STORE pass of frame 0, batched LOAD denoising of all frames, per-frame VAE decode - against the reference's frame-by-frame
    loop restated on the oracle models (oracle/video.py)."""
    mine, ref = _small_pair(seed=1)
    torch.manual_seed(2)
    vcfg = dict(block_out_channels=[32, 64], down_block_types=["DownEncoderBlock2D"] * 2, up_block_types=["UpDecoderBlock2D"] * 2,
                layers_per_block=1, down_filtered_act=[False, True], up_filtered_act=[True, False], up_rescale=[True],
                scaling_factor=0.18215)
    rv = ON.AutoencoderKL(**vcfg).to(DEV).eval()
    jitter(rv, 5)
    mv = AliasFreeAutoencoderKL.from_config(**vcfg).to(DEV).eval()
    mv.load_state_dict(rv.state_dict())
    OA.make_af_vae_from_config(rv)
    frames = [randn(1, 3, 32, 32, seed=20 + i).clamp(-1, 1) for i in range(3)]
    pos, neg, inv = (randn(1, 77, 96, seed=s) for s in (30, 31, 32))
    pipe = VideoEquivariantEditingPipeline(mv, mine, DDIMScheduler.from_config(SD_DDIM))
    before = mine.attn_processors
    out = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
               strength=0.7, guidance_scale=7.5, output_type="pt")
    assert mine.attn_processors == before                          # processors restored (:743)
    lat = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
               strength=0.7, guidance_scale=7.5, output_type="latent").images
    want_lat, want_img = OV.video_edit(ref, rv, ON.DDIMScheduler(**SD_DDIM), frames, pos, neg, inv, 6, 0.7, 7.5)
    assert out.images.shape == (3, 3, 32, 32)
    torch.testing.assert_close(lat.contiguous(), want_lat, rtol=0, atol=2e-3)
    torch.testing.assert_close(out.images, want_img, rtol=0, atol=5e-3)
    # frame-by-frame evaluation (the reference's schedule) gives the same result as the batched one
    one = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
               strength=0.7, guidance_scale=7.5, output_type="latent", frame_batch=1).images
    torch.testing.assert_close(one.contiguous(), lat.contiguous(), rtol=0, atol=1e-4)
    # the same loops replayed from captured steps (device-indexed cross-frame map tables)
    cap = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
               strength=0.7, guidance_scale=7.5, output_type="latent", use_cuda_graph=True).images
    torch.testing.assert_close(cap.contiguous(), lat.contiguous(), rtol=0, atol=1e-5)
    # a second call reuses the captured steps with NEW text conditions (refreshed in place) - still the eager result
    pos2 = pos.flip(1).contiguous()
    eag2 = pipe(frames, prompt_embeds=pos2, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
                strength=0.7, guidance_scale=7.5, output_type="latent").images
    n_cached = len(pipe._captured)
    cap2 = pipe(frames, prompt_embeds=pos2, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
                strength=0.7, guidance_scale=7.5, output_type="latent", use_cuda_graph=True).images
    assert len(pipe._captured) == n_cached and not torch.equal(eag2, lat)
    # the inversion pass stores 1 map per layer, the guided pass 2: both tables must stay allocated, the captured
    # steps address them by pointer
    procs = next(iter(pipe._captured.values()))["procs"]
    assert all(len(p._pairs[0]) == 2 for p in procs.values() if p._pairs[0])
    torch.testing.assert_close(cap2.contiguous(), eag2.contiguous(), rtol=0, atol=1e-5)
    # changed weights invalidate the captured steps (they hold pointers to packed copies of the old ones)
    with torch.no_grad():
        mine.conv_in.weight.mul_(1.05)
    eag3 = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
                strength=0.7, guidance_scale=7.5, output_type="latent").images
    cap3 = pipe(frames, prompt_embeds=pos, negative_prompt_embeds=neg, inv_prompt_embeds=inv, num_inference_steps=6,
                strength=0.7, guidance_scale=7.5, output_type="latent", use_cuda_graph=True).images
    assert not torch.equal(eag3, lat)
    torch.testing.assert_close(cap3.contiguous(), eag3.contiguous(), rtol=0, atol=1e-5)
    assert mine.attn_processors == before
    with pytest.raises(ValueError):
        pipe(frames, prompt="a red car", num_inference_steps=2)    # no text encoder in this build: embeddings required


@pytest.mark.slow
@pytest.mark.parametrize("algo,tmax,tmean", [("simt", 5e-4, 5e-5), ("tf32", 2e-2, 2e-3)])
def test_full_sd15_unet_step_vs_oracle(algo, tmax, tmean):
    """BASELINE config #4 architecture (859.5 M parameters, 64 x 64 latents, CFG batch of 2, 77 x 768 text states)."""
    ops.set_default_conv_algo(algo)
    mine, ref = pair(UNet2DConditionModel.from_config(), OCN.UNet2DConditionModel(), seed=0)
    af_api.make_af_unet(mine)
    OA.make_af_unet(ref)
    assert sum(p.numel() for p in mine.parameters()) == 859_520_964
    x, ehs = randn(2, 4, 64, 64, seed=12), randn(2, 77, 768, seed=13)
    with torch.no_grad():
        want = ref(x, torch.tensor(681, device=DEV), encoder_hidden_states=ehs).sample
        got = mine(x, 681, encoder_hidden_states=ehs).sample
    d = (got - want).abs()
    print(f"[{algo}] SD-1.5 AF-UNet B=2 64x64: max|d eps| {d.max().item():.3e} mean {d.mean().item():.3e} (eps rms {want.pow(2).mean().sqrt().item():.3f})")
    assert d.max().item() < tmax * max(1.0, want.abs().max().item()) and d.mean().item() < tmean
    del mine, ref
    torch.cuda.empty_cache()


@pytest.mark.parametrize("d,nq,nk,bkv", [(80, 1024, 1024, 2), (160, 256, 256, 1), (80, 1024, 77, 2), (160, 64, 77, 2)])
def test_attention_head_dims_80_160(d, nq, nk, bkv):
    """The TF32 tensor-core attention kernel at the head dims of SD-1.5's deeper levels, self- and cross-attention shapes,
    K/V batch smaller than the query batch (cross-frame LOAD)."""
    heads, b = 8, 2
    q, k, v = randn(b, nq, heads * d, seed=1), randn(bkv, nk, heads * d, seed=2), randn(bkv, nk, heads * d, seed=3)
    got = ops.attention(q, k, v, heads, algo="tf32")
    qh = q.view(b, nq, heads, d).transpose(1, 2)
    kh = k.view(bkv, nk, heads, d).transpose(1, 2).repeat_interleave(b // bkv, dim=0)
    vh = v.view(bkv, nk, heads, d).transpose(1, 2).repeat_interleave(b // bkv, dim=0)
    want = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(b, nq, heads * d)
    err = (got - want).abs()
    assert err.max().item() < 8e-3 and err.mean().item() < 8e-4, (err.max().item(), err.mean().item())
