"""GPU parity of the fp16-operand tensor-core convolution (afldm_conv2d_f16in_f32, tcgen05.mma.kind::f16) and of the
fp16-output filtered activation that feeds it (afldm_filtered_act_gn_f16out / _cat_f16out).

Numeric class: products of operands rounded to 11 significant bits (fp16, round to nearest - the width TF32 keeps),
fp32 accumulation in TMEM, fp32 epilogue / output.  Tolerances:
  * vs an fp64 convolution of the SAME fp16-rounded operands: <= 2e-4 abs (only the fp32 accumulation order differs);
  * vs an fp64 convolution of the un-rounded fp32 operands: <= 8e-3 max / 1.5e-3 mean, the TF32-class bound of
    tests/test_gpu_conv_tc.py;
  * filtered activation with fp16 stores == fp16(round to nearest) of the fp32-store kernel's result, bit for bit.
"""
import pytest
import torch
import torch.nn.functional as F

from afldm_b200 import ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ops.F16_CONV, reason="fp16 operand storage switched off (AFLDM_CONV_F16=0)")]
DEV = "cuda"


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.fixture(autouse=True)
def tf32_class():
    prev = ops.default_conv_algo()
    ops.set_default_conv_algo("tf32")
    yield
    ops.set_default_conv_algo(prev)


def run(*args, **kw):
    rec = []
    ops.record_to(rec)
    try:
        out = ops.conv2d(*args, **kw)
    finally:
        ops.record_to(None)
    return out, [r[0] for r in rec]


F16_CASES = [  # B, H, W, Cin, Cout, k: the resnet conv1 / conv2 shapes of the FFHQ UNet (halo pairs, single CTAs, split-K)
    (16, 32, 32, 192, 192, 3), (2, 32, 32, 576, 192, 3), (16, 32, 32, 384, 192, 3), (16, 16, 16, 384, 384, 3),
    (4, 16, 16, 768, 384, 3), (16, 16, 16, 576, 384, 3), (16, 8, 8, 384, 384, 3), (16, 8, 8, 1152, 384, 3),
    (3, 4, 4, 768, 768, 3), (16, 4, 4, 1536, 768, 3), (16, 2, 2, 1536, 768, 3), (16, 2, 2, 768, 768, 3),
    (1, 64, 64, 128, 128, 3), (2, 32, 32, 512, 512, 3), (2, 16, 16, 64, 96, 1), (16, 32, 32, 192, 576, 1), (5, 1, 1, 64, 64, 1),
]


@pytest.mark.parametrize("b,h,w,cin,cout,k", F16_CASES)
def test_conv2d_f16_operands_vs_fp64(b, h, w, cin, cout, k):
    assert ops.conv_f16_supported(b, h, w, cin, cout, k)
    x = randn(b, cin, h, w, seed=cin + h)
    wt = randn(cout, cin, k, k, seed=cout) * (1.0 / (cin * k * k) ** 0.5)
    bias, row, res = randn(cout, seed=3), randn(b, cout, seed=4), randn(b, cout, h, w, seed=5)
    xh, wh = x.half(), wt.half()
    add = bias.double()[None, :, None, None] + row.double()[:, :, None, None] + res.double()
    got, names = run(nhwc(xh), ops.pack_conv_weight(wh), bias, k, row_add=row, residual=nhwc(res))
    assert names == ["conv2d_f16"], names
    assert got.dtype == torch.float32
    got = got.permute(0, 3, 1, 2)
    same_operands = (F.conv2d(xh.double(), wh.double(), None, padding=k // 2) + add).float()
    torch.testing.assert_close(got, same_operands, rtol=0, atol=2e-4)
    full = (F.conv2d(x.double(), wt.double(), None, padding=k // 2) + add).float()
    err = (got - full).abs()
    assert err.max().item() < 8e-3 and err.mean().item() < 1.5e-3, (err.max().item(), err.mean().item())


def test_conv2d_f16_matches_tf32_path_on_exact_inputs():
    """Operands exactly representable in both containers: the kind::f16 and kind::tf32 instantiations agree to the
    fp32 accumulation order."""
    b, h, w, cin, cout = 4, 16, 16, 384, 384
    x = ((randn(b, cin, h, w, seed=1) * 8).round() / 8).clamp(-4, 4)
    wt = (randn(cout, cin, 3, 3, seed=2) * 0.02 * 1024).round() / 1024
    a, na = run(nhwc(x.half()), ops.pack_conv_weight(wt.half()), None, 3)
    c, nc = run(nhwc(x), ops.pack_conv_weight(wt), None, 3, algo="tf32")
    assert na == ["conv2d_f16"] and nc == ["conv2d_tf32"]
    torch.testing.assert_close(a, c, rtol=0, atol=2e-4)


def test_conv2d_f16_groupnorm_partials_and_determinism():
    b, h, cin, cout = 16, 8, 384, 384
    x = nhwc(randn(b, cin, h, h, seed=1)).half()
    wp = ops.pack_conv_weight(randn(cout, cin, 3, 3, seed=2) * (1.0 / (cin * 9) ** 0.5)).half()
    gamma, beta = randn(cout, seed=5) * 0.2 + 1, randn(cout, seed=6) * 0.2
    y = ops.conv2d(x, wp, None, 3, gn_stats=True)
    y2 = ops.conv2d(x, wp, None, 3, gn_stats=True)
    assert torch.equal(y, y2)
    assert getattr(y, "_afldm_gn", None) is not None
    sc, sh = ops.groupnorm_affine(y, 32, 1e-5, gamma, beta)
    plain = y.clone()                                           # no partial sums attached: statistics pass
    sc0, sh0 = ops.groupnorm_affine(plain, 32, 1e-5, gamma, beta)
    torch.testing.assert_close(sc, sc0, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sh, sh0, rtol=1e-4, atol=1e-4)


def test_conv2d_dtype_mismatch_is_an_error():
    x = nhwc(randn(2, 64, 8, 8, seed=1))
    wp = ops.pack_conv_weight(randn(64, 64, 3, 3, seed=2) * 0.05)
    with pytest.raises(RuntimeError):
        ops.conv2d(x.half(), wp, None, 3)
    with pytest.raises(RuntimeError):
        ops.conv2d(x, wp.half(), None, 3)


@pytest.mark.parametrize("n,c", [(32, 192), (16, 384), (8, 384), (4, 768), (2, 768)])
def test_filtered_act_f16_store_is_rounded_f32_result(n, c):
    """The producer side: GroupNorm (from the conv epilogue's partial sums) + filtered activation with fp16 stores
    equals the fp32-store kernel's result rounded to nearest fp16."""
    b = 4
    if n <= 16:     # producer = the ideal down-sampler (emits one partial-sum slot per image)
        y = ops.lpf_down2(nhwc(randn(b, c, 2 * n, 2 * n, seed=n)), gn_stats=True)
    else:           # producer = the tensor-core convolution's epilogue
        x = nhwc(randn(b, 64, n, n, seed=n))
        y = ops.conv2d(x, ops.pack_conv_weight(randn(c, 64, 1, 1, seed=c) * 0.2), None, 1, algo="tf32", gn_stats=True)
    assert getattr(y, "_afldm_gn", None) is not None
    gamma, beta = randn(c, seed=5) * 0.2 + 1, randn(c, seed=6) * 0.2
    f32 = ops.filtered_act_groupnorm(y, 32, 1e-5, gamma, beta, act="silu")
    f16 = ops.filtered_act_groupnorm(y, 32, 1e-5, gamma, beta, act="silu", out_half=True)
    assert f32.dtype == torch.float32 and f16.dtype == torch.float16
    assert torch.equal(f16, f32.half())


def test_filtered_act_cat_f16_store():
    b, n, ca, cb = 2, 16, 384, 192
    x = nhwc(randn(b, 64, n, n, seed=1))
    ya = ops.conv2d(x, ops.pack_conv_weight(randn(ca, 64, 1, 1, seed=2) * 0.2), None, 1, algo="tf32", gn_stats=True)
    yb = ops.conv2d(x, ops.pack_conv_weight(randn(cb, 64, 1, 1, seed=3) * 0.2), None, 1, algo="tf32", gn_stats=True)
    gamma, beta = randn(ca + cb, seed=5) * 0.2 + 1, randn(ca + cb, seed=6) * 0.2
    f32 = ops.filtered_act_groupnorm_cat(ya, yb, 32, 1e-5, gamma, beta)
    f16 = ops.filtered_act_groupnorm_cat(ya, yb, 32, 1e-5, gamma, beta, out_half=True)
    if f32 is None:
        pytest.skip("concat form unavailable under the current switches")
    assert f16.dtype == torch.float16
    assert torch.equal(f16, f32.half())


def test_resnet_block_uses_f16_operands_and_stays_in_class():
    """A surgered ResnetBlock2D in the TF32 class runs conv1 / conv2 with fp16 operands and agrees with the exact-fp32
    class within the block-level TF32 tolerance."""
    from afldm_b200.af_modules.af_blocks import WarpedNonlinearity
    from afldm_b200.models.blocks import ResnetBlock2D
    torch.manual_seed(0)
    blk = ResnetBlock2D(192, 384, 768).to(DEV).eval()
    blk.nonlinearity = WarpedNonlinearity(blk.nonlinearity)
    x = randn(4, 64, 16, 16, seed=1)
    temb = randn(4, 768, seed=2)
    pre = ops.conv2d(nhwc(x), ops.pack_conv_weight(randn(192, 64, 1, 1, seed=3) * 0.2), None, 1, algo="tf32", gn_stats=True)
    rec = []
    with torch.no_grad():
        ops.record_to(rec)
        try:
            got = blk(ops.nchw_view(pre), temb)
        finally:
            ops.record_to(None)
        ops.set_default_conv_algo("simt")
        want = blk(ops.nchw_view(pre.clone()), temb)
    names = [r[0] for r in rec]
    if ops.F16_CONV:
        assert names.count("conv2d_f16") == 2, names
    err = (got - want).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 2e-3, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("n,c", [(2, 768), (4, 768), (8, 384), (16, 384), (32, 64), (64, 64), (128, 32)])
def test_up2_ideal_f16_store_is_rounded_f32_result(n, c):
    x = nhwc(randn(3 if n <= 32 else 1, c, n, n, seed=n))
    f32 = ops.up2_ideal(x)
    f16 = ops.up2_ideal(x, out_half=True)
    assert f16.dtype == torch.float16 and f16.shape == f32.shape
    assert torch.equal(f16, f32.half())


@pytest.mark.parametrize("n,c", [(8, 64), (16, 64), (32, 64), (64, 64), (128, 32)])
def test_filtered_act_plain_f16_store(n, c):
    """afldm_filtered_act_f16out (explicit scale / shift; the three line passes for planes of 64 / 128)."""
    b = 2
    x = nhwc(randn(b, c, n, n, seed=n))
    sc, sh = randn(b, c, seed=1) * 0.1 + 1, randn(b, c, seed=2) * 0.1
    f32 = ops.filtered_act(x, sc, sh)
    f16 = ops.filtered_act(x, sc, sh, out_half=True)
    assert f16.dtype == torch.float16
    assert torch.equal(f16, f32.half())


def test_affine_act_plain_f16_store():
    x = nhwc(randn(2, 128, 64, 64, seed=3))
    sc, sh = randn(2, 128, seed=1) * 0.1 + 1, randn(2, 128, seed=2) * 0.1
    f32 = ops.affine_act(x, sc, sh, act="silu")
    f16 = ops.affine_act(x, sc, sh, act="silu", out_half=True)
    assert f16.dtype == torch.float16 and torch.equal(f16, f32.half())


@pytest.mark.parametrize("hw,c", [(1024, 192), (64, 384), (4, 768)])
def test_groupnorm_act_f16_store(hw, c):
    b, side = 4, int(hw ** 0.5)
    x = nhwc(randn(b, 64, side, side, seed=hw))
    y = ops.conv2d(x, ops.pack_conv_weight(randn(c, 64, 1, 1, seed=c) * 0.2), None, 1, algo="tf32", gn_stats=True)
    if getattr(y, "_afldm_gn", None) is None:
        pytest.skip("no GroupNorm partial sums for this shape")
    gamma, beta = randn(c, seed=5) * 0.2 + 1, randn(c, seed=6) * 0.2
    f32 = ops.groupnorm_act(y, 32, 1e-5, gamma, beta, act="identity")
    f16 = ops.groupnorm_act(y, 32, 1e-5, gamma, beta, act="identity", out_half=True)
    assert f16.dtype == torch.float16
    assert torch.equal(f16, f32.half())


def test_attention_f16_output_and_f16_projection():
    """q | k | v projection with fp16 operands and fp16 result, attention with an fp16 result: each equals the fp32-result
    form rounded to nearest fp16 (attention) / agrees with the TF32-operand projection within the class bound."""
    b, n, c, heads = 16, 16, 384, 16
    xn = nhwc(randn(b, c, n, n, seed=1))
    w = randn(3 * c, c, seed=2) * (1.0 / c ** 0.5)
    bias = randn(3 * c, seed=3) * 0.1
    qkv32 = ops.conv2d_f16out(xn, w.contiguous(), bias, 1)
    qkv16 = ops.conv2d_f16out(xn.half(), w.half().contiguous(), bias, 1)
    assert qkv32 is not None and qkv16 is not None and qkv16.dtype == torch.float16
    want = (xn.half().double() @ w.half().double().t() + bias.double()).float()
    torch.testing.assert_close(qkv16.float(), want, rtol=2e-3, atol=2e-3)       # fp16 result rounding + fp32 sums
    err = (qkv16.float() - qkv32.float()).abs()
    assert err.max().item() < 1.5e-2 and err.mean().item() < 1.5e-3
    qkv = qkv16.view(b, n * n, 3 * c)
    q, k, v = qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:]
    o32 = ops.attention_f16(q, k, v, heads)
    o16 = ops.attention_f16(q, k, v, heads, out_half=True)
    assert o16.dtype == torch.float16
    assert torch.equal(o16, o32.half())


def test_attention_block_f16_chain_stays_in_class():
    from afldm_b200.models.blocks import Attention
    torch.manual_seed(0)
    att = Attention(384, 16, 24).to(DEV).eval()
    x = nhwc(randn(16, 64, 16, 16, seed=1))
    pre = ops.conv2d(x, ops.pack_conv_weight(randn(384, 64, 1, 1, seed=3) * 0.2), None, 1, algo="tf32", gn_stats=True)
    rec = []
    with torch.no_grad():
        ops.record_to(rec)
        try:
            got = att(ops.nchw_view(pre))
        finally:
            ops.record_to(None)
        ops.set_default_conv_algo("simt")
        want = att(ops.nchw_view(pre.clone()))
    names = [r[0] for r in rec]
    if ops.F16_CONV and ops.F16_ATTENTION:
        assert names == ["affine_act", "conv2d_f16", "attention_f16", "conv2d_f16"], names
    err = (got - want).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 2e-3, (err.max().item(), err.mean().item())
