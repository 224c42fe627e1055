"""GPU parity: every C-ABI op (called through afldm_b200.ops -> ctypes -> libafldm_b200.so) against
the oracle on the same seeded inputs, and against the golden vectors produced by the unmodified
reference.  Tolerances are absolute on O(1) data and stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from afldm_b200 import _lib, ops
from oracle import ideal_lpf as OL
from oracle import upfirdn2d as OU

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _exact_torch_reference():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def to_nhwc(x):   # logical NCHW torch tensor -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw(y):
    return y.permute(0, 3, 1, 2).contiguous()


def test_library_is_loaded_and_counts_launches():
    n0 = _lib.launch_count()
    ops.axpby(randn(64), randn(64, seed=1), 0.5, 2.0)
    assert _lib.launch_count() == n0 + 1


# ----------------------------------------------------------------------------- ideal resamplers
GOLDEN_SIZES = ["s2", "s4", "s8", "s16", "s32"]


@pytest.mark.parametrize("name", GOLDEN_SIZES)
def test_resamplers_match_reference_golden(golden, name):
    """Reference outputs (ideal_lpf.py run unmodified) replicated over 32 channels (C % 32 == 0)."""
    g = golden("ideal_ops")
    x = torch.from_numpy(g[f"{name}_x"])                       # [B, c, n, n], c in {3,4,5}
    b, c, n, _ = x.shape
    reps = -(-32 // c)
    xs = x.repeat(1, reps, 1, 1)[:, :32].to(DEV)
    pick = lambda arr: torch.from_numpy(arr).repeat(1, reps, 1, 1)[:, :32].to(DEV)
    got = to_nchw(ops.up2_ideal(to_nhwc(xs)))
    torch.testing.assert_close(got, pick(g[f"{name}_up2"]), rtol=0, atol=5e-6)
    got = to_nchw(ops.filtered_act(to_nhwc(xs)))
    torch.testing.assert_close(got, pick(g[f"{name}_filtered_silu"]), rtol=0, atol=5e-6)
    if n >= 4:
        got = to_nchw(ops.lpf_down2(to_nhwc(xs)))
        torch.testing.assert_close(got, pick(g[f"{name}_lpf_down2"]), rtol=0, atol=5e-6)


@pytest.mark.parametrize("n,c,b", [(2, 768, 16), (4, 768, 3), (8, 384, 2), (16, 384, 2), (32, 192, 2), (32, 576, 16)])
def test_filtered_act_vs_oracle_fft(n, c, b):
    x = randn(b, c, n, n, seed=n + c)
    want = OL.filtered_act_fft(x)
    got = to_nchw(ops.filtered_act(to_nhwc(x)))
    torch.testing.assert_close(got, want, rtol=0, atol=1e-5)


def test_filtered_act_fused_groupnorm_affine_and_inplace():
    b, c, n = 3, 192, 16
    x = randn(b, c, n, n, seed=5) * 2 + 0.5
    gn = torch.nn.GroupNorm(32, c, eps=1e-5).to(DEV)
    with torch.no_grad():
        gn.weight.copy_(randn(c, seed=6) * 0.2 + 1)
        gn.bias.copy_(randn(c, seed=7) * 0.2)
        want = OL.filtered_act_fft(gn(x))
    xn = to_nhwc(x)
    scale, shift = ops.groupnorm_affine(xn, 32, 1e-5, gn.weight, gn.bias)
    got = ops.filtered_act(xn, scale, shift)
    torch.testing.assert_close(to_nchw(got), want, rtol=0, atol=2e-5)
    ops.filtered_act(xn, scale, shift, out=xn)                  # x may alias y
    torch.testing.assert_close(xn, got, rtol=0, atol=0)


def test_filtered_act_identity_is_nyquist_projector():
    """D U removes only the Nyquist bins (SURVEY.md 8(a) identity 4): idempotent, and exact on
    band-limited input."""
    x = to_nhwc(randn(2, 64, 16, 16, seed=9))
    p1 = ops.filtered_act(x, act="identity")
    p2 = ops.filtered_act(p1, act="identity")
    torch.testing.assert_close(p2, p1, rtol=0, atol=2e-6)
    assert (p1 - x).abs().max() > 1e-2


@pytest.mark.parametrize("n,c", [(2, 768), (4, 768), (8, 384), (16, 384), (32, 64)])
def test_up2_and_lpf_down2_vs_oracle(n, c):
    x = randn(2, c, n, n, seed=n)
    up = to_nchw(ops.up2_ideal(to_nhwc(x)))
    torch.testing.assert_close(up, OL.upsample_rfft(x), rtol=0, atol=1e-5)
    torch.testing.assert_close(up[:, :, ::2, ::2], x, rtol=0, atol=0)          # identity 1: even phase is a copy
    y = randn(2, c, 2 * n, 2 * n, seed=n + 1)
    dn = to_nchw(ops.lpf_down2(to_nhwc(y)))
    torch.testing.assert_close(dn, OL.lpf_rfft(y)[:, :, ::2, ::2], rtol=0, atol=1e-5)


@pytest.mark.parametrize("n,c,b", [(16, 192, 16), (8, 384, 3), (4, 384, 2), (2, 768, 16)])
def test_lpf_down2_emits_groupnorm_partials(n, c, b):
    """The down-sampler's optional GroupNorm partial sums (one slot per output plane): same y, and the finalised
    scale / shift equal the statistics pass over y."""
    x = randn(b, 2 * n, 2 * n, c, seed=n)
    y0 = ops.lpf_down2(x)
    y1 = ops.lpf_down2(x, gn_stats=True)
    assert torch.equal(y0, y1) and hasattr(y1, "_afldm_gn")
    gamma, beta = randn(c, seed=5) * 0.2 + 1, randn(c, seed=6) * 0.2
    rec = []
    ops.record_to(rec)
    s1, t1 = ops.groupnorm_affine(y1, 32, 1e-5, gamma, beta)
    ops.record_to(None)
    assert [r[0] for r in rec] == ["groupnorm_finalize"]
    s0, t0 = ops.groupnorm_affine(y0, 32, 1e-5, gamma, beta)
    torch.testing.assert_close(s1, s0, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(t1, t0, rtol=0, atol=2e-5)


@pytest.mark.parametrize("n,c,b", [(64, 64, 2), (128, 32, 1), (64, 32, 3)])
def test_large_plane_resamplers_vs_oracle(n, c, b):
    """n = 64 / 128 (VAE decoder, 64x64-latent UNets): three line passes through a workspace."""
    x = randn(b, c, n, n, seed=n + c)
    torch.testing.assert_close(to_nchw(ops.filtered_act(to_nhwc(x))), OL.filtered_act_fft(x), rtol=0, atol=1e-5)
    up = to_nchw(ops.up2_ideal(to_nhwc(x)))
    torch.testing.assert_close(up, OL.upsample_rfft(x), rtol=0, atol=1e-5)
    torch.testing.assert_close(up[:, :, ::2, ::2], x, rtol=0, atol=0)
    y = randn(b, c, 2 * n, 2 * n, seed=n + 1)
    torch.testing.assert_close(to_nchw(ops.lpf_down2(to_nhwc(y))), OL.lpf_rfft(y)[:, :, ::2, ::2], rtol=0, atol=1e-5)
    # fused GroupNorm affine on the large path
    gamma, beta = randn(c, seed=1) * 0.2 + 1, randn(c, seed=2) * 0.2
    xn = to_nhwc(x)
    scale, shift = ops.groupnorm_affine(xn, 32, 1e-6, gamma, beta)
    want = OL.filtered_act_fft(F.group_norm(x, 32, gamma, beta, 1e-6))
    torch.testing.assert_close(to_nchw(ops.filtered_act(xn, scale, shift)), want, rtol=0, atol=2e-5)


def test_s64_golden_large_plane(golden):
    g = golden("ideal_ops")
    x = torch.from_numpy(g["s64_x"])                           # [1, 3, 64, 64]
    xs = x.repeat(1, 11, 1, 1)[:, :32].to(DEV)
    pick = lambda arr: torch.from_numpy(arr).repeat(1, 11, 1, 1)[:, :32].to(DEV)
    torch.testing.assert_close(to_nchw(ops.up2_ideal(to_nhwc(xs))), pick(g["s64_up2"]), rtol=0, atol=5e-6)
    torch.testing.assert_close(to_nchw(ops.filtered_act(to_nhwc(xs))), pick(g["s64_filtered_silu"]), rtol=0, atol=5e-6)
    torch.testing.assert_close(to_nchw(ops.lpf_down2(to_nhwc(xs))), pick(g["s64_lpf_down2"]), rtol=0, atol=5e-6)


def test_resampler_linearity_at_full_size():
    """Size-independent property at the BASELINE shape (B=16, 576 ch, 32x32)."""
    a, b = to_nhwc(randn(16, 576, 32, 32, seed=1)), to_nhwc(randn(16, 576, 32, 32, seed=2))
    lhs = ops.up2_ideal(a * 0.5 + b * 2.0)
    rhs = ops.up2_ideal(a) * 0.5 + ops.up2_ideal(b) * 2.0
    torch.testing.assert_close(lhs, rhs, rtol=0, atol=2e-5)
    # down2(up2(x)) == x minus its Nyquist component == filtered identity activation
    torch.testing.assert_close(ops.lpf_down2(ops.up2_ideal(a)), ops.filtered_act(a, act="identity"), rtol=0, atol=1e-5)


def test_resampler_rejects_unsupported_shapes():
    with pytest.raises(_lib.AfldmError):
        ops.filtered_act(torch.zeros(1, 8, 4, 32, device=DEV))          # not square (neither does the reference: its
    with pytest.raises(_lib.AfldmError):                                # mask is built from the last dimension only)
        ops.up2_ideal(torch.zeros(1, 8, 4, 32, device=DEV))
    # the C ABI itself reports shapes outside the fused kernels' family (nothing launched) ...
    L = _lib.lib()
    x6, x24 = torch.zeros(1, 6, 6, 32, device=DEV), torch.zeros(1, 8, 8, 24, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert L.afldm_filtered_act_f32(x6.data_ptr(), torch.empty_like(x6).data_ptr(), 1, 6, 6, 32, 1, None, None, None, 0, st) == -3
    assert L.afldm_filtered_act_f32(x24.data_ptr(), torch.empty_like(x24).data_ptr(), 1, 8, 8, 24, 1, None, None, None, 0, st) == -1
    # ... and the Python op then runs the general-plane form (tests/test_gpu_surface.py::test_general_plane_sizes)
    assert ops.filtered_act(x6).shape == x6.shape and ops.filtered_act(x24).shape == x24.shape


# ----------------------------------------------------------------------------- norm / act
@pytest.mark.parametrize("b,c,hw", [(16, 192, 1024), (2, 1536, 4), (3, 576, 256), (1, 128, 40)])
def test_groupnorm_affine(b, c, hw):
    x = randn(b, hw, c, seed=c) * 3 + 1.5
    gamma, beta = randn(c, seed=1) * 0.3 + 1, randn(c, seed=2) * 0.3
    scale, shift = ops.groupnorm_affine(x.view(b, hw, 1, c), 32, 1e-5, gamma, beta)
    got = x * scale[:, None, :] + shift[:, None, :]
    want = F.group_norm(x.transpose(1, 2), 32, gamma, beta, 1e-5).transpose(1, 2)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)
    act = ops.affine_act(x.view(b, hw, 1, c), scale, shift, "silu").view(b, hw, c)
    torch.testing.assert_close(act, F.silu(want), rtol=0, atol=2e-5)
    torch.testing.assert_close(ops.affine_act(x, None, None, "silu"), F.silu(x), rtol=0, atol=2e-6)


# ----------------------------------------------------------------------------- conv
CONV_CASES = [  # B, H, W, Cin, Cout, k
    (2, 32, 32, 192, 192, 3), (16, 2, 2, 1536, 768, 3), (3, 4, 4, 768, 768, 3), (2, 8, 8, 384, 384, 3),
    (2, 32, 32, 4, 192, 3), (2, 32, 32, 192, 4, 3), (2, 16, 16, 576, 384, 1), (1, 7, 5, 20, 33, 3),
    (2, 32, 32, 192, 576, 1), (1, 64, 64, 128, 3, 3), (5, 1, 1, 64, 64, 1),
]


@pytest.mark.parametrize("b,h,w,cin,cout,k", CONV_CASES)
def test_conv2d_simt_vs_torch(b, h, w, cin, cout, k):
    x = randn(b, cin, h, w, seed=cin + h)
    wt = randn(cout, cin, k, k, seed=cout) * (1.0 / (cin * k * k) ** 0.5)
    bias = randn(cout, seed=3)
    row = randn(b, cout, seed=4)
    res = randn(b, cout, h, w, seed=5)
    # fp64 reference: both cuDNN-fp32 and this kernel round differently over K = k*k*Cin terms
    conv64 = F.conv2d(x.double(), wt.double(), None, padding=k // 2)
    want = (conv64 + bias.double()[None, :, None, None] + row.double()[:, :, None, None] + res.double()).float()
    got = ops.conv2d(to_nhwc(x), ops.pack_conv_weight(wt), bias, k, row_add=row, residual=to_nhwc(res), algo="simt")
    torch.testing.assert_close(to_nchw(got), want, rtol=0, atol=2e-5)
    plain = ops.conv2d(to_nhwc(x), ops.pack_conv_weight(wt), None, k, algo="simt")
    torch.testing.assert_close(to_nchw(plain), conv64.float(), rtol=0, atol=2e-5)


def test_conv2d_pitched_views_and_inplace_residual():
    """Input / output / residual as channel slices of wider NHWC buffers; residual aliasing out."""
    b, h, w, cin, cout = 2, 8, 8, 64, 96
    big_in = to_nhwc(randn(b, cin + 32, h, w, seed=1))
    big_out = torch.zeros(b, h, w, cout + 64, device=DEV)
    wt = randn(cout, cin, 3, 3, seed=2) * 0.05
    res0 = randn(b, h, w, cout, seed=3)
    big_out[..., 64:] = res0
    xin = big_in[..., 32:]
    out = big_out[..., 64:]
    ops.conv2d(xin, ops.pack_conv_weight(wt), None, 3, residual=out, out=out, algo="simt")
    want = F.conv2d(to_nchw(xin), wt, None, padding=1) + to_nchw(res0)
    torch.testing.assert_close(to_nchw(big_out[..., 64:]), want, rtol=0, atol=2e-5)
    assert big_out[..., :64].abs().max() == 0


def test_conv2d_split_k_is_deterministic():
    x = to_nhwc(randn(16, 1536, 2, 2, seed=1))
    wp = ops.pack_conv_weight(randn(768, 1536, 3, 3, seed=2) * 0.01)
    a = ops.conv2d(x, wp, None, 3, algo="simt")
    b = ops.conv2d(x, wp, None, 3, algo="simt")
    assert torch.equal(a, b)


# ----------------------------------------------------------------------------- attention
@pytest.mark.parametrize("b,bkv,nq,nk,heads,d", [(4, 4, 1024, 1024, 8, 24), (2, 2, 256, 256, 16, 24), (16, 16, 4, 4, 32, 24),
                                                  (6, 2, 64, 64, 16, 24), (2, 1, 100, 77, 4, 40), (2, 2, 16, 16, 2, 64)])
def test_attention_vs_sdpa(b, bkv, nq, nk, heads, d):
    c = heads * d
    q, k, v = randn(b, nq, c, seed=1), randn(bkv, nk, c, seed=2), randn(bkv, nk, c, seed=3)
    got = ops.attention(q, k, v, heads, algo="simt")
    rep = b // bkv
    kk = k.unsqueeze(1).repeat(1, rep, 1, 1).reshape(b, nk, c)          # cross_frame_attn.py:91-97
    vv = v.unsqueeze(1).repeat(1, rep, 1, 1).reshape(b, nk, c)
    sp = lambda t, n: t.view(b, n, heads, d).transpose(1, 2)
    want = F.scaled_dot_product_attention(sp(q, nq), sp(kk, nk), sp(vv, nk)).transpose(1, 2).reshape(b, nq, c)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)
    if d % 8 == 0:
        # tensor-core variant: TF32 products (10-bit mantissa), fp32 softmax / accumulation
        tc = ops.attention(q, k, v, heads, algo="tf32")
        err = (tc - want).abs()
        assert err.max().item() < 6e-3 and err.mean().item() < 6e-4, (err.max().item(), err.mean().item())


def test_attention_tf32_exact_on_tf32_representable_inputs():
    """With inputs exactly representable in TF32 and a one-hot softmax the tensor-core kernel must
    return the selected V rows exactly: checks fragment layouts / key permutation, not rounding."""
    b, n, heads, d = 2, 192, 3, 24
    c = heads * d
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(0)).to(DEV)
    basis = torch.zeros(n, d, device=DEV)
    basis[torch.arange(n, device=DEV), torch.arange(n, device=DEV) % d] = 1.0
    code = ((torch.arange(n, device=DEV)[:, None] >> torch.arange(d, device=DEV)[None, :]) & 1).float() * 2 - 1  # +-1 codes
    k = (code * 64.0).repeat(1, heads).expand(b, n, c).contiguous()              # key j = 64 * code(j)
    q = (code[idx] * 64.0).repeat(1, heads).expand(b, n, c).contiguous()         # query i matches key idx[i]
    v = ((randn(b, n, c, seed=5) * 16).round() / 16)
    out = ops.attention(q, k, v, heads, algo="tf32")
    torch.testing.assert_close(out, v[:, idx], rtol=0, atol=1e-6)


def test_attention_on_fused_qkv_slices():
    b, n, heads, d = 2, 64, 4, 24
    c = heads * d
    qkv = randn(b, n, 3 * c, seed=4)
    got = ops.attention(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], heads, algo="simt")
    sp = lambda t: t.reshape(b, n, heads, d).transpose(1, 2)
    want = F.scaled_dot_product_attention(sp(qkv[:, :, :c]), sp(qkv[:, :, c:2 * c]), sp(qkv[:, :, 2 * c:]))
    want = want.transpose(1, 2).reshape(b, n, c)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)
    tc = ops.attention(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], heads, algo="tf32")
    assert (tc - want).abs().max().item() < 6e-3


@pytest.mark.parametrize("b,bkv,nq,nk,heads,d", [(4, 4, 1024, 1024, 8, 24), (2, 2, 256, 256, 16, 24), (16, 16, 16, 16, 32, 24),
                                                  (6, 2, 64, 64, 16, 24), (2, 1, 100, 77, 4, 40), (2, 2, 16, 16, 2, 64),
                                                  (1, 1, 200, 130, 3, 32), (2, 2, 8, 4, 2, 8)])
def test_attention_f16_vs_sdpa(b, bkv, nq, nk, heads, d):
    """fp16-operand attention (ldmatrix + mma.m16n8k16): (1) against fp32 SDPA evaluated on the SAME fp16-rounded
    q / k / v the error is P's 11-bit rounding only (<= 1.5e-3 max); (2) against SDPA on the fp32 inputs it is
    in the TF32 class (6e-3 max / 6e-4 mean, the bound of the TF32 kernel)."""
    c = heads * d
    q, k, v = randn(b, nq, c, seed=1), randn(bkv, nk, c, seed=2), randn(bkv, nk, c, seed=3)
    qkv_h = [t.half() for t in (q, k, v)]
    got = ops.attention_f16(*qkv_h, heads)
    rep = b // bkv
    sp = lambda t, n: t.view(b, n, heads, d).transpose(1, 2)
    def sdpa(q_, k_, v_):
        kk = k_.unsqueeze(1).repeat(1, rep, 1, 1).reshape(b, nk, c)
        vv = v_.unsqueeze(1).repeat(1, rep, 1, 1).reshape(b, nk, c)
        return F.scaled_dot_product_attention(sp(q_, nq), sp(kk, nk), sp(vv, nk)).transpose(1, 2).reshape(b, nq, c)
    same = sdpa(*[t.float() for t in qkv_h])
    assert (got - same).abs().max().item() < 1.5e-3
    err = (got - sdpa(q, k, v)).abs()
    assert err.max().item() < 6e-3 and err.mean().item() < 6e-4, (err.max().item(), err.mean().item())


def test_attention_f16_exact_selection_and_qkv_slices():
    """One-hot softmax on +-64 codes: the fp16 kernel must return the selected V rows exactly (fragment layouts of
    ldmatrix / ldmatrix.trans); q, k, v are column slices of one fused fp16 buffer as in the attention block."""
    b, n, heads, d = 2, 192, 3, 24
    c = heads * d
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(0)).to(DEV)
    code = ((torch.arange(n, device=DEV)[:, None] >> torch.arange(d, device=DEV)[None, :]) & 1).float() * 2 - 1
    qkv = torch.empty(b, n, 3 * c, device=DEV, dtype=torch.float16)
    qkv[:, :, :c] = (code[idx] * 8.0).repeat(1, heads)
    qkv[:, :, c:2 * c] = (code * 8.0).repeat(1, heads)
    vv = ((randn(b, n, c, seed=5) * 16).round() / 16)
    qkv[:, :, 2 * c:] = vv
    out = ops.attention_f16(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], heads)
    torch.testing.assert_close(out, vv[:, idx], rtol=0, atol=1e-6)


def test_conv2d_f16out_matches_fp32_epilogue():
    """The fp16-output epilogue of the tcgen05 convolution = fp16 rounding of the fp32 one; split-K shapes decline."""
    x, wt, bias = randn(16, 192, 32, 32, seed=1), randn(576, 192, 1, 1, seed=2) * 0.07, randn(576, seed=3)
    xp, wp = x.permute(0, 2, 3, 1).contiguous(), ops.pack_conv_weight(wt)
    h = ops.conv2d_f16out(xp, wp, bias, 1)
    assert h is not None and h.dtype == torch.float16
    ref = ops.conv2d(xp, wp, bias, 1, algo="tf32")
    assert torch.equal(h, ref.half())
    xs = randn(16, 768, 2, 2, seed=4).permute(0, 2, 3, 1).contiguous()       # 3x3 at the 2x2 level: split-K
    assert ops.conv2d_f16out(xs, ops.pack_conv_weight(randn(768, 768, 3, 3, seed=5) * 0.03), None, 3) is None


def test_attention_large_head_dim_via_gemm():
    b, n, d = 2, 256, 512
    q, k, v = randn(b, n, d, seed=1) * 0.2, randn(b, n, d, seed=2) * 0.2, randn(b, n, d, seed=3)
    got = ops.attention_gemm(q, k, v, 1)
    want = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    torch.testing.assert_close(got, want, rtol=0, atol=3e-5)


# ----------------------------------------------------------------------------- small ops
def test_linear_rows_and_timestep_embedding():
    from oracle.nn import timestep_embedding
    t = torch.tensor([981.0, 501.0, 1.0, 21.0], device=DEV)
    got = ops.timestep_embedding(t, 192)
    want = timestep_embedding(t, 192)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-4)      # fp32 sin/cos of arguments up to ~1e3
    x, w, bias = randn(16, 768, seed=1), randn(1000, 768, seed=2) * 0.03, randn(1000, seed=3)
    torch.testing.assert_close(ops.linear_rows(x, w, bias), F.linear(x, w, bias), rtol=0, atol=2e-5)
    torch.testing.assert_close(ops.linear_rows(x, w, bias, act_in="silu", act_out="silu"),
                               F.silu(F.linear(F.silu(x), w, bias)), rtol=0, atol=2e-5)
    x3 = randn(3, 192, seed=4)
    torch.testing.assert_close(ops.linear_rows(x3, w[:, :192].contiguous(), None), F.linear(x3, w[:, :192]), rtol=0, atol=2e-5)
    # ragged shapes: N not a multiple of the 4 columns a warp owns (vector path), K % 4 != 0 (scalar path), M > 16
    w2, b2 = randn(1001, 768, seed=5) * 0.03, randn(1001, seed=6)
    torch.testing.assert_close(ops.linear_rows(x, w2, b2), F.linear(x, w2, b2), rtol=0, atol=2e-5)
    x5, w5 = randn(5, 190, seed=7), randn(37, 190, seed=8) * 0.05
    torch.testing.assert_close(ops.linear_rows(x5, w5, None), F.linear(x5, w5), rtol=0, atol=2e-5)
    x20 = randn(20, 256, seed=9)
    torch.testing.assert_close(ops.linear_rows(x20, w[:, :256].contiguous(), bias, act_out="silu"),
                               F.silu(F.linear(x20, w[:, :256], bias)), rtol=0, atol=2e-5)


def test_pad_channels_and_tensor_core_conv_in():
    """conv_in on the tensor-core path: latents zero-padded to one 32-channel chunk + zero-padded weights == the
    exact-fp32 4-channel convolution within the TF32 bound."""
    x = randn(16, 32, 32, 4, seed=1)
    xp = ops.pad_channels(x, 32)
    assert xp.shape == (16, 32, 32, 32) and torch.equal(xp[..., :4], x) and xp[..., 4:].abs().max() == 0
    w = randn(192, 4, 3, 3, seed=2) * 0.2
    wp = torch.zeros(192, 32, 3, 3, device=DEV)
    wp[:, :4] = w
    bias = randn(192, seed=3)
    got = ops.conv2d(xp, ops.pack_conv_weight(wp), bias, 3, algo="tf32", gn_stats=True)
    assert hasattr(got, "_afldm_gn")
    want = ops.conv2d(x, ops.pack_conv_weight(w), bias, 3, algo="simt")
    assert (got - want).abs().max().item() < 8e-3


def test_layout_concat_axpby_softmax():
    x = randn(3, 20, 5, 7, seed=1)
    y = ops.nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.to_nchw_contiguous(y), x)
    assert ops.nhwc(ops.nchw_view(y)).data_ptr() == y.data_ptr()              # zero-copy round trip
    a, b = randn(2, 4, 4, 64, seed=2), randn(2, 4, 4, 32, seed=3)
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], dim=-1))
    e = randn(3, 20, 5, 7, seed=4)
    torch.testing.assert_close(ops.axpby(x, e, 1.25, -0.5), 1.25 * x - 0.5 * e, rtol=0, atol=1e-6)
    coef = torch.tensor([0.75, 0.125], device=DEV)
    torch.testing.assert_close(ops.axpby(x, e, coef, None), 0.75 * x + 0.125 * e, rtol=0, atol=1e-6)
    s = randn(37, 100, seed=5)
    torch.testing.assert_close(ops.softmax_rows_(s.clone(), 0.3), torch.softmax(0.3 * s, -1), rtol=0, atol=1e-6)


def test_upfirdn2d_matches_reference_golden(golden):
    """BASELINE config #1 and the reference's general cases (upfirdn2d.py run unmodified)."""
    g = golden("upfirdn2d")
    x, f = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["f1331"]).to(DEV)
    up = ops.upfirdn2d(x, f, up=2, padding=(2, 1, 2, 1), gain=4.0)            # == upsample2d(x, f, up=2)
    assert up.shape == (1, 3, 128, 128)
    np.testing.assert_allclose(up.cpu().numpy(), g["up2"], atol=2e-6)
    dn = ops.upfirdn2d(x, f, down=2, padding=(1, 1, 1, 1))
    np.testing.assert_allclose(dn.cpu().numpy(), g["down2"], atol=2e-6)
    f12 = torch.from_numpy(g["f12"]).to(DEV)
    up12 = ops.upfirdn2d(x, f12, up=2, padding=(6, 5, 6, 5), gain=4.0)
    np.testing.assert_allclose(up12.cpu().numpy(), g["up2_f12"], atol=2e-6)
    xs, fa = torch.from_numpy(g["xs"]).to(DEV), torch.from_numpy(g["fa"]).to(DEV)
    np.testing.assert_allclose(ops.upfirdn2d(xs, fa, 3, 2, (2, 1, 0, 3), True, 1.7).cpu().numpy(), g["gen"], atol=2e-6)
    np.testing.assert_allclose(ops.upfirdn2d(xs, fa, 2, 1, (1, 1, 2, 0), False, 1.0).cpu().numpy(), g["gen_noflip"], atol=2e-6)
    np.testing.assert_allclose(ops.upfirdn2d(xs, fa, 2, 1, (-1, 2, 1, -2)).cpu().numpy(), g["crop"], atol=2e-6)
    # oracle agrees on a fresh random case
    xr = randn(2, 3, 11, 13, seed=8)
    want = OU.upfirdn2d(xr.cpu().numpy(), fa.cpu().numpy(), 2, 3, (3, 2, 1, 4), False, 0.7)
    np.testing.assert_allclose(ops.upfirdn2d(xr, fa, 2, 3, (3, 2, 1, 4), False, 0.7).cpu().numpy(), want, atol=2e-6)


def test_linear_rows_more_than_64_rows():
    """The time-embedding MLP at batch 128 (config #5 on one GPU): rows are processed 64 at a time."""
    x, w, b = randn(130, 192, seed=1), randn(768, 192, seed=2) * 0.1, randn(768, seed=3)
    got = ops.linear_rows(x, w, b, act_out="silu")
    torch.testing.assert_close(got, F.silu(x @ w.t() + b), rtol=0, atol=2e-5)
