"""GPU: the reference-named module surface (afldm.af_libs.ideal_lpf, afldm.shift_utils, afldm.af_libs...upfirdn2d) on the
sm_100a kernels against the UNMODIFIED reference's golden outputs, the explicit tcgen05 filtered-activation entry point,
and the base (aliasing) resamplers against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "afldm", "_compat")):
    if p not in sys.path:
        sys.path.append(p)

from afldm.af_libs import ideal_lpf as IL                          # noqa: E402  (the reference's module path)
from afldm.af_libs.torch_utils.ops import upfirdn2d as U           # noqa: E402
from afldm.shift_utils import flow_utils as FU                     # noqa: E402
from afldm.shift_utils.metrics import mask_psnr                    # noqa: E402
from afldm.shift_utils.shifters import ImageShifter, fourier_shift_batch   # noqa: E402
from afldm_b200 import _lib, ops                                   # noqa: E402
from oracle import ideal_lpf as OL                                 # noqa: E402
from oracle import nn as ON                                        # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True, scope="module")
def _exact_torch_reference():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    prev = ops.default_conv_algo()
    ops.set_default_conv_algo("simt")
    yield
    ops.set_default_conv_algo(prev)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_ideal_lpf_module_surface_vs_reference_goldens(golden):
    g = golden("ideal_ops")
    x12, x16 = dev(g["surf_x12"]), dev(g["surf_x16"])
    cases = [(IL.LPF_RFFT(0.5), x12, "surf_lpf12_c50"), (IL.LPF_RFFT(0.25), x12, "surf_lpf12_c25"),
             (IL.LPF_RECON_RFFT(0.5), x12, "surf_recon12_c50"), (IL.UpsampleRFFT(3), x12, "surf_up3_12"),
             (IL.UpsampleRFFT(2), x12, "surf_up2_12"), (IL.LPF_RFFT(), x16, "surf_lpf16_c50"),
             (IL.UpsampleRFFT(4), x16, "surf_up4_16")]
    for mod, x, key in cases:
        got = mod(x).contiguous().cpu().numpy()
        np.testing.assert_allclose(got, g[key], atol=1e-5, err_msg=key)
    got = IL.subpixel_shift(x16, up=4, shift_x=3, shift_y=-2).cpu().numpy()
    np.testing.assert_allclose(got, g["surf_subpix4_16"], atol=1e-5)
    # the two hot configurations through the same classes (fused kernels)
    for name in ("s8", "s32", "s64"):
        x = dev(g[f"{name}_x"])
        np.testing.assert_allclose(IL.UpsampleRFFT()(x).contiguous().cpu().numpy(), g[f"{name}_up2"], atol=1e-5)
        full = IL.LPF_RFFT(0.5)(x)
        np.testing.assert_allclose(full[:, :, ::2, ::2].contiguous().cpu().numpy(), g[f"{name}_lpf_down2"], atol=1e-5)
    with pytest.raises(_lib.AfldmError):
        IL.LPF_RFFT()(torch.zeros(1, 1, 8, 6, device=DEV))       # the reference's mask is square (ideal_lpf.py:81)


def test_image_shifter_modes_vs_reference_goldens(golden):
    g = golden("shift")
    img, lat = dev(g["img"]), dev(g["lat"])
    bil = ImageShifter()                                           # default = bilinear flow warp
    for k in range(4):
        ti, tj = (float(v) for v in g[f"bil{k}_t"])
        w, m = bil.shift(img, ti, tj)
        np.testing.assert_allclose(w.cpu().numpy(), g[f"bil{k}_img"], atol=1e-5)
        assert m.shape == (2, 1, 24, 24) and np.array_equal(m.cpu().numpy(), g[f"bil{k}_mask"])
    w, m = ImageShifter("ideal", 8).shift(lat, 0.5, -1.375)
    np.testing.assert_allclose(w.cpu().numpy(), g["ideal_img"], atol=2e-5)
    np.testing.assert_allclose(fourier_shift_batch(img, 0.75, -2.3, DEV).cpu().numpy(), g["fourier_img"], atol=2e-5)
    w, m = ImageShifter("fourier_crop").shift(img, 1.5, 0.25)
    np.testing.assert_allclose(w.cpu().numpy(), g["fcrop_img"], atol=2e-5)
    assert np.array_equal(m.cpu().numpy(), g["fcrop_mask"])
    # flow_warp with the uniform flow ImageShifter builds (shifters.py:200-205)
    flow = torch.tensor([-1.75, 3.25]).view(1, 2, 1, 1).repeat(2, 1, 24, 24).to(DEV)
    w2, m2 = FU.flow_warp(img, flow, True)
    np.testing.assert_allclose(w2.cpu().numpy(), g["bil2_img"], atol=1e-5)
    assert np.array_equal(m2.unsqueeze(1).float().cpu().numpy(), g["bil2_mask"])
    # metric on device tensors = the reference's figure
    a, b, mk = dev(g["m_a"]), dev(g["m_b"]), dev(g["m_mask"])
    assert float(mask_psnr(a, b, mk)) == pytest.approx(float(g["mask_psnr"]), rel=1e-5)


def test_upfirdn2d_python_surface_vs_reference_goldens(golden):
    g = golden("upfirdn2d")
    x = dev(g["x"])
    f = U.setup_filter([1, 3, 3, 1], device=DEV)
    np.testing.assert_allclose(f.cpu().numpy(), g["f1331"], atol=0)
    np.testing.assert_allclose(U.upsample2d(x, f, up=2).cpu().numpy(), g["up2"], atol=2e-6)
    np.testing.assert_allclose(U.downsample2d(x, f, down=2).cpu().numpy(), g["down2"], atol=2e-6)
    np.testing.assert_allclose(U.filter2d(x, f).cpu().numpy(), g["filter2d"], atol=2e-6)
    f12 = U.setup_filter([1, 2, 4, 7, 9, 11, 11, 9, 7, 4, 2, 1], device=DEV)
    assert f12.ndim == 1
    np.testing.assert_allclose(U.upsample2d(x, f12, up=2).cpu().numpy(), g["up2_f12"], atol=2e-6)
    xs, fa = dev(g["xs"]), dev(g["fa"])
    got = U.upfirdn2d(xs, fa, up=3, down=2, padding=[2, 1, 0, 3], flip_filter=True, gain=1.7)
    np.testing.assert_allclose(got.cpu().numpy(), g["gen"], atol=2e-6)


@pytest.mark.parametrize("b,c,n", [(2, 24, 12), (1, 8, 20), (2, 5, 6), (1, 3, 48), (1, 4, 256), (2, 24, 8), (1, 7, 32)])
def test_general_plane_sizes(b, c, n):
    """Planes outside the fused kernels' family (not a power of two, above 128, channel counts that are no multiple of
    the kernels' channel groups) run the operator-matrix form (ops._general_resample): same results as the oracle's
    FFT form, which accepts any square size (ideal_lpf.py:69-93, :148-158; af_blocks.py:19-28)."""
    gen = torch.Generator().manual_seed(7 * n + c)
    x = (torch.randn(b, c, n, n, generator=gen) * 1.5).to(DEV)
    sc = (torch.rand(b, c, generator=gen) + 0.5).to(DEV).contiguous()
    sh = (torch.randn(b, c, generator=gen) * 0.3).to(DEV).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    tol = 5e-5 if n >= 256 else 1e-5
    want = OL.filtered_act_fft(x * sc[:, :, None, None] + sh[:, :, None, None])
    got = ops.filtered_act(xn, sc, sh)
    assert got.shape == xn.shape and (got.permute(0, 3, 1, 2) - want).abs().max().item() < tol
    up = ops.up2_ideal(xn)
    assert (up.permute(0, 3, 1, 2) - OL.upsample_rfft(x)).abs().max().item() < tol
    dn = ops.lpf_down2(up, gn_stats=True)
    assert (dn.permute(0, 3, 1, 2) - OL.lpf_rfft(OL.upsample_rfft(x))[..., ::2, ::2]).abs().max().item() < tol
    # module level: the same call a diffusers block makes
    from afldm_b200.af_modules.af_blocks import WarpedNonlinearity
    y = WarpedNonlinearity(torch.nn.SiLU())(x)
    assert (y - OL.filtered_act_fft(x)).abs().max().item() < tol


@pytest.mark.parametrize("b,c,n", [(2, 8, 32), (3, 192, 32), (16, 576, 32), (2, 16, 16), (5, 384, 16)])
def test_filtered_act_tcgen05_entry_point(b, c, n):
    """afldm_filtered_act_tc: the tcgen05 / TMEM form (csrc/fact_tc.cu) against the oracle's FFT form - the same bounds
    as the mma.sync kernel (fp32 accuracy from the 3-term fp16 split), fp32 and fp16 stores, with and without affine."""
    L = _lib.lib()
    gen = torch.Generator().manual_seed(100 * n + c)
    x = (torch.randn(b, c, n, n, generator=gen) * 1.5).to(DEV)
    sc = (torch.rand(b, c, generator=gen) + 0.5).to(DEV).contiguous()
    sh = (torch.randn(b, c, generator=gen) * 0.3).to(DEV).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    want = OL.filtered_act_fft(x * sc[:, :, None, None] + sh[:, :, None, None])
    y = torch.empty_like(xn)
    _lib.check(L.afldm_filtered_act_tc(xn.data_ptr(), y.data_ptr(), 0, b, n, n, c, 1, sc.data_ptr(), sh.data_ptr(), st))
    assert (y.permute(0, 3, 1, 2) - want).abs().max().item() < 1e-5
    yh = torch.empty(xn.shape, dtype=torch.float16, device=DEV)
    _lib.check(L.afldm_filtered_act_tc(xn.data_ptr(), yh.data_ptr(), 1, b, n, n, c, 1, sc.data_ptr(), sh.data_ptr(), st))
    assert torch.equal(yh, y.to(torch.float16))                    # the fp16 store is the rounding of the fp32 result
    yi = torch.empty_like(xn)
    _lib.check(L.afldm_filtered_act_tc(xn.data_ptr(), yi.data_ptr(), 0, b, n, n, c, 0, None, None, st))
    assert (yi.permute(0, 3, 1, 2) - OL.filtered_act_fft(x, act=lambda t: t)).abs().max().item() < 1e-5
    # and against the default kernel of the library on the same input
    ref = ops.filtered_act(xn, sc, sh)
    assert (ref - y).abs().max().item() < 1e-5
    assert L.afldm_filtered_act_tc(xn.data_ptr(), y.data_ptr(), 0, b, 8, 8, c, 1, None, None, st) == -3   # outside the family


def test_base_resamplers_match_oracle():
    """The aliasing originals (diffusers Downsample2D / Upsample2D) that the surgery replaces."""
    from afldm_b200.models import blocks as B
    torch.manual_seed(3)
    for pad in (1, 0):
        ref = ON.Downsample2D(64, True, 64, padding=pad).to(DEV)
        mine = B.Downsample2D(64, True, 64, padding=pad).to(DEV)
        mine.load_state_dict(ref.state_dict())
        x = torch.randn(2, 64, 16, 16, device=DEV)
        with torch.no_grad():
            torch.testing.assert_close(mine(x).contiguous(), ref(x), rtol=0, atol=5e-5)
    ref = ON.Upsample2D(64, True, 64).to(DEV)
    mine = B.Upsample2D(64, True, 64).to(DEV)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(2, 64, 8, 8, device=DEV)
    with torch.no_grad():
        torch.testing.assert_close(mine(x).contiguous(), ref(x), rtol=0, atol=5e-5)


@pytest.mark.parametrize("n,c", [(8, 32), (16, 64), (32, 32)])
def test_backward_of_resamplers_and_filtered_activation(n, c):
    """gradients of WarpedNonlinearity / UpsampleRFFT(2) / LPF + decimate through the sm_100a kernels against
    PyTorch autograd through the oracle's FFT form (what the reference differentiates, ldm_trainer.py:240-272)."""
    from afldm.af_modules.af_blocks import WarpedNonlinearity
    from afldm.af_libs.ideal_lpf import UpsampleRFFT
    from afldm_b200.af_libs.ideal_lpf import LPFDown2
    gen = torch.Generator().manual_seed(n)
    x0 = torch.randn(2, c, n, n, generator=gen).to(DEV)
    cases = [(WarpedNonlinearity(torch.nn.SiLU()), OL.filtered_act_fft, (2, c, n, n)),
             (UpsampleRFFT(2), lambda t: OL.upsample_rfft(t, 2), (2, c, 2 * n, 2 * n)),
             (LPFDown2(), lambda t: OL.lpf_rfft(t, 0.5)[:, :, ::2, ::2], None)]
    for mod, ref, oshape in cases:
        xin = x0 if oshape is not None else torch.randn(2, c, 2 * n, 2 * n, generator=gen).to(DEV)
        w = torch.randn(ref(xin).shape, generator=gen).to(DEV)
        xa = xin.clone().requires_grad_(True)
        ya = mod(xa)
        (ya * w).sum().backward()
        xb = xin.clone().requires_grad_(True)
        yb = ref(xb)
        (yb * w).sum().backward()
        torch.testing.assert_close(ya.detach().contiguous(), yb.detach(), rtol=0, atol=2e-5)
        torch.testing.assert_close(xa.grad, xb.grad, rtol=0, atol=5e-5)


def test_superresolution_x4_degradation(golden):
    """``afldm.af_libs.superresolution.build_sr4x`` (the degradation of scripts/shift_ldm_sr.py:45,102) and the
    ``H / Ht / H_pinv`` methods of ``build_sr_bicubic`` against the unmodified reference's outputs."""
    from afldm.af_libs.superresolution import build_sr4x, build_sr_bicubic, build_sr_pool
    from oracle import superres as OSR
    g = golden("superres")
    for n in (32, 64):
        x = dev(g[f"x_{n}"])
        for filt in ("bicubic", "pool"):
            got = build_sr4x(DEV, filt, n)(x)
            assert got.shape == x.shape
            np.testing.assert_allclose(got.cpu().numpy(), g[f"sr4x_{filt}_{n}"], atol=3e-6)
        hb = build_sr_bicubic(4, DEV, n)
        y = hb.H(x)
        np.testing.assert_allclose(y.cpu().numpy(), g[f"H_{n}"], atol=3e-6)
        np.testing.assert_allclose(hb.Ht(y).cpu().numpy(), g[f"Ht_{n}"], atol=3e-6)
        np.testing.assert_allclose(hb.H_pinv(y).cpu().numpy(), g[f"Hpinv_{n}"], atol=2e-4)
        assert build_sr_pool(4, DEV, n).H(x).shape == (2, 3 * (n // 4) ** 2)
    x = (torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(296)) * 2 - 1)
    y = build_sr4x(DEV, "bicubic", 256)(x.to(DEV))
    np.testing.assert_allclose(y[0, :, ::37, ::41].cpu().numpy(), g["rows_256"], atol=3e-6)
    np.testing.assert_allclose(y.cpu().numpy(), OSR.sr4x(x.numpy()), atol=3e-6)
    assert build_sr4x(DEV, "bicubic", 256)(x[0].to(DEV)).shape == (3, 256, 256)          # 3-D input (:300-303)
    assert build_sr4x(DEV, "bicubic", 256)(x).is_cuda                                    # host input is moved, as :305 does


def test_lanczos_shifter_vs_reference_goldens(golden):
    """``ImageShifter('lanczos')`` on the device against the unmodified reference (shifters.py:159-162)."""
    from afldm.shift_utils.shifters import ImageShifter
    g = golden("shift")
    img = dev(g["img"])
    sh = ImageShifter("lanczos")
    for k in range(5):
        ti, tj = (float(v) for v in g[f"lz{k}_t"])
        w, m = sh.shift(img, ti, tj)
        np.testing.assert_allclose(w.cpu().numpy(), g[f"lz{k}_img"], atol=3e-6)
        assert m.shape == (2, 1, 24, 24) and np.array_equal(m.cpu().numpy(), g[f"lz{k}_mask"])
    ws, ms = sh.shift_batch(img, [tuple(float(v) for v in g[f"lz{k}_t"]) for k in range(5)])
    np.testing.assert_allclose(ws[2].cpu().numpy(), g["lz2_img"], atol=3e-6)
