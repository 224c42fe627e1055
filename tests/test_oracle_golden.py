"""Pin the oracle against golden vectors produced by the unmodified reference
(oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import ideal_lpf as L
from oracle import shift as S
from oracle import upfirdn2d as U

SIZES = ["s2", "s4", "s6", "s8", "s16", "s32", "s64"]


def test_masks_match_reference(golden):
    g = golden("masks")
    for n in (2, 4, 6, 8, 12, 16, 32, 64, 128):
        assert np.array_equal(L.lpf_rect(n).numpy(), g[f"lpf_{n}"])
        assert np.array_equal(L.recon_rect(n).numpy(), g[f"recon_{n}"])
    assert np.array_equal(L.recon_rect(16, 1 / 8).numpy(), g["recon_16_c8"])
    assert np.array_equal(L.recon_rect(64, 1 / 8).numpy(), g["recon_64_c8"])


def test_mask_known_answers():
    # SURVEY.md 8(c) KATs
    assert L.lpf_rect(8)[:, 0].tolist() == [1, 1, 0, 0, 0, 0, 0, 1]
    assert L.lpf_rect(4)[:, 0].tolist() == [1, 0, 0, 0]
    assert L.lpf_rect(6)[:, 0].tolist() == [1, 1, 0, 0, 0, 1]
    assert L.lpf_rect(16)[:, 0].tolist() == [1, 1, 1, 1] + [0] * 9 + [1, 1, 1]
    assert L.recon_rect(8)[:, 0].tolist() == [1, 1, .5, 0, 0, 0, .5, 1]
    assert L.recon_rect(4)[:, 0].tolist() == [1, .5, 0, .5]
    assert L.recon_rect(16, 1 / 8)[:, 0].tolist() == [1, .5] + [0] * 13 + [.5]


@pytest.mark.parametrize("name", SIZES)
def test_fft_form_matches_reference(golden, name):
    g = golden("ideal_ops")
    x = torch.from_numpy(g[f"{name}_x"])
    np.testing.assert_allclose(L.upsample_rfft(x).numpy(), g[f"{name}_up2"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(L.lpf_rfft(x)[:, :, ::2, ::2].numpy(), g[f"{name}_lpf_down2"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(L.filtered_act_fft(x).numpy(), g[f"{name}_filtered_silu"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", [s for s in SIZES if s != "s6"])
def test_separable_form_matches_reference(golden, name):
    """The operator form the CUDA kernels implement == the reference FFT path (even N)."""
    g = golden("ideal_ops")
    x = g[f"{name}_x"]
    np.testing.assert_allclose(L.up2_sep(x), g[f"{name}_up2"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(L.filtered_act_sep(x), g[f"{name}_filtered_silu"], rtol=0, atol=3e-6)
    # LPF+decimate of an arbitrary (not band-limited) 2N signal
    n = x.shape[-1]
    if n >= 4:
        np.testing.assert_allclose(L.lpf_down2_sep(x), g[f"{name}_lpf_down2"], rtol=0, atol=3e-6)


def test_survey_known_answers(golden):
    g = golden("ideal_ops")
    x = torch.from_numpy(g["s64_x"])
    assert abs(float(x[0, 0, 0, 0]) - (-0.111719)) < 1e-5
    up = L.upsample_rfft(x)
    assert abs(float(up.sum()) - 15.25294) < 2e-2
    assert abs(float(up[0, 0, 1, 1]) - 0.5619204) < 1e-5
    assert abs(float(up[0, 2, 63, 127]) - (-0.2661938)) < 1e-5
    fa = L.filtered_act_fft(x)
    assert abs(float(fa[0, 1, 5, 7]) - 0.3288268) < 1e-5
    ld = L.lpf_rfft(x)[..., ::2, ::2]
    assert abs(float(ld[0, 0, 3, 4]) - (-0.1172619)) < 1e-5


def test_up2_matrix_structure():
    u = L.up2_matrix(4)
    assert np.allclose(u[::2], np.eye(4), atol=1e-14)
    assert np.allclose(sorted(u[1]), sorted([0.603553, 0.603553, -0.103553, -0.103553]), atol=1e-6)
    for n in (2, 4, 8, 16, 32, 64):
        u, d = L.up2_matrix(n), L.down2_matrix(n)
        assert np.allclose(u[::2], np.eye(n), atol=1e-13)           # identity 1: even phase is a copy
        d_taps = L.up2_taps(n)
        for i in range(n):
            for j in range(n):
                assert abs(u[2 * i + 1, j] - d_taps[(i - j) % n]) < 1e-13
        # D U = projector removing the Nyquist bin of the length-n signal
        p = d @ u
        assert np.allclose(p @ p, p, atol=1e-12)


def test_up8_and_subpixel(golden):
    g = golden("ideal_ops")
    x = torch.from_numpy(g["up8_x"])
    np.testing.assert_allclose(L.upsample_rfft(x, 8).numpy(), g["up8_y"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(L.subpixel_shift(x).numpy(), g["subpix_y"], rtol=0, atol=3e-6)


def test_upfirdn2d_matches_reference(golden):
    g = golden("upfirdn2d")
    x, f = g["x"], g["f1331"]
    assert np.allclose(U.setup_filter([1, 3, 3, 1]), f)
    assert np.allclose(U.setup_filter([1, 2, 4, 7, 9, 11, 11, 9, 7, 4, 2, 1]), g["f12"])
    np.testing.assert_allclose(U.upsample2d(x, f, up=2), g["up2"], atol=2e-6)
    np.testing.assert_allclose(U.downsample2d(x, f, down=2), g["down2"], atol=2e-6)
    np.testing.assert_allclose(U.upfirdn2d(x, f, padding=(2, 1, 2, 1)), g["filter2d"], atol=2e-6)
    np.testing.assert_allclose(U.upsample2d(x, g["f12"], up=2), g["up2_f12"], atol=2e-6)
    xs, fa = g["xs"], g["fa"]
    np.testing.assert_allclose(U.upfirdn2d(xs, fa, 3, 2, (2, 1, 0, 3), True, 1.7), g["gen"], atol=2e-6)
    np.testing.assert_allclose(U.upfirdn2d(xs, fa, 2, 1, (1, 1, 2, 0), False, 1.0), g["gen_noflip"], atol=2e-6)
    np.testing.assert_allclose(U.upfirdn2d(xs, fa, 2, 1, (-1, 2, 1, -2)), g["crop"], atol=2e-6)
    # SURVEY KAT
    y = U.upsample2d(x, f, up=2)
    assert y.shape == (1, 3, 128, 128) and abs(y[0, 0, 1, 1] - 0.0754179) < 1e-6


def test_shift_harness_matches_reference(golden):
    g = golden("shift")
    lat = torch.from_numpy(g["lat"])
    for k in range(4):
        ti, tj = g[f"shift{k}_t"]
        w, m = S.ideal_shift(lat, float(ti), float(tj), 8, crop=True)
        np.testing.assert_allclose(w.numpy(), g[f"shift{k}_img"], atol=3e-6)
        assert np.array_equal(m.numpy(), g[f"shift{k}_mask"])
    a, b, m = (torch.from_numpy(g[k]) for k in ("m_a", "m_b", "m_mask"))
    assert abs(float(S.mask_mse(a, b, m)) - float(g["mask_mse"])) < 1e-6
    assert abs(float(S.mask_psnr(a, b, m)) - float(g["mask_psnr"])) < 1e-5
    assert abs(float(S.psnr(a, b)) - float(g["psnr"])) < 1e-5


def test_superresolution_operators_match_reference(golden):
    """oracle/superres.py against the unmodified reference's ``build_sr4x`` / ``build_sr_bicubic(...).H``
    (afldm/af_libs/superresolution.py:263-320): the x4 degradation of scripts/shift_ldm_sr.py."""
    from oracle import superres as SR
    g = golden("superres")
    for n in (32, 64):
        x = g[f"x_{n}"]
        for filt in ("bicubic", "pool"):
            np.testing.assert_allclose(SR.sr4x(x, filt), g[f"sr4x_{filt}_{n}"], atol=2e-6)
        np.testing.assert_allclose(SR.degrade(x, "bicubic").reshape(2, -1), g[f"H_{n}"], atol=2e-6)
    x = (torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(296)) * 2 - 1).numpy()
    y = SR.sr4x(x, "bicubic")
    assert abs(y.sum() - g["sum_256"][0]) < 2e-2 and abs(np.abs(y).sum() - g["sum_256"][1]) < 2e-2
    np.testing.assert_allclose(y[0, :, ::37, ::41], g["rows_256"], atol=2e-6)
    s = np.linalg.svd(SR.h_small_bicubic(256), compute_uv=False)
    assert s.min() > 0.3            # nothing is cut at 3e-2 for the script's size: the operator IS the bicubic filter


def test_lanczos_shifter_matches_reference(golden):
    """oracle/shift.lanczos_shift against the unmodified reference's ``ImageShifter('lanczos').shift``
    (shifters.py:159-162 -> StyleGAN3 ``apply_fractional_translation``, af_libs/equivariance.py:68-103), including a
    shift that leaves nothing of the image."""
    g = golden("shift")
    img = torch.from_numpy(g["img"])
    for k in range(5):
        ti, tj = g[f"lz{k}_t"]
        w, m = S.lanczos_shift(img, float(ti), float(tj))
        np.testing.assert_allclose(w.numpy(), g[f"lz{k}_img"], atol=3e-6)
        assert np.array_equal(m.numpy(), g[f"lz{k}_mask"])
    assert g["lz4_mask"].sum() == 0 and np.abs(g["lz4_img"]).max() < 1e-6
