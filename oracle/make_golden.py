"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    PYTHONPATH=/root/reference python oracle/make_golden.py

Imports only reference modules that are importable without diffusers:
afldm.af_libs.ideal_lpf, afldm.af_libs.torch_utils.ops.upfirdn2d, afldm.shift_utils.{shifters,metrics},
afldm.af_libs.superresolution.
`/root/reference` cannot travel to the GPU box, so the vectors are committed as small fixtures.
Test infrastructure only.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")

from afldm.af_libs import ideal_lpf as R                      # noqa: E402
from afldm.af_libs.torch_utils.ops import upfirdn2d as RU     # noqa: E402
from afldm.shift_utils import metrics as RM                   # noqa: E402
from afldm.shift_utils import shifters as RS                  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)


def randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def main():
    # ---- masks (a1, a2)
    masks = {}
    for n in (2, 4, 6, 8, 12, 16, 32, 64, 128):
        masks[f"lpf_{n}"] = R.create_lpf_rect(n, 0.5).numpy()
        masks[f"recon_{n}"] = R.create_recon_rect(n, 0.5).numpy()
    masks["recon_16_c8"] = R.create_recon_rect(16, 1 / 8).numpy()
    masks["recon_64_c8"] = R.create_recon_rect(64, 1 / 8).numpy()
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **masks)

    # ---- ideal ops (a3-a6): input seeds/shapes + outputs
    ops = {}
    cases = {"s2": ((2, 3, 2, 2), 11), "s4": ((2, 5, 4, 4), 12), "s8": ((2, 3, 8, 8), 13),
             "s16": ((1, 4, 16, 16), 14), "s32": ((1, 3, 32, 32), 15), "s64": ((1, 3, 64, 64), 1234),
             "s6": ((1, 2, 6, 6), 16)}
    up, lpf = R.UpsampleRFFT(), R.LPF_RFFT(0.5)
    for name, (shape, seed) in cases.items():
        x = randn(shape, seed)
        ops[f"{name}_seed"] = np.array(seed)
        ops[f"{name}_x"] = x.numpy()
        u = up(x)
        ops[f"{name}_up2"] = u.numpy()
        ops[f"{name}_lpf_down2"] = lpf(x.clone())[:, :, ::2, ::2].contiguous().numpy()
        ops[f"{name}_filtered_silu"] = lpf(torch.nn.functional.silu(u))[:, :, ::2, ::2].contiguous().numpy()
    x8 = randn((1, 4, 8, 8), 21)
    ops["up8_x"] = x8.numpy()
    ops["up8_y"] = R.UpsampleRFFT(8)(x8).numpy()
    ops["subpix_x"] = x8.numpy()
    ops["subpix_y"] = R.subpixel_shift(x8, up=2, shift_x=1, shift_y=1).contiguous().numpy()
    # module surface beyond the two hot configurations (ideal_lpf.py:52-172): full-resolution LPF at several cutoffs,
    # the reconstruction filter, other up factors, odd-friendly sizes, sub-pixel shifts
    x12 = randn((1, 3, 12, 12), 31)
    ops["surf_x12"] = x12.numpy()
    ops["surf_lpf12_c50"] = R.LPF_RFFT(0.5)(x12.clone()).numpy()
    ops["surf_lpf12_c25"] = R.LPF_RFFT(0.25)(x12.clone()).numpy()
    ops["surf_recon12_c50"] = R.LPF_RECON_RFFT(0.5)(x12.clone()).numpy()
    ops["surf_up3_12"] = R.UpsampleRFFT(3)(x12).numpy()
    ops["surf_up2_12"] = R.UpsampleRFFT(2)(x12).numpy()
    x16 = randn((2, 2, 16, 16), 32)
    ops["surf_x16"] = x16.numpy()
    ops["surf_lpf16_c50"] = R.LPF_RFFT(0.5)(x16.clone()).numpy()
    ops["surf_up4_16"] = R.UpsampleRFFT(4)(x16).numpy()
    ops["surf_subpix4_16"] = R.subpixel_shift(x16, up=4, shift_x=3, shift_y=-2).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, "ideal_ops.npz"), **ops)

    # ---- upfirdn2d (a16, BASELINE config #1)
    ud = {}
    x = randn((1, 3, 64, 64), 1234)
    f = RU.setup_filter([1, 3, 3, 1])
    ud["x"] = x.numpy()
    ud["f1331"] = f.numpy()
    ud["up2"] = RU.upsample2d(x, f, up=2, impl="ref").numpy()
    ud["down2"] = RU.downsample2d(x, f, down=2, impl="ref").numpy()
    ud["filter2d"] = RU.filter2d(x, f, impl="ref").numpy()
    f12 = RU.setup_filter([1, 2, 4, 7, 9, 11, 11, 9, 7, 4, 2, 1])      # separable (>= 8 taps)
    ud["f12"] = f12.numpy()
    ud["up2_f12"] = RU.upsample2d(x, f12, up=2, impl="ref").numpy()
    xs = randn((2, 2, 9, 7), 5)
    fa = RU.setup_filter([1, 2, 5], normalize=True)
    ud["xs"] = xs.numpy()
    ud["fa"] = fa.numpy()
    ud["gen"] = RU.upfirdn2d(xs, fa, up=3, down=2, padding=[2, 1, 0, 3], flip_filter=True, gain=1.7, impl="ref").numpy()
    ud["gen_noflip"] = RU.upfirdn2d(xs, fa, up=2, down=1, padding=[1, 1, 2, 0], flip_filter=False, gain=1.0, impl="ref").numpy()
    ud["crop"] = RU.upfirdn2d(xs, fa, up=2, down=1, padding=[-1, 2, 1, -2], impl="ref").numpy()
    np.savez_compressed(os.path.join(OUT, "upfirdn2d.npz"), **ud)

    # ---- shift harness (a15)
    sh = {}
    lat = randn((1, 4, 32, 32), 77)
    sh["lat"] = lat.numpy()
    shifter = RS.ImageShifter("ideal_crop", 8)
    for k, (ti, tj) in enumerate([(0, 1 / 8), (0, 5 / 8), (0.25, 1.0), (-0.375, 2.125)]):
        w, m = shifter.shift(lat, ti, tj)
        sh[f"shift{k}_t"] = np.array([ti, tj])
        sh[f"shift{k}_img"] = w.contiguous().numpy()
        sh[f"shift{k}_mask"] = m.numpy()
    # the default (bilinear flow_warp) shifter on an image-sized tensor, 'ideal' (no crop) and the Fourier shifter
    img = randn((2, 3, 24, 24), 78)
    sh["img"] = img.numpy()
    bil = RS.ImageShifter()
    for k, (ti, tj) in enumerate([(0.0, 1.0), (0.0, 2.5), (1.75, -3.25), (-4.0, 0.5)]):
        w, m = bil.shift(img, ti, tj)
        sh[f"bil{k}_t"] = np.array([ti, tj])
        sh[f"bil{k}_img"] = w.contiguous().numpy()
        sh[f"bil{k}_mask"] = m.numpy()
    w, m = RS.ImageShifter("ideal", 8).shift(lat, 0.5, -1.375)
    sh["ideal_img"], sh["ideal_mask"] = w.contiguous().numpy(), m.numpy()
    sh["fourier_img"] = RS.fourier_shift_batch(img, 0.75, -2.3, device="cpu").numpy()
    w, m = RS.ImageShifter("fourier_crop").shift(img, 1.5, 0.25)
    sh["fcrop_img"], sh["fcrop_mask"] = w.contiguous().numpy(), m.numpy()
    lz = RS.ImageShifter("lanczos")
    for k, (ti, tj) in enumerate([(0.0, 0.5), (2.75, -1.5), (-3.3, 4.9), (1 / 3, -2 / 3), (23.5, -26.0)]):
        w, m = lz.shift(img, ti, tj)
        sh[f"lz{k}_t"] = np.array([ti, tj])
        sh[f"lz{k}_img"] = w.contiguous().numpy()
        sh[f"lz{k}_mask"] = m.numpy()
    a, b = randn((2, 3, 16, 16), 3), randn((2, 3, 16, 16), 4)
    m = RS.gen_valid_mask(a.shape, 2.5, -1.25)
    sh["m_a"], sh["m_b"], sh["m_mask"] = a.numpy(), b.numpy(), m.numpy()
    sh["mask_mse"] = RM.mask_mse(a, b, m).numpy()
    sh["mask_psnr"] = RM.mask_psnr(a, b, m).numpy()
    sh["psnr"] = RM.psnr(a, b).numpy()
    np.savez_compressed(os.path.join(OUT, "shift.npz"), **sh)

    # ---- x4 degradation operators of the super-resolution scripts (scripts/shift_ldm_sr.py:45)
    from afldm.af_libs import superresolution as RSR
    sr = {}
    for n in (32, 64):
        x = torch.rand(2, 3, n, n, generator=torch.Generator().manual_seed(40 + n)) * 2 - 1
        sr[f"x_{n}"] = x.numpy()
        for filt in ("bicubic", "pool"):
            sr[f"sr4x_{filt}_{n}"] = RSR.build_sr4x("cpu", filt, n)(x).numpy()
        hb = RSR.build_sr_bicubic(4, "cpu", n)
        y = hb.H(x)
        sr[f"H_{n}"] = y.numpy()
        sr[f"Ht_{n}"] = hb.Ht(y).numpy()
        sr[f"Hpinv_{n}"] = hb.H_pinv(y).numpy()
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(296)) * 2 - 1       # the script's size: checksums
    y = RSR.build_sr4x("cpu", "bicubic", 256)(x)
    sr["sum_256"] = np.array([y.double().sum().item(), y.double().abs().sum().item()])
    sr["rows_256"] = y[0, :, ::37, ::41].numpy()
    np.savez_compressed(os.path.join(OUT, "superres.npz"), **sr)

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
