"""On-device ideal shifters (afldm_b200/shift_utils): host matrices vs the reference's golden vectors and the oracle
(CPU), the CUDA separable-plane kernel vs both (GPU).

Reference: ImageShifter('ideal' | 'ideal_crop', r).shift - afldm/shift_utils/shifters.py:157-191, gen_valid_mask
:31-49.  Tolerance: 2e-5 abs on O(1) data against the reference's fp32 FFT result (the matrices are fp64-derived, the
kernel accumulates n fp32 FMAs per axis)."""
import os

import numpy as np
import pytest
import torch

from afldm_b200.shift_utils import shifters as SH
from oracle import shift as OS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shift.npz")
SHIFTS = [(0.125, 0.0), (0.5, -0.375), (-1.25, 2.0), (3.0, -2.625), (0.0, 0.0)]


def apply(x, ti, tj, r, crop):
    n = x.shape[-1]
    return np.einsum("ih,bchw,jw->bcij", SH.shift_matrix(n, ti, r, crop), x.double().numpy(), SH.shift_matrix(n, tj, r, crop))


def test_shift_matrices_reproduce_reference_golden_vectors():
    g = np.load(GOLD)
    lat = torch.from_numpy(g["lat"])
    for k in range(4):
        ti, tj = (float(v) for v in g[f"shift{k}_t"])
        got = apply(lat, ti, tj, 8, True)
        np.testing.assert_allclose(got, g[f"shift{k}_img"], atol=2e-5)
        assert np.array_equal(SH.gen_valid_mask(lat.shape, ti, tj).numpy(), g[f"shift{k}_mask"])


@pytest.mark.parametrize("n,r", [(32, 8), (16, 4), (64, 8), (32, 1), (32, 2)])
def test_shift_matrices_match_oracle(n, r):
    x = torch.randn(1, 2, n, n, generator=torch.Generator().manual_seed(n + r))
    for ti, tj in SHIFTS:
        for crop in (True, False):
            want, mask = OS.ideal_shift(x, ti, tj, r, crop=crop)
            np.testing.assert_allclose(apply(x, ti, tj, r, crop), want.numpy(), atol=2e-5)


def test_shifter_rejects_cpu_tensors_and_other_filters():
    with pytest.raises(AssertionError):
        SH.ImageShifter("nearest", 8)
    with pytest.raises(RuntimeError):
        SH.ImageShifter("ideal_crop", 8).shift(torch.zeros(1, 4, 32, 32), 0.5, 0.0)


@pytest.mark.gpu
def test_cuda_shifter_matches_reference_golden_vectors():
    g = np.load(GOLD)
    lat = torch.from_numpy(g["lat"]).cuda()
    sh = SH.ImageShifter("ideal_crop", 8)
    for k in range(4):
        ti, tj = (float(v) for v in g[f"shift{k}_t"])
        w, m = sh.shift(lat, ti, tj)
        np.testing.assert_allclose(w.cpu().numpy(), g[f"shift{k}_img"], atol=2e-5)
        assert np.array_equal(m.cpu().numpy(), g[f"shift{k}_mask"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,r,filt", [((1, 4, 32, 32), 8, "ideal_crop"), ((2, 3, 256, 256), 8, "ideal_crop"),
                                          ((1, 4, 64, 64), 4, "ideal"), ((3, 5, 16, 16), 2, "ideal_crop")])
def test_cuda_shifter_sweep_matches_oracle(shape, r, filt):
    """A whole sweep of shifts in one call (the 16-shift loop of scripts/shift_ldm_ffhq.py:118-150, batched)."""
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(shape[-1]))
    shifts = [(k / 8.0, -k / 8.0 + 0.25) for k in range(-3, 5)]
    w, m = SH.ImageShifter(filt, r).shift_batch(x.cuda(), shifts)
    assert w.shape == (len(shifts),) + shape
    for s, (ti, tj) in enumerate(shifts):
        want, mask = OS.ideal_shift(x, ti, tj, r, crop=filt == "ideal_crop")
        np.testing.assert_allclose(w[s].cpu().numpy(), want.numpy(), atol=3e-5)
        assert torch.equal(m[s].cpu(), mask)
    # integer shifts are exact rolls (up to the mask)
    w0, _ = SH.ImageShifter("ideal", r).shift(x.cuda(), 2.0, -3.0)
    torch.testing.assert_close(w0.cpu(), torch.roll(x, (2, -3), (2, 3)), rtol=0, atol=2e-5)
