"""GPU parity: CheapTrick through the C-ABI vs the reference's own serial CPU build."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: spectral envelope within 1e-4 relative in fp64


def _rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (22050, 1.5), (48000, 2.0)])
def test_cheaptrick_matches_reference(wb, signals, fs, seconds):
    x = signals.synth_speech(fs, seconds, seed=0)
    ref, _ = refbin.run_reference(x, fs, stages="hc")
    wb.randn_reseed()
    ct = wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0))
    assert ct.fft_size == ref["fft_size"]
    sp = ct.compute(x, ref["tpos"], ref["f0"])
    assert sp.shape == ref["sp"].shape
    assert np.all(np.isfinite(sp))
    err = _rel_err(sp, ref["sp"])
    print("cheaptrick fs=%d max rel err %.3e" % (fs, err))
    assert err < RTOL
