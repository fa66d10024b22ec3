"""GPU parity: D4C (Love Train + band aperiodicity) through the C-ABI vs the reference."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: aperiodicity within 1e-4 relative in fp64


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (22050, 1.5), (48000, 2.0)])
def test_d4c_matches_reference(wb, signals, fs, seconds):
    x = signals.synth_speech(fs, seconds, seed=1)
    ref, _ = refbin.run_reference(x, fs, stages="hd")
    wb.randn_reseed()
    d4c = wb.D4C(fs, wb.D4COption(threshold=0.85))
    ap = d4c.compute(x, ref["tpos"], ref["f0"], ref["fft_size"])
    assert ap.shape == ref["ap"].shape
    assert np.all(np.isfinite(ap))
    # the voiced/unvoiced decision of Love Train must agree exactly: rows left at 1 - 1e-12
    unv_ref = np.all(ref["ap"] == 1.0 - 1e-12, axis=1)
    unv_got = np.all(ap == 1.0 - 1e-12, axis=1)
    assert np.array_equal(unv_ref, unv_got)
    assert (~unv_ref).sum() > 10
    err = float(np.max(np.abs(ap - ref["ap"]) / np.abs(ref["ap"])))
    print("d4c fs=%d analysed frames %d max rel err %.3e" % (fs, (~unv_ref).sum(), err))
    assert err < RTOL


def test_d4c_after_cheaptrick_stream_position(wb, signals):
    """Stage order CheapTrick -> D4C consumes the randn stream like one reference process."""
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=2)
    ref, _ = refbin.run_reference(x, fs, stages="hcd")
    wb.randn_reseed()
    sp = wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0)).compute(x, ref["tpos"], ref["f0"])
    ap = wb.D4C(fs).compute(x, ref["tpos"], ref["f0"], ref["fft_size"])
    assert np.max(np.abs(sp - ref["sp"]) / ref["sp"]) < RTOL
    assert np.max(np.abs(ap - ref["ap"]) / ref["ap"]) < RTOL
