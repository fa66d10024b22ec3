"""GPU parity of the parameter modification between analysis and synthesis against
the reference demo's own ParameterModification (test/test.cpp:201-243, run through oracle/_ref/refmod)."""
import os

import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shift,ratio", [(1.5, None), (0.8, 1.3), (1.25, 0.7), (1.0, 1.0), (2.0, 0.31)])
def test_host_api_matches_reference(wb, shift, ratio):
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_16k_1s.npz"))
    fs, fft = int(g["fs"]), int(g["fft_size"])
    f0, sp = g["f0"].copy(), g["sp"].copy()
    if refbin.available() and os.path.exists(refbin.REFMOD):
        rf0, rsp = refbin.run_modification(f0, sp, fs, fft, shift, ratio)
    else:
        from oracle import world_np
        rf0, rsp = world_np.parameter_modification(f0, sp, fs, fft, shift, ratio)
    of0, osp = wb.ParameterModification(f0, sp, fs, fft, f0_shift=shift, ratio=ratio)
    assert of0 is f0 and osp is sp   # in place, like the demo
    assert np.array_equal(of0, rf0)  # one multiplication: bit-exact
    assert np.max(np.abs(osp - rsp) / rsp) < 1e-12


def test_golden_vectors(wb):
    m = np.load(os.path.join(ROOT, "tests", "golden", "mod_16k.npz"))
    for tag in "abc":
        shift, ratio = m["args_" + tag]
        f0, sp = wb.ParameterModification(m["f0_in"].copy(), m["sp_in"].copy(), int(m["fs"]), int(m["fft_size"]),
                                          f0_shift=float(shift), ratio=None if np.isnan(ratio) else float(ratio))
        assert np.array_equal(f0, m["f0_" + tag]), tag
        assert np.max(np.abs(sp - m["sp_" + tag]) / m["sp_" + tag]) < 1e-12, tag


def test_pipeline_applies_modification_between_analysis_and_synthesis(wb, signals):
    """The resident chain with a modification equals the class-API chain with the modification applied by
    hand between D4C and Synthesis (same randn positions -> bit-identical waveform)."""
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=31)
    hopt, copt, dopt = wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85)
    shift, ratio = 1.3, 0.85
    wb.randn_reseed()
    pl = wb.Pipeline(fs, hopt, copt, dopt)
    pl.set_modification(f0_shift=shift, ratio=ratio)
    out = pl.run(x)
    wb.randn_reseed()
    hv, ct, d4 = wb.Harvest(fs, hopt), wb.CheapTrick(fs, copt), wb.D4C(fs, dopt)
    tpos, f0 = hv.compute(x)
    sp = ct.compute(x, tpos, f0)
    ap = d4.compute(x, tpos, f0, ct.fft_size)
    wb.ParameterModification(f0, sp, fs, ct.fft_size, f0_shift=shift, ratio=ratio)
    y = wb.Synthesis(fs, ct.fft_size, 5.0).compute(f0, sp, ap, len(out["y"]))
    assert np.array_equal(out["f0"], f0) and np.array_equal(out["sp"], sp) and np.array_equal(out["ap"], ap)
    assert np.array_equal(out["y"], y)
    assert np.abs(y).max() > 0.05
    # and the analysis half is the reference's: modification of the reference's parameters
    ref, _ = refbin.run_reference(x, fs, stages="hcd")
    rf0, rsp = refbin.run_modification(ref["f0"], ref["sp"], fs, ct.fft_size, shift, ratio)
    v = rf0 > 0
    assert np.array_equal(out["f0"] > 0, v)
    assert np.max(np.abs(out["f0"][v] - rf0[v]) / rf0[v]) < 1e-4
    assert np.max(np.abs(out["sp"] - rsp) / rsp) < 1e-4


def test_bad_arguments(wb):
    f0, sp = np.ones(4), np.ones((4, 513))
    with pytest.raises(wb.WorldB200Error):
        wb.ParameterModification(f0, sp, 16000, 1024, ratio=1e-4)   # band edge below the first bin
    pl = wb.Pipeline(16000)
    with pytest.raises(wb.WorldB200Error):
        pl.set_modification(f0_shift=-1.0)
