"""GPU parity: Synthesis through the C-ABI vs the reference (fed the reference's own
f0 / spectrogram / aperiodicity so that only Synthesis is under test)."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: synthesized waveform within 1e-4 relative (to the waveform's peak) in fp64


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (22050, 1.5), (48000, 2.0)])
def test_synthesis_matches_reference(wb, signals, fs, seconds):
    x = signals.synth_speech(fs, seconds, seed=3)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    # fresh process semantics: CheapTrick and D4C consumed randn() before Synthesis; replay their
    # consumption by running them (their parity is covered elsewhere), then synthesize from the
    # REFERENCE parameters.
    wb.randn_reseed()
    wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0)).compute(x, ref["tpos"], ref["f0"])
    wb.D4C(fs).compute(x, ref["tpos"], ref["f0"], ref["fft_size"])
    syn = wb.Synthesis(fs, ref["fft_size"], 5.0)
    y = syn.compute(ref["f0"], ref["sp"], ref["ap"], len(ref["y"]))
    assert np.all(np.isfinite(y))
    peak = np.abs(ref["y"]).max()
    err = float(np.abs(y - ref["y"]).max() / peak)
    print("synthesis fs=%d peak %.3f max err/peak %.3e" % (fs, peak, err))
    assert err < RTOL
