"""GPU parity of the codec round trip (BASELINE configs[4]) against the reference's codec.cpp."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (48000, 5.2)])
def test_codec_matches_reference(wb, signals, fs, seconds):
    """48 kHz / 5.2 s gives 1041 frames >= the 1024-frame batch of configs[4]; nd = 60."""
    x = signals.synth_speech(fs, seconds, seed=9)
    ref, _ = refbin.run_reference(x, fs, stages="hcdk", codec_nd=60)
    fft_size = ref["fft_size"]
    n_ap = wb.GetNumberOfAperiodicities(fs)
    assert n_ap == ref["cap"].shape[1]
    csp = wb.CodeSpectralEnvelope(ref["sp"], fs, fft_size, 60)
    cap = wb.CodeAperiodicity(ref["ap"], fs, fft_size)
    dsp = wb.DecodeSpectralEnvelope(ref["csp"], fs, fft_size, 60)
    dap = wb.DecodeAperiodicity(ref["cap"], fs, fft_size)
    scale = np.abs(ref["csp"]).max(axis=0, keepdims=True)
    assert np.max(np.abs(csp - ref["csp"]) / scale) < RTOL          # cepstra cross zero: relative to each coefficient's range
    assert np.max(np.abs(cap - ref["cap"]) / np.maximum(np.abs(ref["cap"]), 1e-3)) < RTOL
    assert np.max(np.abs(dsp - ref["dsp"]) / ref["dsp"]) < RTOL
    assert np.max(np.abs(dap - ref["dap"]) / ref["dap"]) < RTOL
    # rows the reference leaves at 1 - 1e-12 (CheckVUV) must be identical
    assert np.array_equal(np.all(dap == 1.0 - 1e-12, axis=1), np.all(ref["dap"] == 1.0 - 1e-12, axis=1))
    print("codec fs=%d frames %d: csp %.2e cap %.2e dsp %.2e dap %.2e" % (
        fs, len(ref["f0"]), np.max(np.abs(csp - ref["csp"]) / scale),
        np.max(np.abs(cap - ref["cap"]) / np.maximum(np.abs(ref["cap"]), 1e-3)),
        np.max(np.abs(dsp - ref["dsp"]) / ref["dsp"]), np.max(np.abs(dap - ref["dap"]) / ref["dap"])))
