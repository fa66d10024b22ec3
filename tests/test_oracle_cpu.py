"""CPU tests pinning the numpy/C restatement (oracle/world_np.py, oracle/world_c.c) to the REFERENCE:
the golden vectors in tests/golden/ were produced by the reference's own sources (oracle/_ref/refrun,
see tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import world_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_16k_1s.npz"))
FS = int(G["fs"])
FFT = int(G["fft_size"])


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_randn_c_matches_python_restatement():
    a, state = world_np.randn_stream(500)
    s = world_np.RandnStream()
    b = s.take(500)
    assert np.array_equal(a, b) and tuple(s.state) == state


@pytest.fixture(scope="module")
def chain():
    """CheapTrick -> D4C -> Synthesis restated, fed the golden f0, drawing from ONE randn stream in
    the reference's stage order (the golden run is Harvest -> CheapTrick -> D4C -> Synthesis; Harvest
    draws no random numbers)."""
    rng = world_np.RandnStream()
    sp = world_np.cheaptrick(G["x"], FS, G["tpos"], G["f0"], rng)
    ap = world_np.d4c(G["x"], FS, G["tpos"], G["f0"], FFT, rng)
    y = world_np.synthesis(G["f0"], G["sp"], G["ap"], FS, FFT, 5.0, len(G["y"]), rng)
    return sp, ap, y


def test_cheaptrick_restatement_matches_reference(chain):
    assert world_np.cheaptrick_fft_size(FS) == FFT
    assert _rel(chain[0], G["sp"]) < 1e-9


def test_d4c_restatement_matches_reference(chain):
    ap = chain[1]
    assert np.array_equal(np.all(ap == 1.0 - 1e-12, axis=1), np.all(G["ap"] == 1.0 - 1e-12, axis=1))
    assert _rel(ap, G["ap"]) < 1e-8


def test_synthesis_restatement_matches_reference(chain):
    y = chain[2]
    assert np.max(np.abs(y - G["y"])) / np.abs(G["y"]).max() < 1e-10


def test_codec_restatement_matches_reference():
    csp = world_np.code_spectral_envelope(G["sp"], FS, FFT, 60)
    cap = world_np.code_aperiodicity(G["ap"], FS, FFT)
    dsp = world_np.decode_spectral_envelope(G["csp"], FS, FFT, 60)
    dap = world_np.decode_aperiodicity(G["cap"], FS, FFT)
    assert np.max(np.abs(csp - G["csp"])) / np.abs(G["csp"]).max() < 1e-11
    assert np.max(np.abs(cap - G["cap"])) < 1e-9
    assert _rel(dsp, G["dsp"]) < 1e-10
    assert _rel(dap, G["dap"]) < 1e-12


def test_fft_wrapper_convention():
    # forward = e^{+i}, unnormalised; c2r(r2c(x)) = N x
    n = 64
    x = np.zeros(n)
    x[1] = 1.0
    k = np.arange(n // 2 + 1)
    assert np.allclose(world_np.fft_r2c(x), np.exp(2j * np.pi * k / n))
    z = np.random.default_rng(0).standard_normal(n)
    assert np.allclose(world_np.fft_c2r(world_np.fft_r2c(z), n), n * z)


def test_golden_48k_voicing_matches_second_fixture():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_48k_1s.npz"))
    assert int(g["fft_size"]) == 2048 and g["sp_every8"].shape[1] == 1025
    assert (g["f0"] > 0).sum() > 100


def test_parameter_modification_restatement_matches_reference():
    """oracle/world_np.parameter_modification vs the reference demo's own function (golden from oracle/_ref/refmod)."""
    m = np.load(os.path.join(ROOT, "tests", "golden", "mod_16k.npz"))
    for tag in "abc":
        shift, ratio = m["args_" + tag]
        f0, sp = world_np.parameter_modification(m["f0_in"], m["sp_in"], int(m["fs"]), int(m["fft_size"]), shift,
                                                 None if np.isnan(ratio) else float(ratio))
        assert np.array_equal(f0, m["f0_" + tag]), tag
        assert _rel(sp, m["sp_" + tag]) < 1e-14, tag
