"""The C++ boundary EXECUTED on the GPU: programs that use include/{harvest,cheaptrick,d4c,
synthesis,codec,world_common,world_matlabfunctions,world_fft}.hpp the way the reference's consumer does, linked
against libworldb200.so only (built by __graft_entry__.build()), compared with the reference.

 * class_api_main (ours, tests/dropin/class_api_main.cpp): the call sequence of test/test.cpp:76-264 on raw arrays
   -> every array against one reference process;
 * demo_dropin: the reference's OWN demo program (test/test.cpp, compiled where it lies against the drop-in headers,
   wav in -> wav out, with F0 scaling and formant shift) against the same program linked with the reference's sources;
 * helpers_main all: decimate and the FFT structs (GPU-backed in the drop-in) against the reference's."""
import os
import subprocess

import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "dropin", "_bin")
RTOL = 1e-4


def _need(*paths):
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        pytest.skip("not built: %s (run __graft_entry__.build() where /root/reference exists)" % ", ".join(missing))


def test_cpp_class_api_on_the_gpu_matches_reference(tmp_path, signals):
    exe = os.path.join(BIN, "class_api_main")
    _need(exe)
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=0)
    xin, prefix = str(tmp_path / "x.f64"), str(tmp_path / "o")
    x.tofile(xin)
    r = subprocess.run([exe, xin, str(fs), prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref, _ = refbin.run_reference(x, fs, stages="hcdsk")   # one process: the codec runs last, like in class_api_main
    L, bins = len(ref["f0"]), ref["fft_size"] // 2 + 1
    got = {k: np.fromfile(prefix + "_%s.f64" % k) for k in ("tpos", "f0", "sp", "ap", "y", "coded_sp", "coded_ap")}
    assert np.array_equal(got["tpos"], ref["tpos"])
    assert np.array_equal(got["f0"] > 0, ref["f0"] > 0)
    v = ref["f0"] > 0
    assert np.max(np.abs(got["f0"][v] - ref["f0"][v]) / ref["f0"][v]) < RTOL
    assert np.max(np.abs(got["sp"].reshape(L, bins) - ref["sp"]) / ref["sp"]) < RTOL
    assert np.max(np.abs(got["ap"].reshape(L, bins) - ref["ap"]) / ref["ap"]) < RTOL
    assert np.max(np.abs(got["y"] - ref["y"])) / np.abs(ref["y"]).max() < RTOL
    if "csp" in ref:
        assert np.max(np.abs(got["coded_sp"].reshape(L, -1) - ref["csp"])) < 1e-6 * max(1.0, np.abs(ref["csp"]).max())
        assert np.max(np.abs(got["coded_ap"].reshape(L, -1) - ref["cap"])) < 1e-6 * max(1.0, np.abs(ref["cap"]).max())


def test_reference_demo_program_runs_on_the_gpu(tmp_path, wb, signals):
    """test/test.cpp itself: analysis, ParameterModification (F0 x 1.2, formants x 0.9), synthesis, 16-bit wav out."""
    ours, theirs = os.path.join(BIN, "demo_dropin"), os.path.join(ROOT, "oracle", "_ref", "ref_demo")
    _need(ours, theirs)
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=3)
    wav_in = str(tmp_path / "in.wav")
    wb.wavwrite(x, fs, 16, wav_in)
    outs = []
    for exe, tag in ((ours, "gpu"), (theirs, "ref")):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([exe, wav_in, "out", "1.2", "0.9"], capture_output=True, text=True, cwd=str(d))
        assert r.returncode == 0 and "complete." in r.stdout, r.stdout + r.stderr
        y, fs_out, nbit = wb.wavread(str(d / "out_1.wav"))
        assert fs_out == fs and nbit == 16
        outs.append(y)
    a, b = outs
    assert len(a) == len(b) and np.abs(b).max() > 0.05
    # 16-bit samples: the two programs may round a sample to neighbouring codes
    assert np.max(np.abs(a - b)) <= 1.5 / 32768.0
    assert np.mean(a != b) < 0.01


def test_helpers_with_gpu_backed_parts_match_reference(tmp_path):
    ours, theirs = os.path.join(BIN, "helpers_main"), os.path.join(ROOT, "oracle", "_ref", "refhelpers")
    _need(ours, theirs)
    a, b = str(tmp_path / "ours.f64"), str(tmp_path / "ref.f64")
    for exe, out in ((ours, a), (theirs, b)):
        r = subprocess.run([exe, "all", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    va, vb = np.fromfile(a), np.fromfile(b)
    assert len(va) == len(vb) > 15000 and np.all(np.isfinite(va))
    # transforms of 2048 points agree to rounding relative to the size of the spectra; the rest is identical
    scale = np.maximum(np.abs(vb), 1.0)
    assert np.max(np.abs(va - vb) / scale) < 1e-9
