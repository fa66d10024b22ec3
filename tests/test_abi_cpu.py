"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/worldb200.h declares, the pure host-side size formulas match the reference's expressions,
and computing without a GPU fails loudly (there is no CPU fallback)."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "worldb200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(wb):
    names = _declared_functions()
    assert len(names) >= 50
    lib = ctypes.CDLL(wb.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in worldb200.h but not exported: %s" % missing


def test_python_mirror_binds_every_symbol(wb):
    L = wb.lib()
    for n in _declared_functions():
        fn = getattr(L, n)
        assert fn.argtypes is not None, "%s has no ctypes signature in world-class_b200/__init__.py" % n


def test_no_torch_or_cuda_types_in_the_abi():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # declarations only
    assert "cudaStream_t" not in code and "torch" not in code and "at::" not in code and "#include" not in code


def test_size_formulas_match_reference_expressions(wb):
    L = wb.lib()
    for fs in (8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000):
        for n in (1, fs // 3, fs, 10 * fs + 7):
            for fp in (1.0, 5.0, 10.0):
                assert L.wb_harvest_get_samples(fs, n, fp) == int(1000.0 * n / fs / fp) + 1        # harvest.cpp:173-176
        for floor in (40.0, 71.0, 90.0):
            ref = int(math.pow(2.0, 1.0 + int(math.log(3.0 * fs / floor + 1) / 0.69314718055994529)))  # cheaptrick.cpp:97-100
            assert L.wb_cheaptrick_get_fft_size(fs, floor) == ref
        assert L.wb_cheaptrick_get_f0_floor(fs, 2048) == 3 * fs / (2048 - 3.0)                       # cheaptrick.cpp:102-105
        assert L.wb_get_number_of_aperiodicities(fs) == int(min(15000.0, fs / 2.0 - 3000.0) / 3000.0)  # d4c.cpp:65-67
    assert L.wb_cheaptrick_get_fft_size(48000, 71.0) == 2048 and L.wb_cheaptrick_get_fft_size(16000, 71.0) == 1024


def test_option_defaults(wb):
    h, c, d = wb.HarvestOption(), wb.CheapTrickOption(), wb.D4COption()
    assert (h.f0_floor, h.f0_ceil, h.frame_period, h.target_fs, h.channels_in_octave, h.use_cos_table) == \
        (71.0, 800.0, 5.0, 8000.0, 40.0, 0)                                                         # harvest.cpp:52-56
    assert (c.q1, c.f0_floor, c.fft_size) == (-0.15, 71.0, 0)                                        # cheaptrick.cpp:22-24
    assert d.threshold == 0.85                                                                       # d4c.cpp:31-33


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_compute_without_gpu_fails_loudly(wb):
    with pytest.raises(wb.WorldB200Error):
        wb.CheapTrick(16000)
    with pytest.raises(wb.WorldB200Error):
        wb.randn(4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference sources not present")
def test_reference_test_cpp_compiles_unchanged_against_our_headers(tmp_path):
    """Drop-in check: the reference's own demo compiles and links, unmodified,
    against include/*.hpp + libworldb200.so -- its wav I/O included (include/audioio.hpp replaces tools/audioio.hpp;
    no reference translation unit other than test.cpp itself is compiled)."""
    exe = tmp_path / "test_dropin"
    cmd = ["/usr/bin/g++", "-std=c++11", "-w", "-I" + os.path.join(ROOT, "include"),
           "-o", str(exe), "/root/reference/test/test.cpp",
           "-L" + os.path.join(ROOT, "world-class_b200"), "-lworldb200",
           "-Wl,-rpath," + os.path.join(ROOT, "world-class_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert exe.exists()


def test_synthetic_input_is_deterministic(signals):
    a = signals.synth_speech(16000, 1.0, seed=0)
    b = signals.synth_speech(16000, 1.0, seed=0)
    assert np.array_equal(a, b) and a.dtype == np.float64 and len(a) == 16000
    assert np.abs(a).max() < 1.0
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_16k_1s.npz"))
    assert np.array_equal(a, g["x"]), "generator drifted from the input the golden vectors were made from"
