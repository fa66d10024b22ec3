"""The host-side helpers of the drop-in headers (include/world_common.hpp, include/world_matlabfunctions.hpp) against
the reference's own: tests/dropin/helpers_main.cpp is built twice -- with the reference's headers and sources
(oracle/_ref/refhelpers) and with this repository's headers (tests/dropin/_bin/helpers_main) -- and both are run on
the same seeded inputs.  histc / interp1 (with its extrapolation) / interp1Q / fftshift / diff / matlab_round /
GetSuitableFFTSize / DCCorrection / LinearSmoothing / NuttallWindow need no GPU; test/test.cpp:228 depends on interp1."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "tests", "dropin", "_bin", "helpers_main")
REF = os.path.join(ROOT, "oracle", "_ref", "refhelpers")


@pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="run __graft_entry__.build() where /root/reference exists")
def test_header_helpers_equal_the_references(tmp_path):
    a, b = str(tmp_path / "ours.f64"), str(tmp_path / "ref.f64")
    for exe, out in ((OURS, a), (REF, b)):
        r = subprocess.run([exe, "host", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    va, vb = np.fromfile(a), np.fromfile(b)
    assert len(va) == len(vb) > 7000
    assert np.all(np.isfinite(vb))
    # the same expressions in the same order: identical on this compiler; the bound is what a different
    # contraction of the interpolation products could cost
    assert np.max(np.abs(va - vb) / np.maximum(np.abs(vb), 1e-300)) < 1e-12
