"""The reference's file formats: tools/parameterio.cpp and tools/audioio.cpp.

Host code, no GPU needed.  Checked three ways: against files WRITTEN BY THE REFERENCE's own tools and committed
as a fixture (tests/golden/io_files.npz, made by tests/golden/make_io_golden.py), live against oracle/_ref/refio
where it is built, and through the source-compatible C++ shims (include/audioio.hpp, include/parameterio.hpp)."""
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "io_files.npz")
REFIO = os.path.join(ROOT, "oracle", "_ref", "refio")


def _maker():
    spec = importlib.util.spec_from_file_location("make_io_golden", os.path.join(ROOT, "tests", "golden", "make_io_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("tag,nd", [("full", 0), ("coded", 7)])
def test_writers_are_byte_identical_to_the_reference(wb, gold, tmp_path, tag, nd):
    tpos, f0, sp, ap, x = _maker().case_arrays(nd=nd)
    L, fft, fs, fp = len(f0), 64, 16000, 5.0
    p = lambda n: str(tmp_path / n)
    wb.WriteF0(p("f0.bin"), L, fp, tpos, f0, 0)
    wb.WriteF0(p("f0.txt"), L, fp, tpos, f0, 1)
    wb.WriteSpectralEnvelope(p("sp.bin"), fs, L, fp, fft, nd, sp)
    wb.WriteAperiodicity(p("ap.bin"), fs, L, fp, fft, nd, ap)
    wb.wavwrite(x, fs, 16, p("x.wav"))
    for ours, ref in (("f0.bin", "ref_f0.bin"), ("f0.txt", "ref_f0.txt"), ("sp.bin", "ref_sp.bin"),
                      ("ap.bin", "ref_ap.bin"), ("x.wav", "ref_x.wav")):
        assert open(p(ours), "rb").read() == gold["%s/%s" % (tag, ref)].tobytes(), ours
    # contiguous-matrix variant writes the same bytes (also from a wider, strided matrix)
    wide = np.zeros((L, sp.shape[1] + 5))
    wide[:, :sp.shape[1]] = sp
    rc = wb.lib().wb_write_parameter_matrix(0, p("sp2.bin").encode(), fs, L, fp, fft, nd, wide.ctypes.data, wide.shape[1])
    assert rc == 0 and open(p("sp2.bin"), "rb").read() == gold["%s/ref_sp.bin" % tag].tobytes()


@pytest.mark.parametrize("tag,nd", [("full", 0), ("coded", 7)])
def test_readers_on_files_written_by_the_reference(wb, gold, tmp_path, tag, nd):
    tpos, f0, sp, ap, x = _maker().case_arrays(nd=nd)
    p = lambda n: str(tmp_path / n)
    for n in ("ref_f0.bin", "ref_sp.bin", "ref_ap.bin", "ref_x.wav"):
        open(p(n), "wb").write(gold["%s/%s" % (tag, n)].tobytes())
    hdr = json.loads(gold["%s/header" % tag].tobytes().decode())
    for key, tagname, f in (("NOF", "NOF ", "ref_f0.bin"), ("FP", "FP  ", "ref_sp.bin"), ("FFT", "FFT ", "ref_sp.bin"),
                            ("NOD", "NOD ", "ref_sp.bin"), ("FS", "FS  ", "ref_ap.bin")):
        assert wb.GetHeaderInformation(p(f), tagname) == hdr[key]
    assert wb.GetHeaderInformation(p("ref_f0.bin"), "FFT ") == 0      # not in an F0 file (parameterio.hpp:53-54)
    t2, f2 = wb.ReadF0(p("ref_f0.bin"))
    assert np.array_equal(f2, f0) and np.array_equal(t2, gold["%s/tpos_read" % tag])
    assert np.array_equal(wb.ReadSpectralEnvelope(p("ref_sp.bin")), sp)
    assert np.array_equal(wb.ReadAperiodicity(p("ref_ap.bin")), ap)
    m = np.full((len(f0), sp.shape[1] + 3), -1.0)
    assert wb.lib().wb_read_parameter_matrix(1, p("ref_ap.bin").encode(), m.ctypes.data, m.shape[1]) == 0
    assert np.array_equal(m[:, :sp.shape[1]], ap) and np.all(m[:, sp.shape[1]:] == -1.0)
    xr, fs, nbit = wb.wavread(p("ref_x.wav"))
    assert (fs, nbit) == (hdr["wav_fs"], hdr["wav_nbit"]) and wb.GetAudioLength(p("ref_x.wav")) == hdr["wav_length"]
    assert np.array_equal(xr, gold["%s/x_read" % tag])
    pcm, fs2 = wb.wavread_pcm16(p("ref_x.wav"))
    assert fs2 == fs and np.array_equal(pcm.astype(np.float64) / 32768.0, xr)
    # wavwrite's quantisation (audioio.cpp:176-180): truncation toward zero, clamped
    want = np.clip(np.trunc(x * 32767), -32768, 32767).astype(np.int16)
    assert np.array_equal(pcm, want)


@pytest.mark.parametrize("nbit", [8, 16, 24, 32])
@pytest.mark.parametrize("extra", [False, True])
def test_wavread_bit_depths_match_the_reference(wb, gold, tmp_path, nbit, extra):
    key = "wav%d%s" % (nbit, "x" if extra else "")
    f = str(tmp_path / "w.wav")
    open(f, "wb").write(gold[key + "/bytes"].tobytes())
    hdr = json.loads(gold[key + "/header"].tobytes().decode())
    assert wb.GetAudioLength(f) == hdr["wav_length"] == 301
    x, fs, nb = wb.wavread(f)
    assert (fs, nb) == (hdr["wav_fs"], hdr["wav_nbit"]) == (22050, nbit)
    assert np.array_equal(x, gold[key + "/x_read"])
    assert x[0] == -1.0 and x[2] == 0.0


def test_error_behaviour(wb, tmp_path):
    missing = str(tmp_path / "nope.wav")
    assert wb.GetAudioLength(missing) == 0                       # audioio.cpp:188-190
    bad = str(tmp_path / "bad.wav")
    open(bad, "wb").write(b"RIFX" + b"\0" * 60)
    assert wb.GetAudioLength(bad) == -1                          # audioio.cpp:192-195
    stereo = bytearray(_maker().wav_bytes([1, 2, 3], 16, 8000))
    stereo[22] = 2
    open(bad, "wb").write(bytes(stereo))
    assert wb.GetAudioLength(bad) == -1
    with pytest.raises(wb.WorldB200Error):
        wb.wavread(missing)
    assert wb.GetHeaderInformation(missing, "NOF ") == 0         # parameterio.cpp:123-126
    # a file of the wrong kind is refused by the magic check (parameterio.cpp:48-58)
    f = str(tmp_path / "f0.bin")
    wb.WriteF0(f, 3, 5.0, np.zeros(3), np.ones(3), 0)
    rows = np.zeros((3, 4))
    with pytest.raises(wb.WorldB200Error):
        wb.ReadSpectralEnvelope(f)
    assert wb.lib().wb_read_parameter_matrix(0, f.encode(), rows.ctypes.data, 4) != 0
    assert wb.lib().wb_write_f0(str(tmp_path / "no" / "dir.bin").encode(), 3, 5.0, rows.ctypes.data, rows.ctypes.data, 0) != 0
    # empty contour: header only
    wb.WriteF0(f, 0, 5.0, np.zeros(0), np.zeros(0), 0)
    assert os.path.getsize(f) == 4 + 8 + 12 and wb.GetHeaderInformation(f, "NOF ") == 0


@pytest.mark.skipif(not os.path.exists(REFIO), reason="oracle/_ref/refio not built (needs /root/reference)")
def test_reference_readers_accept_our_files(wb, tmp_path):
    tpos, f0, sp, ap, x = _maker().case_arrays(seed=77, L=31, fft=128, nd=0, nx=1234)
    d = str(tmp_path)
    wb.WriteF0(os.path.join(d, "our_f0.bin"), len(f0), 2.5, tpos, f0, 0)
    wb.WriteSpectralEnvelope(os.path.join(d, "our_sp.bin"), 44100, len(f0), 2.5, 128, 0, sp)
    wb.WriteAperiodicity(os.path.join(d, "our_ap.bin"), 44100, len(f0), 2.5, 128, 0, ap)
    wb.wavwrite(x, 44100, 16, os.path.join(d, "our_x.wav"))
    r = subprocess.run([REFIO, "read", d, "our"], check=True, capture_output=True, text=True)
    hdr = json.loads(r.stdout.strip())
    assert hdr["ok"] == [1, 1, 1] and (hdr["NOF"], hdr["FP"], hdr["FFT"], hdr["NOD"], hdr["FS"]) == (31, 2.5, 128, 0, 44100)
    assert (hdr["wav_length"], hdr["wav_fs"], hdr["wav_nbit"]) == (1234, 44100, 16)
    rd = lambda n: np.fromfile(os.path.join(d, "our_%s.rd.f64" % n))
    assert np.array_equal(rd("f0"), f0) and np.array_equal(rd("tpos"), np.arange(31) / 1000.0 * 2.5)
    assert np.array_equal(rd("sp").reshape(sp.shape), sp) and np.array_equal(rd("ap").reshape(ap.shape), ap)
    assert np.array_equal(rd("x"), wb.wavread(os.path.join(d, "our_x.wav"))[0])


def test_cpp_shims_are_source_compatible(tmp_path):
    """oracle/refio.cpp is written against the reference's tools/*.hpp; it must compile unchanged against
    include/audioio.hpp + include/parameterio.hpp and produce the reference's bytes."""
    exe = str(tmp_path / "ourio")
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-w", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(ROOT, "oracle", "refio.cpp"), "-L" + os.path.join(ROOT, "world-class_b200"),
                        "-lworldb200", "-Wl,-rpath," + os.path.join(ROOT, "world-class_b200")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    gold = np.load(GOLD)
    tpos, f0, sp, ap, x = _maker().case_arrays(nd=7)
    d = str(tmp_path)
    for n, a in (("tpos", tpos), ("f0", f0), ("sp", sp), ("ap", ap), ("x", x)):
        np.ascontiguousarray(a, np.float64).tofile(os.path.join(d, n + ".f64"))
    subprocess.run([exe, "write", d, "16000", str(len(f0)), "64", "7", "5.0", str(len(x))], check=True)
    for n in ("ref_f0.bin", "ref_f0.txt", "ref_sp.bin", "ref_ap.bin", "ref_x.wav"):
        assert open(os.path.join(d, n), "rb").read() == gold["coded/" + n].tobytes(), n
    rr = subprocess.run([exe, "read", d, "ref"], check=True, capture_output=True, text=True)
    assert json.loads(rr.stdout.strip()) == json.loads(gold["coded/header"].tobytes().decode())
