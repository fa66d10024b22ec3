"""Generates tests/golden/io_files.npz: the bytes of parameter files and a wav file WRITTEN BY THE REFERENCE's own
tools (oracle/_ref/refio = tools/parameterio.cpp + tools/audioio.cpp compiled where they lie), the arrays they
were written from, and what the reference's own readers return for wav files of 8 / 16 / 24 / 32 bit.
Run from the repo root (needs /root/reference):

    python tests/golden/make_io_golden.py
"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REFIO = os.path.join(ROOT, "oracle", "_ref", "refio")


def case_arrays(seed=5, L=23, fft=64, nd=0, nx=777):
    rng = np.random.default_rng(seed)
    dims = nd if nd else fft // 2 + 1
    tpos = np.arange(L) * 0.005
    f0 = np.where(rng.random(L) < 0.3, 0.0, 80 + 300 * rng.random(L))
    sp = np.exp(rng.normal(size=(L, dims)) * 3 - 8)
    ap = rng.uniform(0.001, 0.999, size=(L, dims))
    x = np.clip(rng.normal(size=nx) * 0.4, -1.3, 1.3)     # a few samples clip in wavwrite
    x[:4] = [1.0, -1.0, 0.99999, -1.00002]
    return tpos, f0, sp, ap, x


def wav_bytes(samples, nbit, fs, extra_chunk=False):
    """A mono PCM wav of `nbit` per sample; optionally a LIST chunk containing a stray 'd' before "data"."""
    qb = nbit // 8
    body = b"".join(int(v).to_bytes(qb, "little", signed=True) for v in samples)
    fmt = struct.pack("<IHHIIHH", 16, 1, 1, fs, fs * qb, qb, nbit)
    extra = (b"LIST" + struct.pack("<I", 10) + b"dadat dxyz") if extra_chunk else b""
    riff = b"WAVE" + b"fmt " + fmt + extra + b"data" + struct.pack("<I", len(body)) + body
    return b"RIFF" + struct.pack("<I", len(riff)) + riff


def main():
    out = {}
    for tag, nd in (("full", 0), ("coded", 7)):
        tpos, f0, sp, ap, x = case_arrays(nd=nd)
        L, fft, fs, fp = len(f0), 64, 16000, 5.0
        with tempfile.TemporaryDirectory() as d:
            for n, a in (("tpos", tpos), ("f0", f0), ("sp", sp), ("ap", ap), ("x", x)):
                np.ascontiguousarray(a, np.float64).tofile(os.path.join(d, n + ".f64"))
            subprocess.run([REFIO, "write", d, str(fs), str(L), str(fft), str(nd), repr(fp), str(len(x))], check=True)
            for n in ("ref_f0.bin", "ref_f0.txt", "ref_sp.bin", "ref_ap.bin", "ref_x.wav"):
                out["%s/%s" % (tag, n)] = np.frombuffer(open(os.path.join(d, n), "rb").read(), dtype=np.uint8)
            r = subprocess.run([REFIO, "read", d, "ref"], check=True, capture_output=True, text=True)
            out["%s/header" % tag] = np.frombuffer(r.stdout.strip().encode(), dtype=np.uint8)
            out["%s/x_read" % tag] = np.fromfile(os.path.join(d, "ref_x.rd.f64"))
            out["%s/tpos_read" % tag] = np.fromfile(os.path.join(d, "ref_tpos.rd.f64"))
    # the reference's wavread on 8 / 16 / 24 / 32-bit files (and one with an extra chunk before "data")
    rng = np.random.default_rng(9)
    for nbit in (8, 16, 24, 32):
        lim = 1 << (nbit - 1)
        s = rng.integers(-lim, lim, size=301)
        s[:3] = [-lim, lim - 1, 0]
        for extra in (False, True):
            raw = wav_bytes(s, nbit, 22050, extra)
            with tempfile.TemporaryDirectory() as d:
                open(os.path.join(d, "w_x.wav"), "wb").write(raw)
                # the read command wants the parameter files too: give it the smallest valid ones
                tpos, f0, sp, ap, x = case_arrays(L=2, nx=4)
                for n, a in (("tpos", tpos), ("f0", f0), ("sp", sp), ("ap", ap), ("x", x)):
                    np.ascontiguousarray(a, np.float64).tofile(os.path.join(d, n + ".f64"))
                subprocess.run([REFIO, "write", d, "16000", "2", "64", "0", "5.0", "4"], check=True)
                for n in ("f0.bin", "sp.bin", "ap.bin"):
                    os.rename(os.path.join(d, "ref_" + n), os.path.join(d, "w_" + n))
                r = subprocess.run([REFIO, "read", d, "w"], check=True, capture_output=True, text=True)
                key = "wav%d%s" % (nbit, "x" if extra else "")
                out[key + "/bytes"] = np.frombuffer(raw, dtype=np.uint8)
                out[key + "/x_read"] = np.fromfile(os.path.join(d, "w_x.rd.f64"))
                out[key + "/header"] = np.frombuffer(r.stdout.strip().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "io_files.npz"), **out)
    print("wrote io_files.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
