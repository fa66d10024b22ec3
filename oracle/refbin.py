"""TEST INFRASTRUCTURE: run the reference's own CPU implementation (oracle/_ref/refrun,
built from /root/reference by oracle/Makefile) on raw f64 arrays.

One call == one fresh process == one fresh randn() stream.  Only tests/, smoke() and
bench.py's CPU baseline may import this module.
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFRUN = os.path.join(HERE, "_ref", "refrun")
REFRUN_OMP = os.path.join(HERE, "_ref", "refrun_omp")
REFMOD = os.path.join(HERE, "_ref", "refmod")


def available(omp=False):
    return os.path.exists(REFRUN_OMP if omp else REFRUN)


def run_reference(x, fs, stages="hcds", f0=None, frame_period=5.0, harvest_f0_floor=40.0,
                  harvest_f0_ceil=800.0, ct_f0_floor=71.0, d4c_threshold=0.85, codec_nd=60,
                  omp=False, repeat=1, write=True, taskset=None):
    """Returns (outputs: dict name -> ndarray, timings: list of dict per repetition)."""
    exe = REFRUN_OMP if omp else REFRUN
    if not os.path.exists(exe):
        raise FileNotFoundError("%s missing: run `make -C oracle` where /root/reference exists" % exe)
    x = np.ascontiguousarray(x, dtype=np.float64)
    with tempfile.TemporaryDirectory(prefix="wbref_") as d:
        xin = os.path.join(d, "x.f64")
        x.tofile(xin)
        cmd = [exe, "--in", xin, "--fs", str(int(fs)), "--out", os.path.join(d, "o"), "--stages", stages,
               "--frame-period", repr(float(frame_period)), "--harvest-f0-floor", repr(float(harvest_f0_floor)),
               "--harvest-f0-ceil", repr(float(harvest_f0_ceil)), "--ct-f0-floor", repr(float(ct_f0_floor)),
               "--d4c-threshold", repr(float(d4c_threshold)), "--codec-nd", str(int(codec_nd)),
               "--repeat", str(int(repeat))]
        if not write:
            cmd.append("--no-write")
        if "h" not in stages:
            if f0 is None:
                raise ValueError("f0 is required when Harvest is not run")
            f0in = os.path.join(d, "f0.f64")
            np.ascontiguousarray(f0, dtype=np.float64).tofile(f0in)
            cmd += ["--f0-in", f0in]
        if taskset is not None:
            cmd = ["taskset", "-c", str(taskset)] + cmd
        env = dict(os.environ)
        if omp:
            # the OpenMP baseline uses every host thread (torchrun pins OMP_NUM_THREADS=1 for its workers)
            env.pop("OMP_NUM_THREADS", None)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("refrun failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        timings = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
        out = {}
        if write:
            info = timings[0]
            L, bins = info["f0_length"], info["fft_size"] // 2 + 1
            shapes = {"tpos": (L,), "f0": (L,), "sp": (L, bins), "ap": (L, bins), "y": (info["y_length"],),
                      "csp": (L, codec_nd), "cap": (L, max(info["n_ap"], 0)), "dsp": (L, bins), "dap": (L, bins)}
            for name, shape in shapes.items():
                p = os.path.join(d, "o." + name)
                if os.path.exists(p):
                    out[name] = np.fromfile(p, dtype=np.float64).reshape(shape)
            out["fft_size"] = info["fft_size"]
        return out, timings


def run_modification(f0, sp, fs, fft_size, shift=None, ratio=None):
    """The reference demo's ParameterModification (test/test.cpp:201-243) through oracle/_ref/refmod.
    shift / ratio = None leaves the argument off the demo's command line.  Returns (f0, sp)."""
    if not os.path.exists(REFMOD):
        raise FileNotFoundError("%s missing: run `make -C oracle` where /root/reference exists" % REFMOD)
    f0 = np.ascontiguousarray(f0, dtype=np.float64)
    sp = np.ascontiguousarray(sp, dtype=np.float64)
    L, bins = sp.shape
    assert bins == fft_size // 2 + 1 and len(f0) == L
    with tempfile.TemporaryDirectory(prefix="wbmod_") as d:
        pf, ps, po = os.path.join(d, "f0.f64"), os.path.join(d, "sp.f64"), os.path.join(d, "o")
        f0.tofile(pf)
        sp.tofile(ps)
        cmd = [REFMOD, pf, ps, str(int(fs)), str(int(fft_size)), str(L),
               "-" if shift is None else repr(float(shift)), "-" if ratio is None else repr(float(ratio)), po]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("refmod failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        return np.fromfile(po + ".f0", dtype=np.float64), np.fromfile(po + ".sp", dtype=np.float64).reshape(L, bins)
