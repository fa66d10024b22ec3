"""TEST INFRASTRUCTURE: run oracle/_ref/refdump_harvest and load its intermediate dumps."""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "_ref", "refdump_harvest")


def harvest_intermediates(x, fs, f0_floor=40.0, f0_ceil=800.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    with tempfile.TemporaryDirectory(prefix="wbdump_") as d:
        xin = os.path.join(d, "x.f64")
        x.tofile(xin)
        r = subprocess.run([EXE, xin, str(int(fs)), os.path.join(d, "o"), repr(float(f0_floor)), repr(float(f0_ceil))],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("refdump_harvest failed: " + r.stderr[-2000:])
        info = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
        Lb, nch, mc = info["Lb"], info["nch"], info["max_candidates"]
        out = dict(info)
        shapes = {"y": (info["y_length"],), "raw": (nch, Lb), "cand0": (Lb, mc), "cand1": (Lb, mc), "score1": (Lb, mc),
                  "cand2": (Lb, mc), "score2": (Lb, mc), "base": (Lb,), "step1": (Lb,), "step2": (Lb,),
                  "step3": (Lb,), "step4": (Lb,), "f0": (Lb,)}
        for name, shape in shapes.items():
            out[name] = np.fromfile(os.path.join(d, "o." + name), dtype=np.float64).reshape(shape)
        return out
