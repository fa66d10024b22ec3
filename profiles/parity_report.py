"""Worst-case deviations of the CUDA chain from the reference's serial CPU build (oracle/_ref) on BASELINE
configs[0] (16 kHz / 1 s) and configs[1] (48 kHz / 10 s).  Prints one line per configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import worldb200 as wb
from worldb200 import signals
from oracle import refbin

wb._check(wb.lib().wb_init(0), "wb_init")
for fs, sec in ((16000, 1.0), (48000, 10.0)):
    x = signals.synth_speech(fs, sec, seed=0)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    wb.randn_reseed()
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85))
    out = pl.run(x)
    v = ref["f0"] > 0
    print("fs %d, %.0f s: frames %d, voiced %d, voicing identical %s, f0 rel %.2e, sp rel %.2e, ap rel %.2e, y / peak %.2e" % (
        fs, sec, len(v), int(v.sum()), bool(np.array_equal(out["f0"] > 0, v)),
        np.max(np.abs(out["f0"][v] - ref["f0"][v]) / ref["f0"][v]),
        np.max(np.abs(out["sp"] - ref["sp"]) / ref["sp"]), np.max(np.abs(out["ap"] - ref["ap"]) / ref["ap"]),
        np.max(np.abs(out["y"] - ref["y"])) / np.abs(ref["y"]).max()))
