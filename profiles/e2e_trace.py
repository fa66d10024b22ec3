"""Host-side breakdown of the end-to-end class-API step (BASELINE configs[1]); run with WB_TRACE=1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import worldb200 as wb
from worldb200 import signals

fs = 48000
x = signals.synth_speech(fs, 10.0, seed=0)
wb._check(wb.lib().wb_init(0), "wb_init")
xp = torch.from_numpy(x).pin_memory().numpy()
hopt = wb.HarvestOption(f0_floor=40.0, frame_period=5.0)
hv, ct, d4 = wb.Harvest(fs, hopt), wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0)), wb.D4C(fs, wb.D4COption(threshold=0.85))
sy = wb.Synthesis(fs, ct.fft_size, 5.0)
L = hv.getSamples(fs, len(x)); bins = ct.fft_size // 2 + 1; ny = len(x)
pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
if os.environ.get("PAGEABLE"):
    pin = lambda *shape: np.empty(shape)
tp, f0, sp, ap, y = pin(L), pin(L), pin(L, bins), pin(L, bins), pin(ny)
for it in range(6):
    t = [time.perf_counter()]
    hv.compute(xp, tp, f0); t.append(time.perf_counter())
    ct.compute(xp, tp, f0, sp); t.append(time.perf_counter())
    d4.compute(xp, tp, f0, ct.fft_size, ap); t.append(time.perf_counter())
    sy.compute(f0, sp, ap, ny, y); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("iter %d: harvest %.3f cheaptrick %.3f d4c %.3f synthesis %.3f total %.3f ms" % (it, d[0], d[1], d[2], d[3], d.sum()), file=sys.stderr)
