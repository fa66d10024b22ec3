"""Per-kernel timing of the configs[2]-style batch (22.05 kHz / 5 s utterances) and of one such utterance alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import worldb200 as wb
from worldb200 import signals

fs, sec = 22050, 5.0
wb._check(wb.lib().wb_init(0), "wb_init")
xs = [torch.from_numpy(signals.synth_speech(fs, sec, seed=1000 + i)).cuda() for i in range(32)]
opts = dict(harvest_option=wb.HarvestOption(f0_floor=40.0, frame_period=5.0), cheaptrick_option=wb.CheapTrickOption(f0_floor=71.0),
            d4c_option=wb.D4COption(threshold=0.85))
pl = wb.Pipeline(fs, opts["harvest_option"], opts["cheaptrick_option"], opts["d4c_option"])
pl.set_fresh_rng(True)
n = xs[0].numel()
for _ in range(3):
    pl.run_dev(xs[0].data_ptr(), n)
wb.device_synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pl.run_dev(xs[0].data_ptr(), n)
wb.device_synchronize()
print("single utterance, eager, one pipeline: %.3f ms" % ((time.perf_counter() - t0) * 100))
wb.profile_reset(); wb.profile(True)
for _ in range(5):
    pl.run_dev(xs[0].data_ptr(), n)
res = wb.profile_results(); wb.profile(False)
for k, v in sorted(res.items(), key=lambda kv: -kv[1][0])[:14]:
    print("  %-26s %.4f ms" % (k, v[0] / 5))
for ns in (1, 2, 4, 8):
    bp = wb.BatchPipeline(fs, n_streams=ns, **opts)
    for _ in range(2):
        bp.run(xs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        bp.run(xs)
    torch.cuda.synchronize()
    print("batch of 32, %d streams: %.2f ms per batch" % (ns, (time.perf_counter() - t0) / 3 * 1e3))
