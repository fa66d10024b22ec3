"""Runs a few device-resident whole-chain steps (BASELINE configs[1]) -- the command wrapped by ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import worldb200 as wb  # noqa: E402
from worldb200 import signals  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fs = 48000
x = signals.synth_speech(fs, 10.0, seed=0)
wb._check(wb.lib().wb_init(0), "wb_init")
d_x = torch.from_numpy(x).cuda()
pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0),
                 wb.D4COption(threshold=0.85))
for _ in range(steps):
    pl.run_dev(d_x.data_ptr(), len(x))
wb.device_synchronize()
print("done", wb.launch_count())
