"""Where the end-to-end step goes: host wall time of each of the four class-API compute() calls (pinned caller
buffers, bench.py's workload) next to the raw PCIe copy times of the same byte counts.  Run on the GPU box:
    python profiles/micro/e2e_breakdown.py"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import worldb200 as wb  # noqa: E402


def main():
    fs, fp = bench.FS, bench.FRAME_PERIOD
    from worldb200 import signals
    wb._check(wb.lib().wb_init(0), "wb_init")
    x = signals.synth_speech(fs, bench.SECONDS, seed=0)
    hopt = wb.HarvestOption(f0_floor=40.0, f0_ceil=800.0, frame_period=fp)
    harvest, cheaptrick = wb.Harvest(fs, hopt), wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0))
    d4c = wb.D4C(fs, wb.D4COption(threshold=0.85))
    synthesis = wb.Synthesis(fs, cheaptrick.fft_size, fp)
    L = harvest.getSamples(fs, len(x))
    bins = cheaptrick.fft_size // 2 + 1
    ny = wb.synthesis_length(L, fp, fs)

    def pinned(*shape):
        return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()

    xp = pinned(len(x)); xp[:] = x
    h_tpos, h_f0, h_sp, h_ap, h_y = pinned(L), pinned(L), pinned(L, bins), pinned(L, bins), pinned(ny)
    calls = [("harvest", lambda: harvest.compute(xp, h_tpos, h_f0)),
             ("cheaptrick", lambda: cheaptrick.compute(xp, h_tpos, h_f0, h_sp)),
             ("d4c", lambda: d4c.compute(xp, h_tpos, h_f0, cheaptrick.fft_size, h_ap)),
             ("synthesis", lambda: synthesis.compute(h_f0, h_sp, h_ap, ny, h_y))]
    for _ in range(5):
        for _, f in calls:
            f()
    reps = 30
    acc = {k: 0.0 for k, _ in calls}
    for _ in range(reps):
        for k, f in calls:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f()
            acc[k] += (time.perf_counter() - t0) * 1e3
    out = {"calls_ms": {k: v / reps for k, v in acc.items()}}
    out["calls_ms"]["sum"] = sum(out["calls_ms"].values())
    # raw copies of the same sizes
    d = torch.empty(L * bins, dtype=torch.float64, device="cuda")
    hp = torch.from_numpy(h_sp.reshape(-1))
    s = torch.cuda.Stream()
    for name, fn in (("d2h_matrix", lambda: hp.copy_(d, non_blocking=True)), ("h2d_matrix", lambda: d.copy_(hp, non_blocking=True))):
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            s.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            s.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / 10
        out[name] = {"ms": ms, "GB/s": L * bins * 8 / ms / 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
