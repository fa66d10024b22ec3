#!/usr/bin/env python
"""bench.py -- headline benchmark of the WORLD hot path on B200.

Metric (BASELINE.json): frames/s (and x real-time) of the full Harvest -> CheapTrick -> D4C ->
Synthesis chain at 48 kHz / 5 ms frame period.  Workload = BASELINE.json configs[1]: one 10 s
48 kHz utterance (2001 frames) per GPU per step, synthetic sinusoid-plus-noise speech
(SURVEY.md section 8d).  With N GPUs every rank analyses/synthesises its own utterance (the path
shards by utterance with no data-path collective): weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` is timed with the inputs already resident in HBM;
`e2e` goes through the reference-facing class API (four compute() calls, HOST buffers, H2D/D2H
inside the timed region); `roofline` is the dominant kernel's algorithmic FFT-I/O bytes over its
live CUDA-event duration; `cpu_baseline` is the reference's own OpenMP build on this host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
SECONDS = 10.0
FRAME_PERIOD = 5.0
METRIC = "frames/sec, full Harvest->CheapTrick->D4C->Synthesis @48kHz/5ms"
WORKLOAD = "single 10 s utterance @48 kHz, 5 ms frame period, FFT size 2048 (BASELINE configs[1])"


def bench_config(frames):
    """The `config` object of the JSON line: identical in both arms (ours / --impl reference)."""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": frames, "utterances_per_step_per_gpu": 1,
            "parallelism": "one utterance per GPU, no data-path collective",
            "l2": "256 MiB device memset between timed iterations (flush); inputs 3.8 MB"}


def r2c_bytes(n):  # fp64 input + output of one real transform
    return 8 * n + 16 * (n // 2 + 1)


def c2c_bytes(n):
    return 32 * n


def r2c_flops(n):  # butterflies of one real transform (honesty note ii: the second bound is fp64)
    return 2.5 * n * np.log2(n)


def c2c_flops(n):
    return 5.0 * n * np.log2(n)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in r.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_run(x, repeat, omp=True):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference) on this host."""
    from oracle import refbin
    if not refbin.available(omp=omp):
        return None
    _, timings = refbin.run_reference(x, FS, stages="hcds", omp=omp, repeat=repeat, write=False)
    per = [t["harvest_ms"] + t["cheaptrick_ms"] + t["d4c_ms"] + t["synthesis_ms"] for t in timings]
    return {"ms": per, "threads": timings[0]["threads"], "frames": timings[0]["f0_length"], "stages": timings}


def run_reference_arm(args):
    """--impl reference: the reference's OpenMP build, all host threads, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import worldb200  # noqa: F401
    from worldb200 import signals
    x = signals.synth_speech(FS, SECONDS, seed=0)
    res = cpu_reference_run(x, repeat=args.warmup + args.steps, omp=True)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/refrun_omp is not built (run make -C oracle where /root/reference exists)"}))
        return 0
    ms = res["ms"][args.warmup:]
    ms_per_step = float(np.mean(ms))
    fps = res["frames"] / (ms_per_step / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "x_realtime": SECONDS / (ms_per_step / 1e3),
        "config": bench_config(res["frames"]),
        "impl_note": "reference OpenMP build (Makefile flags -O3 -mavx -fopenmp) of /root/reference, one utterance per step, host cores only",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": res["threads"], "kind": "reference",
                         "sample": "%d x one 10 s / 48 kHz utterance through refrun_omp (all host threads)" % args.steps},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (development)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    # one JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION prints to stdout) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the worldb200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import worldb200 as wb
    from worldb200 import parallel, signals
    # one process per GPU, on the CPUs of the GPU's NUMA node (before any page-locked buffer is allocated)
    cpus_at_start = os.sched_getaffinity(0)
    host_binding = parallel.bind_to_gpu_node(local_rank)
    wb._check(wb.lib().wb_init(local_rank), "wb_init")
    lib_stream = torch.cuda.ExternalStream(wb.stream_handle())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: every rank gets its own utterance (seed = rank)
    x_host = signals.synth_speech(FS, SECONDS, seed=rank)
    x_pinned = torch.from_numpy(x_host).pin_memory()
    d_x = x_pinned.cuda()
    hopt = wb.HarvestOption(f0_floor=40.0, frame_period=FRAME_PERIOD)     # test/test.cpp:83-87
    copt = wb.CheapTrickOption(f0_floor=71.0)                             # test/test.cpp:130
    dopt = wb.D4COption(threshold=0.85)                                   # test/test.cpp:181
    pl = wb.Pipeline(FS, hopt, copt, dopt)
    pl.set_graph(True)   # the ~45 launches of one step are replayed as one CUDA graph
    n = len(x_host)
    L, ny, fft_size = pl.f0_length(n), pl.out_length(n), pl.fft_size
    bins = fft_size // 2 + 1
    d_y = torch.empty(ny, dtype=torch.float64, device="cuda")
    d_f0 = torch.empty(L, dtype=torch.float64, device="cuda")
    d_tpos = torch.empty(L, dtype=torch.float64, device="cuda")
    d_sp = torch.empty((L, bins), dtype=torch.float64, device="cuda")
    d_ap = torch.empty((L, bins), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_resident():
        pl.run_dev(d_x.data_ptr(), n, d_y=d_y.data_ptr(), y_length=ny, d_tpos=d_tpos.data_ptr(), d_f0=d_f0.data_ptr(),
                   d_sp=d_sp.data_ptr(), d_ap=d_ap.data_ptr())

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step_resident()
    wb.device_synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = wb.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the timed span)
        torch.cuda.synchronize()
        ev[k][0].record(lib_stream)
        step_resident()
        ev[k][1].record(lib_stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = wb.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * L / (ms_per_step / 1e3)

    # ---- end to end through the reference-facing class API with HOST buffers (`e2e`)
    harvest = wb.Harvest(FS, hopt)
    cheaptrick = wb.CheapTrick(FS, copt)
    d4c = wb.D4C(FS, dopt)
    synthesis = wb.Synthesis(FS, cheaptrick.fft_size, FRAME_PERIOD)
    x_np = x_pinned.numpy()

    # caller-owned host buffers, allocated once like test/test.cpp:102-104,146-149,170-173 -- page-locked
    # (torch pin_memory), each matrix one block: the library then DMAs straight from / into them
    def pinned(*shape):
        return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()

    h_tpos, h_f0 = pinned(L), pinned(L)
    h_sp, h_ap = pinned(L, bins), pinned(L, bins)
    h_y = pinned(ny)

    def step_e2e():
        harvest.compute(x_np, h_tpos, h_f0)
        cheaptrick.compute(x_np, h_tpos, h_f0, h_sp)
        d4c.compute(x_np, h_tpos, h_f0, cheaptrick.fft_size, h_ap)
        return synthesis.compute(h_f0, h_sp, h_ap, ny, h_y)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e2e_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y_host = step_e2e()                # synchronous: returns after the D2H of the waveform
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * L / (e2e_ms / args.steps / 1e3)
    # ---- the same end to end through ONE call (wb_pipeline_run: the resident chain, page-locked outputs downloaded
    # inside the chain as their stages finish) -- reported beside `e2e`, which stays the reference-shaped class API
    one_call = wb.Pipeline(FS, hopt, copt, dopt)
    one_call.set_graph(True)
    out_bufs = dict(tpos=h_tpos, f0=h_f0, sp=h_sp, ap=h_ap, y=h_y)
    for _ in range(max(args.warmup, 3)):
        one_call.run(x_np, out=out_bufs)
    barrier()
    one_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one_call.run(x_np, out=out_bufs)
        one_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([one_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        one_ms = float(t.item())
    e2e_one_call = {"value": world * L / (one_ms / args.steps / 1e3), "unit": "frames/s", "ms_per_step": one_ms / args.steps,
                    "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * (2 * L + 2 * L * bins + ny),
                    "path": "wb_pipeline_run (one call: x up, tpos / f0 / sp / ap / y down into page-locked caller buffers)"}
    del one_call
    h2d = 8 * (3 * n + 2 * 2 * L + L + 2 * L * bins)      # x x3, (tpos,f0) x2, f0, sp+ap rows for Synthesis
    d2h = 8 * (2 * L + 2 * L * bins + ny)                 # tpos,f0; sp; ap; y
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # ---- per-kernel live timing for the roofline line (separate pass, CUDA events per launch)
    roofline = None
    kernel_table = {}
    if rank == 0:
        wb.profile_reset()
        wb.profile(True)
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            step_resident()
        kernel_table = wb.profile_results()
        wb.profile(False)
        f0_h = d_f0.cpu().numpy()
        ap_h = d_ap.cpu().numpy()
        voiced_lt = int(np.sum(f0_h != 0))
        voiced_body = int(np.sum(np.any(ap_h != 1.0 - 1e-12, axis=1)))
        n_ap = wb.lib().wb_get_number_of_aperiodicities(FS)
        n_d4c = 1 << int(np.floor(np.log2(4.0 * FS / 47.0 + 1)) + 1)
        n_lt = 1 << int(np.floor(np.log2(3.0 * FS / 40.0 + 1)) + 1)
        # Harvest: counts from the last run's device buffers
        r = min(max(int(FS / 8000.0 + 0.5), 1), 12)
        afs = FS / r
        Lb = int(1000.0 * n / FS / 1) + 1
        nch = 1 + int(np.log(800.0 * 1.1 / (40.0 * 0.9)) / 0.69314718055994529 * 40.0)
        mc = int(nch // 10) * 7
        own_cap = mc // 7
        nc = int(pl.debug_read("hv_nc", (4,), dtype=np.int32)[0])
        own = pl.debug_read("hv_own", (Lb, own_cap))[:, :max(nc, 1)]
        # overlapF0Candidates replicates every own candidate to frames k-3..k+3 (harvest.cpp:987-1000)
        refine_bytes = 0
        refine_flops = 0.0
        f = own[own > 0]
        if f.size:
            hw = (1.5 * afs / f + 1.0).astype(np.int64)
            n_i = 1 << (2 + np.floor(np.log2(2 * hw + 1)).astype(np.int64))
            per = 2 * (8 * n_i + 16 * (n_i // 2 + 1))
            refine_bytes = int(7 * per.sum())   # edge frames lose a few copies: <0.1 %
            refine_flops = float(7 * (2 * 2.5 * n_i * np.log2(n_i)).sum())
        y_len = 1 + n // r
        n_h = 1 << int(np.floor(np.log2(y_len + 4 * int(1.0 + afs / (40.0 * 0.9 * 2 ** (1 / 40.0)) / 2.0))) + 1)
        n_pulses = int(pl.debug_read("syn_np", (1,), dtype=np.int32)[0])
        alg = {
            "ct_frame_kernel": 3 * L * r2c_bytes(fft_size),
            "lt_frame_kernel": voiced_lt * r2c_bytes(n_lt),
            "d4c_body_kernel": voiced_body * (5 + n_ap) * r2c_bytes(n_d4c),
            "channel_kernel": 2 * nch * r2c_bytes(n_h),          # reference: r2c + c2r of N_h per channel
            "yspec_kernel": r2c_bytes(n_h),
            "refine_kernel": refine_bytes,
            "response_kernel": n_pulses * (5 * r2c_bytes(fft_size) + 2 * c2c_bytes(fft_size)),
        }
        alg_flops = {
            "ct_frame_kernel": 3 * L * r2c_flops(fft_size),
            "lt_frame_kernel": voiced_lt * r2c_flops(n_lt),
            "d4c_body_kernel": voiced_body * (5 + n_ap) * r2c_flops(n_d4c),
            "channel_kernel": 2 * nch * r2c_flops(n_h),
            "yspec_kernel": r2c_flops(n_h),
            "refine_kernel": refine_flops,
            "response_kernel": n_pulses * (5 * r2c_flops(fft_size) + 2 * c2c_flops(fft_size)),
        }
        fp64_peak = None
        try:
            fp64_peak = wb.measure_fp64_peak()      # TFLOP/s, measured on this GPU now (MEASURED_PEAKS.json has no fp64 entry)
        except Exception:
            pass
        dom = max(kernel_table.items(), key=lambda kv: kv[1][0])[0] if kernel_table else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_stale = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj.get(dom)
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            from summarize import csrc_sha1
            # the ncu capture the figure comes from was made on other kernel sources than the ones timed here
            traffic_stale = tj.get("_csrc_sha1_by_kernel", {}).get(dom) != csrc_sha1(dom)
        except Exception:
            pass
        if dom is not None:
            tot_ms, cnt = kernel_table[dom]
            avg_ms = tot_ms / max(cnt, 1)
            a_bytes = alg.get(dom)
            achieved = (a_bytes / (avg_ms / 1e3) / 1e9) if a_bytes else None
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_stale": traffic_stale,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                        "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": a_bytes,
                        "kernel_share_of_step": tot_ms / max(sum(v[0] for v in kernel_table.values()), 1e-9),
                        "algorithmic_bytes_per_step_all_kernels": int(sum(alg.values())),
                        "whole_chain_frac": (sum(alg.values()) / (ms_per_step / 1e3) / 1e9) / peak}
            if fp64_peak:
                f_dom = float(alg_flops.get(dom) or 0.0)
                f_all = float(sum(alg_flops.values()))
                roofline["fp64"] = {
                    "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured now: independent DFMA chains on every SM (wb_measure_fp64_peak)",
                    "achieved": f_dom / (avg_ms / 1e3) / 1e12, "frac": f_dom / (avg_ms / 1e3) / 1e12 / fp64_peak,
                    "whole_chain_achieved": f_all / (ms_per_step / 1e3) / 1e12,
                    "whole_chain_frac": f_all / (ms_per_step / 1e3) / 1e12 / fp64_peak,
                    "algorithmic_gflop_per_step": f_all / 1e9,
                    "model": "butterflies of the transforms the ALGORITHM executes: 2.5 N log2 N per real, 5 N log2 N per complex transform (SURVEY.md 8d)"}

    # ---- extra (not the headline): BASELINE configs[2] -- 256 independent 22.05 kHz / 5 s utterances on this GPU
    # (each utterance == one reference process; 32 distinct signals, each used eight times: the generator costs
    # 0.35 s of host time per utterance)
    batch_extra = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            n_utt, n_distinct, bfs, bsec, n_streams, reps = 256, 32, 22050, 5.0, 16, 6
            distinct = [torch.from_numpy(signals.synth_speech(bfs, bsec, seed=1000 + i)).cuda() for i in range(n_distinct)]
            bxs = [distinct[i % n_distinct] for i in range(n_utt)]
            bp = wb.BatchPipeline(bfs, n_streams=n_streams, harvest_option=wb.HarvestOption(f0_floor=40.0, frame_period=FRAME_PERIOD),
                                  cheaptrick_option=copt, d4c_option=dopt)
            for _ in range(2):
                outs = bp.run(bxs)
            torch.cuda.synchronize()
            frames = sum(int(o["f0"].numel()) for o in outs)
            del outs
            bms = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs = bp.run(bxs)
                e1.record()
                torch.cuda.synchronize()
                bms.append(e0.elapsed_time(e1))
                del outs
            med = float(np.median(bms))
            batch_extra = {"workload": "%d x %.0f s utterances @%d Hz (%d distinct signals), %d concurrent pipelines, graph replay, results copied out, 1 GPU"
                                       % (n_utt, bsec, bfs, n_distinct, n_streams),
                           "ms_per_batch": med, "ms_per_batch_all": [round(v, 3) for v in bms],
                           "spread": (max(bms) - min(bms)) / med, "frames_per_s": frames / (med / 1e3),
                           "x_realtime": n_utt * bsec / (med / 1e3)}
            del bp, bxs, distinct
        except Exception as exc:  # the headline line must not depend on the extra
            batch_extra = {"error": repr(exc)}

    # ---- extra (not the headline): BASELINE configs[4] -- codec round trip (codec.cpp:216-325) of 1024 frames at
    # 48 kHz / fft 2048, 60 mel-cepstral dimensions, device resident; HBM roofline with SURVEY 8d's bytes per frame
    codec_extra = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            from worldb200 import tensors as wt
            n_ap_c = wb.lib().wb_get_number_of_aperiodicities(FS)
            res = {}
            for n_fr in (1024, 65536):
                g = torch.Generator(device="cuda").manual_seed(5)
                sp_c = torch.exp(torch.rand((n_fr, bins), dtype=torch.float64, device="cuda", generator=g) * 18.0 - 16.0)  # log-uniform
                ap_c = torch.rand((n_fr, bins), dtype=torch.float64, device="cuda", generator=g) * 0.998 + 0.001

                def round_trip():
                    with torch.cuda.stream(lib_stream):
                        c_sp = wt.codec("code_sp", sp_c, FS, fft_size, 60)
                        d_sp_c = wt.codec("decode_sp", c_sp, FS, fft_size, 60)
                        c_ap = wt.codec("code_ap", ap_c, FS, fft_size)
                        d_ap_c = wt.codec("decode_ap", c_ap, FS, fft_size)
                    return d_sp_c, d_ap_c

                for _ in range(3):
                    round_trip()
                torch.cuda.synchronize()
                reps = 20 if n_fr <= 4096 else 5
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(lib_stream)
                for _ in range(reps):
                    round_trip()
                e1.record(lib_stream)
                torch.cuda.synchronize()
                cms = e0.elapsed_time(e1) / reps
                bytes_per_frame = 2 * (2 * 8 * bins + 8 * (60 + n_ap_c))
                gbs = n_fr * bytes_per_frame / (cms / 1e3) / 1e9
                peak_c = 6538.0
                try:
                    peak_c = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6538.0))
                except Exception:
                    pass
                res[str(n_fr)] = {"ms_per_round_trip": cms, "frames_per_s": n_fr / (cms / 1e3), "algorithmic_GB_per_s": gbs,
                                  "roofline_frac_hbm": gbs / peak_c}
                del sp_c, ap_c
            codec_extra = {"workload": "CodeSpectralEnvelope -> Decode -> CodeAperiodicity -> Decode, 1025 bins, 60 dimensions, %d band aperiodicities; 1024 frames (configs[4]) and 65536 frames (bandwidth regime)" % n_ap_c,
                           "bytes_per_frame": 2 * (2 * 8 * bins + 8 * (60 + n_ap_c)), "by_frames": res}
        except Exception as exc:
            codec_extra = {"error": repr(exc)}

    # ---- extra (not the headline): BASELINE configs[3] -- one continuous 48 kHz stream, frames sharded over ALL
    # ranks (strong scaling: the same 600 s at every N) with the exact exchange steps of worldb200/parallel.py; plus a
    # 30 s stream through the same sharded path checked against ONE reference process on rank 0
    stream_extra = None
    try:
        if args.no_extras:
            raise RuntimeError("skipped (--no-extras)")
        import hashlib
        from worldb200 import parallel
        from oracle import refbin
        total_shards = 16
        k_shards = max(1, total_shards // world)
        s_seconds = 600
        block = signals.synth_speech(FS, 60.0, seed=0)
        sx = torch.from_numpy(np.tile(block, s_seconds // 60)).cuda()
        keep, best, best_t = {}, None, None
        for rep in range(3):               # first repetition = warm-up (allocations, plan tables)
            barrier()
            t = {}
            so = parallel.process_stream_exact(sx, FS, hopt, copt, dopt, segment_seconds=120, halo_seconds=2,
                                               shards_per_rank=k_shards, keep_rows=False, timings=t, state=keep)
            tot = torch.tensor([t["total"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            if rep > 0 and (best is None or float(tot.item()) < best):
                best, best_t = float(tot.item()), dict(t)
        sha_f0 = hashlib.sha1(so["f0"].cpu().numpy().tobytes()).hexdigest()
        sha_y = hashlib.sha1(so["y"].cpu().numpy().tobytes()).hexdigest()
        s_frames = so["plan"].f0_length
        del sx, so, keep
        # parity of the sharded path: 30 s through all ranks vs one reference process (serial build: its waveform is
        # reproducible, the OpenMP build's is not)
        cx = block[:30 * FS].copy()
        co = parallel.process_stream_exact(torch.from_numpy(cx).cuda(), FS, hopt, copt, dopt, segment_seconds=5, halo_seconds=2,
                                           shards_per_rank=max(1, 6 // world), keep_rows=True)
        check = None
        if rank == 0 and refbin.available():
            ref, _ = refbin.run_reference(cx, FS, stages="hcds")
            f0c, yc = co["f0"].cpu().numpy(), co["y"].cpu().numpy()
            v = ref["f0"] > 0
            fb_c, fe_c = co["frames"]
            check = {"seconds": 30, "voicing_identical": bool(np.array_equal(f0c > 0, v)),
                     "f0_rel": float(np.max(np.abs(f0c[v] - ref["f0"][v]) / ref["f0"][v])),
                     "sp_rel_rank0_rows": float(np.max(np.abs(co["sp"].cpu().numpy() - ref["sp"][fb_c:fe_c]) / ref["sp"][fb_c:fe_c])),
                     "ap_rel_rank0_rows": float(np.max(np.abs(co["ap"].cpu().numpy() - ref["ap"][fb_c:fe_c]) / ref["ap"][fb_c:fe_c])),
                     "y_over_peak": float(np.max(np.abs(yc - ref["y"])) / np.abs(ref["y"]).max())}
        del co
        if rank == 0:
            exposed = sum(best_t.get(k2, 0.0) for k2 in ("gather_f0", "wait_ap0", "stitch_tail"))
            stream_extra = {"workload": "one continuous %d s stream @%d Hz, exact sharding over %d GPU(s), %d shards per rank (strong scaling: same stream at every N)"
                                        % (s_seconds, FS, world, k_shards),
                            "n_gpus": world, "ms": best, "frames_per_s": s_frames / (best / 1e3), "x_realtime": s_seconds / (best / 1e3),
                            "phases_ms_rank0": {k2: round(v2, 3) for k2, v2 in best_t.items()},
                            "exposed_collective_ms_rank0": exposed, "exposed_collective_frac": exposed / best_t["total"],
                            "sha1_f0": sha_f0, "sha1_y": sha_y, "check_vs_reference": check}
    except Exception as exc:  # the headline line must not depend on the extra
        if rank == 0:
            stream_extra = {"error": repr(exc)}

    # ---- CPU baseline: the reference's OpenMP build on this host, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # (N = 1 only, as the contract says)
        os.sched_setaffinity(0, cpus_at_start)    # the reference's OpenMP build gets every host core
        res = cpu_reference_run(x_host, repeat=3, omp=True)
        if res is not None:
            ms = float(np.median(res["ms"]))
            cpu_baseline = {"value": res["frames"] / (ms / 1e3), "unit": "frames/s", "cores": res["threads"],
                            "kind": "reference", "ms_per_utterance": ms,
                            "sample": "3 x the same 10 s / 48 kHz utterance through oracle/_ref/refrun_omp (reference OpenMP build), median"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "x_realtime": world * SECONDS / (ms_per_step / 1e3),
            "config": bench_config(L),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "x_realtime": world * SECONDS / (e2e_ms / args.steps / 1e3),
                    "path": "Harvest/CheapTrick/D4C/Synthesis compute() with host buffers (caller-owned page-locked x, f0, sp, ap, y), 4 calls per step"},
            "e2e_one_call": e2e_one_call,
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline, "host_binding": host_binding,
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in sorted(kernel_table.items(), key=lambda kv: -kv[1][0])},
            "wall_s_timed_region": t_wall,
            "batch_config3_extra": batch_extra,
            "stream_config4_extra": stream_extra,
            "codec_config5_extra": codec_extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
