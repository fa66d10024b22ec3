"""BASELINE configs[3]: one continuous 48 kHz stream, frames sharded over the GPUs of one box with the exact
exchange steps of world-class_b200/parallel.py (f0 all-gather, Love Train decisions all-gather, waveform stitch).

    python profiles/stream_bench.py --seconds 600                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/stream_bench.py --seconds 600                          # 2 GPUs (strong scaling: same stream)

Prints one JSON line: frames/s and x real time of the whole job (max over ranks, device-synchronised phases),
the per-phase milliseconds of rank 0 and SHA-1 digests of f0 and of the stitched waveform -- the digests of runs
with different world sizes must agree from the f0 gather on (given the same f0 the rest is bit-identical by
construction; Harvest itself is cut at different places)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--fs", type=int, default=48000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--shards-per-rank", type=int, default=1)
    ap.add_argument("--segment-seconds", type=int, default=120)
    ap.add_argument("--pipelines", type=int, default=1, help="groups of the rank's shards in flight (one pipeline + stream each)")
    ap.add_argument("--profile", action="store_true", help="one extra run with per-kernel CUDA-event timing (stderr)")
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import worldb200 as wb
    from worldb200 import parallel, signals
    wb._check(wb.lib().wb_init(local_rank), "wb_init")
    # long streams: a 60 s synthetic block repeated (the generator costs ~0.2 s of host time per second of audio)
    block = signals.synth_speech(args.fs, min(args.seconds, 60.0), seed=0)
    n_total = int(round(args.seconds * args.fs))
    x = np.tile(block, -(-n_total // len(block)))[:n_total].copy()
    d_x = torch.from_numpy(x).cuda()
    hopt = wb.HarvestOption(f0_floor=40.0, frame_period=5.0)
    copt, dopt = wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85)
    best, out, timings = None, None, None
    keep = {}     # pipeline objects / device workspaces live across the repetitions (steady state)
    for rep in range(args.reps + 1):          # first repetition = warm-up (allocations, plan tables)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = {}
        t0 = time.perf_counter()
        out = parallel.process_stream_exact(d_x, args.fs, hopt, copt, dopt, segment_seconds=args.segment_seconds, halo_seconds=2, pipelines_in_flight=args.pipelines,
                                            shards_per_rank=args.shards_per_rank, keep_rows=False, timings=t, state=keep)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([t["total"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0 and (best is None or float(ms.item()) < best):
            best, timings = float(ms.item()), dict(t, wall_ms=wall)
    if args.profile and rank == 0:
        wb.profile_reset()
        wb.profile(True)
        parallel.process_stream_exact(d_x, args.fs, hopt, copt, dopt, segment_seconds=args.segment_seconds, halo_seconds=2, pipelines_in_flight=args.pipelines,
                                      shards_per_rank=args.shards_per_rank, keep_rows=False, state=keep)
        table = wb.profile_results()
        wb.profile(False)
        for name, (ms, cnt) in sorted(table.items(), key=lambda kv: -kv[1][0])[:24]:
            sys.stderr.write("%-28s %10.3f ms %6d launches\n" % (name, ms, cnt))
    if rank == 0:
        plan = out["plan"]
        line = {"metric": "frames/sec, full Harvest->CheapTrick->D4C->Synthesis @48kHz/5ms", "workload":
                "one continuous %.0f s stream @%d Hz sharded over %d GPU(s) (BASELINE configs[3] shape), exact exchange steps" % (args.seconds, args.fs, world),
                "n_gpus": world, "scaling": "strong", "ms": best, "value": plan.f0_length / (best / 1e3), "unit": "frames/s",
                "x_realtime": args.seconds / (best / 1e3), "phases_ms_rank0": timings, "shards_per_rank": args.shards_per_rank,
                "sha1_f0": hashlib.sha1(out["f0"].cpu().numpy().tobytes()).hexdigest(),
                "sha1_y": hashlib.sha1(out["y"].cpu().numpy().tobytes()).hexdigest(),
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
