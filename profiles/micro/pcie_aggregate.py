"""What the host link gives N ranks at once: every rank copies a page-locked matrix (bench.py's sp / ap size) to and
from its GPU in a loop, first alone (the others wait), then all together.  The per-rank rate under contention bounds
the end-to-end (`e2e`) step of bench.py at N GPUs: that step moves h2d_bytes_per_step + d2h_bytes_per_step per rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        profiles/micro/pcie_aggregate.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def rate(fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 2001 * 1025
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h2 = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    d2 = torch.empty(n, dtype=torch.float64, device="cuda")
    s2 = torch.cuda.Stream()

    def h2d():
        d.copy_(h, non_blocking=True)

    def d2h():
        h.copy_(d, non_blocking=True)

    def both():
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    nbytes = n * 8
    for name, fn, b in (("h2d", h2d, nbytes), ("d2h", d2h, nbytes), ("both_directions", both, 2 * nbytes)):
        alone = None
        for r in range(world):           # one rank at a time
            barrier()
            if r == rank and rank == 0:
                alone = rate(fn, b)
            barrier()
            if r == 0:
                break                    # (rank 0's solo rate is enough)
        barrier()
        together = rate(fn, b)           # every rank at once
        t = torch.tensor([together], dtype=torch.float64, device="cuda")
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            rates = [float(x.item()) for x in lst]
        else:
            rates = [together]
        out[name] = {"rank0_alone_GB_s": alone, "per_rank_together_GB_s": rates, "aggregate_GB_s": sum(rates)}
    if rank == 0:
        out["n_gpus"] = world
        out["matrix_bytes"] = nbytes
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
