"""Sharding of the WORLD hot path over the GPUs of one box (one process per GPU).

Two cases:

* batches of independent utterances (BASELINE configs[2]): round-robin over ranks, no data-path
  collective;
* one long stream (BASELINE configs[3]): the stream is cut into segments that overlap by a halo,
  every rank analyses / re-synthesises its segments as independent utterances, the halos are
  discarded and ONE all-gather stitches the segment cores back into the output stream.  Every
  segment reproduces the reference run on that segment (fresh randn() stream per segment); the
  halo (default 1 s >> the longest analysis window of 3/40 s and Harvest's 300 ms contour
  padding) makes the cores independent of where the cuts fall.

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin assignment of item indices to `rank`."""
    return list(range(rank, n_items, world))


def plan_segments(n_samples, fs, segment_seconds=30.0, halo_seconds=1.0):
    """Cuts [0, n_samples) into segment cores of `segment_seconds`; returns a list of dicts
    core=(a, b) and padded=(a - halo, b + halo) clipped to the stream, sample indices."""
    seg = max(1, int(round(segment_seconds * fs)))
    halo = max(0, int(round(halo_seconds * fs)))
    out = []
    a = 0
    while a < n_samples:
        b = min(n_samples, a + seg)
        out.append({"core": (a, b), "padded": (max(0, a - halo), min(n_samples, b + halo))})
        a = b
    return out


def extract_core(y_padded, seg):
    """Drops the halos of a re-synthesised padded segment (lengths may differ by the frame grid:
    the synthesised length of a segment is floor-aligned to the frame period)."""
    (a, b), (pa, _pb) = seg["core"], seg["padded"]
    core = np.zeros(b - a, dtype=np.float64)
    src = y_padded[a - pa:a - pa + (b - a)]
    core[:len(src)] = src
    return core


def process_stream(x, fs, process_segment, segment_seconds=30.0, halo_seconds=1.0, group=None):
    """Shards a long stream over the ranks of `group` and returns the stitched output on every rank.

    process_segment(x_padded) -> y_padded must map a padded segment to its re-synthesis (same
    sample grid).  Works without torch.distributed being initialised (single process).
    """
    try:
        import torch
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    except Exception:  # pragma: no cover
        distributed = False
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    segs = plan_segments(len(x), fs, segment_seconds, halo_seconds)
    mine = shard_indices(len(segs), rank, world)
    seg_len = max(s["core"][1] - s["core"][0] for s in segs)
    per_rank = (len(segs) + world - 1) // world
    # local cores in a fixed-size [per_rank, seg_len] block so that one all-gather suffices
    local = np.zeros((per_rank, seg_len), dtype=np.float64)
    for slot, k in enumerate(mine):
        s = segs[k]
        y_pad = process_segment(x[s["padded"][0]:s["padded"][1]])
        core = extract_core(np.asarray(y_pad, dtype=np.float64), s)
        local[slot, :len(core)] = core
    if distributed:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t_local = torch.from_numpy(local).to(dev)
        gathered = torch.empty((world,) + t_local.shape, dtype=t_local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), t_local.view(-1), group=group) if hasattr(
            dist, "all_gather_into_tensor") and backend == "nccl" else dist.all_gather(
            [gathered[r] for r in range(world)], t_local, group=group)
        blocks = gathered.cpu().numpy()
    else:
        blocks = local[None]
    y = np.zeros(len(x), dtype=np.float64)
    for k, s in enumerate(segs):
        r, slot = k % world, k // world
        a, b = s["core"]
        y[a:b] = blocks[r, slot, :b - a]
    return y


def process_batch(items, process_item, group=None):
    """Round-robin shards a list of independent utterances; returns {index: result} of the local
    shard (no collective on the data path; gather the results with `gather_objects` if wanted)."""
    try:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    except Exception:  # pragma: no cover
        distributed = False
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    return {i: process_item(items[i]) for i in shard_indices(len(items), rank, world)}
