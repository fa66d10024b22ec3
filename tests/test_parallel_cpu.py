"""World-size-2 `gloo` test of the sharding plumbing (CPU): segment planning, halo discard and the
single all-gather stitch, with a stand-in per-segment function (the CUDA path needs a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_vocoder(x_pad):
    """A local (finite-support) operator: 3-tap smoothing.  Its result inside a core does not depend
    on where the stream was cut as long as the halo is >= 1 sample -- like the real chain with its halo."""
    xp = np.pad(x_pad, 1, mode="edge")
    return (xp[:-2] + 2.0 * xp[1:-1] + xp[2:]) / 4.0


def _worker(rank, world, port, n, fs, out_dir):
    sys.path.insert(0, ROOT)
    import worldb200  # noqa: F401
    from worldb200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    x = np.sin(np.arange(n) * 0.01) + 0.1 * np.cos(np.arange(n) * 0.37)
    y = parallel.process_stream(x, fs, _fake_vocoder, segment_seconds=1.0, halo_seconds=0.01)
    np.save(os.path.join(out_dir, "y%d.npy" % rank), y)
    items = [np.full(4, i, dtype=np.float64) for i in range(7)]
    mine = parallel.process_batch(items, lambda a: float(a.sum()))
    np.save(os.path.join(out_dir, "b%d.npy" % rank), np.array(sorted(mine.items())))
    dist.barrier()
    dist.destroy_process_group()


def test_stream_sharding_world2(tmp_path):
    n, fs, world = 10500, 1000, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, fs, str(tmp_path)), nprocs=world, join=True)
    x = np.sin(np.arange(n) * 0.01) + 0.1 * np.cos(np.arange(n) * 0.37)
    expect = _fake_vocoder(x)
    for r in range(world):
        y = np.load(tmp_path / ("y%d.npy" % r))
        assert y.shape == (n,)
        assert np.allclose(y, expect, rtol=0, atol=1e-15), "rank %d stitched stream differs" % r
    b0, b1 = np.load(tmp_path / "b0.npy"), np.load(tmp_path / "b1.npy")
    assert [int(i) for i in b0[:, 0]] == [0, 2, 4, 6] and [int(i) for i in b1[:, 0]] == [1, 3, 5]
    assert np.allclose(b0[:, 1], [0, 8, 16, 24]) and np.allclose(b1[:, 1], [4, 12, 20])


def test_segment_plan_covers_stream_once():
    sys.path.insert(0, ROOT)
    import worldb200  # noqa: F401
    from worldb200 import parallel
    for n, fs, seg, halo in [(480000, 48000, 2.0, 0.5), (1000, 1000, 0.3, 0.05), (7, 10, 1.0, 1.0)]:
        segs = parallel.plan_segments(n, fs, seg, halo)
        cover = np.zeros(n, dtype=int)
        for s in segs:
            a, b = s["core"]
            pa, pb = s["padded"]
            assert 0 <= pa <= a < b <= pb <= n
            cover[a:b] += 1
        assert np.all(cover == 1)
    assert parallel.shard_indices(10, 1, 4) == [1, 5, 9]


def test_single_process_fallback_matches():
    sys.path.insert(0, ROOT)
    import worldb200  # noqa: F401
    from worldb200 import parallel
    x = np.random.default_rng(0).standard_normal(5000)
    y = parallel.process_stream(x, 1000, _fake_vocoder, segment_seconds=0.7, halo_seconds=0.01)
    assert np.allclose(y, _fake_vocoder(x), atol=1e-15)
