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


# ---- exact stream sharding: planning and the exchange step (no GPU) ----------------------------------------
@pytest.mark.parametrize("n,fs,world,seg", [(480000, 48000, 2, 4), (60 * 48000 + 1234, 48000, 3, 30), (22050 * 7 + 5, 22050, 2, 3),
                                            (16000, 16000, 1, 30), (3600 * 48000, 48000, 8, 30), (16000 * 3, 16000, 8, 1),
                                            (3600 * 48000, 48000, 16, 120), (600 * 48000, 48000, 16, 120)])
def test_stream_plan_covers_the_stream_and_keeps_the_analysis_grid(n, fs, world, seg):
    import worldb200  # noqa: F401
    from worldb200 import parallel as P
    fft = 2048 if fs == 48000 else 1024
    pl = P.StreamPlan(n, fs, world, 5.0, fft, segment_seconds=seg, halo_seconds=2)
    assert pl.f0_length == int(1000.0 * n / fs / 5.0) + 1 and pl.out_length == int((pl.f0_length - 1) * 5.0 / 1000.0 * fs) + 1
    # frames and samples are partitioned in rank order
    assert pl.frames[0][0] == 0 and pl.frames[-1][1] == pl.f0_length and pl.samples[0][0] == 0 and pl.samples[-1][1] == pl.out_length
    for k in range(world - 1):
        assert pl.frames[k][1] == pl.frames[k + 1][0] and pl.samples[k][1] == pl.samples[k + 1][0]
    r = P.decimation_ratio(fs)
    for k in range(world):
        fb, fe = pl.frames[k]
        covered = fb
        for s in pl.segments[k]:
            pa, pb = s["padded"]
            cfb, cfe = s["frames"]
            assert cfb == covered and cfe > cfb
            covered = cfe
            assert pa % fs == 0 and pa % r == 0 and (pb - n) % r == 0 and 0 <= pa < pb <= n      # grid + decimation phase
            assert s["frame_offset"] * 5.0 / 1000.0 == pa / fs                                   # same frame times
            local_len = int(1000.0 * (pb - pa) / fs / 5.0) + 1                                   # Harvest::getSamples of the segment
            assert 0 <= cfb - s["frame_offset"] and cfe - s["frame_offset"] <= local_len
            # the core keeps >= 1 s of context on both sides unless it touches the stream's end
            assert pa == 0 or (cfb / 200.0 - pa / fs) >= 1.0
            assert pb >= n - r or (pb / fs - cfe / 200.0) >= 1.0
        assert covered == fe
        ra, rb = pl.rows[k]
        sa, sb = pl.samples[k]
        assert ra <= fb and rb >= fe
        # every pulse reaching into [sa, sb) sits at a sample in (sa - fft, sb + fft): its two frames are rows
        assert ra == 0 or ra <= int((sa - fft) / fs * 200.0)
        assert rb == pl.f0_length or rb >= int(np.ceil((sb + fft) / fs * 200.0)) + 1


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import worldb200  # noqa: F401
    from worldb200 import parallel as P
    ranges = [(0, 5), (5, 6), (6, 13)][:world] if world == 3 else [(0, 7), (7, 13)]
    b, e = ranges[rank]
    full = torch.arange(13, dtype=torch.float64) * 1.5 + 0.25
    got = P.gather_ranges(full[b:e].clone(), ranges, 13)
    # the same exchange in flight (the stream path issues it, enqueues other work, and collects it later), two at once
    pg1 = P.PendingGather(full[b:e].clone(), ranges)
    pg2 = P.PendingGather(-full[b:e].clone(), ranges)
    out2 = torch.zeros(13, dtype=torch.float64)
    pg2.finish_into(out2)
    got1 = pg1.finish(13)
    q.put((rank, bool(torch.equal(got, full) and torch.equal(got1, full) and torch.equal(out2, -full))))
    dist.destroy_process_group()


def test_gather_ranges_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_stream_plan_properties_hypothesis():
    """Randomised: any stream length / world size keeps the partition, grid-alignment and coverage invariants."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import worldb200  # noqa: F401
    from worldb200 import parallel as P

    @settings(max_examples=60, deadline=None)
    @given(fs=st.sampled_from([16000, 22050, 24000, 32000, 44100, 48000]), seconds=st.floats(0.6, 400.0), world=st.integers(1, 9),
           seg=st.integers(1, 40), halo=st.integers(1, 3), extra=st.integers(0, 11))
    def check(fs, seconds, world, seg, halo, extra):
        n = int(seconds * fs) + extra
        fft = 2048 if fs >= 44100 else 1024
        pl = P.StreamPlan(n, fs, world, 5.0, fft, segment_seconds=seg, halo_seconds=halo)
        r = P.decimation_ratio(fs)
        assert pl.frames[0][0] == 0 and pl.frames[-1][1] == pl.f0_length
        assert pl.samples[0][0] == 0 and pl.samples[-1][1] == pl.out_length
        for k in range(world):
            fb, fe = pl.frames[k]
            assert fb <= fe and (k == 0 or fb == pl.frames[k - 1][1])
            sa, sb = pl.samples[k]
            assert sa <= sb and (k == 0 or sa == pl.samples[k - 1][1])
            covered = fb
            for s in pl.segments[k]:
                pa, pb = s["padded"]
                assert s["frames"][0] == covered and pa % fs == 0 and (pb - n) % r == 0 and 0 <= pa < pb <= n
                assert s["frames"][1] - s["frame_offset"] <= int(1000.0 * (pb - pa) / fs / 5.0) + 1
                covered = s["frames"][1]
            assert covered == fe
            ra, rb = pl.rows[k]
            assert 0 <= ra <= fb and fe <= rb <= pl.f0_length
    check()


def test_numa_binding_helper_never_raises():
    """bind_to_gpu_node() is best effort: without NVML / a GPU (this container) it reports bound=False and why."""
    import os
    from worldb200 import parallel
    before = os.sched_getaffinity(0)
    info = parallel.bind_to_gpu_node(0)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_stream_plan_cuts_a_shard_into_the_fewest_even_segments():
    """One hour over 16 shards: 225 s each, Harvest segments of at most 120 s -> two pieces of 113 and 112 s (not 120 +
    105): every piece pays its halo, and the longest one bounds the scratch of a Harvest call."""
    import worldb200  # noqa: F401
    from worldb200 import parallel as P
    pl = P.StreamPlan(3600 * 48000, 48000, 16, 5.0, 2048, segment_seconds=120, halo_seconds=2)
    for segs in pl.segments:
        cores = [(s["frames"][1] - s["frames"][0]) / 200.0 for s in segs]
        assert len(segs) == 2 and max(cores) <= 120.0 + 0.005 and abs(cores[0] - cores[1]) <= 1.0 + 0.005
    one = P.StreamPlan(600 * 48000, 48000, 16, 5.0, 2048, segment_seconds=120, halo_seconds=2)
    assert all(len(segs) == 1 for segs in one.segments)
