"""Exact sharding of one long stream (BASELINE configs[3]) on one GPU with VIRTUAL ranks:
the rows / samples computed rank by rank -- whole-stream randn() positions, whole-stream phase sum, exchanged
Love Train decisions -- are bit-identical to an unsharded run, and Harvest on aligned, padded segments
reproduces the whole-stream contour."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu


def _options(wb):
    return wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85)


@pytest.mark.parametrize("fs,seconds,world", [(16000, 7.3, 3), (48000, 4.0, 2), (22050, 5.0, 4)])
def test_ranges_are_bit_identical_to_the_unsharded_run(wb, signals, fs, seconds, world):
    import torch
    from worldb200 import parallel
    x = signals.synth_speech(fs, seconds, seed=31)
    hopt, copt, dopt = _options(wb)
    whole = wb.Pipeline(fs, hopt, copt, dopt)
    whole.set_fresh_rng(True)
    ref = whole.run(x)
    d_x = torch.from_numpy(x).cuda()
    out = parallel.simulate_stream_ranks(d_x, fs, world, hopt, copt, dopt, segment_seconds=2, halo_seconds=1,
                                         d_f0_all=torch.from_numpy(ref["f0"]).cuda())
    assert out["plan"].f0_length == len(ref["f0"]) and out["plan"].out_length == len(ref["y"])
    assert np.array_equal(out["sp"].cpu().numpy(), ref["sp"])
    assert np.array_equal(out["ap"].cpu().numpy(), ref["ap"])
    y = out["y"].cpu().numpy()
    bad = np.flatnonzero(y != ref["y"])
    assert bad.size == 0, "first / last / count of differing samples: %d %d %d, samples per rank %r" % (
        bad[0], bad[-1], bad.size, out["plan"].samples)
    # every rank builds its pulse list for its own samples only, yet knows the first pulse of the whole stream (the
    # origin of the excitation's randn() positions) and the number of draws the whole stream makes
    first = int(whole.debug_read("syn_np", (2,), dtype=np.int32)[1])
    draws = int(whole.debug_read("syn_ncount", (1,), dtype=np.uint64)[0])
    pulses = int(whole.debug_read("syn_np", (2,), dtype=np.int32)[0])
    for w in out["workers"]:
        np_w = w.pipe.debug_read("syn_np", (2,), dtype=np.int32)
        assert int(np_w[1]) == first and int(w.pipe.debug_read("syn_ncount", (1,), dtype=np.uint64)[0]) == draws
        assert 0 < int(np_w[0]) < pulses, "a rank's pulse list covers its own window of the stream"


def test_sharded_harvest_and_chain_match_the_reference_process(wb, signals):
    """Everything sharded, Harvest included (3 virtual ranks, 2 s sub-segments, 2 s halo), against ONE reference
    process run on the whole stream."""
    import torch
    from worldb200 import parallel
    fs = 16000
    x = signals.synth_speech(fs, 9.0, seed=32)
    hopt, copt, dopt = _options(wb)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    out = parallel.simulate_stream_ranks(torch.from_numpy(x).cuda(), fs, 3, hopt, copt, dopt, segment_seconds=2, halo_seconds=2)
    f0 = out["f0"].cpu().numpy()
    assert np.array_equal(f0 > 0, ref["f0"] > 0)                       # voicing decisions: exact
    v = ref["f0"] > 0
    assert np.max(np.abs(f0[v] - ref["f0"][v]) / ref["f0"][v]) < 1e-4
    assert np.max(np.abs(out["sp"].cpu().numpy() - ref["sp"]) / ref["sp"]) < 1e-4
    assert np.max(np.abs(out["ap"].cpu().numpy() - ref["ap"]) / ref["ap"]) < 1e-4
    y = out["y"].cpu().numpy()
    assert np.max(np.abs(y - ref["y"])) / np.abs(ref["y"]).max() < 1e-4


def test_sharded_harvest_matches_the_whole_stream_contour_48k(wb, signals):
    import torch
    from worldb200 import parallel
    fs = 48000
    x = signals.synth_speech(fs, 12.0, seed=33)
    hopt, copt, dopt = _options(wb)
    tpos, f0 = wb.Harvest(fs, hopt).compute(x)
    plan = parallel.StreamPlan(len(x), fs, 2, 5.0, 2048, segment_seconds=3, halo_seconds=2)
    d_x = torch.from_numpy(x).cuda()
    got = torch.cat([parallel.StreamWorker(plan, k, hopt, copt, dopt).harvest_local(d_x) for k in range(2)]).cpu().numpy()
    assert len(got) == len(f0) and np.array_equal(got > 0, f0 > 0)
    v = f0 > 0
    assert np.max(np.abs(got[v] - f0[v]) / f0[v]) < 1e-9


def test_single_process_driver_with_several_shards(wb, signals):
    """process_stream_exact() without torch.distributed, the rank's share cut into 3 shards (bounded scratch):
    same bits as the unsharded run given the same f0, and all three exchange steps degenerate gracefully."""
    import torch
    from worldb200 import parallel
    fs = 16000
    x = signals.synth_speech(fs, 6.0, seed=34)
    hopt, copt, dopt = _options(wb)
    whole = wb.Pipeline(fs, hopt, copt, dopt)
    whole.set_fresh_rng(True)
    ref = whole.run(x)
    t = {}
    out = parallel.process_stream_exact(torch.from_numpy(x).cuda(), fs, hopt, copt, dopt, segment_seconds=2, halo_seconds=2,
                                        d_f0_all=torch.from_numpy(ref["f0"]).cuda(), shards_per_rank=3, timings=t)
    assert np.array_equal(out["y"].cpu().numpy(), ref["y"]) and np.array_equal(out["sp"].cpu().numpy(), ref["sp"])
    assert np.array_equal(out["ap"].cpu().numpy(), ref["ap"]) and t["total"] > 0
    # and with its own (sharded) Harvest: decisions identical to the whole-stream contour
    out2 = parallel.process_stream_exact(torch.from_numpy(x).cuda(), fs, hopt, copt, dopt, segment_seconds=2, halo_seconds=2, shards_per_rank=2)
    f0 = out2["f0"].cpu().numpy()
    assert np.array_equal(f0 > 0, ref["f0"] > 0) and np.max(np.abs(f0 - ref["f0"])) < 1e-7
