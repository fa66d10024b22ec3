"""GPU parity of the whole chain (test/test.cpp call sequence) against one reference process."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _check_chain(out, ref):
    assert np.array_equal(out["tpos"], ref["tpos"])
    assert np.array_equal(out["f0"] > 0, ref["f0"] > 0)           # voicing decisions: exact
    v = ref["f0"] > 0
    assert np.max(np.abs(out["f0"][v] - ref["f0"][v]) / ref["f0"][v]) < RTOL
    assert np.max(np.abs(out["sp"] - ref["sp"]) / ref["sp"]) < RTOL
    assert np.max(np.abs(out["ap"] - ref["ap"]) / ref["ap"]) < RTOL
    peak = np.abs(ref["y"]).max()
    assert np.max(np.abs(out["y"] - ref["y"])) / peak < RTOL


@pytest.mark.parametrize("fs,seconds,seed", [(16000, 1.0, 0), (22050, 5.0, 1000), (48000, 10.0, 0)])
def test_pipeline_matches_reference(wb, signals, fs, seconds, seed):
    """configs[0] (16 kHz / 1 s), one utterance of configs[2] (22.05 kHz / 5 s) and configs[1] (48 kHz / 10 s)."""
    x = signals.synth_speech(fs, seconds, seed=seed)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    wb.randn_reseed()
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0),
                     wb.D4COption(threshold=0.85))
    out = pl.run(x)
    assert pl.fft_size == ref["fft_size"]
    _check_chain(out, ref)


def test_class_api_chain_matches_reference(wb, signals):
    """The four stage objects called like test.cpp does, host buffers throughout."""
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=7)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    wb.randn_reseed()
    tpos, f0 = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0)).compute(x)
    ct = wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0))
    sp = ct.compute(x, tpos, f0)
    ap = wb.D4C(fs, wb.D4COption(threshold=0.85)).compute(x, tpos, f0, ct.fft_size)
    y = wb.Synthesis(fs, ct.fft_size, 5.0).compute(f0, sp, ap, wb.synthesis_length(len(f0), 5.0, fs))
    _check_chain(dict(tpos=tpos, f0=f0, sp=sp, ap=ap, y=y), ref)


def test_silence_edges(wb, signals):
    """Digital silence at both ends (SURVEY section 8d parity-only case): envelope and aperiodicity of
    silent frames are pure functions of the randn() stream, so this pins the exact noise placement."""
    fs = 16000
    x = signals.silence_edge(fs, 1.0, seed=8)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    wb.randn_reseed()
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
    out = pl.run(x)
    _check_chain(out, ref)


def test_graph_replay_is_identical(wb, signals):
    """run_dev through a captured CUDA graph gives bit-identical results to ordinary launches."""
    import torch
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=12)
    d_x = torch.from_numpy(x).cuda()
    outs = []
    for use_graph in (False, True):
        pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
        pl.set_fresh_rng(True)
        pl.set_graph(use_graph)
        n, ny = len(x), pl.out_length(len(x))
        d_y = torch.zeros(ny, dtype=torch.float64, device="cuda")
        d_f0 = torch.zeros(pl.f0_length(n), dtype=torch.float64, device="cuda")
        ys = []
        for _ in range(4):   # run 1: warm, run 2: capture, runs 3-4: replay
            d_y.zero_()
            pl.run_dev(d_x.data_ptr(), n, d_y=d_y.data_ptr(), y_length=ny, d_f0=d_f0.data_ptr())
            wb.device_synchronize()
            ys.append(d_y.cpu().numpy().copy())
        assert all(np.array_equal(ys[0], y) for y in ys[1:])
        outs.append((ys[-1], d_f0.cpu().numpy().copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][0]).max() > 0.1


def test_graph_survives_workspace_growth(wb, signals):
    """A captured graph must not be replayed after a longer input made the workspace reallocate its buffers
    (ADVICE round 1): A, A (capture), B (longer, grows), A again."""
    import torch
    fs = 16000
    xa = signals.synth_speech(fs, 0.6, seed=14)
    xb = signals.synth_speech(fs, 1.5, seed=15)
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
    pl.set_fresh_rng(True)
    pl.set_graph(True)
    d_a, d_b = torch.from_numpy(xa).cuda(), torch.from_numpy(xb).cuda()

    d_ya = torch.zeros(pl.out_length(len(xa)), dtype=torch.float64, device="cuda")
    d_yb = torch.zeros(pl.out_length(len(xb)), dtype=torch.float64, device="cuda")

    def run(d_x, n, d_y):   # (same pointers every time: the graph key matches)
        d_y.zero_()
        pl.run_dev(d_x.data_ptr(), n, d_y=d_y.data_ptr(), y_length=d_y.numel())
        wb.device_synchronize()
        return d_y.cpu().numpy().copy()

    ya = [run(d_a, len(xa), d_ya) for _ in range(3)]
    yb = run(d_b, len(xb), d_yb)
    ya2 = run(d_a, len(xa), d_ya)
    assert np.abs(yb).max() > 0.1
    assert all(np.array_equal(ya[0], y) for y in ya[1:]) and np.array_equal(ya[0], ya2)


def test_page_locked_outputs_are_downloaded_inside_the_chain(wb, signals):
    """wb_pipeline_run with page-locked caller buffers sends every result home as soon as its stage is done (f0 after
    Harvest, the spectrogram beside D4C, the aperiodicity beside the impulse responses): same bits as the ordinary
    pageable path, with and without graph replay, also when the caller rotates between two sets of buffers."""
    import torch
    fs = 16000
    x = signals.synth_speech(fs, 1.3, seed=77)
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85))
    pl.set_fresh_rng(True)
    want = pl.run(x)
    L, bins, ny = len(want["f0"]), want["sp"].shape[1], len(want["y"])
    pin = lambda *shape: torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()
    sets = [dict(tpos=pin(L), f0=pin(L), sp=pin(L, bins), ap=pin(L, bins), y=pin(ny)) for _ in range(2)]
    for graph in (False, True):
        pl.set_graph(graph)
        for k in (0, 0, 0, 1, 0, 1, 1, 1):          # repeats (graph captured and replayed) and rotations (plain launches)
            for a in sets[k].values():
                a[...] = -1.0
            got = pl.run(x, out=sets[k])
            for name in ("tpos", "f0", "sp", "ap", "y"):
                assert np.array_equal(got[name], want[name]), (graph, k, name)
    pl.set_graph(False)
    only_y = pl.run(x, want_params=False, out={"y": sets[0]["y"]})
    assert np.array_equal(only_y["y"], want["y"])
