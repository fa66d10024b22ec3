"""GPU parity: Synthesis through the C-ABI vs the reference (fed the reference's own
f0 / spectrogram / aperiodicity so that only Synthesis is under test)."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: synthesized waveform within 1e-4 relative (to the waveform's peak) in fp64


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (22050, 1.5), (48000, 2.0)])
def test_synthesis_matches_reference(wb, signals, fs, seconds):
    x = signals.synth_speech(fs, seconds, seed=3)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    # fresh process semantics: CheapTrick and D4C consumed randn() before Synthesis; replay their
    # consumption by running them (their parity is covered elsewhere), then synthesize from the
    # REFERENCE parameters.
    wb.randn_reseed()
    wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0)).compute(x, ref["tpos"], ref["f0"])
    wb.D4C(fs).compute(x, ref["tpos"], ref["f0"], ref["fft_size"])
    syn = wb.Synthesis(fs, ref["fft_size"], 5.0)
    y = syn.compute(ref["f0"], ref["sp"], ref["ap"], len(ref["y"]))
    assert np.all(np.isfinite(y))
    peak = np.abs(ref["y"]).max()
    err = float(np.abs(y - ref["y"]).max() / peak)
    print("synthesis fs=%d peak %.3f max err/peak %.3e" % (fs, peak, err))
    assert err < RTOL


# ---- streaming synthesis -----------------------------------------------------------
def _analysis(wb, signals, fs, seconds, seed):
    x = signals.synth_speech(fs, seconds, seed=seed)
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
    pl.set_fresh_rng(True)
    out = pl.run(x)
    return out["f0"], out["sp"], out["ap"], pl.fft_size, len(out["y"])


@pytest.mark.parametrize("fs,pieces", [(16000, "random"), (16000, "single frames"), (48000, "random"), (22050, "two")])
def test_streaming_synthesis_equals_the_one_shot_call_bit_for_bit(wb, signals, fs, pieces):
    f0, sp, ap, fft_size, ny = _analysis(wb, signals, fs, 2.0 if fs != 48000 else 1.5, 41)
    wb.randn_reseed()
    want = wb.Synthesis(fs, fft_size, 5.0).compute(f0, sp, ap, ny)
    state_after = wb.randn_get_state()
    rng = np.random.default_rng(5)
    L = len(f0)
    if pieces == "random":
        cuts = np.unique(np.concatenate([[0, L], rng.integers(1, L, size=12)]))
    elif pieces == "single frames":
        cuts = np.arange(L + 1)
    else:
        cuts = np.array([0, L // 3, L])
    wb.randn_reseed()
    st = wb.SynthesisStream(fs, fft_size, 5.0, 1000.0)
    got = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        piece = st.push(f0[a:b], sp[a:b], ap[a:b])
        # nothing is emitted before it is final, and never more than the frames so far can determine
        assert sum(len(g) for g in got) + len(piece) <= int((b - 1) * 5.0 / 1000.0 * fs) + 1
        got.append(piece)
    got.append(st.finish(ny))
    y = np.concatenate(got)
    assert len(y) == ny
    bad = np.flatnonzero(y != want)
    assert bad.size == 0, "first / count of differing samples: %d %d" % (bad[0], bad.size)
    assert wb.randn_get_state() == state_after            # the randn() stream was consumed exactly as by compute()
    emitted_early = sum(len(g) for g in got[:-1])
    assert emitted_early > 0.8 * ny - 2 * fft_size          # it really streams: most samples leave before finish()


def test_long_waveform_pulse_scan_matches_the_piecewise_one(wb):
    """90 s at 48 kHz in one compute(): 4.32 M samples = 16 875 pulse-count blocks, past the 16 384 items one CTA
    scans (wb_scan.cuh: grid-wide above).  The streaming synthesis sees the same frames in pieces of 1500 (its scans
    stay on one CTA) and must give the same waveform and randn() state bit for bit."""
    fs, fft_size, L = 48000, 2048, 18001
    bins = fft_size // 2 + 1
    rng = np.random.default_rng(5)
    t = np.arange(L) * 0.005
    f0 = 140.0 + 50.0 * np.sin(2 * np.pi * 0.31 * t) + 20.0 * np.sin(2 * np.pi * 2.3 * t)
    f0[(t % 3.7) > 2.9] = 0.0                                   # unvoiced stretches
    env = np.exp(-np.arange(bins) / 300.0)[None, :]
    sp = (env * (0.5 + rng.random((L, 1)))) ** 2 + 1e-12
    ap = np.clip(0.05 + 0.9 * (np.arange(bins)[None, :] / bins) * (0.5 + 0.5 * rng.random((L, 1))), 0.001, 0.999)
    ny = wb.synthesis_length(L, 5.0, fs)
    wb.randn_reseed()
    want = wb.Synthesis(fs, fft_size, 5.0).compute(f0, sp, ap, ny)
    state_after = wb.randn_get_state()
    wb.randn_reseed()
    st = wb.SynthesisStream(fs, fft_size, 5.0, 1000.0)
    got = [st.push(f0[a:a + 1500], sp[a:a + 1500], ap[a:a + 1500]) for a in range(0, L, 1500)]
    got.append(st.finish(ny))
    y = np.concatenate(got)
    assert len(y) == ny and np.abs(want).max() > 1e-3
    bad = np.flatnonzero(y != want)
    assert bad.size == 0, "first / count of differing samples: %d %d" % (bad[0], bad.size)
    assert wb.randn_get_state() == state_after
    wb.randn_reseed()


def test_streaming_synthesis_small_output_buffers_and_errors(wb, signals):
    fs = 16000
    f0, sp, ap, fft_size, ny = _analysis(wb, signals, fs, 1.0, 42)
    wb.randn_reseed()
    want = wb.Synthesis(fs, fft_size, 5.0).compute(f0, sp, ap, ny)
    wb.randn_reseed()
    st = wb.SynthesisStream(fs, fft_size, 5.0)
    got = []
    for a in range(0, len(f0), 40):
        got.append(st.push(f0[a:a + 40], sp[a:a + 40], ap[a:a + 40], out_capacity=333))   # fewer than a piece yields
        assert len(got[-1]) <= 333
    got.append(st.finish(ny, out_capacity=1000))
    assert np.array_equal(np.concatenate(got), want)
    with pytest.raises(wb.WorldB200Error):
        st.push(f0[:2], sp[:2], ap[:2])                    # finished
    st2 = wb.SynthesisStream(fs, fft_size, 5.0)
    st2.push(f0[:100], sp[:100], ap[:100])
    with pytest.raises(wb.WorldB200Error):
        st2.finish(10)                                     # shorter than what is already determined
    with pytest.raises(wb.WorldB200Error):
        wb.SynthesisStream(fs, 1000, 5.0)                  # not a power of two
    wb.randn_reseed()


def test_pinned_contiguous_matrices_take_the_direct_pipelined_path(wb, signals):
    """Page-locked, contiguous caller matrices are DMAed directly and Synthesis::compute renders sample ranges under the
    remaining uploads (four ranges, each with its own pulses at their whole-waveform noise positions): the waveform and
    the randn() state afterwards equal the ordinary path's (separately allocated / pageable rows) bit for bit; the same
    for the outputs of CheapTrick::compute / D4C::compute written straight into page-locked memory."""
    import torch
    fs = 48000
    x = signals.synth_speech(fs, 2.0, seed=9)
    tpos, f0 = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0)).compute(x)
    ct = wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0))
    d4 = wb.D4C(fs, wb.D4COption(threshold=0.85))
    L, bins = len(f0), ct.fft_size // 2 + 1
    ny = wb.synthesis_length(L, 5.0, fs)
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    res = []
    for pinned in (False, True):
        wb.randn_reseed()
        sp = pin(L, bins) if pinned else np.empty((L, bins))
        ap = pin(L, bins) if pinned else np.empty((L, bins))
        y = pin(ny) if pinned else np.empty(ny)
        ct.compute(x, tpos, f0, sp)
        d4.compute(x, tpos, f0, ct.fft_size, ap)
        wb.Synthesis(fs, ct.fft_size, 5.0).compute(f0, sp, ap, ny, y)
        res.append((sp.copy(), ap.copy(), y.copy(), wb.randn_get_state()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2])
    assert res[0][3] == res[1][3]
    assert np.abs(res[0][2]).max() > 0.05
