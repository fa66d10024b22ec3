"""GPU parity sweep over sample rates, frame periods, options and edge cases (the reference ships no
tests; these are the configurations its headers make reachable)."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _chain(wb, x, fs, frame_period=5.0, h_floor=40.0, h_ceil=800.0, ct_floor=71.0, thr=0.85):
    wb.randn_reseed()
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=h_floor, f0_ceil=h_ceil, frame_period=frame_period),
                     wb.CheapTrickOption(f0_floor=ct_floor), wb.D4COption(threshold=thr))
    return pl.run(x)


def _compare(out, ref, what):
    assert np.array_equal(out["tpos"], ref["tpos"]), what
    assert np.array_equal(out["f0"] > 0, ref["f0"] > 0), what + ": voicing"
    v = ref["f0"] > 0
    if v.any():
        assert np.max(np.abs(out["f0"][v] - ref["f0"][v]) / ref["f0"][v]) < RTOL, what
    assert np.max(np.abs(out["sp"] - ref["sp"]) / ref["sp"]) < RTOL, what + ": sp"
    assert np.max(np.abs(out["ap"] - ref["ap"]) / ref["ap"]) < RTOL, what + ": ap"
    peak = max(np.abs(ref["y"]).max(), 1e-12)
    assert np.max(np.abs(out["y"] - ref["y"])) / peak < RTOL, what + ": y"


@pytest.mark.parametrize("fs", [8000, 12000, 22050, 24000, 32000, 44100, 96000])
def test_sample_rates(wb, signals, fs):
    x = signals.synth_speech(fs, 0.6, seed=20)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    _compare(_chain(wb, x, fs), ref, "fs=%d" % fs)


@pytest.mark.parametrize("frame_period", [1.0, 2.5, 10.0])
def test_frame_periods(wb, signals, frame_period):
    fs = 16000
    x = signals.synth_speech(fs, 0.8, seed=21)
    ref, _ = refbin.run_reference(x, fs, stages="hcds", frame_period=frame_period)
    _compare(_chain(wb, x, fs, frame_period=frame_period), ref, "frame_period=%g" % frame_period)


def test_default_harvest_options_and_conventional_d4c(wb, signals):
    """Library defaults (f0_floor 71) and D4C threshold 0 (the 'conventional D4C' of test.cpp:176-180)."""
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=22)
    ref, _ = refbin.run_reference(x, fs, stages="hcds", harvest_f0_floor=71.0, d4c_threshold=0.0)
    _compare(_chain(wb, x, fs, h_floor=71.0, thr=0.0), ref, "defaults")


def test_narrow_f0_range(wb, signals):
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=23)
    ref, _ = refbin.run_reference(x, fs, stages="hcds", harvest_f0_floor=90.0, harvest_f0_ceil=300.0)
    _compare(_chain(wb, x, fs, h_floor=90.0, h_ceil=300.0), ref, "narrow range")


def test_loud_input_takes_the_truncating_dc_path(wb, signals):
    """|y| >= 1 after decimation: harvest.cpp:238-241 subtracts an int-truncated running sum."""
    fs = 16000
    x = 4.0 * signals.synth_speech(fs, 0.7, seed=24) + 0.8
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    _compare(_chain(wb, x, fs), ref, "loud input")


def test_long_utterance(wb, signals):
    """90 s in one piece: 113 overlap-save blocks in Harvest, contour buffers beyond shared memory, and 18 001 frames --
    past the 16 384 up to which the frames' randn() positions are scanned by one CTA (wb_scan.cuh: grid-wide above)."""
    fs = 16000
    x = signals.synth_speech(fs, 90.0, seed=25)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    _compare(_chain(wb, x, fs), ref, "90 s utterance")


def test_harvest_beyond_128_overlap_save_blocks(wb, signals):
    """150 s in one Harvest call (188 overlap-save blocks): the per-block edge runs are concatenated by a prefix sum
    over up to 1024 blocks; past that (a little over 16 minutes) the call refuses before launching anything."""
    fs = 16000
    x = signals.synth_speech(fs, 150.0, seed=26)
    ref, _ = refbin.run_reference(x, fs, stages="h")
    h = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, f0_ceil=800.0, frame_period=5.0))
    tpos, f0 = h.compute(x)
    assert np.array_equal(tpos, ref["tpos"])
    assert np.array_equal(f0 > 0, ref["f0"] > 0), "voicing decisions differ"
    np.testing.assert_allclose(f0, ref["f0"], rtol=1e-4, atol=0)   # BASELINE.json tolerance (1e-4 relative)
    too_long = np.zeros(8000 * 20 * 60)
    with pytest.raises(wb.WorldB200Error):
        wb.Harvest(8000).compute(too_long)


def test_ragged_lengths(wb, signals):
    """Lengths that are not multiples of the decimation ratio / frame hop."""
    fs = 48000
    for n in (12345, 23999, 26001):   # (shorter inputs have no voiced section: the reference itself crashes)
        x = signals.synth_speech(fs, 0.6, seed=24)[:n]
        ref, _ = refbin.run_reference(x, fs, stages="hcds")
        _compare(_chain(wb, x, fs), ref, "n=%d" % n)


def test_noise_only_input_is_all_unvoiced(wb):
    """No voiced section at all: the reference's contour merge reads an empty array (undefined, it may
    crash); we define the result as unvoiced.  CheapTrick / D4C / Synthesis then follow the reference
    given that f0."""
    fs = 16000
    x = 0.01 * np.random.default_rng(5).standard_normal(8000)
    hv = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0))
    tpos, f0 = hv.compute(x)
    assert np.all(f0 == 0.0)
    ref, _ = refbin.run_reference(x, fs, stages="cd", f0=f0)
    wb.randn_reseed()
    ct = wb.CheapTrick(fs, wb.CheapTrickOption(f0_floor=71.0))
    sp = ct.compute(x, tpos, f0)
    ap = wb.D4C(fs).compute(x, tpos, f0, ct.fft_size)
    assert np.max(np.abs(sp - ref["sp"]) / ref["sp"]) < RTOL
    assert np.array_equal(ap, ref["ap"])   # every row stays at 1 - 1e-12


def test_empty_and_bad_arguments(wb):
    L = wb.lib()
    assert L.wb_harvest_get_samples(16000, 0, 5.0) == 1
    hv = wb.Harvest(16000)
    with pytest.raises(wb.WorldB200Error):
        hv.compute(np.zeros(0))                      # x_length <= 0 -> WB_ERR_ARG (the reference: undefined)
    with pytest.raises(wb.WorldB200Error):
        wb.Harvest(16000, wb.HarvestOption(use_cos_table=1))
    with pytest.raises(wb.WorldB200Error):
        wb.Harvest(16000, wb.HarvestOption(f0_floor=0.0))
    ct = wb.CheapTrick(16000)
    sp = ct.compute(np.zeros(100), np.zeros(0), np.zeros(0))   # zero frames: nothing to do
    assert sp.shape == (0, ct.fft_size // 2 + 1)


def test_create_time_validation_and_async_error_reporting(wb, signals):
    """Options are checked when a handle is created (NaN / zero periods, floors above ceilings, FFT sizes that are
    not powers of two); conditions a kernel of an ASYNCHRONOUS call cannot handle are flagged on the device and
    readable afterwards."""
    import torch
    from worldb200 import tensors as wt
    fs = 16000
    for bad in (wb.HarvestOption(frame_period=0.0), wb.HarvestOption(frame_period=float("nan")),
                wb.HarvestOption(f0_floor=900.0, f0_ceil=800.0), wb.HarvestOption(f0_floor=-1.0)):
        with pytest.raises(wb.WorldB200Error):
            wb.Pipeline(fs, bad)
        with pytest.raises(wb.WorldB200Error):
            wb.Harvest(fs, bad)
    with pytest.raises(wb.WorldB200Error):
        wb.Pipeline(fs, None, wb.CheapTrickOption(fft_size=1000))
    with pytest.raises(wb.WorldB200Error):
        wb.CheapTrick(fs, wb.CheapTrickOption(fft_size=1000))
    with pytest.raises(wb.WorldB200Error):
        wb.Pipeline(fs, None, None, wb.D4COption(threshold=float("nan")))
    # a too small f0 bound makes the synthesis kernels drop pulses: the asynchronous tensor call cannot return that ...
    x = signals.synth_speech(fs, 1.0, seed=40)
    out = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0)).run(x)
    f0, sp, ap = (torch.from_numpy(out[k]).cuda() for k in ("f0", "sp", "ap"))
    wt.check_errors()                                            # nothing pending
    y_ok = wt.synthesis(f0, sp, ap, fs, 5.0, f0_upper_bound=900.0)
    wt.check_errors()
    f0_high = torch.full_like(f0, 760.0)                         # more pulses than a bound of 500 Hz (the unvoiced rate) allows for
    wt.synthesis(f0_high, sp, ap, fs, 5.0, f0_upper_bound=100.0)
    with pytest.raises(wb.WorldB200Error):
        wt.check_errors()                                        # ... but it is there to be asked for
    wt.check_errors()                                            # and cleared by the query
    assert float(y_ok.abs().max()) > 0.05
    # debug_read is bounded by the buffer it names
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0))
    pl.run(x)
    with pytest.raises(wb.WorldB200Error):
        pl.debug_read("no_such_buffer", (4,))
    with pytest.raises(wb.WorldB200Error):
        pl.debug_read("hv_nc", (1 << 20,), dtype=np.int32)
