"""Data formats either side of the chain on the GPU: the wav -> wav flow of the
reference demo with 16-bit PCM crossing PCIe, fp32-narrowed outputs, and the sample-format conversions of
tools/audioio.cpp as device kernels."""
import numpy as np
import pytest

from oracle import refbin

pytestmark = pytest.mark.gpu


def _quantise(x):   # wavwrite, tools/audioio.cpp:176-180
    return np.clip(np.trunc(np.asarray(x, dtype=np.float64) * 32767), -32768, 32767).astype(np.int16)


def test_wav_to_wav_matches_the_reference_demo_flow(wb, signals, tmp_path):
    """wavread -> Harvest -> CheapTrick -> D4C -> Synthesis -> wavwrite (test/test.cpp:288-384): the reference
    sees x = pcm / 32768 (audioio.cpp:232-249) and its waveform is quantised by wavwrite."""
    fs = 16000
    wav = str(tmp_path / "in.wav")
    wb.wavwrite(signals.synth_speech(fs, 1.0, seed=21), fs, 16, wav)
    pcm, fs_read = wb.wavread_pcm16(wav)
    x, _, _ = wb.wavread(wav)
    assert fs_read == fs and np.array_equal(x, pcm / 32768.0)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
    pl.set_fresh_rng(True)
    out_pcm = pl.run_pcm16(pcm)
    y64 = pl.run(x, want_params=False)["y"]
    assert out_pcm.dtype == np.int16 and len(out_pcm) == len(ref["y"])
    assert np.array_equal(out_pcm, _quantise(y64))                     # the device conversion is wavwrite's, bit for bit
    want = _quantise(ref["y"])
    diff = np.abs(out_pcm.astype(np.int32) - want.astype(np.int32))
    # the waveforms agree to ~1e-10 of peak; a sample sitting on a truncation boundary may land one step apart
    assert diff.max() <= 1 and np.count_nonzero(diff) <= len(want) // 1000
    # and the file written from it is what the reference's writer would have written
    wb.wavwrite(out_pcm / 32767.0 + (np.sign(out_pcm) * 0.25 / 32767.0), fs, 16, str(tmp_path / "out.wav"))
    assert np.array_equal(wb.wavread_pcm16(str(tmp_path / "out.wav"))[0], out_pcm)


def test_fp32_outputs_are_the_rounded_fp64_results(wb, signals):
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=22)
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
    pl.set_fresh_rng(True)
    full = pl.run(x)
    narrow = pl.run_f32(x)
    for k in ("f0", "sp", "ap", "y"):
        assert narrow[k].dtype == np.float32 and narrow[k].shape == full[k].shape
        assert np.array_equal(narrow[k], full[k].astype(np.float32)), k
    only_params = pl.run_f32(x, want_y=False)
    assert "y" not in only_params and np.array_equal(only_params["sp"], narrow["sp"])


def test_device_sample_format_conversions(wb):
    import torch
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32768, size=100003).astype(np.int16)
    pcm[:3] = [-32768, 32767, 0]
    d_pcm = torch.from_numpy(pcm).cuda()
    d_x = torch.empty(len(pcm), dtype=torch.float64, device="cuda")
    assert wb.lib().wb_pcm16_to_f64_dev(d_pcm.data_ptr(), len(pcm), d_x.data_ptr(), None) == 0
    wb.device_synchronize()
    assert np.array_equal(d_x.cpu().numpy(), pcm / 32768.0)
    x = rng.normal(size=100003) * 0.5
    x[:6] = [1.0, -1.0, 1.5, -1.5, 1e300, -1e300]
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty(len(x), dtype=torch.int16, device="cuda")
    assert wb.lib().wb_f64_to_pcm16_dev(d_in.data_ptr(), len(x), d_out.data_ptr(), None) == 0
    d_f32 = torch.empty(len(x), dtype=torch.float32, device="cuda")
    assert wb.lib().wb_f64_to_f32_dev(d_in.data_ptr(), len(x), d_f32.data_ptr(), None) == 0
    wb.device_synchronize()
    assert np.array_equal(d_out.cpu().numpy(), _quantise(np.clip(x, -4, 4)))
    with np.errstate(over="ignore"):
        assert np.array_equal(d_f32.cpu().numpy(), x.astype(np.float32))


def test_torch_tensor_front_end_matches_the_host_api(wb, signals):
    """worldb200.tensors: the stages on CUDA tensors (device-pointer ABI on torch's current stream) give the bits
    of the host-pointer class API."""
    import torch
    from worldb200 import tensors as wt
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=23)
    hopt = wb.HarvestOption(f0_floor=40.0, frame_period=5.0)
    wb.randn_reseed()
    tpos, f0 = wb.Harvest(fs, hopt).compute(x)
    ct = wb.CheapTrick(fs)
    sp = ct.compute(x, tpos, f0)
    ap = wb.D4C(fs).compute(x, tpos, f0, ct.fft_size)
    y = wb.Synthesis(fs, ct.fft_size, 5.0).compute(f0, sp, ap, wb.synthesis_length(len(f0), 5.0, fs))
    csp = wb.CodeSpectralEnvelope(sp, fs, ct.fft_size, 24)
    wb.randn_reseed()
    d_x = torch.from_numpy(x).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):                      # a non-default torch stream: the calls must follow it
        d_t, d_f0 = wt.harvest(d_x, fs, hopt)
        d_sp = wt.cheaptrick(d_x, fs, d_t, d_f0)
        d_ap = wt.d4c(d_x, fs, d_t, d_f0, ct.fft_size)
        d_y = wt.synthesis(d_f0, d_sp, d_ap, fs, 5.0)
        d_csp = wt.codec("code_sp", d_sp, fs, ct.fft_size, 24)
        d_sp32 = wt.to_float32(d_sp)
    s.synchronize()
    assert np.array_equal(d_t.cpu().numpy(), tpos) and np.array_equal(d_f0.cpu().numpy(), f0)
    assert np.array_equal(d_sp.cpu().numpy(), sp) and np.array_equal(d_ap.cpu().numpy(), ap)
    assert np.array_equal(d_y.cpu().numpy(), y) and np.array_equal(d_csp.cpu().numpy(), csp)
    assert np.array_equal(d_sp32.cpu().numpy(), sp.astype(np.float32))
    with pytest.raises(ValueError):
        wt.harvest(torch.zeros(10), fs)             # not a CUDA tensor
