"""torch-tensor front end of the device-pointer C-ABI: the four stages and the codec on
contiguous CUDA float64 tensors, no host copies.  torch is plumbing only (device memory and stream ordering); every
call goes to libworldb200.so (`wb_*_compute_dev`), on torch's CURRENT stream.

    import worldb200
    from worldb200 import tensors as wt
    tpos, f0 = wt.harvest(x, fs)                       # x: 1-D float64 CUDA tensor
    sp = wt.cheaptrick(x, fs, tpos, f0)                # [frames][fft_size/2+1]
    ap = wt.d4c(x, fs, tpos, f0, fft_size=(sp.shape[1] - 1) * 2)
    y = wt.synthesis(f0, sp, ap, fs, frame_period=5.0)

The randn() stream is the library's process-global one, consumed in call order like the reference's
(src/world_matlabfunctions.cpp:243-264); `worldb200.randn_reseed()` restarts it."""
import ctypes

import worldb200 as wb

_objects = {}


def _cached(kind, key, make):
    k = (kind,) + key
    if k not in _objects:
        _objects[k] = make()
    return _objects[k]


def _check_vec(t, name):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous float64 CUDA tensor" % name)
    return t


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _opt_key(o):
    return tuple(getattr(o, f[0]) for f in o._fields_) if o is not None else ()


def harvest(x, fs, option=None):
    """-> (temporal_positions, f0); include/harvest.hpp:39-41"""
    import torch
    x = _check_vec(x, "x")
    opt = option if option is not None else wb.HarvestOption()
    h = _cached("harvest", (int(fs), x.device.index) + _opt_key(opt), lambda: wb.Harvest(fs, opt))
    n = h.getSamples(fs, x.numel())
    tpos = torch.empty(n, dtype=torch.float64, device=x.device)
    f0 = torch.empty(n, dtype=torch.float64, device=x.device)
    wb._check(wb.lib().wb_harvest_compute_dev(h._h, x.data_ptr(), x.numel(), tpos.data_ptr(), f0.data_ptr(), _stream()),
              "wb_harvest_compute_dev")
    return tpos, f0


def cheaptrick(x, fs, temporal_positions, f0, option=None):
    """-> spectrogram [frames][fft_size/2+1]; include/cheaptrick.hpp:30-32"""
    import torch
    x, tpos, f0 = _check_vec(x, "x"), _check_vec(temporal_positions, "temporal_positions"), _check_vec(f0, "f0")
    h = _cached("cheaptrick", (int(fs), x.device.index) + _opt_key(option), lambda: wb.CheapTrick(fs, option))
    sp = torch.empty((f0.numel(), h.fft_size // 2 + 1), dtype=torch.float64, device=x.device)
    wb._check(wb.lib().wb_cheaptrick_compute_dev(h._h, x.data_ptr(), x.numel(), tpos.data_ptr(), f0.data_ptr(), f0.numel(),
                                                 sp.data_ptr(), _stream()), "wb_cheaptrick_compute_dev")
    return sp


def d4c(x, fs, temporal_positions, f0, fft_size, option=None):
    """-> aperiodicity [frames][fft_size/2+1]; include/d4c.hpp:30-33"""
    import torch
    x, tpos, f0 = _check_vec(x, "x"), _check_vec(temporal_positions, "temporal_positions"), _check_vec(f0, "f0")
    h = _cached("d4c", (int(fs), x.device.index) + _opt_key(option), lambda: wb.D4C(fs, option))
    ap = torch.empty((f0.numel(), int(fft_size) // 2 + 1), dtype=torch.float64, device=x.device)
    wb._check(wb.lib().wb_d4c_compute_dev(h._h, x.data_ptr(), x.numel(), tpos.data_ptr(), f0.data_ptr(), f0.numel(),
                                          int(fft_size), ap.data_ptr(), _stream()), "wb_d4c_compute_dev")
    return ap


def synthesis(f0, spectrogram, aperiodicity, fs, frame_period=5.0, out_length=None, f0_upper_bound=0.0):
    """-> waveform; include/synthesis.hpp:47-51.  f0_upper_bound > 0 (e.g. Harvest's f0_ceil * 1.25) avoids one
    device->host read of the pulse count."""
    import torch
    f0, sp, ap = _check_vec(f0, "f0"), _check_vec(spectrogram, "spectrogram"), _check_vec(aperiodicity, "aperiodicity")
    if sp.shape != ap.shape or sp.shape[0] != f0.numel():
        raise ValueError("spectrogram / aperiodicity must be [len(f0)][fft_size/2+1]")
    fft_size = (sp.shape[1] - 1) * 2
    h = _cached("synthesis", (int(fs), fft_size, float(frame_period), f0.device.index),
                lambda: wb.Synthesis(fs, fft_size, frame_period))
    ny = wb.synthesis_length(f0.numel(), frame_period, fs) if out_length is None else int(out_length)
    y = torch.empty(ny, dtype=torch.float64, device=f0.device)
    wb._check(wb.lib().wb_synthesis_compute_dev(h._h, f0.data_ptr(), f0.numel(), sp.data_ptr(), ap.data_ptr(), ny, y.data_ptr(),
                                                float(f0_upper_bound), _stream()), "wb_synthesis_compute_dev")
    return y


def codec(kind, rows, fs, fft_size, number_of_dimensions=0):
    """kind: 'code_ap' | 'decode_ap' | 'code_sp' | 'decode_sp' on contiguous [frames][cols] tensors (include/codec.hpp:23-88)"""
    import torch
    rows = _check_vec(rows, "rows")
    k = {"code_ap": 0, "decode_ap": 1, "code_sp": 2, "decode_sp": 3}[kind]
    bins = int(fft_size) // 2 + 1
    cols = {0: wb.GetNumberOfAperiodicities(fs), 1: bins, 2: int(number_of_dimensions), 3: bins}[k]
    out = torch.empty((rows.shape[0], cols), dtype=torch.float64, device=rows.device)
    wb._check(wb.lib().wb_codec_dev(k, rows.data_ptr(), rows.shape[0], int(fs), int(fft_size), int(number_of_dimensions),
                                    out.data_ptr(), _stream()), "wb_codec_dev")
    return out


def to_float32(t):
    """fp64 -> fp32 narrowing on the device (round to nearest), for consumers that do not want doubles"""
    import torch
    t = _check_vec(t, "t")
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    wb._check(wb.lib().wb_f64_to_f32_dev(t.data_ptr(), t.numel(), out.data_ptr(), _stream()), "wb_f64_to_f32_dev")
    return out


def check_errors():
    """The functions above are asynchronous on torch's current stream and cannot report what a kernel flags on the
    device (e.g. synthesis() with an f0_upper_bound below the contour's maximum drops pulses).  This waits for the
    current stream and raises if any of the cached stage objects flagged an error since the last check."""
    kinds = {"harvest": "wb_harvest_last_error", "cheaptrick": "wb_cheaptrick_last_error", "d4c": "wb_d4c_last_error",
             "synthesis": "wb_synthesis_last_error"}
    for k, h in list(_objects.items()):
        kind = k[0]
        fn = kinds.get(kind)
        if fn is not None:
            wb._check(getattr(wb.lib(), fn)(h._h, _stream()), "device-side error of an earlier %s call" % kind)
