"""worldb200: host-side mirror of the world-class C++ interface over the C-ABI library.

Class and method names follow the reference's headers (/root/reference/include/*.hpp):
Harvest / CheapTrick / D4C / Synthesis objects with `compute(...)` plus their option
structs.  Everything computes on the GPU through libworldb200.so (hand-written CUDA for
sm_100a); there is NO CPU fallback -- importing works without a GPU, computing does not.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libworldb200.so")

WB_OK = 0
_ERR = {1: "CUDA error (or no GPU: this library has no CPU fallback)", 2: "bad argument", 3: "unsupported configuration"}


class WorldB200Error(RuntimeError):
    pass


class HarvestOption(ctypes.Structure):
    """include/harvest.hpp:16-28"""
    _fields_ = [("f0_floor", ctypes.c_double), ("f0_ceil", ctypes.c_double), ("frame_period", ctypes.c_double),
                ("target_fs", ctypes.c_double), ("channels_in_octave", ctypes.c_double), ("use_cos_table", ctypes.c_int)]

    def __init__(self, **kw):
        super().__init__()
        lib().wb_harvest_option_default(ctypes.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


class CheapTrickOption(ctypes.Structure):
    """include/cheaptrick.hpp:14-20"""
    _fields_ = [("q1", ctypes.c_double), ("f0_floor", ctypes.c_double), ("fft_size", ctypes.c_int)]

    def __init__(self, **kw):
        super().__init__()
        lib().wb_cheaptrick_option_default(ctypes.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


class D4COption(ctypes.Structure):
    """include/d4c.hpp:16-20"""
    _fields_ = [("threshold", ctypes.c_double)]

    def __init__(self, **kw):
        super().__init__()
        lib().wb_d4c_option_default(ctypes.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


_lib = None
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_double_pp = ctypes.POINTER(_c_double_p)


def _declare(L):
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    sig = {
        "wb_init": (ci, [ci]),
        "wb_version": (ctypes.c_char_p, []),
        "wb_device_synchronize": (ci, []),
        "wb_harvest_option_default": (None, [ctypes.POINTER(HarvestOption)]),
        "wb_cheaptrick_option_default": (None, [ctypes.POINTER(CheapTrickOption)]),
        "wb_d4c_option_default": (None, [ctypes.POINTER(D4COption)]),
        "wb_randn_reseed": (ci, []),
        "wb_randn_get_state": (ci, [ctypes.POINTER(ctypes.c_uint * 4)]),
        "wb_randn_set_state": (ci, [ctypes.POINTER(ctypes.c_uint * 4)]),
        "wb_randn_skip": (ci, [ctypes.c_ulonglong]),
        "wb_randn_fill": (ci, [vp, ci]),
        "wb_fft_r2c": (ci, [vp, ci, ci, vp]),
        "wb_fft_c2r": (ci, [vp, ci, ci, vp]),
        "wb_fft_c2c": (ci, [vp, ci, ci, ci, vp]),
        "wb_harvest_get_samples": (ci, [ci, ci, cd]),
        "wb_harvest_create": (ci, [ci, ctypes.POINTER(HarvestOption), ctypes.POINTER(vp)]),
        "wb_harvest_destroy": (None, [vp]),
        "wb_harvest_compute": (ci, [vp, vp, ci, vp, vp]),
        "wb_harvest_compute_dev": (ci, [vp, vp, ci, vp, vp, vp]),
        "wb_harvest_debug_read": (ci, [vp, ctypes.c_char_p, vp, ctypes.c_ulonglong]),
        "wb_cheaptrick_get_fft_size": (ci, [ci, cd]),
        "wb_cheaptrick_get_f0_floor": (cd, [ci, ci]),
        "wb_cheaptrick_create": (ci, [ci, ctypes.POINTER(CheapTrickOption), ctypes.POINTER(vp)]),
        "wb_cheaptrick_destroy": (None, [vp]),
        "wb_cheaptrick_fft_size": (ci, [vp]),
        "wb_cheaptrick_compute": (ci, [vp, vp, ci, vp, vp, ci, vp]),
        "wb_cheaptrick_compute_dev": (ci, [vp, vp, ci, vp, vp, ci, vp, vp]),
        "wb_get_number_of_aperiodicities": (ci, [ci]),
        "wb_code_aperiodicity": (ci, [vp, ci, ci, ci, vp]),
        "wb_decode_aperiodicity": (ci, [vp, ci, ci, ci, vp]),
        "wb_code_spectral_envelope": (ci, [vp, ci, ci, ci, ci, vp]),
        "wb_decode_spectral_envelope": (ci, [vp, ci, ci, ci, ci, vp]),
        "wb_codec_dev": (ci, [ci, vp, ci, ci, ci, ci, vp, vp]),
        "wb_pipeline_create": (ci, [ci, ctypes.POINTER(HarvestOption), ctypes.POINTER(CheapTrickOption),
                                    ctypes.POINTER(D4COption), ctypes.POINTER(vp)]),
        "wb_pipeline_destroy": (None, [vp]),
        "wb_pipeline_set_fresh_rng": (ci, [vp, ci]),
        "wb_pipeline_set_graph": (ci, [vp, ci]),
        "wb_pipeline_set_modification": (ci, [vp, cd, cd]),
        "wb_parameter_modification": (ci, [vp, ci, vp, ci, ci, cd, cd]),
        "wb_parameter_modification_dev": (ci, [vp, ci, vp, ci, ci, cd, cd, vp]),
        "wb_pipeline_fft_size": (ci, [vp]),
        "wb_pipeline_f0_length": (ci, [vp, ci]),
        "wb_pipeline_out_length": (ci, [vp, ci]),
        "wb_pipeline_run_dev": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, ci, vp]),
        "wb_pipeline_run": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, ci]),
        "wb_pipeline_debug_read": (ci, [vp, ctypes.c_char_p, vp, ctypes.c_ulonglong]),
        "wb_pipeline_stream_begin_dev": (ci, [vp, vp, ci, ci, vp]),
        "wb_pipeline_stream_begin_range_dev": (ci, [vp, vp, ci, ci, ci, ci, vp]),
        "wb_pipeline_stream_envelope_dev": (ci, [vp, vp, ci, vp, ci, ci, ci, vp, vp, vp]),
        "wb_decimate_length": (ci, [ci, ci]),
        "wb_decimate": (ci, [ctypes.POINTER(cd), ci, ci, ctypes.POINTER(cd)]),
        "wb_harvest_last_error": (ci, [vp, vp]),
        "wb_cheaptrick_last_error": (ci, [vp, vp]),
        "wb_d4c_last_error": (ci, [vp, vp]),
        "wb_synthesis_last_error": (ci, [vp, vp]),
        "wb_pipeline_last_error": (ci, [vp, vp]),
        "wb_pipeline_set_stream_f0_bound": (ci, [vp, cd]),
        "wb_pipeline_stream_lovetrain_dev": (ci, [vp, vp, ci, vp, ci, ci, ci, vp, vp]),
        "wb_pipeline_stream_cheaptrick_dev": (ci, [vp, vp, ci, vp, ci, ci, ci, vp, vp]),
        "wb_pipeline_stream_aperiodicity_dev": (ci, [vp, vp, ci, vp, vp, ci, ci, ci, vp, vp]),
        "wb_pipeline_stream_synthesis_dev": (ci, [vp, ci, vp, vp, ci, ci, ci, ci, ci, vp, vp]),
        "wb_pipeline_stream_end_dev": (ci, [vp, vp]),
        "wb_synthesis_stream_create": (ci, [ci, ci, cd, cd, ctypes.POINTER(vp)]),
        "wb_synthesis_stream_destroy": (None, [vp]),
        "wb_synthesis_stream_push": (ci, [vp, vp, vp, vp, ci, vp, ci, ctypes.POINTER(ci)]),
        "wb_synthesis_stream_finish": (ci, [vp, ci, vp, ci, ctypes.POINTER(ci)]),
        "wb_pipeline_run_pcm16": (ci, [vp, vp, ci, vp, ci]),
        "wb_pipeline_run_f32": (ci, [vp, vp, ci, vp, vp, vp, vp, ci]),
        "wb_pcm16_to_f64_dev": (ci, [vp, ci, vp, vp]),
        "wb_f64_to_pcm16_dev": (ci, [vp, ci, vp, vp]),
        "wb_f64_to_f32_dev": (ci, [vp, ctypes.c_ulonglong, vp, vp]),
        "wb_write_f0": (ci, [ctypes.c_char_p, ci, cd, vp, vp, ci]),
        "wb_read_f0": (ci, [ctypes.c_char_p, vp, vp]),
        "wb_get_header_information": (cd, [ctypes.c_char_p, ctypes.c_char_p]),
        "wb_write_spectral_envelope": (ci, [ctypes.c_char_p, ci, ci, cd, ci, ci, vp]),
        "wb_read_spectral_envelope": (ci, [ctypes.c_char_p, vp]),
        "wb_write_aperiodicity": (ci, [ctypes.c_char_p, ci, ci, cd, ci, ci, vp]),
        "wb_read_aperiodicity": (ci, [ctypes.c_char_p, vp]),
        "wb_write_parameter_matrix": (ci, [ci, ctypes.c_char_p, ci, ci, cd, ci, ci, vp, ctypes.c_longlong]),
        "wb_read_parameter_matrix": (ci, [ci, ctypes.c_char_p, vp, ctypes.c_longlong]),
        "wb_wavwrite": (ci, [vp, ci, ci, ci, ctypes.c_char_p]),
        "wb_get_audio_length": (ci, [ctypes.c_char_p]),
        "wb_wavread": (ci, [ctypes.c_char_p, ctypes.POINTER(ci), ctypes.POINTER(ci), vp]),
        "wb_wavread_pcm16": (ci, [ctypes.c_char_p, ctypes.POINTER(ci), vp]),
        "wb_launch_count": (ctypes.c_ulonglong, []),
        "wb_measure_fp64_peak": (ci, [ctypes.POINTER(cd)]),
        "wb_stream": (vp, []),
        "wb_profile_enable": (None, [ci]),
        "wb_profile_reset": (None, []),
        "wb_profile_collect": (ci, []),
        "wb_profile_query": (ci, [ctypes.c_char_p, ctypes.POINTER(cd), ctypes.POINTER(ci)]),
        "wb_profile_names": (ci, [ctypes.c_char_p, ci]),
        "wb_synthesis_create": (ci, [ci, ci, cd, ctypes.POINTER(vp)]),
        "wb_synthesis_destroy": (None, [vp]),
        "wb_synthesis_compute": (ci, [vp, vp, ci, vp, vp, ci, vp]),
        "wb_synthesis_compute_dev": (ci, [vp, vp, ci, vp, vp, ci, vp, cd, vp]),
        "wb_d4c_create": (ci, [ci, ctypes.POINTER(D4COption), ctypes.POINTER(vp)]),
        "wb_d4c_destroy": (None, [vp]),
        "wb_d4c_compute": (ci, [vp, vp, ci, vp, vp, ci, ci, vp]),
        "wb_d4c_compute_dev": (ci, [vp, vp, ci, vp, vp, ci, ci, vp, vp]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(L, name):
            continue  # tests/test_abi.py checks the export list against include/worldb200.h
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    """Load libworldb200.so (built by build.py / __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WorldB200Error("%s is missing: run `python __graft_entry__.py` (build()) first; "
                                 "there is no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _check(rc, what):
    if rc != WB_OK:
        raise WorldB200Error("%s failed: %s (status %d)" % (what, _ERR.get(rc, "unknown"), rc))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _out_array(arr, shape):
    """A caller-owned output array (checked) or a fresh one."""
    if arr is None:
        return np.empty(shape, dtype=np.float64)
    if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"] and
            arr.shape == tuple(shape)):
        raise ValueError("output array must be C-contiguous float64 of shape %r" % (tuple(shape),))
    return arr


_row_pointer_cache = {}


def _row_pointers(mat):
    """A C `double **` (array of row pointers) into a C-contiguous 2-D float64 array.  The pointer table is a
    function of (base address, rows, row stride) only: callers that reuse their matrices (test/test.cpp allocates
    them once) get the cached table."""
    assert mat.flags["C_CONTIGUOUS"] and mat.dtype == np.float64 and mat.ndim == 2
    key = (mat.ctypes.data, mat.shape[0], mat.strides[0])
    rows = _row_pointer_cache.get(key)
    if rows is None:
        if len(_row_pointer_cache) >= 32:
            _row_pointer_cache.clear()
        addr = key[0] + np.arange(mat.shape[0], dtype=np.uint64) * np.uint64(mat.strides[0])
        rows = _row_pointer_cache[key] = np.ascontiguousarray(addr, dtype=np.uint64)
    return rows


# ---- randn stream -----------------------------------------------------------------------
def randn_reseed():
    _check(lib().wb_randn_reseed(), "wb_randn_reseed")


def randn_get_state():
    s = (ctypes.c_uint * 4)()
    _check(lib().wb_randn_get_state(ctypes.byref(s)), "wb_randn_get_state")
    return tuple(int(v) for v in s)


def randn_set_state(state):
    s = (ctypes.c_uint * 4)(*state)
    _check(lib().wb_randn_set_state(ctypes.byref(s)), "wb_randn_set_state")


def randn_skip(n):
    _check(lib().wb_randn_skip(int(n)), "wb_randn_skip")


def randn(n):
    out = np.empty(int(n), dtype=np.float64)
    _check(lib().wb_randn_fill(out.ctypes.data, int(n)), "wb_randn_fill")
    return out


# ---- stand-alone FFT (include/world_fft.hpp) ---------------------------------------------
FFT_FORWARD, FFT_BACKWARD = 1, 2


def fft_r2c(x):
    """x: [batch][n] real -> [batch][n/2+1] complex, reference convention (e^{+i}, unnormalised)."""
    x = np.atleast_2d(_f64(x))
    b, n = x.shape
    out = np.empty((b, n // 2 + 1), dtype=np.complex128)
    _check(lib().wb_fft_r2c(x.ctypes.data, n, b, out.ctypes.data), "wb_fft_r2c")
    return out


def fft_c2r(X, n):
    X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.complex128)
    b = X.shape[0]
    assert X.shape[1] == n // 2 + 1
    out = np.empty((b, n), dtype=np.float64)
    _check(lib().wb_fft_c2r(X.ctypes.data, n, b, out.ctypes.data), "wb_fft_c2r")
    return out


def fft_c2c(X, sign):
    X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.complex128)
    b, n = X.shape
    out = np.empty((b, n), dtype=np.complex128)
    _check(lib().wb_fft_c2c(X.ctypes.data, n, b, int(sign), out.ctypes.data), "wb_fft_c2c")
    return out


# ---- Harvest (include/harvest.hpp:31-44) ---------------------------------------------------
class Harvest:
    def __init__(self, fs, option=None):
        self._h = ctypes.c_void_p()
        self.fs = int(fs)
        self.option = option if option is not None else HarvestOption()
        _check(lib().wb_harvest_create(self.fs, ctypes.byref(self.option), ctypes.byref(self._h)), "wb_harvest_create")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_harvest_destroy(self._h)
            self._h = None

    def getSamples(self, fs, x_length, frame_period=None):
        fp = self.option.frame_period if frame_period is None else frame_period
        return lib().wb_harvest_get_samples(int(fs), int(x_length), float(fp))

    def compute(self, x, temporal_positions=None, f0=None):
        """-> (temporal_positions, f0).  Like the reference (test/test.cpp:102-104) the caller may own the
        output arrays: pass float64 arrays of getSamples() entries to have them filled in place."""
        x = _f64(x)
        n = self.getSamples(self.fs, len(x))
        tpos = _out_array(temporal_positions, (n,))
        f0 = _out_array(f0, (n,))
        _check(lib().wb_harvest_compute(self._h, x.ctypes.data, len(x), tpos.ctypes.data, f0.ctypes.data),
               "wb_harvest_compute")
        return tpos, f0

    def check_errors(self, stream=0):
        """After asynchronous wb_harvest_compute_dev calls on `stream`: waits for it and raises if a kernel flagged
        a condition it could not handle; the flag is cleared."""
        _check(lib().wb_harvest_last_error(self._h, stream or None), "device-side error of an earlier Harvest call")

    def debug_read(self, name, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        _check(lib().wb_harvest_debug_read(self._h, name.encode(), out.ctypes.data, out.nbytes), "wb_harvest_debug_read")
        return out


# ---- CheapTrick (include/cheaptrick.hpp:23-38) ---------------------------------------------
class CheapTrick:
    def __init__(self, fs, option=None):
        self._h = ctypes.c_void_p()
        self.fs = int(fs)
        opt = ctypes.byref(option) if option is not None else None
        _check(lib().wb_cheaptrick_create(self.fs, opt, ctypes.byref(self._h)), "wb_cheaptrick_create")
        self.fft_size = lib().wb_cheaptrick_fft_size(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_cheaptrick_destroy(self._h)
            self._h = None

    @staticmethod
    def getFFTSizeForCheapTrick(fs, f0_floor):
        return lib().wb_cheaptrick_get_fft_size(int(fs), float(f0_floor))

    @staticmethod
    def getF0FloorForCheapTrick(fs, fft_size):
        return lib().wb_cheaptrick_get_f0_floor(int(fs), int(fft_size))

    def compute(self, x, temporal_positions, f0, spectrogram=None):
        """-> spectrogram [f0_length][fft_size/2+1] (host arrays in, host array out; `spectrogram` may be a
        caller-owned output array, test/test.cpp:146-149)."""
        x, tpos, f0 = _f64(x), _f64(temporal_positions), _f64(f0)
        sp = _out_array(spectrogram, (len(f0), self.fft_size // 2 + 1))
        rows = _row_pointers(sp)
        _check(lib().wb_cheaptrick_compute(self._h, x.ctypes.data, len(x), tpos.ctypes.data, f0.ctypes.data,
                                           len(f0), rows.ctypes.data), "wb_cheaptrick_compute")
        return sp


# ---- D4C (include/d4c.hpp:23-36) ---------------------------------------------------------
class D4C:
    def __init__(self, fs, option=None):
        self._h = ctypes.c_void_p()
        self.fs = int(fs)
        opt = ctypes.byref(option) if option is not None else None
        _check(lib().wb_d4c_create(self.fs, opt, ctypes.byref(self._h)), "wb_d4c_create")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_d4c_destroy(self._h)
            self._h = None

    def compute(self, x, temporal_positions, f0, fft_size, aperiodicity=None):
        """-> aperiodicity [f0_length][fft_size/2+1] (`aperiodicity` may be a caller-owned output array,
        test/test.cpp:170-173)."""
        x, tpos, f0 = _f64(x), _f64(temporal_positions), _f64(f0)
        ap = _out_array(aperiodicity, (len(f0), int(fft_size) // 2 + 1))
        rows = _row_pointers(ap)
        _check(lib().wb_d4c_compute(self._h, x.ctypes.data, len(x), tpos.ctypes.data, f0.ctypes.data,
                                    len(f0), int(fft_size), rows.ctypes.data), "wb_d4c_compute")
        return ap


# ---- Synthesis (include/synthesis.hpp:29-51) -----------------------------------------------
class Synthesis:
    def __init__(self, fs, fft_size, frame_period):
        self._h = ctypes.c_void_p()
        self.fs, self.fft_size, self.frame_period = int(fs), int(fft_size), float(frame_period)
        _check(lib().wb_synthesis_create(self.fs, self.fft_size, self.frame_period, ctypes.byref(self._h)),
               "wb_synthesis_create")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_synthesis_destroy(self._h)
            self._h = None

    def compute(self, f0, spectrogram, aperiodicity, out_length, out=None):
        """-> waveform of out_length samples (`out` may be a caller-owned output array)"""
        f0 = _f64(f0)
        sp, ap = _f64(spectrogram), _f64(aperiodicity)
        assert sp.shape == ap.shape == (len(f0), self.fft_size // 2 + 1)
        out = _out_array(out, (int(out_length),))
        rs, ra = _row_pointers(sp), _row_pointers(ap)
        _check(lib().wb_synthesis_compute(self._h, f0.ctypes.data, len(f0), rs.ctypes.data, ra.ctypes.data,
                                          len(out), out.ctypes.data), "wb_synthesis_compute")
        return out


class SynthesisStream:
    """Streaming Synthesis: push frames in pieces, get the samples that became final; the concatenated output
    equals Synthesis.compute on all frames bit for bit (phase sum, pulses and randn() position are carried)."""

    def __init__(self, fs, fft_size, frame_period, f0_upper_bound=0.0):
        self._h = ctypes.c_void_p()
        self.fs, self.fft_size, self.frame_period = int(fs), int(fft_size), float(frame_period)
        _check(lib().wb_synthesis_stream_create(self.fs, self.fft_size, self.frame_period, float(f0_upper_bound),
                                                ctypes.byref(self._h)), "wb_synthesis_stream_create")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_synthesis_stream_destroy(self._h)
            self._h = None

    def push(self, f0, spectrogram, aperiodicity, out_capacity=None):
        f0, sp, ap = _f64(f0), _f64(spectrogram), _f64(aperiodicity)
        n = len(f0)
        assert sp.shape == ap.shape == (n, self.fft_size // 2 + 1)
        cap = int(out_capacity) if out_capacity is not None else int((n + 2) * self.frame_period / 1000.0 * self.fs) + 2 * self.fft_size
        out = np.empty(cap, dtype=np.float64)
        got = ctypes.c_int()
        _check(lib().wb_synthesis_stream_push(self._h, f0.ctypes.data, sp.ctypes.data, ap.ctypes.data, n, out.ctypes.data, cap,
                                              ctypes.byref(got)), "wb_synthesis_stream_push")
        return out[:got.value]

    def finish(self, out_length, out_capacity=None):
        pieces = []
        cap = int(out_capacity) if out_capacity is not None else max(1, int(out_length))
        while True:
            out = np.empty(cap, dtype=np.float64)
            got = ctypes.c_int()
            _check(lib().wb_synthesis_stream_finish(self._h, int(out_length), out.ctypes.data, cap, ctypes.byref(got)),
                   "wb_synthesis_stream_finish")
            if got.value == 0:
                break
            pieces.append(out[:got.value])
        return np.concatenate(pieces) if pieces else np.empty(0)


def synthesis_length(f0_length, frame_period, fs):
    """test/test.cpp:362-363"""
    return int((f0_length - 1) * frame_period / 1000.0 * fs) + 1


# ---- parameter modification (test/test.cpp:201-243) -------------------------------------------
def ParameterModification(f0, spectrogram, fs, fft_size, f0_shift=None, ratio=None):
    """In place on host arrays, like the reference demo: f0 *= f0_shift, spectrogram stretched by `ratio`
    (None = leave alone).  Returns (f0, spectrogram)."""
    f0 = _out_array(f0, (len(f0),))
    sp = _out_array(spectrogram, (len(f0), int(fft_size) // 2 + 1))
    rows = _row_pointers(sp)
    _check(lib().wb_parameter_modification(f0.ctypes.data, len(f0), rows.ctypes.data, int(fs), int(fft_size),
                                           float("nan") if f0_shift is None else float(f0_shift),
                                           0.0 if ratio is None else float(ratio)), "wb_parameter_modification")
    return f0, sp


# ---- whole chain, device resident (test/test.cpp:288-384) ------------------------------------
class Pipeline:
    """Harvest -> CheapTrick -> D4C -> Synthesis with every intermediate kept in HBM."""

    def __init__(self, fs, harvest_option=None, cheaptrick_option=None, d4c_option=None):
        self._h = ctypes.c_void_p()
        self.fs = int(fs)
        ref = lambda o: ctypes.byref(o) if o is not None else None
        _check(lib().wb_pipeline_create(self.fs, ref(harvest_option), ref(cheaptrick_option), ref(d4c_option),
                                        ctypes.byref(self._h)), "wb_pipeline_create")
        self.fft_size = lib().wb_pipeline_fft_size(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.wb_pipeline_destroy(self._h)
            self._h = None

    def set_fresh_rng(self, fresh=True):
        """Batch mode: every run starts its own randn() stream at the reference's seed."""
        _check(lib().wb_pipeline_set_fresh_rng(self._h, 1 if fresh else 0), "wb_pipeline_set_fresh_rng")

    def set_graph(self, use_graph=True):
        """Replay a captured CUDA graph for repeated run_dev calls with identical arguments."""
        _check(lib().wb_pipeline_set_graph(self._h, 1 if use_graph else 0), "wb_pipeline_set_graph")

    def set_modification(self, f0_shift=None, ratio=None):
        """Apply the demo's ParameterModification (test/test.cpp:201-243) between analysis and synthesis:
        F0 scaling by `f0_shift`, spectral stretching by `ratio` (None = off).  The returned f0 / sp are the
        modified ones."""
        _check(lib().wb_pipeline_set_modification(self._h, float("nan") if f0_shift is None else float(f0_shift),
                                                  0.0 if ratio is None else float(ratio)), "wb_pipeline_set_modification")

    def f0_length(self, x_length):
        return lib().wb_pipeline_f0_length(self._h, int(x_length))

    def out_length(self, x_length):
        return lib().wb_pipeline_out_length(self._h, int(x_length))

    def run(self, x, want_params=True, out=None):
        """host x -> dict(tpos, f0, sp, ap, y) on the host.  `out`: a dict of caller-owned C-contiguous float64 arrays
        to fill instead (keys y and, with want_params, tpos, f0, sp, ap).  When all of them are page-locked (e.g.
        torch pin_memory().numpy()) every result is downloaded as soon as its stage is done, beside the rest of the
        chain."""
        x = _f64(x)
        L, bins, ny = self.f0_length(len(x)), self.fft_size // 2 + 1, self.out_length(len(x))
        ptr = lambda a: a.ctypes.data if a is not None else None
        if out is not None:
            shapes = {"y": (ny,)}
            if want_params:
                shapes.update(tpos=(L,), f0=(L,), sp=(L, bins), ap=(L, bins))
            out = {k: _out_array(out[k], shp) for k, shp in shapes.items()}
            y = out["y"]
        else:
            y = np.empty(ny, dtype=np.float64)
            out = {"y": y}
            if want_params:
                out.update(tpos=np.empty(L), f0=np.empty(L), sp=np.empty((L, bins)), ap=np.empty((L, bins)))
        _check(lib().wb_pipeline_run(self._h, x.ctypes.data, len(x), ptr(out.get("tpos")), ptr(out.get("f0")),
                                     ptr(out.get("sp")), ptr(out.get("ap")), y.ctypes.data, ny), "wb_pipeline_run")
        return out

    def run_pcm16(self, pcm):
        """wav in -> wav out: host int16 samples in, re-synthesised host int16 samples out (the sample-format
        conversions of the reference's wavread / wavwrite run on the device)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ny = self.out_length(len(pcm))
        out = np.empty(ny, dtype=np.int16)
        _check(lib().wb_pipeline_run_pcm16(self._h, pcm.ctypes.data, len(pcm), out.ctypes.data, ny), "wb_pipeline_run_pcm16")
        return out

    def run_f32(self, x, want_params=True, want_y=True):
        """host x -> dict(f0, sp, ap, y) as float32 host arrays (narrowed on the device; arithmetic stays fp64)"""
        x = _f64(x)
        L, bins = self.f0_length(len(x)), self.fft_size // 2 + 1
        ny = self.out_length(len(x)) if want_y else 0
        out = {}
        if want_params:
            out.update(f0=np.empty(L, np.float32), sp=np.empty((L, bins), np.float32), ap=np.empty((L, bins), np.float32))
        if want_y:
            out["y"] = np.empty(ny, np.float32)
        ptr = lambda a: a.ctypes.data if a is not None else None
        _check(lib().wb_pipeline_run_f32(self._h, x.ctypes.data, len(x), ptr(out.get("f0")), ptr(out.get("sp")),
                                         ptr(out.get("ap")), ptr(out.get("y")), ny), "wb_pipeline_run_f32")
        return out

    def debug_read(self, name, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        _check(lib().wb_pipeline_debug_read(self._h, name.encode(), out.ctypes.data, out.nbytes), "wb_pipeline_debug_read")
        return out

    def run_dev(self, d_x, x_length, d_y=0, y_length=None, stream=0, d_tpos=0, d_f0=0, d_sp=0, d_ap=0):
        """device pointers (ints); asynchronous on `stream` (0 = the library's stream)"""
        ny = self.out_length(x_length) if y_length is None else int(y_length)
        _check(lib().wb_pipeline_run_dev(self._h, d_x, int(x_length), d_tpos or None, d_f0 or None, d_sp or None,
                                         d_ap or None, d_y or None, ny, stream or None), "wb_pipeline_run_dev")

    def check_errors(self, stream=0):
        """After asynchronous calls (run_dev, the stream_* calls): waits for `stream` and raises if a kernel flagged a
        condition it could not handle (e.g. more pulses than the f0 bound allows for); the flag is cleared."""
        _check(lib().wb_pipeline_last_error(self._h, stream or None), "device-side error of an earlier asynchronous call")

    def set_stream_f0_bound(self, f0_upper_bound):
        _check(lib().wb_pipeline_set_stream_f0_bound(self._h, float(f0_upper_bound)), "wb_pipeline_set_stream_f0_bound")


# ---- measurement hooks ------------------------------------------------------------------------
def launch_count():
    return int(lib().wb_launch_count())


def measure_fp64_peak():
    """fp64 multiply-add throughput of the current GPU in TFLOP/s (measured, a few milliseconds)"""
    v = ctypes.c_double()
    _check(lib().wb_measure_fp64_peak(ctypes.byref(v)), "wb_measure_fp64_peak")
    return v.value


def stream_handle():
    return lib().wb_stream()


def device_synchronize():
    _check(lib().wb_device_synchronize(), "wb_device_synchronize")


def profile(enable):
    lib().wb_profile_enable(1 if enable else 0)


def profile_reset():
    lib().wb_profile_reset()


def profile_results():
    """{kernel name: (total_ms, launches)} of everything launched while profiling was enabled."""
    _check(lib().wb_profile_collect(), "wb_profile_collect")
    buf = ctypes.create_string_buffer(1 << 16)
    _check(lib().wb_profile_names(buf, len(buf)), "wb_profile_names")
    res = {}
    for name in [n for n in buf.value.decode().split(";") if n]:
        ms, cnt = ctypes.c_double(), ctypes.c_int()
        _check(lib().wb_profile_query(name.encode(), ctypes.byref(ms), ctypes.byref(cnt)), "wb_profile_query")
        res[name] = (ms.value, cnt.value)
    return res


# ---- codec (include/codec.hpp:23-88) ---------------------------------------------------------
def GetNumberOfAperiodicities(fs):
    return lib().wb_get_number_of_aperiodicities(int(fs))


def _codec(fn, name, src, out_cols, *args):
    src = _f64(src)
    out = np.empty((src.shape[0], out_cols), dtype=np.float64)
    ri, ro = _row_pointers(src), _row_pointers(out)
    _check(fn(ri.ctypes.data, src.shape[0], *args, ro.ctypes.data), name)
    return out


def CodeAperiodicity(aperiodicity, fs, fft_size):
    return _codec(lib().wb_code_aperiodicity, "wb_code_aperiodicity", aperiodicity, GetNumberOfAperiodicities(fs),
                  int(fs), int(fft_size))


def DecodeAperiodicity(coded_aperiodicity, fs, fft_size):
    return _codec(lib().wb_decode_aperiodicity, "wb_decode_aperiodicity", coded_aperiodicity, int(fft_size) // 2 + 1,
                  int(fs), int(fft_size))


def CodeSpectralEnvelope(spectrogram, fs, fft_size, number_of_dimensions):
    return _codec(lib().wb_code_spectral_envelope, "wb_code_spectral_envelope", spectrogram, int(number_of_dimensions),
                  int(fs), int(fft_size), int(number_of_dimensions))


def DecodeSpectralEnvelope(coded_spectral_envelope, fs, fft_size, number_of_dimensions):
    return _codec(lib().wb_decode_spectral_envelope, "wb_decode_spectral_envelope", coded_spectral_envelope,
                  int(fft_size) // 2 + 1, int(fs), int(fft_size), int(number_of_dimensions))


# ---- the reference's file formats (tools/parameterio.hpp, tools/audioio.hpp) ----------------------------
def WriteF0(filename, f0_length, frame_period, temporal_positions, f0, text_flag=0):
    tpos, f0 = _f64(temporal_positions), _f64(f0)
    _check(lib().wb_write_f0(os.fsencode(filename), int(f0_length), float(frame_period), tpos.ctypes.data,
                             f0.ctypes.data, int(text_flag)), "wb_write_f0")


def GetHeaderInformation(filename, parameter):
    return lib().wb_get_header_information(os.fsencode(filename), parameter.encode())


def ReadF0(filename):
    """-> (temporal_positions, f0)"""
    n = int(GetHeaderInformation(filename, "NOF "))
    tpos, f0 = np.empty(n), np.empty(n)
    _check(lib().wb_read_f0(os.fsencode(filename), tpos.ctypes.data, f0.ctypes.data), "wb_read_f0")
    return tpos, f0


def _write_matrix(fn, name, filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, mat):
    mat = _f64(mat)
    rows = _row_pointers(mat)
    _check(fn(os.fsencode(filename), int(fs), int(f0_length), float(frame_period), int(fft_size),
              int(number_of_dimensions), rows.ctypes.data), name)


def _read_matrix(fn, name, filename):
    n, fft, nod = (int(GetHeaderInformation(filename, k)) for k in ("NOF ", "FFT ", "NOD "))
    mat = np.empty((n, nod if nod else fft // 2 + 1))
    rows = _row_pointers(mat)
    _check(fn(os.fsencode(filename), rows.ctypes.data), name)
    return mat


def WriteSpectralEnvelope(filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, spectrogram):
    _write_matrix(lib().wb_write_spectral_envelope, "wb_write_spectral_envelope", filename, fs, f0_length, frame_period,
                  fft_size, number_of_dimensions, spectrogram)


def ReadSpectralEnvelope(filename):
    return _read_matrix(lib().wb_read_spectral_envelope, "wb_read_spectral_envelope", filename)


def WriteAperiodicity(filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, aperiodicity):
    _write_matrix(lib().wb_write_aperiodicity, "wb_write_aperiodicity", filename, fs, f0_length, frame_period,
                  fft_size, number_of_dimensions, aperiodicity)


def ReadAperiodicity(filename):
    return _read_matrix(lib().wb_read_aperiodicity, "wb_read_aperiodicity", filename)


def wavwrite(x, fs, nbit, filename):
    x = _f64(x)
    _check(lib().wb_wavwrite(x.ctypes.data, len(x), int(fs), int(nbit), os.fsencode(filename)), "wb_wavwrite")


def GetAudioLength(filename):
    return lib().wb_get_audio_length(os.fsencode(filename))


def wavread(filename):
    """-> (x, fs, nbit)"""
    n = GetAudioLength(filename)
    if n <= 0:
        raise WorldB200Error("wavread: cannot read %r (GetAudioLength = %d)" % (filename, n))
    x = np.empty(n)
    fs, nbit = ctypes.c_int(), ctypes.c_int()
    _check(lib().wb_wavread(os.fsencode(filename), ctypes.byref(fs), ctypes.byref(nbit), x.ctypes.data), "wb_wavread")
    return x, fs.value, nbit.value


def wavread_pcm16(filename):
    """-> (int16 samples, fs)"""
    n = GetAudioLength(filename)
    if n <= 0:
        raise WorldB200Error("wavread_pcm16: cannot read %r (GetAudioLength = %d)" % (filename, n))
    pcm = np.empty(n, dtype=np.int16)
    fs = ctypes.c_int()
    _check(lib().wb_wavread_pcm16(os.fsencode(filename), ctypes.byref(fs), pcm.ctypes.data), "wb_wavread_pcm16")
    return pcm, fs.value


class BatchPipeline:
    """Independent utterances (BASELINE configs[2]) on `n_streams` concurrent pipelines of one GPU.

    Each utterance is processed exactly like one reference process (fresh randn() stream).  Inputs and
    outputs are torch CUDA tensors; utterance i runs on stream i % n_streams, so the small,
    latency-bound kernels of one utterance overlap with the wide kernels of the others.

    A pipeline keeps persistent input / output buffers per utterance length, so that every utterance of a
    length it has seen before is ONE replay of the captured CUDA graph of the whole chain (~50 kernels) between a
    device copy of the waveform in and device copies of the results out: the batch is bound by the GPU, not by
    the host's launch rate."""

    def __init__(self, fs, n_streams=8, harvest_option=None, cheaptrick_option=None, d4c_option=None):
        import torch
        self._torch = torch
        self.pipes = [Pipeline(fs, harvest_option, cheaptrick_option, d4c_option) for _ in range(n_streams)]
        for p in self.pipes:
            p.set_fresh_rng(True)
            p.set_graph(True)
        self.streams = [torch.cuda.Stream() for _ in range(n_streams)]
        self.fft_size = self.pipes[0].fft_size
        self._bufs = [dict() for _ in range(n_streams)]   # per pipeline: utterance length -> persistent tensors

    def _slot(self, k, n, device):
        torch = self._torch
        b = self._bufs[k].get(n)
        if b is None:
            pl = self.pipes[k]
            L, ny, bins = pl.f0_length(n), pl.out_length(n), self.fft_size // 2 + 1
            f = dict(dtype=torch.float64, device=device)
            b = {"x": torch.empty(n, **f), "tpos": torch.empty(L, **f), "f0": torch.empty(L, **f),
                 "sp": torch.empty((L, bins), **f), "ap": torch.empty((L, bins), **f), "y": torch.empty(ny, **f)}
            self._bufs[k][n] = b
        return b

    def run(self, xs, copy_out=True):
        """xs: list of 1-D float64 CUDA tensors -> list of dict(tpos, f0, sp, ap, y) of CUDA tensors.
        copy_out=False returns views of the pipelines' persistent buffers instead of copies: they are only valid
        until the same pipeline processes its next utterance (for consumers that reduce the results on the fly)."""
        torch = self._torch
        outs = []
        cur = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(cur)
        for i, x in enumerate(xs):
            k = i % len(self.pipes)
            pl, st = self.pipes[k], self.streams[k]
            n = x.numel()
            b = self._slot(k, n, x.device)
            with torch.cuda.stream(st):
                b["x"].copy_(x, non_blocking=True)
                pl.run_dev(b["x"].data_ptr(), n, d_y=b["y"].data_ptr(), y_length=b["y"].numel(), stream=st.cuda_stream,
                           d_tpos=b["tpos"].data_ptr(), d_f0=b["f0"].data_ptr(), d_sp=b["sp"].data_ptr(),
                           d_ap=b["ap"].data_ptr())
                if copy_out:
                    outs.append({key: b[key].clone() for key in ("tpos", "f0", "sp", "ap", "y")})
                else:
                    outs.append({key: b[key] for key in ("tpos", "f0", "sp", "ap", "y")})
        for st in self.streams:
            cur.wait_stream(st)
        return outs

    def check_errors(self):
        """Waits for the pipelines' streams and raises if a kernel of an earlier run() flagged an error."""
        for pl, st in zip(self.pipes, self.streams):
            pl.check_errors(st.cuda_stream)
