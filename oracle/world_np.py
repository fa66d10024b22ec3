"""TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference algorithms in numpy.

Parity status: every function here is pinned against the reference itself
(oracle/_ref/refrun, the reference's own sources compiled by oracle/Makefile) by
tests/test_oracle_cpu.py and the committed fixtures in tests/golden/.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module; the
product (world-class_b200/) never does.

All citations are file:line under /root/reference.
"""
import numpy as np

M32 = 0xFFFFFFFF
RANDN_SEED = (123456789, 362436069, 521288629, 88675123)  # world_matlabfunctions.cpp:244-247


def randn_stream(n, state=RANDN_SEED):
    """The next n values of randn() and the state afterwards.

    src/world_matlabfunctions.cpp:243-264: one partial shift (x<-y, y<-z, z<-w; the first
    `t` is dead) and then 12 xorshift128 steps whose (w >> 4) are summed in uint32.
    """
    x, y, z, w = state
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        x, y, z = y, z, w
        tmp = 0
        for _ in range(12):
            t = (x ^ (x << 11)) & M32
            x, y, z = y, z, w
            w = ((w ^ (w >> 19)) ^ (t ^ (t >> 8))) & M32
            tmp = (tmp + (w >> 4)) & M32
        out[i] = tmp / 268435456.0 - 6.0
    return out, (x, y, z, w)


# ---- FFT wrapper semantics (src/world_fft.cpp:31-77) ---------------------------------------
def fft_r2c(x):
    """fft_plan_dft_r2c_1d + fft_execute: X[k] = sum x[n] e^{+2 pi i n k / N}, k = 0..N/2."""
    return np.conj(np.fft.rfft(x, axis=-1))


def fft_c2r(X, n):
    """fft_plan_dft_c2r_1d + fft_execute: unnormalised, e^{-i}; Im X[0], Im X[N/2] ignored."""
    X = np.array(X, dtype=np.complex128, copy=True)
    X[..., 0] = X[..., 0].real
    X[..., -1] = X[..., -1].real
    return n * np.fft.irfft(np.conj(X), n=n, axis=-1)


def fft_c2c(x, sign):
    """fft_plan_dft_1d: FFT_FORWARD (1) = e^{+i}, FFT_BACKWARD (2) = e^{-i}; unnormalised."""
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[-1]
    return n * np.fft.ifft(x, axis=-1) if sign == 1 else np.fft.fft(x, axis=-1)


def matlab_round(x):
    """src/world_matlabfunctions.cpp:212-214 (half away from zero, via int truncation)."""
    return int(x + 0.5) if x > 0 else int(x - 0.5)


# =============================================================================================
# C helper (oracle/world_c.c): sequential generator and phase sum
# =============================================================================================
import ctypes as _ct
import os as _os
import subprocess as _sp

_HERE = _os.path.dirname(_os.path.abspath(__file__))
_CLIB = None


def build_c_helper(force=False):
    """gcc -O2 -shared oracle/world_c.c -> oracle/libworldoracle.so (test infrastructure)."""
    src, lib = _os.path.join(_HERE, "world_c.c"), _os.path.join(_HERE, "libworldoracle.so")
    if force or not _os.path.exists(lib) or _os.path.getmtime(lib) < _os.path.getmtime(src):
        _sp.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", lib, src, "-lm"])
    return lib


def _clib():
    global _CLIB
    if _CLIB is None:
        _CLIB = _ct.CDLL(build_c_helper())
        _CLIB.oracle_randn_fill.argtypes = [_ct.POINTER(_ct.c_uint32 * 4), _ct.c_long, _ct.c_void_p]
        _CLIB.oracle_pulse_locations.argtypes = [_ct.c_void_p, _ct.c_int, _ct.c_int, _ct.c_void_p, _ct.c_void_p]
        _CLIB.oracle_pulse_locations.restype = _ct.c_int
    return _CLIB


class RandnStream:
    """The process-global randn() stream of the reference (one object == one process)."""

    def __init__(self, state=RANDN_SEED):
        self.state = (_ct.c_uint32 * 4)(*state)
        self.calls = 0

    def take(self, n):
        out = np.empty(int(n), dtype=np.float64)
        _clib().oracle_randn_fill(_ct.byref(self.state), int(n), out.ctypes.data)
        self.calls += int(n)
        return out


# =============================================================================================
# MATLAB-like helpers (src/world_matlabfunctions.cpp)
# =============================================================================================
def interp1(x, y, xi):
    """:136-182: histc index = first knot above xi, clamped to [1, len(x)-1]; linear extrapolation."""
    x, y, xi = np.asarray(x, float), np.asarray(y, float), np.asarray(xi, float)
    k = np.clip(np.searchsorted(x, xi, side="right"), 1, len(x) - 1)
    s = (xi - x[k - 1]) / (x[k] - x[k - 1])
    return y[k - 1] + s * (y[k] - y[k - 1])


def interp1Q(x, shift, y, xi):
    """:220-241"""
    y, xi = np.asarray(y, float), np.asarray(xi, float)
    base = ((xi - x) / shift).astype(np.int64)       # C cast: truncation toward zero
    frac = (xi - x) / shift - base
    dy = np.append(np.diff(y), 0.0)
    return y[base] + dy[base] * frac


def dc_correction(inp, f0, fs, fft_size):
    """src/world_common.cpp:61-80 (in place on a copy)"""
    out = np.array(inp, dtype=float, copy=True)
    upper_limit = 2 + int(f0 * fft_size / fs)
    low_axis = np.arange(upper_limit, dtype=float) * fs / fft_size
    rep = interp1Q(f0 - low_axis[0], -float(fs) / fft_size, inp[:upper_limit + 1], low_axis[:upper_limit - 1])
    out[:upper_limit - 1] = inp[:upper_limit - 1] + rep
    return out


def linear_smoothing(inp, width, fs, fft_size):
    """src/world_common.cpp:27-52, :82-116"""
    nc = fft_size // 2
    boundary = int(width * fft_size / fs) + 1
    ms = np.concatenate([inp[boundary:0:-1], inp[:nc], inp[nc - np.arange(0, boundary + 1)]])
    seg = np.cumsum(ms * fs / fft_size)               # sequential running sum like the reference
    fa = np.arange(nc + 1, dtype=float) / fft_size * fs - width / 2.0
    origin = -(boundary - 0.5) * fs / fft_size
    interval = float(fs) / fft_size
    low = interp1Q(origin, interval, seg, fa)
    high = interp1Q(origin, interval, seg, fa + width)
    return (high - low) / width


def nuttall_window(n):
    """src/world_common.cpp:118-126"""
    t = np.arange(n) / (n - 1.0)
    return 0.355768 - 0.487396 * np.cos(2.0 * np.pi * t) + 0.144232 * np.cos(4.0 * np.pi * t) - 0.012604 * np.cos(6.0 * np.pi * t)


def _round_arr(v):
    return np.where(v > 0, (v + 0.5).astype(np.int64), (v - 0.5).astype(np.int64))


# =============================================================================================
# CheapTrick (src/cheaptrick.cpp)
# =============================================================================================
def cheaptrick_fft_size(fs, f0_floor=71.0):
    return int(2 ** (1 + int(np.log(3.0 * fs / f0_floor + 1) / 0.69314718055994529)))  # :97-100


def cheaptrick(x, fs, tpos, f0, rng, q1=-0.15, f0_floor=71.0, fft_size=0):
    """:48-95 (serial loop order: frame by frame, window noise then per-bin noise)"""
    x = np.asarray(x, float)
    N = fft_size or cheaptrick_fft_size(fs, f0_floor)
    floor_internal = 3 * fs / (N - 3.0)                                           # :102-105
    nc = N // 2
    sp = np.empty((len(f0), nc + 1))
    k = np.arange(1, nc + 1)
    for i in range(len(f0)):
        cf0 = 500.0 if f0[i] <= floor_internal else f0[i]                         # :76
        hw = matlab_round(1.5 * fs / cf0)                                         # :141
        base = np.arange(-hw, hw + 1)
        origin = matlab_round(tpos[i] * fs + 0.001)                               # :177
        safe = np.clip(origin + base, 0, len(x) - 1)
        w = 0.5 * np.cos(np.pi * (base / 1.5 / fs) * cf0) + 0.5                   # :186-189
        w = w / np.sqrt(np.sum(w * w))
        wave = x[safe] * w + rng.take(2 * hw + 1) * 0.000000000000001             # :153
        wave = wave - w * (np.sum(wave) / np.sum(w))                              # :154-163
        buf = np.zeros(N)
        buf[:2 * hw + 1] = wave
        X = fft_r2c(buf)
        P = dc_correction(X.real ** 2 + X.imag ** 2, cf0, fs, N)                  # :198-218
        P = linear_smoothing(P, cf0 * 2.0 / 3.0, fs, N)                           # :124-125
        P = P + np.abs(rng.take(nc + 1)) * 2.2204460492503131e-16                 # :220-228
        logp = np.log(P)
        mirrored = np.concatenate([logp, logp[nc - 1:0:-1]])                      # :255-258
        C = fft_r2c(mirrored).real
        quef = k / float(fs)
        sl = np.concatenate([[1.0], np.sin(np.pi * cf0 * quef) / (np.pi * cf0 * quef)])
        cl = np.concatenate([[(1.0 - 2.0 * q1) + 2.0 * q1], (1.0 - 2.0 * q1) + 2.0 * q1 * np.cos(2.0 * np.pi * quef * cf0)])
        sp[i] = np.exp(fft_c2r((C * sl * cl / N).astype(np.complex128), N)[:nc + 1])   # :262-272
    return sp


# =============================================================================================
# D4C (src/d4c.cpp)
# =============================================================================================
def _d4c_window(x, fs, f0, pos, kind, ratio, rng):
    """:246-303; kind 1 = Hanning, 2 = Blackman"""
    hw = matlab_round(ratio * fs / f0 / 2.0)
    base = np.arange(-hw, hw + 1)
    origin = matlab_round(pos * fs + 0.001)
    safe = np.clip(origin + base, 0, len(x) - 1)
    position = (2.0 / ratio / fs) * base
    c2 = np.pi * f0
    if kind == 1:
        w = 0.5 * np.cos(c2 * position) + 0.5
    else:
        w = 0.42 + 0.5 * np.cos(c2 * position) + 0.08 * np.cos(c2 * position * 2)
    wave = x[safe] * w + rng.take(2 * hw + 1) * 0.000000000001
    return wave - w * (np.sum(wave) / np.sum(w))


def d4c(x, fs, tpos, f0, fft_size, rng, threshold=0.85):
    """:113-173 (Love Train over all frames first, then the body frames)"""
    x = np.asarray(x, float)
    L = len(f0)
    bins = fft_size // 2 + 1
    N = int(2 ** (1 + int(np.log(4.0 * fs / 47.0 + 1) / 0.69314718055994529)))     # :63-64
    N_lt = int(2 ** (1 + int(np.log(3.0 * fs / 40.0 + 1) / 0.69314718055994529)))  # :102-103
    n_ap = int(min(15000.0, fs / 2.0 - 3000.0) / 3000.0)                          # :65-67
    wl = int(3000.0 * N / fs) * 2 + 1                                             # :70
    nutt = nuttall_window(wl)
    ap = np.full((L, bins), 1.0 - 0.000000000001)                                 # :127-132
    # ---- Love Train (:181-240)
    b0, b1, b2 = (int(np.ceil(v * N_lt / fs)) for v in (100.0, 4000.0, 7900.0))
    ap0 = np.zeros(L)
    for i in range(L):
        if f0[i] == 0.0:
            continue
        cf0 = max(f0[i], 40.0)
        wave = _d4c_window(x, fs, cf0, tpos[i], 2, 3.0, rng)
        buf = np.zeros(N_lt)
        buf[:len(wave)] = wave
        X = fft_r2c(buf)
        ps = np.zeros(N_lt)
        ps[b0 + 1:N_lt // 2 + 1] = (X.real ** 2 + X.imag ** 2)[b0 + 1:]
        cum = np.cumsum(ps[:b2 + 1]) if b2 < N_lt else np.cumsum(ps)
        ap0[i] = cum[b1] / cum[min(b2, len(cum) - 1)]
    coarse_axis = np.concatenate([np.arange(n_ap + 1) * 3000.0, [fs / 2.0]])
    freq_axis = np.arange(bins, dtype=float) * fs / fft_size
    center = [int(3000.0 * (b + 1) * N / fs) for b in range(n_ap)]
    boundary = matlab_round(N * 8.0 / wl)
    nb = N // 2 + 1
    ramp = np.arange(1, N + 1, dtype=float)
    for i in range(L):
        if f0[i] == 0 or ap0[i] <= threshold:
            continue
        cf0 = max(47.0, f0[i])
        # static centroid (:339-405)
        sc = np.zeros(nb)
        for pos in (tpos[i] - 0.25 / cf0, tpos[i] + 0.25 / cf0):
            wave = _d4c_window(x, fs, cf0, pos, 2, 4.0, rng)
            wave = wave / np.sqrt(np.sum(wave * wave))
            buf = np.zeros(N)
            buf[:len(wave)] = wave
            X1 = fft_r2c(buf)
            X2 = fft_r2c(buf * ramp)
            sc = sc + (X2.real * X1.real + X1.imag * X2.imag)
        sc = dc_correction(sc, cf0, fs, N)
        # smoothed power (:411-434)
        wave = _d4c_window(x, fs, cf0, tpos[i], 1, 4.0, rng)
        buf = np.zeros(N)
        buf[:len(wave)] = wave
        X = fft_r2c(buf)
        spw = linear_smoothing(dc_correction(X.real ** 2 + X.imag ** 2, cf0, fs, N), cf0, fs, N)
        # static group delay (:440-460)
        sgd = linear_smoothing(sc / spw, cf0 / 2.0, fs, N)
        sgd = sgd - linear_smoothing(sgd, cf0, fs, N)
        # coarse aperiodicity (:466-503)
        coarse = np.empty(n_ap + 2)
        coarse[0], coarse[-1] = -60.0, -0.000000000001
        for b in range(n_ap):
            buf = np.zeros(N)
            buf[:wl] = sgd[center[b] - wl // 2:center[b] - wl // 2 + wl] * nutt
            X = fft_r2c(buf)
            cum = np.cumsum(np.sort(X.real ** 2 + X.imag ** 2))
            ca = 10 * np.log10(cum[nb - boundary - 2] / cum[nb - 1])
            coarse[b + 1] = min(0.0, ca + (cf0 - 100) / 50.0)                     # :325-327
        ap[i] = 10.0 ** (interp1(coarse_axis, coarse, freq_axis) / 20.0)          # :160-168
    return ap


# =============================================================================================
# Synthesis (src/synthesis.cpp) + MinimumPhaseAnalysis (src/world_common.cpp:192-233)
# =============================================================================================
def minimum_phase(log_spectrum_half, N):
    nc = N // 2
    full = np.concatenate([log_spectrum_half, log_spectrum_half[nc - 1:0:-1]])
    c = fft_r2c(full)
    cep = np.zeros(N, dtype=np.complex128)
    cep[0] = complex(c[0].real, -c[0].imag)
    cep[1:nc] = 2.0 * c[1:nc].real - 2.0j * c[1:nc].imag
    cep[nc] = complex(c[nc].real, -c[nc].imag)
    m = fft_c2c(cep, 1)[:nc + 1]
    return np.exp(m.real / N) * (np.cos(m.imag / N) + 1j * np.sin(m.imag / N))


def synthesis(f0, sp, ap, fs, fft_size, frame_period_ms, out_length, rng):
    """:77-177 (serial build: noise drawn pulse by pulse)"""
    f0 = np.asarray(f0, float)
    L, N, nc = len(f0), fft_size, fft_size // 2
    fp = frame_period_ms / 1000.0
    lowest_f0 = fs // fft_size + 1.0                                              # :97 (integer division)
    # ---- time base (:180-243)
    cf0 = np.where(f0 < lowest_f0, 0.0, f0)
    cvuv = np.where(cf0 == 0.0, 0.0, 1.0)
    ct = np.arange(L + 1) * fp
    cf0 = np.append(cf0, cf0[-1] * 2 - cf0[-2])
    cvuv = np.append(cvuv, cvuv[-1] * 2 - cvuv[-2])
    t = np.arange(out_length) / float(fs)
    i_f0 = interp1(ct, cf0, t)
    i_vuv = (interp1(ct, cvuv, t) > 0.5).astype(float)
    i_f0 = np.where(i_vuv == 0.0, 500.0, i_f0)
    pidx = np.empty(out_length, dtype=np.int32)
    pshift = np.empty(out_length)
    i_f0 = np.ascontiguousarray(i_f0)
    P = _clib().oracle_pulse_locations(i_f0.ctypes.data, out_length, fs, pidx.ctypes.data, pshift.ctypes.data)
    pidx, pshift = pidx[:P], pshift[:P]
    # dc remover (:290-303)
    r = 0.5 - 0.5 * np.cos(2.0 * np.pi / (1.0 + N) * (np.arange(nc) + 1.0))
    dcr = r / (np.sum(r) * 2)
    y = np.zeros(out_length)
    karr = np.arange(nc + 1)
    for p in range(P):
        noise_size = int(pidx[min(P - 1, p + 1)] - pidx[p])
        cur_time = pidx[p] / float(fs)
        vuv = i_vuv[pidx[p]]
        tf = cur_time / fp
        fl, cl = min(L - 1, int(np.floor(tf))), min(L - 1, int(np.ceil(tf)))
        w = tf - fl
        safe = lambda a: np.maximum(0.001, np.minimum(0.999999999999, a))
        if fl == cl:
            se, ar = np.abs(sp[fl]), safe(ap[fl]) ** 2
        else:
            se = (1.0 - w) * np.abs(sp[fl]) + w * np.abs(sp[cl])
            ar = ((1.0 - w) * safe(ap[fl]) + w * safe(ap[cl])) ** 2
        # periodic response (:403-474)
        if vuv <= 0.5 or ar[0] > 0.999:
            periodic = np.zeros(N)
        else:
            mp = minimum_phase(np.log(se * (1.0 - ar) + 0.000000000001) / 2.0, N)
            coef = 2.0 * np.pi * pshift[p] * fs / N
            re2 = np.cos(coef * karr)
            im2 = np.sqrt(1.0 - re2 * re2)
            spec = (mp.real * re2 - mp.imag * im2) + 1j * (mp.real * im2 + mp.imag * re2)
            wave = fft_c2r(spec, N)
            dc = np.sum(wave[:nc])
            periodic = np.concatenate([-dc * dcr, wave[:nc] - dc * dcr])          # Q9
        # aperiodic response (:479-530)
        if noise_size > 0:
            nz = rng.take(noise_size)
            buf = np.zeros(N)
            buf[:noise_size] = nz - np.sum(nz) / noise_size
            NS = fft_r2c(buf)
            mp = minimum_phase(np.log(se * ar) / 2.0 if vuv != 0.0 else np.log(se) / 2.0, N)
            wave = fft_c2r(mp * NS, N)
            aper = np.concatenate([wave[nc:], wave[:nc]])
            resp = (periodic * np.sqrt(float(noise_size)) + aper) / N
        else:
            resp = np.zeros(N)
        index = int(pidx[p]) - nc
        if index + N < 0 or index + 1 >= out_length:
            continue
        b = abs(index + 1) if index + 1 < 0 else 0
        e = out_length - index - 1 if index + N >= out_length else N
        y[index + 1 + b:index + 1 + e] += resp[b:e]
    return y


# =============================================================================================
# codec (src/codec.cpp)
# =============================================================================================
def _mel(f):
    return 1127.01048 * np.log(f / 700.0 + 1.0)


def code_aperiodicity(ap, fs, fft_size):
    n_ap = int(min(15000.0, fs / 2.0 - 3000.0) / 3000.0)
    axis = 3000.0 * (np.arange(n_ap) + 1.0)
    return np.stack([interp1Q(0, float(fs) / fft_size, 20 * np.log10(row), axis) for row in ap])   # :216-235


def decode_aperiodicity(coded, fs, fft_size):
    n_ap = coded.shape[1]
    bins = fft_size // 2 + 1
    out = np.full((len(coded), bins), 1.0 - 0.000000000001)
    axis = np.concatenate([np.arange(n_ap + 1) * 3000.0, [fs / 2.0]])
    fa = float(fs) / fft_size * np.arange(bins)
    for i, c in enumerate(coded):
        if np.sum(c) / n_ap > -0.5:                                               # CheckVUV :30-40
            continue
        out[i] = 10.0 ** (interp1(axis, np.concatenate([[-60.0], c, [-0.000000000001]]), fa) / 20.0)
    return out


def code_spectral_envelope(sp, fs, fft_size, nd):
    M = fft_size // 2
    floor_mel, ceil_mel = _mel(40.0), _mel(min(fs / 2.0, 20000.0))
    mel_axis = (ceil_mel - floor_mel) * np.arange(M) / M + floor_mel
    i = np.arange(M)
    w = 2.0 * np.cos(i * np.pi / fft_size) / np.sqrt(fft_size) + 1j * (2.0 * np.sin(i * np.pi / fft_size) / np.sqrt(fft_size))
    w[0] = complex(w[0].real / np.sqrt(2.0), w[0].imag)
    fa = np.append(_mel(np.arange(M, dtype=float) * fs / fft_size), 0.0)          # [M] is never set (Q12)
    out = np.empty((len(sp), nd))
    for f, row in enumerate(sp):
        mel = interp1(fa[:M], np.log(row)[:M], mel_axis) if np.all(mel_axis < fa[M - 1]) else None
        assert mel is not None, "mel axis reaches the unset knot (SURVEY Q12): not restated"
        buf = np.concatenate([mel[0::2], mel[M - 1 - 2 * np.arange(M // 2)]])       # :76-81
        X = fft_r2c(buf)[:nd]
        out[f] = (X.real * w[:nd].real - X.imag * w[:nd].imag) / np.sqrt(M)
    return out


def decode_spectral_envelope(coded, fs, fft_size, nd):
    M = fft_size // 2
    bins = M + 1
    floor_mel, ceil_mel = _mel(40.0), _mel(min(fs / 2.0, 20000.0))
    i = np.arange(nd)
    w = np.cos(i * np.pi / fft_size) * np.sqrt(fft_size) + 1j * (np.sin(i * np.pi / fft_size) * np.sqrt(fft_size))
    w[0] = complex(w[0].real / np.sqrt(2.0), w[0].imag)
    mel_axis = np.concatenate([[0.0], 700.0 * (np.exp(((ceil_mel - floor_mel) * np.arange(M) / M + floor_mel) / 1127.01048) - 1.0), [fs / 2.0]])
    fa = np.arange(bins, dtype=float) * fs / fft_size
    out = np.empty((len(coded), bins))
    for f, c in enumerate(coded):
        z = np.zeros(M, dtype=np.complex128)
        z[:nd] = c * w.real * np.sqrt(M) - 1j * (c * w.imag * np.sqrt(M))
        o = fft_c2c(z, 2).real
        ms = np.empty(M + 2)
        ms[1:M + 1:2] = o[:M // 2]
        ms[2:M + 2:2] = o[M - 1 - np.arange(M // 2)]
        ms[0], ms[M + 1] = ms[1], ms[M]
        out[f] = np.exp(interp1(mel_axis, ms, fa) / M)
    return out


# ---- the demo's parameter modification (test/test.cpp:201-243) -------------------------------------
def parameter_modification(f0, sp, fs, fft_size, shift=None, ratio=None):
    """F0 scaling (:205-209) and spectral stretching (:211-236): every row is log'ed, interp1'ed from the
    axis ratio * i / fft_size * fs to the axis i / fft_size * fs (linear extrapolation beyond the last knot),
    exp'ed, and for ratio < 1 the bins from int(fft_size / 2 * ratio) on repeat the bin before them."""
    f0 = np.array(f0, dtype=float, copy=True)
    sp = np.array(sp, dtype=float, copy=True)
    if shift is not None:
        f0 = f0 * shift
    if ratio is None:
        return f0, sp
    bins = fft_size // 2 + 1
    idx = np.arange(bins)
    axis1 = ratio * idx / fft_size * fs
    axis2 = idx.astype(float) / fft_size * fs
    cut = int(fft_size / 2.0 * ratio)
    for i in range(sp.shape[0]):
        row = np.exp(interp1(axis1, np.log(sp[i]), axis2))
        if ratio < 1.0:
            row[cut:] = row[cut - 1]
        sp[i] = row
    return f0, sp
