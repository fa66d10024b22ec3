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
