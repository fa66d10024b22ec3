"""Synthetic sinusoid-plus-noise speech used by the parity tests and bench.py.

SURVEY.md section 8d: harmonic source f0(t) = 140 + 20 sin(2 pi 1.5 t) Hz, harmonics
k = 1..39 below 0.95 fs/2 with amplitude 1/k x (three Lorentzian formants + 0.02),
peak 0.6, additive Gaussian noise sigma = 0.003 everywhere (keeps every frame off
exact-zero spectra), 0.15 s noise-only lead-in/out with 20 ms ramps and, for inputs
longer than 2 s, a 0.2 s unvoiced gap every 2 s.  numpy only; fp64; deterministic.
"""
import numpy as np

_FORMANTS = ((700.0, 150.0), (1200.0, 200.0), (2600.0, 300.0))


def _formant_gain(freq):
    g = np.full_like(freq, 0.02)
    for fc, bw in _FORMANTS:
        g += 1.0 / (1.0 + ((freq - fc) / (0.5 * bw)) ** 2)
    return g


def _ramp_mask(n, fs, on_spans, ramp_s=0.020):
    """1.0 inside the voiced spans with raised-cosine ramps of ramp_s at both ends."""
    t = np.arange(n) / fs
    m = np.zeros(n)
    for (a, b) in on_spans:
        up = np.clip((t - a) / ramp_s, 0.0, 1.0)
        dn = np.clip((b - t) / ramp_s, 0.0, 1.0)
        m = np.maximum(m, 0.5 - 0.5 * np.cos(np.pi * np.minimum(up, dn)))
    return m


def synth_speech(fs, seconds, seed=0, noise_sigma=0.003):
    """Return `int(fs*seconds)` fp64 samples in (-1, 1)."""
    n = int(round(fs * seconds))
    t = np.arange(n) / fs
    f0 = 140.0 + 20.0 * np.sin(2.0 * np.pi * 1.5 * t)
    phase = 2.0 * np.pi * np.cumsum(f0) / fs
    x = np.zeros(n)
    for k in range(1, 40):
        fk = k * f0
        if fk.max() >= 0.95 * fs / 2.0:
            break
        x += _formant_gain(fk) / k * np.sin(k * phase)
    # voiced spans: lead-in/out 0.15 s; for long inputs a 0.2 s unvoiced gap every 2 s
    lead = 0.15
    spans = []
    if seconds > 2.0:
        a = lead
        while a < seconds - lead:
            b = min(a + 1.8, seconds - lead)
            if b - a > 0.05:
                spans.append((a, b))
            a = b + 0.2
    else:
        spans.append((lead, seconds - lead))
    x *= _ramp_mask(n, fs, spans)
    peak = np.max(np.abs(x))
    if peak > 0:
        x *= 0.6 / peak
    rng = np.random.default_rng(seed)
    x += noise_sigma * rng.standard_normal(n)
    return np.ascontiguousarray(x, dtype=np.float64)


def silence_edge(fs, seconds, seed=0):
    """Parity-only case: exact digital silence for the first and last 0.2 s."""
    x = synth_speech(fs, seconds, seed)
    k = int(0.2 * fs)
    x[:k] = 0.0
    x[-k:] = 0.0
    return x
