"""Generates the committed golden vectors by running the REFERENCE ITSELF (oracle/_ref/refrun, the
reference's own sources compiled by oracle/Makefile) in this container.  Run from the repo root:

    python tests/golden/make_golden.py

The reference ships no fixtures, so these are the pinned known answers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import worldb200  # noqa: E402,F401
from worldb200 import signals  # noqa: E402
from oracle import refbin  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # BASELINE configs[0]: 1 s mono 16 kHz synthetic vowel through the whole chain + codec round trip
    fs = 16000
    x = signals.synth_speech(fs, 1.0, seed=0)
    ref, _ = refbin.run_reference(x, fs, stages="hcdsk", codec_nd=60)
    np.savez_compressed(os.path.join(HERE, "cfg1_16k_1s.npz"), x=x, fs=fs, tpos=ref["tpos"], f0=ref["f0"], sp=ref["sp"],
                        ap=ref["ap"], y=ref["y"], csp=ref["csp"], cap=ref["cap"], dsp=ref["dsp"], dap=ref["dap"],
                        fft_size=ref["fft_size"])
    # a short 48 kHz case (FFT 2048 / D4C 4096): every 8th frame of sp/ap to keep the file small
    fs = 48000
    x = signals.synth_speech(fs, 1.0, seed=11)
    ref, _ = refbin.run_reference(x, fs, stages="hcds")
    np.savez_compressed(os.path.join(HERE, "cfg2_48k_1s.npz"), x=x, fs=fs, tpos=ref["tpos"], f0=ref["f0"],
                        sp_every8=ref["sp"][::8], ap_every8=ref["ap"][::8], y=ref["y"], fft_size=ref["fft_size"])
    # the demo's ParameterModification (test/test.cpp:201-243) through oracle/_ref/refmod on 12 frames of the
    # 16 kHz case: F0 scaling alone, stretching up, stretching down
    g = np.load(os.path.join(HERE, "cfg1_16k_1s.npz"))
    sel = slice(60, 72)
    mod = {"f0_in": g["f0"][sel], "sp_in": g["sp"][sel], "fs": 16000, "fft_size": int(g["fft_size"])}
    for tag, (shift, ratio) in {"a": (1.5, None), "b": (0.8, 1.3), "c": (1.25, 0.7)}.items():
        f0m, spm = refbin.run_modification(mod["f0_in"], mod["sp_in"], 16000, int(g["fft_size"]), shift, ratio)
        mod["f0_" + tag], mod["sp_" + tag] = f0m, spm
        mod["args_" + tag] = np.array([shift, np.nan if ratio is None else ratio])
    np.savez_compressed(os.path.join(HERE, "mod_16k.npz"), **mod)
    # first values of the randn() stream (a tiny C program would do the same: the generator is public
    # in world_matlabfunctions.cpp:243-264); obtained here from Synthesis' noise is not possible, so
    # the stream is pinned through the waveform parity instead.
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
