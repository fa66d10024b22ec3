"""GPU parity of the batch mode (BASELINE configs[2]: independent 22.05 kHz / 5 s utterances):
utterances run concurrently on several streams, each one must equal its own reference process."""
import numpy as np
import pytest
import torch

from oracle import refbin

pytestmark = pytest.mark.gpu


def test_batch_of_utterances_matches_reference(wb, signals):
    # nine utterances on three pipelines: each pipeline sees a first run (ordinary launches), the run that captures
    # its CUDA graph and a replay; one utterance of another length in between (ragged batch: its own buffers, a
    # new capture afterwards)
    fs, seconds, n_utt = 22050, 5.0, 9
    xs = [signals.synth_speech(fs, 3.0 if i == 4 else seconds, seed=1000 + i) for i in range(n_utt)]   # SURVEY 8d seeds
    bp = wb.BatchPipeline(fs, n_streams=3, harvest_option=wb.HarvestOption(f0_floor=40.0, frame_period=5.0),
                          cheaptrick_option=wb.CheapTrickOption(f0_floor=71.0), d4c_option=wb.D4COption(threshold=0.85))
    outs = bp.run([torch.from_numpy(x).cuda() for x in xs])
    torch.cuda.synchronize()
    for i, (x, o) in enumerate(zip(xs, outs)):
        ref, _ = refbin.run_reference(x, fs, stages="hcds")
        f0, sp, ap, y = (o[k].cpu().numpy() for k in ("f0", "sp", "ap", "y"))
        assert np.array_equal(f0 > 0, ref["f0"] > 0), "utterance %d voicing" % i
        v = ref["f0"] > 0
        assert np.max(np.abs(f0[v] - ref["f0"][v]) / ref["f0"][v]) < 1e-4
        assert np.max(np.abs(sp - ref["sp"]) / ref["sp"]) < 1e-4
        assert np.max(np.abs(ap - ref["ap"]) / ref["ap"]) < 1e-4
        assert np.max(np.abs(y - ref["y"])) / np.abs(ref["y"]).max() < 1e-4
