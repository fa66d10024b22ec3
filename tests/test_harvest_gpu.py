"""GPU parity: Harvest through the C-ABI vs the reference, stage by stage and end to end."""
import numpy as np
import pytest

from oracle import refbin, refdump

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: f0 values within 1e-4 relative; voicing decisions exact


def _cmp(name, got, ref, tol=1e-8):
    assert got.shape == ref.shape, name
    nz_g, nz_r = got != 0, ref != 0
    n_mis = int(np.sum(nz_g != nz_r))
    both = nz_g & nz_r
    err = float(np.max(np.abs(got[both] - ref[both]) / np.abs(ref[both]))) if both.any() else 0.0
    print("%-8s nonzero ref %d, zero/nonzero mismatches %d, max rel err %.3e" % (name, nz_r.sum(), n_mis, err))
    assert n_mis == 0, name
    assert err < tol, name


@pytest.mark.parametrize("fs,seconds", [(16000, 1.0), (48000, 2.0)])
def test_harvest_stages(wb, signals, fs, seconds):
    x = signals.synth_speech(fs, seconds, seed=4)
    ref = refdump.harvest_intermediates(x, fs, f0_floor=40.0)
    hv = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, frame_period=1.0))
    tpos, f0 = hv.compute(x)
    Lb, nch, mc = ref["Lb"], ref["nch"], ref["max_candidates"]
    assert len(f0) == Lb
    y = hv.debug_read("hv_y", (ref["y_length"],))
    assert np.max(np.abs(y - ref["y"])) / np.max(np.abs(ref["y"])) < 1e-12
    _cmp("raw", hv.debug_read("hv_raw", (nch, Lb)), ref["raw"])
    nc = hv.debug_read("hv_nc", (4,), dtype=np.int32)
    assert nc[0] == ref["nc"]
    cand_score = hv.debug_read("hv_candA", (2, Lb, mc))   # refined candidates | scores
    _cmp("cand1", cand_score[0], ref["cand1"])
    _cmp("score1", cand_score[1], ref["score1"], tol=1e-6)
    _cmp("cand2", hv.debug_read("hv_candB", (Lb, mc)), ref["cand2"])
    cont = hv.debug_read("tl_contours", (5, Lb))
    for k, name in enumerate(["base", "step1", "step2", "step3", "step4"]):
        _cmp(name, cont[k], ref[name])
    _cmp("f0", f0, ref["f0"])
    assert np.array_equal(tpos, np.arange(Lb) * 1 / 1000.0)


@pytest.mark.parametrize("fs,seconds,seed", [(16000, 1.0, 0), (22050, 2.5, 5), (48000, 3.0, 6)])
def test_harvest_matches_reference(wb, signals, fs, seconds, seed):
    x = signals.synth_speech(fs, seconds, seed=seed)
    ref, _ = refbin.run_reference(x, fs, stages="h")
    hv = wb.Harvest(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0))
    tpos, f0 = hv.compute(x)
    assert np.array_equal(tpos, ref["tpos"])
    # voiced/unvoiced decisions bit-exact
    assert np.array_equal(f0 > 0, ref["f0"] > 0)
    v = ref["f0"] > 0
    assert v.sum() > 20
    err = float(np.max(np.abs(f0[v] - ref["f0"][v]) / ref["f0"][v]))
    print("harvest fs=%d voiced %d/%d max rel err %.3e" % (fs, v.sum(), len(v), err))
    assert err < RTOL


def test_harvest_is_a_function_of_its_input_alone(wb, signals):
    """Grow-only workspaces keep stale data beyond the current sizes, and the refinement hands out work through
    atomics: the contour must not depend on either (same segment on a fresh object, twice, after a longer and after
    a shorter input)."""
    import torch
    fs = 16000
    a = torch.from_numpy(signals.synth_speech(fs, 2.0, seed=30)).cuda()
    b = torch.from_numpy(signals.synth_speech(fs, 3.5, seed=31)).cuda()
    c = torch.from_numpy(signals.synth_speech(fs, 0.7, seed=32)).cuda()
    opt = wb.HarvestOption(f0_floor=40.0, frame_period=5.0)

    def run(h, x):
        n = h.getSamples(fs, x.numel())
        t = torch.empty(n, dtype=torch.float64, device="cuda")
        f = torch.empty(n, dtype=torch.float64, device="cuda")
        wb._check(wb.lib().wb_harvest_compute_dev(h._h, x.data_ptr(), x.numel(), t.data_ptr(), f.data_ptr(), None), "harvest")
        wb.device_synchronize()
        return f.cpu().numpy()

    h1 = wb.Harvest(fs, opt)
    f1, f1b = run(h1, a), run(h1, a)
    h2 = wb.Harvest(fs, opt)
    run(h2, b)
    f2 = run(h2, a)
    h3 = wb.Harvest(fs, opt)
    run(h3, c)
    f3 = run(h3, a)
    assert (f1 > 0).sum() > 50
    assert np.array_equal(f1, f1b) and np.array_equal(f1, f2) and np.array_equal(f1, f3)
