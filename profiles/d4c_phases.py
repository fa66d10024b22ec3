"""Phase cost of the D4C body kernel by elimination (WB_D4C_SKIP leaves phases out; results are then wrong)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch, worldb200 as wb
    from worldb200 import signals
    fs = 48000
    x = signals.synth_speech(fs, 10.0, seed=0)
    wb._check(wb.lib().wb_init(0), "wb_init")
    d_x = torch.from_numpy(x).cuda()
    pl = wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0), wb.D4COption(threshold=0.85))
    for _ in range(3):
        pl.run_dev(d_x.data_ptr(), len(x))
    wb.device_synchronize()
    wb.profile_reset(); wb.profile(True)
    for _ in range(5):
        pl.run_dev(d_x.data_ptr(), len(x))
    r = wb.profile_results()
    print("skip=%s d4c_body %.4f ms" % (os.environ.get("WB_D4C_SKIP", "0"), r["d4c_body_kernel"][0] / 5))
else:
    for skip in (0, 1, 2, 4, 8, 15):
        env = dict(os.environ, WB_D4C_SKIP=str(skip))
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
