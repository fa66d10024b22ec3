import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import worldb200 as wb
from worldb200 import signals
fs=48000
x=signals.synth_speech(fs,10.0,seed=0)
d_x=torch.from_numpy(x).cuda()
pl=wb.Pipeline(fs, wb.HarvestOption(f0_floor=40.0, frame_period=5.0), wb.CheapTrickOption(f0_floor=71.0))
for _ in range(3):
    pl.run_dev(d_x.data_ptr(), len(x))
wb.device_synchronize()
c=pl.debug_read("tl_clocks",(16,),dtype=np.int64)
d=np.diff(c[:9])
print("phase cycles:", d.tolist(), "total", c[8]-c[0])
names=["fixStep1","fixStep2(+bound)","bound+meta+sections","extend","extendSub+sort","merge","fixStep4(+bound)","smooth setup"]
for n,v in zip(names,d): print("%-18s %8d cyc  %6.1f us" % (n, v, v/1.9e3))
