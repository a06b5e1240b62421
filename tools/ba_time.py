"""Time of one local-BA window (20 free + 40 fixed keyframes, 3000 points, 400 lines): wall time of the C-ABI call and the kernel's
own time (CUPTI through torch.profiler).  python tools/ba_time.py [kitti]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import plslam_b200 as pl
from plslam_b200 import synth

kw = {}
if len(sys.argv) > 1 and sys.argv[1] == "kitti":
    kw = dict(K=(718.856, 718.856, 607.1928, 185.2157), w=1241, h=376)
p = synth.synth_ba_problem(4, n_free=20, n_fixed=40, n_pt=3000, n_ln=400, **kw)
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); g = pl.LocalBundleAdjustmentWithLine(p); torch.cuda.synchronize()
    print("local BA wall ms", round(1000 * (time.perf_counter() - t), 1), "its", g["its"])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pl.LocalBundleAdjustmentWithLine(p); torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0 or "emcpy" in e.key:
        print(f"{e.key[:60]:60s} n={e.count} device_ms={e.device_time_total / 1000:.2f} cpu_ms={e.cpu_time_total / 1000:.2f}")
