import time, numpy as np, sys
sys.path.insert(0, ".")
import plslam_b200 as pl
from plslam_b200 import synth
p = synth.synth_ba_problem(11, n_free=20, n_fixed=40, n_pt=3000, n_ln=400)
for i in range(4):
    t=time.perf_counter(); g = pl.LocalBundleAdjustmentWithLine(p); print("local BA ms", round(1000*(time.perf_counter()-t),1), g["its"])
