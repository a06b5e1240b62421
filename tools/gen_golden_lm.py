"""Freeze oracle outputs of the pose-only LM on seeded synthetic problems -> tests/golden/lm_oracle.npz."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200  # noqa
from plslam_b200 import synth
import oracle
seeds = [3, 7, 21, 42]
T, po, lo, inl, its = [], [], [], [], []
for s in seeds:
    p = synth.synth_pose_problem(s)
    n, t, a, b, i = oracle.pose_optimization(0, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"], p["line_func"], p["line_Xw"])
    T.append(t); po.append(a); lo.append(b); inl.append(n); its.append(i)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lm_oracle.npz")
np.savez_compressed(out, count=len(seeds), seeds=np.array(seeds), T=np.array(T), po=np.array(po), lo=np.array(lo), inliers=np.array(inl), its=np.array(its))
print(inl, its)
