"""CPU tests of the local-BA oracle: no expected values exist in the reference (parity unpinned); known-answer =
recovery of the ground-truth window from noise-free observations, plus literal-quirk and stop-flag checks."""
import numpy as np
import oracle
from plslam_b200 import synth


def _pose_err(T, Tt):
    return np.abs(T.reshape(-1, 4, 4)[:, :3, 3] - Tt[:, :3, 3]).max()


def test_recovers_ground_truth_without_noise():
    p = synth.synth_ba_problem(4, n_free=8, n_fixed=10, n_pt=600, n_ln=80, noise_px=0.0, outlier_frac=0.0)
    o = oracle.local_ba(p)
    assert _pose_err(p["kf_Tcw"], p["kf_Tcw_true"]) > 5e-3
    assert _pose_err(o["kf_Tcw"], p["kf_Tcw_true"]) < 1e-4
    assert np.abs(o["pt_Xw"] - p["pt_Xw_true"]).mean() < 5e-3
    assert not o["pe_erase"].any() and not o["le_erase"].any()
    fixed = p["kf_fixed"].astype(bool)
    assert np.array_equal(o["kf_Tcw"][fixed], p["kf_Tcw"][fixed])       # fixed keyframes untouched
    # end points may slide along their 3-D line (point-to-line residual): check the distance to the true line
    A, B = p["ln_Xw_true"][:, :3], p["ln_Xw_true"][:, 3:]
    d = (B - A) / np.linalg.norm(B - A, axis=1)[:, None]
    v = o["ln_Xw"][:, :3] - A
    assert np.linalg.norm(v - (v * d).sum(1)[:, None] * d, axis=1).mean() < 0.02


def test_outliers_are_gated_and_quirks_kept():
    p = synth.synth_ba_problem(6, n_free=6, n_fixed=8, n_pt=400, n_ln=60, outlier_frac=0.08)
    o = oracle.local_ba(p)
    assert 5 <= o["its"] <= 15 and o["pe_erase"].sum() > 10
    assert np.array_equal(o["le_erase_kf"], p["le_kf"][np.arange(len(p["le_kf"])) // 2])   # vpLineEdgeKF double push
    pts_only = dict(p); pts_only.update(le_kf=p["le_kf"][:0], le_ln=p["le_ln"][:0], le_func=p["le_func"][:0])
    o2 = oracle.local_ba(pts_only)                                                          # Optimizer::LocalBundleAdjustment
    assert _pose_err(o2["kf_Tcw"], p["kf_Tcw_true"]) < 0.05


def test_stop_flag():
    p = synth.synth_ba_problem(6, n_free=6, n_fixed=8, n_pt=400, n_ln=60)
    o = oracle.local_ba(p, stop_flag=np.array([1], np.int32))
    assert o["its"] == 0 and np.array_equal(o["kf_Tcw"], p["kf_Tcw"])        # returns before optimising (Optimizer.cc:1951)
