"""GPU parity tests: pose-only LM through the C ABI vs the fp64 CPU oracle.
Tolerance (BASELINE.json north_star): translation within 1e-4 relative; outlier masks identical."""
import os
import numpy as np
import pytest
import oracle
import plslam_b200 as pl
from plslam_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-4


def _close(T, To):
    t, to = T[:3, 3].astype(np.float64), To[:3, 3].astype(np.float64)
    assert np.linalg.norm(t - to) <= REL_TOL * max(np.linalg.norm(to), 1e-9), (t, to)
    assert np.abs(T[:3, :3] - To[:3, :3]).max() <= 1e-5


@pytest.mark.parametrize("seed", [3, 7, 21, 42, 100, 101])
def test_pose_optimization_matches_oracle(seed):
    p = synth.synth_pose_problem(seed)
    a = (p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"], p["line_func"], p["line_Xw"])
    n, T, po, lo, its = pl.Optimizer.PoseOptimization(*a)
    on, oT, opo, olo, oits = oracle.pose_optimization(0, *a)
    _close(T, oT)
    assert n == on and np.array_equal(po, opo) and np.array_equal(lo, olo) and its == oits
    assert np.abs(T - p["Tcw_true"]).max() < 0.02


def test_points_only_and_lines_only():
    p = synth.synth_pose_problem(9, n_points=500, n_lines=120)
    n, T, po, lo, its = pl.Optimizer.PoseOptimizationWithPoints(p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"])
    on, oT, opo, olo, oits = oracle.pose_optimization(1, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"],
                                                      np.zeros((0, 3)), np.zeros((0, 6)))
    _close(T, oT); assert n == on and np.array_equal(po, opo) and its == oits
    n, T, po, lo, its = pl.Optimizer.PoseOptimizationWithLines(p["Tcw0"], p["K"], p["line_func"], p["line_Xw"])
    on, oT, opo, olo, oits = oracle.pose_optimization(2, p["Tcw0"], p["K"], np.zeros((0, 2)), np.zeros(0), np.zeros((0, 3)),
                                                      p["line_func"], p["line_Xw"])
    _close(T, oT); assert n == on and np.array_equal(lo, olo) and its == oits


def test_golden_and_degenerate():
    g = np.load(os.path.join(G, "lm_oracle.npz"))
    for k in range(int(g["count"])):
        p = synth.synth_pose_problem(int(g["seeds"][k]))
        n, T, po, lo, its = pl.Optimizer.PoseOptimization(p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"],
                                                          p["line_func"], p["line_Xw"])
        _close(T, g["T"][k])
        assert n == g["inliers"][k] and np.array_equal(po, g["po"][k]) and np.array_equal(lo, g["lo"][k])
    p = synth.synth_pose_problem(1)
    n, T, po, lo, its = pl.Optimizer.PoseOptimization(p["Tcw0"], p["K"], p["pt_obs"][:2], p["pt_inv_sigma2"][:2], p["pt_Xw"][:2],
                                                      p["line_func"], p["line_Xw"])
    assert n == 0 and its == 0 and np.array_equal(T, p["Tcw0"])      # <3 correspondences: untouched (Optimizer.cc:846)
    a = (p["Tcw0"], p["K"], p["pt_obs"][:8], p["pt_inv_sigma2"][:8], p["pt_Xw"][:8])
    n, T, po, lo, its = pl.Optimizer.PoseOptimizationWithPoints(*a)
    on, oT, opo, olo, oits = oracle.pose_optimization(1, *a, np.zeros((0, 3)), np.zeros((0, 6)))
    _close(T, oT); assert n == on and its == oits                       # <10 edges: one round only (Optimizer.cc:961)


def test_large_outlier_fraction_and_far_start():
    p = synth.synth_pose_problem(13, outlier_frac=0.35, pert_t=0.08, pert_deg=3.0)
    a = (p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"], p["line_func"], p["line_Xw"])
    n, T, po, lo, its = pl.Optimizer.PoseOptimization(*a)
    on, oT, opo, olo, oits = oracle.pose_optimization(0, *a)
    _close(T, oT)
    assert n == on and np.array_equal(po, opo) and np.array_equal(lo, olo)
