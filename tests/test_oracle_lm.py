"""CPU tests of the pose-only LM oracle.  The reference stores no expected values for this path (parity
unpinned); the known-answer checks are ground-truth pose recovery on the testOpt.cpp recipe
(Examples/TestDebug/testOpt.cpp:31-48,401-407) and on the TUM-shaped synthetic problems, plus golden vectors
frozen from the oracle itself (tests/golden/lm_oracle.npz, tools/gen_golden_lm.py)."""
import os
import numpy as np
import oracle
from plslam_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def _testopt_problem(noise=1.0, seed=0):
    """Geometry of Examples/TestDebug/testOpt.cpp: fixed 3-D points/segments, K, ground-truth pose."""
    rng = np.random.default_rng(seed)
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0.5, 0.5, 1], [2, 1, 0.5], [-1, 0.5, 0.2], [0.3, -1, 0.8],
                    [1.5, 1.5, 1.2]], np.float64) * 2
    segs = np.array([[0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0], [1, 1, 0, 1, 0, 0], [1, 1, 0, 0, 1, 0], [0, 0, 0, .5, .5, 1],
                     [1, 1, 0, .5, .5, 1]], np.float64) * 2
    K = np.array([535.4, 539.2, 320.1, 247.6], np.float32)
    R = synth._rot(0.05, -0.1, 0.3); t = np.array([1.0, -2.0, 10.0])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t

    def proj(X):
        Xc = X @ R.T + t
        return np.stack([Xc[:, 0] / Xc[:, 2] * K[0] + K[2], Xc[:, 1] / Xc[:, 2] * K[1] + K[3]], 1)
    obs = proj(pts) + rng.uniform(-noise, noise, (len(pts), 2))
    a = proj(segs[:, :3]) + rng.uniform(-noise, noise, (len(segs), 2)); b = proj(segs[:, 3:]) + rng.uniform(-noise, noise, (len(segs), 2))
    l = np.cross(np.c_[a, np.ones(len(a))], np.c_[b, np.ones(len(b))]); l /= np.hypot(l[:, 0], l[:, 1])[:, None]
    T0 = T.copy(); T0[:3, 3] += [0.5, 0.5, -0.4]
    return dict(T=T, T0=T0.astype(np.float32), K=K, obs=obs.astype(np.float32), w=np.ones(len(pts), np.float32),
                X=pts.astype(np.float32), lf=l, lX=segs)


def test_testopt_recipe_recovers_ground_truth():
    p = _testopt_problem(noise=0.0)
    n, T, po, lo, its = oracle.pose_optimization(0, p["T0"], p["K"], p["obs"], p["w"], p["X"], p["lf"], p["lX"])
    assert n == 9 and not po.any() and not lo.any()
    assert np.abs(T - p["T"]).max() < 2e-4
    p = _testopt_problem(noise=1.0)
    n, T, po, lo, its = oracle.pose_optimization(0, p["T0"], p["K"], p["obs"], p["w"], p["X"], p["lf"], p["lX"])
    assert np.abs(T[:3, 3] - p["T"][:3, 3]).max() < 0.15      # +-1 px noise at 10 m depth


def test_tum_shaped_problems():
    for seed in range(5):
        p = synth.synth_pose_problem(seed)
        for mode in (0, 1, 2):
            n, T, po, lo, its = oracle.pose_optimization(mode, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"],
                                                         p["pt_Xw"], p["line_func"], p["line_Xw"])
            assert np.abs(T - p["Tcw_true"]).max() < (0.02 if mode != 2 else 0.05), (seed, mode)
            assert 4 <= its <= 40
            if mode != 2:
                assert n == (~po).sum() and (po & p["pt_is_outlier"]).sum() >= 0.9 * p["pt_is_outlier"].sum()


def test_degenerate_counts():
    p = synth.synth_pose_problem(1)
    # fewer than 3 point correspondences: returns 0, pose untouched (Optimizer.cc:846-847)
    n, T, po, lo, its = oracle.pose_optimization(0, p["Tcw0"], p["K"], p["pt_obs"][:2], p["pt_inv_sigma2"][:2],
                                                 p["pt_Xw"][:2], p["line_func"], p["line_Xw"])
    assert n == 0 and its == 0 and np.array_equal(T, p["Tcw0"])
    # fewer than 10 edges in total: only the first of the 4 rounds runs (Optimizer.cc:961)
    n, T, po, lo, its = oracle.pose_optimization(1, p["Tcw0"], p["K"], p["pt_obs"][:8], p["pt_inv_sigma2"][:8],
                                                 p["pt_Xw"][:8], p["line_func"], p["line_Xw"])
    assert its <= 10


def test_golden():
    g = np.load(os.path.join(G, "lm_oracle.npz"))
    for k in range(int(g["count"])):
        p = synth.synth_pose_problem(int(g["seeds"][k]))
        n, T, po, lo, its = oracle.pose_optimization(0, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"],
                                                     p["line_func"], p["line_Xw"])
        assert n == g["inliers"][k] and its == g["its"][k]
        assert np.allclose(T, g["T"][k], rtol=0, atol=1e-6)
        assert np.array_equal(po, g["po"][k]) and np.array_equal(lo, g["lo"][k])
