"""An INDEPENDENT check of the pose-only LM oracle (the g2o path cannot be pinned to the reference's compiled code: Optimizer.cc needs
g2o + Eigen): when no edge is ever classified as an outlier, the reference's procedure ends with ten plain least-squares iterations
(the Huber kernel is dropped for the last round, Optimizer.cc:789-790 / :1010-1011), so the pose it returns must be the minimiser of
   sum_i  w_i |obs_i - proj(T X_i)|^2  +  sum_j  (l_j . proj_h(T S_j))^2 + (l_j . proj_h(T E_j))^2 .
scipy.optimize.least_squares - another algorithm (trust-region reflective), another parametrisation (rotation vector + translation,
finite-difference Jacobian), none of the oracle's code - must land on the same pose from the same start."""
import numpy as np
import pytest
import oracle
from plslam_b200 import synth

scipy_opt = pytest.importorskip("scipy.optimize")
from scipy.spatial.transform import Rotation  # noqa: E402


def _residuals(x, T0, p):
    R = Rotation.from_rotvec(x[:3]).as_matrix() @ T0[:3, :3]
    t = T0[:3, 3] + x[3:]
    fx, fy, cx, cy = [float(v) for v in p["K"]]

    def proj(X):
        Xc = X @ R.T + t
        return np.stack([Xc[:, 0] / Xc[:, 2] * fx + cx, Xc[:, 1] / Xc[:, 2] * fy + cy], 1)
    r = [(np.sqrt(p["pt_inv_sigma2"].astype(np.float64))[:, None] * (p["pt_obs"].astype(np.float64) - proj(p["pt_Xw"].astype(np.float64)))).ravel()]
    if len(p["line_func"]):
        l = p["line_func"]
        for e in (p["line_Xw"][:, :3], p["line_Xw"][:, 3:]):
            uv = proj(e)
            r.append(l[:, 0] * uv[:, 0] + l[:, 1] * uv[:, 1] + l[:, 2])
    return np.concatenate(r)


@pytest.mark.parametrize("seed,mode", [(3, 0), (4, 0), (5, 1), (6, 2), (7, 0)])
def test_final_pose_is_the_least_squares_minimiser(seed, mode):
    p = synth.synth_pose_problem(seed, outlier_frac=0.0, noise_px=0.4)
    n, T, po, lo, its = oracle.pose_optimization(mode, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"], p["line_func"], p["line_Xw"])
    q = dict(p)
    if mode == 1:
        q["line_func"] = p["line_func"][:0]; q["line_Xw"] = p["line_Xw"][:0]        # PoseOptimizationWithPoints
    if mode == 2:
        q["pt_obs"] = p["pt_obs"][:0]; q["pt_inv_sigma2"] = p["pt_inv_sigma2"][:0]; q["pt_Xw"] = p["pt_Xw"][:0]   # ...WithLines
    if po.any() or lo.any():
        pytest.skip("an edge was classified as an outlier: the last round then runs on a subset the oracle does not report")
    T0 = p["Tcw0"].astype(np.float64)
    sol = scipy_opt.least_squares(_residuals, np.zeros(6), args=(T0, q), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    Rs = Rotation.from_rotvec(sol.x[:3]).as_matrix() @ T0[:3, :3]; ts = T0[:3, 3] + sol.x[3:]
    To = np.asarray(T, np.float64).reshape(4, 4)
    assert np.abs(Rs - To[:3, :3]).max() < 5e-6            # the oracle returns fp32 poses
    assert np.linalg.norm(ts - To[:3, 3]) < 2e-5 * max(1.0, np.linalg.norm(ts))
    # and the two costs agree to a relative 1e-7
    xo = np.concatenate([Rotation.from_matrix(To[:3, :3] @ T0[:3, :3].T).as_rotvec(), To[:3, 3] - T0[:3, 3]])
    co, cs = (_residuals(xo, T0, q) ** 2).sum(), (sol.fun ** 2).sum()
    assert abs(co - cs) <= 1e-6 * cs
