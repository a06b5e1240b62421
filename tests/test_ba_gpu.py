"""GPU parity: local BA through the C ABI vs the fp64 CPU oracle (1e-4 relative on translations and structure,
identical erase masks)."""
import numpy as np
import pytest
import oracle
import plslam_b200 as pl
from plslam_b200 import synth

pytestmark = pytest.mark.gpu


def _check(p, stop=None):
    g = pl.LocalBundleAdjustmentWithLine(p, stop)
    o = oracle.local_ba(p)
    t, to = g["kf_Tcw"].reshape(-1, 4, 4)[:, :3, 3].astype(np.float64), o["kf_Tcw"].reshape(-1, 4, 4)[:, :3, 3].astype(np.float64)
    assert np.linalg.norm(t - to, axis=1).max() <= 1e-4 * np.linalg.norm(to, axis=1).max()
    assert np.abs(g["kf_Tcw"] - o["kf_Tcw"]).max() < 1e-4
    assert np.abs(g["pt_Xw"] - o["pt_Xw"]).max() <= 1e-4 * np.abs(o["pt_Xw"]).max()
    if len(o["ln_Xw"]):
        assert np.abs(g["ln_Xw"] - o["ln_Xw"]).max() <= 1e-3 * np.abs(o["ln_Xw"]).max()   # end points slide along the line: weakly determined
    assert g["its"] == o["its"]
    assert np.array_equal(g["pe_erase"], o["pe_erase"]) and np.array_equal(g["le_erase"], o["le_erase"])
    assert np.array_equal(g["le_erase_kf"], o["le_erase_kf"])
    return g, o


@pytest.mark.parametrize("seed,nf,nx,npt,nln", [(4, 8, 10, 600, 80), (6, 6, 8, 400, 60), (9, 12, 20, 1500, 200)])
def test_local_ba_matches_oracle(seed, nf, nx, npt, nln):
    p = synth.synth_ba_problem(seed, n_free=nf, n_fixed=nx, n_pt=npt, n_ln=nln)
    g, o = _check(p)
    assert g["pe_erase"].sum() > 0


def test_points_only_noise_free_and_fixed():
    p = synth.synth_ba_problem(4, n_free=8, n_fixed=10, n_pt=600, n_ln=80, noise_px=0.0, outlier_frac=0.0)
    g, o = _check(p)
    assert np.abs(g["kf_Tcw"].reshape(-1, 4, 4)[:, :3, 3] - p["kf_Tcw_true"][:, :3, 3]).max() < 1e-4
    fixed = p["kf_fixed"].astype(bool)
    assert np.array_equal(g["kf_Tcw"][fixed], p["kf_Tcw"][fixed])
    pts_only = dict(p); pts_only.update(le_kf=p["le_kf"][:0], le_ln=p["le_ln"][:0], le_func=p["le_func"][:0])
    _check(pts_only)


def test_full_size_window():
    """SURVEY.md §8d config 4 sizes: 20 free + 40 fixed keyframes, 3000 points, 400 lines."""
    p = synth.synth_ba_problem(11, n_free=20, n_fixed=40, n_pt=3000, n_ln=400)
    _check(p)


# ---------------------------------------------------------------------------------------------- global BA (Optimizer.cc:275-638)
def _check_global(p, its, robust):
    g = pl.GlobalBundleAdjustemnt(p, its, robust)
    o = oracle.global_ba(p, its, robust)
    assert g["its"] == o["its"]
    t, to = g["kf_Tcw"].reshape(-1, 4, 4)[:, :3, 3].astype(np.float64), o["kf_Tcw"].reshape(-1, 4, 4)[:, :3, 3].astype(np.float64)
    assert np.linalg.norm(t - to, axis=1).max() <= 1e-4 * np.linalg.norm(to, axis=1).max()
    assert np.abs(g["kf_Tcw"] - o["kf_Tcw"]).max() < 1e-4
    assert np.abs(g["pt_Xw"] - o["pt_Xw"]).max() <= 1e-4 * np.abs(o["pt_Xw"]).max()
    if len(o["ln_Xw"]):
        assert np.abs(g["ln_Xw"] - o["ln_Xw"]).max() <= 1e-3 * np.abs(o["ln_Xw"]).max()
    return g, o


@pytest.mark.parametrize("seed,nkf,npt,nln,robust", [(5, 8, 400, 60, True), (7, 14, 900, 120, False), (3, 40, 2500, 300, True)])
def test_global_ba_matches_oracle(seed, nkf, npt, nln, robust):
    p = synth.synth_ba_problem(seed, n_free=nkf, n_fixed=0, n_pt=npt, n_ln=nln)
    g, o = _check_global(p, 6, robust)
    assert g["its"] >= 2 and g["solve_ms"] > 0


def test_global_ba_points_only_single_iteration_and_unobserved_point():
    p = synth.synth_ba_problem(9, n_free=10, n_fixed=0, n_pt=500, n_ln=0)
    p["pt_Xw"] = np.concatenate([p["pt_Xw"], np.array([[9.0, 9.0, 9.0]], np.float32)])
    g, o = _check_global(p, 1, True)
    assert g["its"] == 1 and np.array_equal(g["pt_Xw"][-1], p["pt_Xw"][-1])
    s = pl.GlobalBundleAdjustemnt(p, 5, True, stop_flag=np.array([1], np.int32))
    assert s["its"] == 0 and np.abs(s["kf_Tcw"] - p["kf_Tcw"]).max() < 1e-6


def test_global_ba_is_bit_reproducible():
    """No atomics anywhere: the reduced system is summed per block in landmark order, the edge sums on a fixed grid."""
    p = synth.synth_ba_problem(12, n_free=30, n_fixed=0, n_pt=2000, n_ln=250)
    a = pl.GlobalBundleAdjustemnt(p, 5, True); b = pl.GlobalBundleAdjustemnt(p, 5, True)
    assert a["its"] == b["its"]
    for k in ("kf_Tcw", "pt_Xw", "ln_Xw"):
        assert np.array_equal(a[k], b[k])


def test_global_ba_whole_map_size():
    """300 keyframes, 20000 points, 2500 lines (a TUM-sequence map): the reduced system is 1794 x 1794 (57 Cholesky tile
    columns); noise-free observations, so the plain reprojection chi2 has to fall to (numerically) nothing from the perturbed start."""
    from test_oracle_ba import reprojection_chi2
    p = synth.synth_ba_problem(21, n_free=300, n_fixed=0, n_pt=20000, n_ln=2500, noise_px=0.0, outlier_frac=0.0)
    c0 = reprojection_chi2(p, p["kf_Tcw"], p["pt_Xw"], p["ln_Xw"])
    g = pl.GlobalBundleAdjustemnt(p, 10, False)
    c1 = reprojection_chi2(p, g["kf_Tcw"], g["pt_Xw"], g["ln_Xw"])
    assert c1 < 1e-3 * c0 and g["its"] >= 3
    print(f"global BA 300 KF / 20000 pts / 2500 lines: {g['its']} iterations, {g['solve_ms']:.1f} ms on the device")
