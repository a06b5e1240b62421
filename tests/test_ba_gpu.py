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
