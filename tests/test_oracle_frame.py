"""CPU tests of the Frame-glue oracle (oracle/oracle_frame.cpp) against cv2 golden vectors
(tests/golden/frame_cv2.npz, tools/gen_golden_frame.py): initUndistortRectifyMap + remap (reference src/Frame.cc:220-222),
undistortPoints (Frame.cc:915-945, :947-985) and the fp32 gemm order behind every Rcw*P+tcw on the path."""
import os
import numpy as np
import pytest
import oracle
from plslam_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "frame_cv2.npz"))
CAMS = {"tum1": (synth.TUM1_K, synth.TUM1_DIST, 640, 480, 1), "euroc": (synth.EUROC_K, synth.EUROC_DIST, 752, 480, 5)}


@pytest.mark.parametrize("cam", ["tum1", "euroc"])
def test_undistort_map_and_remap_match_cv2(cam):
    K, D, w, h, seed = CAMS[cam]
    mx, my = oracle.undistort_map(K, D, w, h)
    assert np.array_equal(mx[::7, ::7], G[f"{cam}_mx_s"]) and np.array_equal(my[::7, ::7], G[f"{cam}_my_s"])
    img = synth.synth_frame(w, h, seed)
    assert np.array_equal(oracle.remap_linear(img, mx, my), G[f"{cam}_und"])
    assert np.array_equal(oracle.undistort_remap(img, K, D), G[f"{cam}_und"])


@pytest.mark.parametrize("cam", ["tum1", "euroc"])
def test_undistort_points_match_cv2(cam):
    K, D, w, h, _ = CAMS[cam]
    pts, want = G[f"{cam}_pts"], G[f"{cam}_upts"]
    kps = np.zeros(len(pts), oracle.KP_DTYPE); kps["x"], kps["y"] = pts[:, 0], pts[:, 1]; kps["octave"] = 3; kps["angle"] = 17.5
    out = oracle.undistort_keypoints(kps, K, D)
    assert np.array_equal(np.stack([out["x"], out["y"]], 1), want)
    assert np.array_equal(out["octave"], kps["octave"]) and np.array_equal(out["angle"], kps["angle"])
    # ComputeImageBounds uses the last four golden points (the image corners)
    c = want[-4:]
    b = oracle.image_bounds(K, D, w, h)
    assert np.array_equal(b, np.array([min(c[0, 0], c[2, 0]), min(c[0, 1], c[1, 1]), max(c[1, 0], c[3, 0]), max(c[2, 1], c[3, 1])], np.float32))


def test_no_distortion_is_identity():
    kps = np.zeros(5, oracle.KP_DTYPE); kps["x"] = np.arange(5) * 10.5; kps["y"] = 3
    assert np.array_equal(oracle.undistort_keypoints(kps, synth.TUM1_K, (0, 0, 0, 0, 0)), kps)
    assert np.array_equal(oracle.image_bounds(synth.TUM1_K, (0, 0, 0, 0, 0), 640, 480), np.array([0, 0, 640, 480], np.float32))


def test_gemm_accumulation_order_matches_cv2():
    """cv::gemm on 3x3 * 3x1 + 3x1 (CV_32F) == ((a0*b0 + a1*b1) + a2*b2) + c in fp32: the restatement used by
    isInFrustum / SearchByProjection; pinned through isInFrustum's own outputs (u, v of the projected point)."""
    A, x, c, want = G["gemm_A"], G["gemm_x"], G["gemm_c"], G["gemm_out"]
    f = np.float32
    got = np.empty_like(want)
    for i in range(3):
        got[:, i, 0] = f(f(f(f(A[:, i, 0] * x[:, 0, 0]) + f(A[:, i, 1] * x[:, 1, 0])) + f(A[:, i, 2] * x[:, 2, 0])) + c[:, i, 0])
    assert np.array_equal(got, want)
    # the oracle's gemm through isInFrustum: K = (1,1,0,0) and bounds wide open => proj = Pc.xy / Pc.z
    for i in range(0, 400):
        T = np.eye(4, dtype=np.float32); T[:3, :3] = A[i]; T[:3, 3] = c[i, :, 0]
        Pc = want[i, :, 0]
        iv, proj, _, _ = oracle.is_in_frustum_points(T, np.zeros(3), (1, 1, 0, 0), (-1e30, -1e30, 1e30, 1e30), np.log(1.2), 8, -2.0,
                                                     x[i].reshape(1, 3), np.array([[0, 0, 1]], f), [0.0], [1e30])
        if Pc[2] < 0:
            assert iv[0] == 0
        else:
            invz = f(1.0) / Pc[2]
            assert iv[0] == 1 and proj[0, 0] == f(f(Pc[0]) * invz) and proj[0, 1] == f(f(Pc[1]) * invz)


def test_is_in_frustum_points_semantics():
    v = synth.synth_map_view(7, 4000)
    b = oracle.image_bounds(synth.TUM1_K, synth.TUM1_DIST, 640, 480)
    iv, proj, lvl, vc = oracle.is_in_frustum_points(v["Tcw"], v["Ow"], synth.TUM1_K, b, np.log(1.2), 8, 0.5, v["pos"], v["normal"],
                                                    v["min_dist"], v["max_dist"])
    assert 100 < iv.sum() < 3900
    # fp64 recomputation agrees away from the decision boundaries
    P = v["pos"].astype(np.float64); T = v["Tcw"].astype(np.float64)
    Pc = P @ T[:3, :3].T + T[:3, 3]
    with np.errstate(all="ignore"):
        u = synth.TUM1_K[0] * Pc[:, 0] / Pc[:, 2] + synth.TUM1_K[2]; vv = synth.TUM1_K[1] * Pc[:, 1] / Pc[:, 2] + synth.TUM1_K[3]
    d = np.linalg.norm(P - v["Ow"], axis=1); cosv = ((P - v["Ow"]) * v["normal"]).sum(1) / d
    ok = (Pc[:, 2] >= 0) & (u >= b[0]) & (u <= b[2]) & (vv >= b[1]) & (vv <= b[3]) & (d >= np.float32(0.8) * v["min_dist"]) & (d <= np.float32(1.2) * v["max_dist"]) & (cosv >= 0.5)
    assert (ok != iv.astype(bool)).sum() <= 2
    m = ok & iv.astype(bool)
    assert np.abs(proj[m, 0] - u[m]).max() < 1e-2 and np.abs(vc[m] - cosv[m]).max() < 1e-5
    assert lvl[m].min() >= 0 and lvl[m].max() <= 7 and lvl[m].max() > 0
    assert not proj[~iv.astype(bool)].any()


def test_is_in_frustum_lines_semantics():
    v = synth.synth_map_view(9, 3000, lines=True)
    v["min_dist"] = (v["max_dist"] / 1.2 ** 12).astype(np.float32)
    b = oracle.image_bounds(synth.TUM1_K, synth.TUM1_DIST, 640, 480)
    iv, proj, lvl, vc = oracle.is_in_frustum_lines(v["Tcw"], v["Ow"], synth.TUM1_K, b, np.log(1.2), 0.5, v["pos"], v["normal"],
                                                   v["min_dist"], v["max_dist"])
    assert 50 < iv.sum() < 2950
    m = iv.astype(bool)
    assert (proj[m, 0] >= b[0]).all() and (proj[m, 2] <= b[2]).all() and (vc[m] >= 0.5).all()
    # the line variant does NOT clamp the predicted level (MapLine.cpp:395-404: ceil(log(ratio)/logScaleFactor) bare)
    assert lvl[m].max() > 7


def test_predict_scale_known_answers():
    """MapPoint::PredictScale uses the RAW mfMaxDistance (MapPoint.cc:396-428), not GetMaxDistanceInvariance() = 1.2 * mfMaxDistance:
    a point seen at mfMaxDistance / 1.2^k (plus a hair) is predicted on level k = ceil(log(ratio) / log 1.2); the invariance factors only widen the range test."""
    K = synth.TUM1_K
    b = np.array([0, 0, 640, 480], np.float32)
    Tcw = np.eye(4, dtype=np.float32); Ow = np.zeros(3, np.float32)
    maxd = np.float32(10.0)
    logsf = float(np.float32(np.log(np.float32(1.2))))
    for k in range(8):
        d = float(maxd) / 1.2 ** k * 1.001
        pos = np.array([[0.0, 0.0, d]], np.float32); normal = np.array([[0.0, 0.0, 1.0]], np.float32)
        iv, proj, lvl, vc = oracle.is_in_frustum_points(Tcw, Ow, K, b, logsf, 8, 0.5, pos, normal, np.array([0.1], np.float32), np.array([maxd], np.float32))
        assert iv[0] == 1 and lvl[0] == k, (k, lvl[0])
    # the range test uses 0.8 * min and 1.2 * max: a point at 1.15 * mfMaxDistance is in view (level 0), one at 1.25 * is not
    for d, expect in ((11.5, 1), (12.5, 0), (0.085, 1), (0.075, 0)):
        pos = np.array([[0.0, 0.0, d]], np.float32); normal = np.array([[0.0, 0.0, 1.0]], np.float32)
        iv, _, lvl, _ = oracle.is_in_frustum_points(Tcw, Ow, K, b, logsf, 8, 0.5, pos, normal, np.array([0.1], np.float32), np.array([maxd], np.float32))
        assert iv[0] == expect, (d, iv[0])
