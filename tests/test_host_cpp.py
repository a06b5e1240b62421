"""The C++ drop-in classes (pl-slam_b200/host): they compile against the C ABI without OpenCV (CPU test) and, on a GPU,
produce exactly what the oracle produces when driven like Frame::ExtractORB / Frame::ExtractLSD drive the reference."""
import os
import struct
import subprocess
import numpy as np
import pytest
import oracle
import plslam_b200 as pl
from plslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "host", "host_demo")


def test_host_classes_compile_and_link():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s"])
    assert os.path.exists(DEMO)


@pytest.mark.gpu
def test_host_classes_match_oracle(tmp_path):
    if not os.path.exists(DEMO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s"])
    img = synth.synth_frame(640, 480, 1)
    raw, out = tmp_path / "frame.raw", tmp_path / "out.bin"
    raw.write_bytes(img.tobytes())
    subprocess.check_call([DEMO, str(raw), "640", "480", str(out)])
    b = out.read_bytes()
    n = struct.unpack_from("<i", b, 0)[0]; off = 4
    kps = np.frombuffer(b, pl.KP_DTYPE, n, off); off += 28 * n
    desc = np.frombuffer(b, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
    nl = struct.unpack_from("<i", b, off)[0]; off += 4
    kl = np.frombuffer(b, pl.KEYLINE_DTYPE, nl, off); off += 68 * nl
    ldesc = np.frombuffer(b, np.uint8, 32 * nl, off).reshape(nl, 32)
    okps, odesc = oracle.OrbOracle(1000, 1.2, 8, 20, 7).extract(img)
    assert kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    okl, oldesc, _ = oracle.line_extract(img)
    assert nl == len(okl) and kl.tobytes() == okl.tobytes() and np.array_equal(ldesc, oldesc)


PIPE = os.path.join(ROOT, "tests", "host", "host_pipeline")


def test_host_matcher_classes_compile_and_link():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s", "host_pipeline"])
    assert os.path.exists(PIPE)


@pytest.mark.gpu
def test_host_pipeline_matches_oracle(tmp_path):
    """ORB_SLAM2::FrameUndistorter / ORBmatcher / LSDmatcher / Optimizer (pl-slam_b200/host/Matchers.h) driven from C++ like
    Frame::Frame, MonocularInitialization and TrackWithMotionModel drive the reference; compared with the oracle."""
    if not os.path.exists(PIPE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s", "host_pipeline"])
    K, D = synth.TUM1_K, synth.TUM1_DIST
    f = synth.synth_sequence(2, 640, 480, seed=9)
    p = synth.synth_pose_problem(91)
    (tmp_path / "f1.raw").write_bytes(f[0].tobytes()); (tmp_path / "f2.raw").write_bytes(f[1].tobytes())
    np_, nl_ = len(p["pt_obs"]), len(p["line_func"])
    blob = (p["Tcw0"].astype(np.float32).tobytes() + np.asarray(p["K"], np.float32).tobytes() + struct.pack("<i", np_) +
            p["pt_obs"].astype(np.float32).tobytes() + p["pt_inv_sigma2"].astype(np.float32).tobytes() + p["pt_Xw"].astype(np.float32).tobytes() +
            struct.pack("<i", nl_) + p["line_func"].astype(np.float64).tobytes() + p["line_Xw"].astype(np.float64).tobytes())
    (tmp_path / "prob.bin").write_bytes(blob)
    out = tmp_path / "out.bin"
    subprocess.check_call([PIPE, str(tmp_path / "f1.raw"), str(tmp_path / "f2.raw"), "640", "480", str(tmp_path / "prob.bin"), str(out)])
    b = out.read_bytes(); off = 0

    def rd(fmt, n=1):
        nonlocal off
        v = np.frombuffer(b, fmt, n, off); off += v.nbytes
        return v
    n1, nm = rd("<i4", 2); m12 = rd("<i4", n1)
    nl1, nlm = rd("<i4", 2); lm = rd("<i4", nl1)
    n2 = rd("<i4")[0]; ku2 = rd(pl.KP_DTYPE, n2); bounds = rd("<f4", 4)
    inl = rd("<i4")[0]; T = rd("<f4", 16).reshape(4, 4); po = rd("u1", np_); lo = rd("u1", nl_)
    # oracle, same sequence of calls
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    feats = []
    for k in range(2):
        kp, de = o.extract(f[k])
        _, ld, _ = oracle.line_extract(oracle.undistort_remap(f[k], K, D))
        feats.append((oracle.undistort_keypoints(kp, K, D), de, ld))
    ob = oracle.image_bounds(K, D, 640, 480)
    assert np.array_equal(bounds, ob) and ku2.tobytes() == feats[1][0].tobytes()
    pm = np.stack([feats[0][0]["x"], feats[0][0]["y"]], 1).astype(np.float32)
    onm, om, _ = oracle.search_for_initialization(feats[0][0], feats[0][1], feats[1][0], feats[1][1], ob, pm, 100, 0.9, True)
    assert n1 == len(feats[0][0]) and nm == onm and np.array_equal(m12, om)
    # the line path is byte-exact end to end (tests/test_line_gpu.py), so the line matches are the oracle's, exactly
    onl, olm = oracle.search_double(feats[0][2], feats[1][2], 0.7)
    assert nl1 == len(feats[0][2]) and nlm == onl and np.array_equal(lm, olm)
    on, oT, opo, olo, _ = oracle.pose_optimization(0, p["Tcw0"], p["K"], p["pt_obs"], p["pt_inv_sigma2"], p["pt_Xw"], p["line_func"], p["line_Xw"])
    assert inl == on and np.array_equal(po, opo) and np.array_equal(lo, olo)
    assert np.linalg.norm(T[:3, 3] - oT[:3, 3]) <= 1e-4 * np.linalg.norm(oT[:3, 3])


GLUE = os.path.join(ROOT, "tests", "host", "glue_track")


def test_reference_signature_glue_compiles_and_links():
    """reference_glue.cc: ORBmatcher / LSDmatcher / Optimizer with the reference's own signatures (Frame&, KeyFrame*, Map*),
    compiled against the mock map classes of reference_mock.h."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s", "glue_track"])
    assert os.path.exists(GLUE)


@pytest.mark.gpu
def test_glue_track_with_motion_model_matches_oracle(tmp_path):
    """Tracking::TrackWithMotionModel's call sequence (Tracking.cc:1345-1372) through the reference-signature classes on mock
    Frame / MapPoint / MapLine objects; every call recomputed by the CPU oracle from the flat inputs the driver dumps."""
    if not os.path.exists(GLUE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s", "glue_track"])
    K = np.asarray(synth.TUM1_K, np.float32)
    f = synth.synth_sequence(2, 640, 480, seed=9)
    (tmp_path / "f1.raw").write_bytes(f[0].tobytes()); (tmp_path / "f2.raw").write_bytes(f[1].tobytes())
    out = tmp_path / "glue.bin"
    subprocess.check_call([GLUE, str(tmp_path / "f1.raw"), str(tmp_path / "f2.raw"), "640", "480", str(out)])
    b = out.read_bytes(); off = 0

    def rd(fmt, n=1):
        nonlocal off
        v = np.frombuffer(b, fmt, n, off); off += v.nbytes
        return v
    NL_, NLL, NC, NCL = rd("<i4", 4)
    bounds = rd("<f4", 4); Tcw = rd("<f4", 16).reshape(4, 4); Ow = rd("<f4", 3)
    lk = rd(pl.KP_DTYPE, NL_); lku = rd(pl.KP_DTYPE, NL_); ldesc = rd("u1", 32 * NL_).reshape(NL_, 32)
    mp = rd(np.dtype([("valid", "u1"), ("has", "u1"), ("X", "<f4", 3)]), NL_)
    lkl = rd(pl.KEYLINE_DTYPE, NLL); lld = rd("u1", 32 * NLL).reshape(NLL, 32)
    ml = rd(np.dtype([("cand", "u1"), ("has", "u1"), ("P", "<f8", 6), ("n", "<f8", 3), ("md", "<f4", 2)]), NLL)
    cku = rd(pl.KP_DTYPE, NC); cdesc = rd("u1", 32 * NC).reshape(NC, 32)
    ckl = rd(pl.KEYLINE_DTYPE, NCL); cld = rd("u1", 32 * NCL).reshape(NCL, 32); clf = rd("<f8", 3 * NCL).reshape(NCL, 3)
    nmatches, lmatches = rd("<i4", 2)
    cur_mp = rd("<i4", NC); cur_ml = rd("<i4", NCL)
    inl = rd("<i4")[0]; Tout = rd("<f4", 16).reshape(4, 4); pout = rd("u1", NC).astype(bool); lout = rd("u1", NCL).astype(bool)
    assert off == len(b)
    sf = np.cumprod(np.r_[np.float32(1.0), np.full(7, np.float32(1.2))]).astype(np.float32)     # mvScaleFactor: cumulative fp32 products
    # --- ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, 15, mono)
    onm, om = oracle.search_by_projection_last(cku, cdesc, bounds, Tcw, K, sf, mp["valid"], mp["X"], ldesc, lk["octave"], lku["angle"], 15.0, True,
                                               np.zeros(NC, np.uint8))
    assert nmatches == onm and nmatches >= 20
    assert np.array_equal(cur_mp, om)              # mnId of a mock MapPoint == its last-frame keypoint index
    # --- LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, 15): isInFrustum on the candidates, then the search
    cand = ml["cand"].astype(bool)
    iv, proj, lvl, vc = oracle.is_in_frustum_lines(Tcw, Ow, K, bounds, float(np.float32(np.log(np.float32(1.2)))), 0.5, ml["P"][cand], ml["n"][cand],
                                                   ml["md"][cand, 0], ml["md"][cand, 1])
    valid = np.zeros(NLL, np.uint8); lproj = np.zeros((NLL, 4), np.float32)
    valid[np.nonzero(cand)[0]] = iv; lproj[np.nonzero(cand)[0]] = proj
    onl, olm = oracle.line_search_by_projection_last(ckl, clf, cld, bounds, valid, lproj, lld, lkl["lineLength"], 15.0, np.zeros(NCL, np.uint8))
    assert lmatches == onl and lmatches > 5 and np.array_equal(cur_ml, olm)
    # --- Optimizer::PoseOptimization(&mCurrentFrame) on exactly these correspondences
    pi = np.nonzero(cur_mp >= 0)[0]; li = np.nonzero(cur_ml >= 0)[0]
    inv_sigma2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    on, oT, opo, olo, _ = oracle.pose_optimization(0, Tcw, K, np.stack([cku["x"][pi], cku["y"][pi]], 1), inv_sigma2[cku["octave"][pi]], mp["X"][cur_mp[pi]],
                                                   clf[li], ml["P"][cur_ml[li]])
    assert inl == on and np.array_equal(pout[pi], opo) and np.array_equal(lout[li], olo)
    assert not pout[cur_mp < 0].any() and not lout[cur_ml < 0].any()
    assert np.linalg.norm(Tout[:3, 3] - oT[:3, 3]) <= 1e-4 * max(np.linalg.norm(oT[:3, 3]), 1e-3)
