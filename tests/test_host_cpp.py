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
