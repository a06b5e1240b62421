"""Wire formats: the reference's monocular trajectory writers (src/System.cc:396-464) and the binary
replay dump.  Golden lines are hand-derived from the iostream format (`fixed`, setprecision(6/7/9))."""
import os
import numpy as np
import pytest
from plslam_b200 import trajectory as tr, synth
traj = tr


def _pose(rx, ry, rz, t):
    T = np.eye(4, dtype=np.float32); T[:3, :3] = synth._rot(rx, ry, rz).astype(np.float32); T[:3, 3] = t
    return T


def test_quaternion_branches_match_rotation():
    for ang in [(0.1, -0.2, 0.3), (3.0, 0.1, 0.0), (0.0, 3.1, 0.2), (0.2, 0.0, 3.0), (2.2, 2.2, 0.1), (0, 0, 0)]:
        R = synth._rot(*ang).astype(np.float32)
        x, y, z, w = tr.to_quaternion(R).astype(np.float64)
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(Rq - R).max() < 2e-6 and abs(x * x + y * y + z * z + w * w - 1) < 1e-6


def test_tum_and_kitti_lines():
    T0 = np.eye(4, dtype=np.float32); T0[:3, 3] = [1.5, -2.25, 0.125]
    s = tr.format_keyframe_trajectory_tum([1305031102.175304], [T0])
    assert s == "1305031102.175304 -1.5000000 2.2500000 -0.1250000 0.0000000 0.0000000 0.0000000 1.0000000\n"
    k = tr.format_keyframe_trajectory_mono_kitti([T0])
    assert k == ("1.000000000 0.000000000 0.000000000 -1.500000000 0.000000000 1.000000000 0.000000000 2.250000000 "
                 "0.000000000 0.000000000 1.000000000 -0.125000000\n")
    # a 90 degree turn about z: Rcw = Rz(90), camera centre = -Rcw^T t
    T1 = _pose(0, 0, np.pi / 2, [1, 2, 3])
    line = tr.format_keyframe_trajectory_tum([0.5], [T1]).split()
    assert line[0] == "0.500000" and np.allclose([float(v) for v in line[1:4]], [-2, 1, -3], atol=1e-6)
    assert np.allclose([float(v) for v in line[4:]], [0, 0, -np.sqrt(0.5), np.sqrt(0.5)], atol=1e-6)
    # bad keyframes are skipped (System.cc:414, :451)
    assert tr.format_keyframe_trajectory_tum([0.0, 1.0], [T0, T1], bad=[True, False]).count("\n") == 1


def test_files_and_binary_dump(tmp_path):
    poses = [_pose(0.01 * i, -0.02 * i, 0.03 * i, [0.1 * i, 0, 1]) for i in range(5)]
    p = tmp_path / "KeyFrameTrajectory.txt"
    tr.SaveKeyFrameTrajectoryTUM(str(p), np.arange(5) / 30.0, poses)
    rows = [l.split() for l in open(p)]
    assert len(rows) == 5 and all(len(r) == 8 for r in rows) and all(len(r[1].split(".")[1]) == 7 for r in rows)
    tr.SaveKeyFrameTrajectoryMonoKitti(str(tmp_path / "k.txt"), poses)
    assert all(len(l.split()) == 12 for l in open(tmp_path / "k.txt"))
    B = 2
    out = dict(kps=np.zeros((B, 4), synth.KP_DTYPE), desc=np.arange(B * 4 * 32, dtype=np.uint8).reshape(B, 4, 32),
               n=np.array([3, 4], np.int32), keylines=np.zeros((B, 2, 17), np.float32), ldesc=np.ones((B, 2, 32), np.uint8),
               linefunc=np.full((B, 2, 3), 0.25), nl=np.array([1, 2], np.int32), pt_matches=np.full((B, 4), -1, np.int32),
               n_pt_matches=np.zeros(B, np.int32), line_matches=np.zeros((B, 2), np.int32), n_line_matches=np.zeros(B, np.int32),
               poses=np.random.default_rng(0).normal(size=(2, B, 16)).astype(np.float32), inliers=np.array([[5, 6], [7, 8]], np.int32))
    out["kps"]["x"] = [[1, 2, 3, 4], [5, 6, 7, 8]]
    f = tmp_path / "replay.bin"
    tr.dump_frontend(str(f), out, B)
    B2, back = tr.load_frontend(str(f))
    assert B2 == B and all(back[k].tobytes() == np.ascontiguousarray(out[k]).tobytes() and back[k].shape == out[k].shape for k in out)


@pytest.mark.gpu
def test_c_abi_writers_equal_the_python_formatters(tmp_path):
    """pl_trajectory_format_* (pose records from k_pose_records on the GPU, C formatting) byte for byte against the Python
    restatement of System::SaveKeyFrameTrajectoryTUM / MonoKitti, including the isBad() skip and both quaternion branches."""
    import plslam_b200 as pl
    rng = np.random.default_rng(5)
    poses = []
    for i in range(300):
        ang = rng.uniform(-np.pi, np.pi, 3) * (1.0 if i % 3 else 0.05)          # small and large rotations: trace > 0 and the other branch
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = rng.normal(0, 3, 3)
        poses.append(T.astype(np.float32))
    poses = np.array(poses)
    ts = 1305031102.175304 + np.arange(len(poses)) * 0.0333
    bad = (rng.random(len(poses)) < 0.1)
    assert traj.c_format_keyframe_trajectory_tum(ts, poses, bad) == traj.format_keyframe_trajectory_tum(ts, poses, bad)
    assert traj.c_format_keyframe_trajectory_mono_kitti(poses, bad) == traj.format_keyframe_trajectory_mono_kitti(poses, bad)
    assert traj.c_format_keyframe_trajectory_tum(ts[:0], poses[:0]) == ""
    traj.c_save(tmp_path / "kf_tum.txt", poses, ts); traj.c_save(tmp_path / "kf_kitti.txt", poses)
    assert (tmp_path / "kf_tum.txt").read_text() == traj.format_keyframe_trajectory_tum(ts, poses)
    assert (tmp_path / "kf_kitti.txt").read_text() == traj.format_keyframe_trajectory_mono_kitti(poses)


@pytest.mark.gpu
def test_frontend_dump_round_trip(tmp_path):
    import plslam_b200 as pl
    from plslam_b200 import synth
    B = 2
    frames = synth.synth_sequence(B, 640, 480, seed=3)
    fe = pl.Frontend(640, 480, max_batch=B, lm_caps=(320, 88))
    fe.set_pose_problems([synth.synth_pose_problem(40 + k) for k in range(B)])
    out = fe.run(frames)
    fe.dump(B, tmp_path / "step.bin")
    Bq, q = traj.load_frontend(tmp_path / "step.bin")
    assert Bq == B
    for k in ("kps", "desc", "n", "keylines", "ldesc", "linefunc", "nl", "pt_matches", "n_pt_matches", "line_matches", "n_line_matches", "poses", "inliers"):
        assert q[k].tobytes() == np.ascontiguousarray(out[k]).tobytes(), k
