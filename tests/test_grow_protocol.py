"""CPU test of the speculative LSD region growing PROTOCOL (no GPU): the lane state machine of
pl-slam_b200/csrc/lsd_grow_core.cuh, compiled for the host, is run by tools/grow_sim.cpp for W warps x 32 lanes under
random / adversarial interleavings (neighbourhood snapshots and their use scheduled separately, lanes stalling, commits
rare and late) and must reproduce the oracle's sequential cv::LineSegmentDetector segment list bit for bit."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import oracle
from oracle import binding as ob
from plslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tools", "bin", "libgrowsim.so")


@pytest.fixture(scope="module")
def sim():
    src = os.path.join(ROOT, "tools", "grow_sim.cpp")
    core = os.path.join(ROOT, "pl-slam_b200", "csrc", "lsd_grow_core.cuh")
    os.makedirs(os.path.dirname(SIM), exist_ok=True)
    if not os.path.exists(SIM) or os.path.getmtime(SIM) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", SIM, src])
    L = C.CDLL(SIM)
    L.grow_sim.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return L


def _run(L, img, warps, seed, mode, lane_cap, window):
    scaled = ob.lsd_stages(img)[0]
    sh, sw = scaled.shape
    out = np.zeros((20000, 4), np.float32)
    stats = (C.c_long * 16)()
    n = L.grow_sim(scaled.ctypes.data, sw, sh, warps, seed, mode, lane_cap, out.ctypes.data, 20000, stats, window)
    return n, out[:max(n, 0)], list(stats)


# mode 0: lock step like a GPU warp; 1: lanes in random order, snapshot / use split, commits rare; 2: + lanes stall at random
@pytest.mark.parametrize("warps,mode,lane_cap,window", [(1, 0, 1024, 0), (1, 1, 64, 4096), (4, 2, 1024, 1024), (64, 0, 1024, 8192),
                                                        (64, 1, 32, 0), (16, 2, 64, 256)])
def test_speculative_growing_equals_sequential_lsd(sim, warps, mode, lane_cap, window):
    K, D = synth.TUM1_K, synth.TUM1_DIST
    frames = synth.synth_sequence(3, 640, 480, seed=1)
    for k, img in enumerate((frames[0], oracle.undistort_remap(frames[1], K, D))):
        ref = np.asarray(ob.lsd_detect(img, 1), np.float32).reshape(-1, 4)
        n, segs, st = _run(sim, img, warps, 17 * k + warps, mode, lane_cap, window)
        assert n == len(ref) and segs.tobytes() == ref.tobytes(), (k, n, len(ref))
        assert st[0] == st[5]                       # every seed was committed exactly once


def test_degenerate_images(sim):
    rng = np.random.Generator(np.random.PCG64(11))
    flat = np.full((480, 640), 128, np.uint8)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    edge = np.zeros((480, 640), np.uint8); edge[:, 320:] = 255
    for img in (flat, noise, edge):
        ref = np.asarray(ob.lsd_detect(img, 1), np.float32).reshape(-1, 4)
        n, segs, st = _run(sim, img, 8, 5, 1, 64, 2048)
        assert n == len(ref) and segs.tobytes() == ref.tobytes()
