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
    eq = np.array([kl[i].tobytes() == okl[i].tobytes() for i in range(nl)])
    assert nl == len(okl) and eq.mean() > 0.99 and np.array_equal(ldesc[eq], oldesc[eq])
