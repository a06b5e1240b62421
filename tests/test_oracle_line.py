"""CPU tests of the line oracle.  LSD is pinned bit-exactly (full ordered segment lists) against OpenCV: live cv2 when
importable and committed cv2-4.13 goldens otherwise; the LBD prefilter (5x5 blur, Sobel pair) likewise; the LBD band
arithmetic follows the vendored spec copy (no opencv_contrib build exists to pin it: "parity unpinned" for that part)."""
import os
import numpy as np
import pytest
import oracle
from plslam_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"640x480_s1": (640, 480, 1), "640x480_s2": (640, 480, 2), "752x480_s5": (752, 480, 5), "1241x376_s4": (1241, 376, 4)}


@pytest.mark.parametrize("name", list(CASES))
def test_lsd_vs_cv2_golden(name):
    w, h, seed = CASES[name]
    g = np.load(os.path.join(G, f"lsd_cv2_{name}.npz"))
    img = synth.synth_frame(w, h, seed)
    assert np.array_equal(oracle.lsd_detect(img), g["segments"])


def test_lsd_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for seed in (11, 12):
        img = synth.synth_frame(640, 480, seed)
        ref = cv2.createLineSegmentDetector().detect(img)[0].reshape(-1, 4)
        assert np.array_equal(oracle.lsd_detect(img), ref)
    flat = np.full((480, 640), 100, np.uint8)
    assert len(oracle.lsd_detect(flat)) == 0
    sc, mg, an = oracle.lsd_stages(synth.synth_frame(640, 480, 11))
    b = cv2.GaussianBlur(synth.synth_frame(640, 480, 11), (7, 7), 0.75)
    assert np.array_equal(sc, cv2.resize(b, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT))


def test_lbd_prefilter_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    img = synth.synth_frame(640, 480, 3)
    dx, dy = oracle.lbd_sobel(img)
    b = cv2.GaussianBlur(img, (5, 5), 1)
    assert np.array_equal(dx, cv2.Sobel(b, cv2.CV_16S, 1, 0, ksize=3)) and np.array_equal(dy, cv2.Sobel(b, cv2.CV_16S, 0, 1, ksize=3))


def test_line_extract_semantics():
    img = synth.synth_frame(640, 480, 1)
    kl, desc, lf = oracle.line_extract(img, nfeatures=200)
    assert len(kl) == 201                                  # size > nFeatures keeps nFeatures+1 (LineExtractor.cpp:44-67)
    assert (np.diff(kl["response"]) <= 0).all() and list(kl["class_id"]) == list(range(201))
    segs = oracle.lsd_detect(img)
    kl2, desc2, lf2 = oracle.line_extract(img, nfeatures=len(segs) + 50)
    assert len(kl2) == len(segs) + 1                       # size <= nFeatures appends one default KeyLine
    assert kl2[-1]["lineLength"] == 0 and not desc2[-1].any() and np.isnan(lf2[-1]).all()
    # line equation: unit normal, passes through both end points
    assert np.allclose(np.hypot(lf[:, 0], lf[:, 1]), 1)
    assert np.abs(lf[:, 0] * kl["startPointX"] + lf[:, 1] * kl["startPointY"] + lf[:, 2]).max() < 1e-3
    # mask: lines with both end points outside the valid box are dropped
    mask = np.zeros((480, 640), np.uint8); mask[14:465, 14:625] = 255
    klm, _, _ = oracle.line_extract(img, mask=mask, nfeatures=5000)
    inside = lambda x, y: (x >= 14) & (x < 625) & (y >= 14) & (y < 465)
    k = klm[:-1]
    assert (inside(k["startPointX"].astype(int), k["startPointY"].astype(int)) | inside(k["endPointX"].astype(int), k["endPointY"].astype(int))).all()
    # descriptors of the same line in a slightly shifted frame stay close (LBD is a usable descriptor)
    img2 = np.roll(img, 2, axis=1)
    kl3, desc3, _ = oracle.line_extract(img2, nfeatures=200)
    nm, m = oracle.search_double(desc, desc3, 0.7)
    assert nm > 100


def test_line_golden():
    g = np.load(os.path.join(G, "line_oracle_640x480_s1.npz"))
    kl, desc, lf = oracle.line_extract(synth.synth_frame(640, 480, 1))
    assert kl.tobytes() == g["kl"].tobytes() and np.array_equal(desc, g["desc"]) and np.array_equal(lf, g["lf"], equal_nan=True)
