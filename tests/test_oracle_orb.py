"""CPU tests: the ORB oracle against (a) committed cv2-4.13 golden vectors, (b) live cv2 when importable,
(c) its own committed end-to-end goldens (regression pin for the seeded synthetic frames)."""
import os
import numpy as np
import pytest
import oracle
from plslam_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def test_primitives_vs_cv2_golden():
    g = np.load(os.path.join(G, "orb_cv2_primitives.npz"))
    small = g["small"]
    assert np.array_equal(oracle.resize_linear_u8(small, 107, 80), g["resize_107x80"])
    assert np.array_equal(oracle.blur_u8(small, 7), g["blur7"])
    assert np.array_equal(oracle.blur_u8(small, 5), g["blur5"])
    for th in (20, 7):
        k = oracle.fast_detect(small, th)
        mine = np.stack([k["x"], k["y"], k["response"]], 1).astype(np.float32).reshape(-1, 3)
        assert np.array_equal(mine, g[f"fast{th}"])
    at = np.array([oracle.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(at, g["atan"])


def test_primitives_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    img = synth.synth_frame(640, 480, 3)
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    o.extract(img)
    prev = img
    for l in range(1, 8):
        w, h = o.level_dims(l)
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, o.level(l)), f"resize level {l}"
        assert np.array_equal(cv2.GaussianBlur(ref, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101), o.blurred(l))
        prev = ref
    assert np.array_equal(cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101), o.level(0, True))
    # per-cell FAST on random windows, both thresholds
    rng = np.random.default_rng(0)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
        for _ in range(60):
            w = rng.integers(7, 45); h = rng.integers(4, 45)
            x0 = rng.integers(0, 640 - w); y0 = rng.integers(0, 480 - h)
            sub = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
            r = [(k.pt[0], k.pt[1], k.response) for k in det.detect(sub)]
            m = [(float(k["x"]), float(k["y"]), float(k["response"])) for k in oracle.fast_detect(sub, th)]
            assert r == m


def test_scale_tables_and_quotas():
    t = oracle.OrbOracle(1000, 1.2, 8, 20, 7).tables()
    assert list(t["per_level"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(oracle.OrbOracle(2000, 1.2, 8, 20, 7).tables()["per_level"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    o.extract(synth.synth_frame(640, 480, 1))
    assert [o.level_dims(l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                   (257, 193), (214, 161), (179, 134)]


@pytest.mark.parametrize("name", ["640x480_n1000", "640x480_n2000", "752x480_n1000", "1241x376_n2000"])
def test_end_to_end_golden(name):
    g = np.load(os.path.join(G, f"orb_oracle_{name}.npz"))
    w, h, seed, nf = [int(v) for v in g["params"]]
    img = synth.synth_frame(w, h, seed)
    assert int(img.astype(np.int64).sum()) == int(g["img_sum"])  # the generator is bit-stable
    kps, desc = oracle.OrbOracle(nf, 1.2, 8, 20, 7).extract(img)
    assert kps.tobytes() == g["kps"].tobytes()
    assert np.array_equal(desc, g["desc"])


def test_quadtree_properties():
    # every selected keypoint is a candidate; per level the count is >= min(N, #candidates-ish) and <= N+3
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    o.extract(synth.synth_frame(640, 480, 2))
    per = o.tables()["per_level"]
    for l in range(8):
        c, s = o.candidates(l), o.selected(l)
        cs = {(int(k["x"]), int(k["y"])) for k in c}
        assert all((int(k["x"]) - 16, int(k["y"]) - 16) in cs for k in s)
        assert len(s) <= per[l] + 3
        assert len(s) >= min(per[l], len(cs) and 1)


def test_degenerate_images():
    o = oracle.OrbOracle(1000, 1.2, 8, 20, 7)
    kps, desc = o.extract(np.full((480, 640), 128, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
