"""CPU tests of the matching oracle: pinned against cv2's BFMatcher (live when importable, and a committed golden),
against the SWAR popcount definition, and against brute-force numpy restatements of the grid queries."""
import os
import numpy as np
import pytest
import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def rand_desc(rng, n):
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(0)
    a, b = rand_desc(rng, 200), rand_desc(rng, 200)
    ref = np.unpackbits(a ^ b, axis=1).sum(1)
    assert [oracle.descriptor_distance(a[i], b[i]) for i in range(200)] == list(ref)
    assert oracle.descriptor_distance(a[0], a[0]) == 0
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_bf_knn2_vs_cv2_golden_and_live():
    g = np.load(os.path.join(G, "match_cv2_knn.npz"))
    idx, dist = oracle.bf_knn2(g["d1"], g["d2"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    # few distinct bytes -> many distance ties, which exercises the tie rule (lower train index first)
    d1 = rng.integers(0, 2, (150, 32), dtype=np.uint8) * 255
    d2 = rng.integers(0, 2, (180, 32), dtype=np.uint8) * 255
    mm = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(d1, d2, 2)
    ref_i = np.array([[m[0].trainIdx, m[1].trainIdx] for m in mm]); ref_d = np.array([[m[0].distance, m[1].distance] for m in mm])
    idx, dist = oracle.bf_knn2(d1, d2)
    assert np.array_equal(idx, ref_i) and np.array_equal(dist, ref_d.astype(np.int32))


def test_grid_and_area_query_vs_numpy():
    rng = np.random.default_rng(1)
    n = 800
    keys = np.zeros(n, oracle.KP_DTYPE)
    keys["x"] = rng.uniform(0, 640, n).astype(np.float32); keys["y"] = rng.uniform(0, 480, n).astype(np.float32)
    keys["octave"] = rng.integers(0, 8, n)
    bounds = [0, 0, 640, 480]
    start, items = oracle.assign_grid(keys, bounds)
    px = np.round((keys["x"] - np.float32(0)) * np.float32(64 / 640)).astype(int)  # np.round is half-even: avoid .5 exactly
    py = np.round((keys["y"] - np.float32(0)) * np.float32(48 / 480)).astype(int)
    ok = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert len(items) == ok.sum()
    for c in rng.integers(0, 64 * 48, 50):
        got = list(items[start[c]:start[c + 1]])
        assert got == [i for i in range(n) if ok[i] and px[i] * 48 + py[i] == c]
    assert sorted(items) == sorted(np.nonzero(ok)[0])


def test_frame_bf_match_properties():
    rng = np.random.default_rng(2)
    d2 = rand_desc(rng, 120)
    d1 = d2[rng.permutation(120)[:100]].copy()
    flip = rng.integers(0, 32, 100)
    d1[np.arange(100), flip] ^= 0x0f       # 4 bits of noise
    nm, m = oracle.search_double(d1, d2, 0.7)
    assert nm == (m >= 0).sum() and nm > 90
    for i in np.nonzero(m >= 0)[0]:
        assert np.unpackbits(d1[i] ^ d2[m[i]]).sum() <= 4
    # degenerate sizes: <2 train rows -> no matches (reference reads out of bounds there)
    assert (oracle.frame_bf_match(d1, d2[:1]) == -1).all()
    assert oracle.search_double(d1[:0], d2)[0] == 0
