"""LocalMapping helpers: MapPoint::ComputeDistinctiveDescriptors and the search half of LSDmatcher::Fuse.
CPU tests: the oracle against an independent numpy restatement / construction properties; GPU tests: kernel == oracle."""
import numpy as np
import pytest
import oracle
from plslam_b200 import synth


def _mp_lists(seed, n_mp=300):
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = rng.integers(0, 40, n_mp); counts[:5] = [0, 1, 2, 3, 64]
    off = np.zeros(n_mp + 1, np.int32); off[1:] = np.cumsum(counts)
    base = rng.integers(0, 256, (n_mp, 32), dtype=np.uint8)
    desc = np.zeros((off[-1], 32), np.uint8)
    for m in range(n_mp):
        for k in range(counts[m]):
            d = base[m].copy()
            flips = rng.integers(0, 256, rng.integers(0, 60))
            for f in flips:
                d[f >> 3] ^= 1 << (f & 7)
            desc[off[m] + k] = d
    return desc, off


def _numpy_distinctive(desc, off):
    out = np.full(len(off) - 1, -1, np.int32)
    for m in range(len(off) - 1):
        d = desc[off[m]:off[m + 1]]
        N = len(d)
        if N == 0:
            continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (N - 1))]
        out[m] = int(np.argmin(med))          # first minimum, as the reference's strict <
    return out


def test_oracle_distinctive_descriptors_equals_numpy():
    desc, off = _mp_lists(3)
    assert np.array_equal(oracle.distinctive_descriptors(desc, off), _numpy_distinctive(desc, off))


def _line_fuse_problem(seed, behind=None):
    import plslam_b200 as pl
    rng = np.random.Generator(np.random.PCG64(seed))
    K = synth.TUM1_K
    v = synth.synth_map_view(seed, 700, lines=True)
    T = v["Tcw"].astype(np.float64)
    P = v["pos"]
    S = P[:, :3] @ T[:3, :3].T + T[:3, 3]; E = P[:, 3:] @ T[:3, :3].T + T[:3, 3]
    ok = (S[:, 2] > 0.2) & (E[:, 2] > 0.2)
    if behind is None:                         # no map line behind the camera: the loop runs to the end
        keep = np.nonzero(ok)[0]
    else:                                      # keep the lines in front, and put ONE behind-camera line at index `behind`
        front = np.nonzero(ok)[0]; back = np.nonzero(~ok & ((S[:, 2] < 0) | (E[:, 2] < 0)))[0]
        keep = np.concatenate([front[:behind], back[:1], front[behind:]])
    for k in ("pos", "normal", "min_dist", "max_dist"):
        v[k] = v[k][keep]
    S, E = S[keep], E[keep]
    n = len(keep)
    u1 = K[0] * S[:, 0] / S[:, 2] + K[2]; v1 = K[1] * S[:, 1] / S[:, 2] + K[3]
    u2 = K[0] * E[:, 0] / E[:, 2] + K[2]; v2 = K[1] * E[:, 1] / E[:, 2] + K[3]
    mid = 0.5 * (P[keep, :3] + P[keep, 3:]); d = np.linalg.norm(mid - v["Ow"], axis=1)
    v["max_dist"] = (d * rng.uniform(0.85, 1.35, n)).astype(np.float32)       # predicted level 0 or 1 (and a few -1 / 2)
    v["min_dist"] = (v["max_dist"] / 1.2 ** 6).astype(np.float32)
    v["normal"] = ((mid - v["Ow"]) / d[:, None] + rng.normal(0, 0.15, (n, 3))).astype(np.float64)
    # keyframe lines: noisy projections of every third map line, plus clutter
    src = np.arange(0, n, 3)
    nk = len(src) + 40
    kl = np.zeros(nk, pl.KEYLINE_DTYPE)
    for j, i in enumerate(src):
        e = np.array([u1[i], v1[i], u2[i], v2[i]]) + rng.normal(0, 0.4, 4)
        kl["startPointX"][j], kl["startPointY"][j], kl["endPointX"][j], kl["endPointY"][j] = e
    cl = rng.uniform([0, 0, 0, 0], [640, 480, 640, 480], (40, 4))
    kl["startPointX"][len(src):], kl["startPointY"][len(src):], kl["endPointX"][len(src):], kl["endPointY"][len(src):] = cl.T
    kl["ptx"] = 0.5 * (kl["startPointX"] + kl["endPointX"]); kl["pty"] = 0.5 * (kl["startPointY"] + kl["endPointY"])
    kl["octave"] = rng.integers(0, 2, nk)
    n_pdesc = nk - 25                           # the last line indices have no row in the POINT descriptor matrix (the :966 quirk)
    pdesc = rng.integers(0, 256, (n_pdesc, 32), dtype=np.uint8)
    ml_desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for j, i in enumerate(src):
        if j < n_pdesc:
            ml_desc[i] = pdesc[j]
            for f in rng.integers(0, 256, rng.integers(0, 30)):
                ml_desc[i, f >> 3] ^= 1 << (f & 7)
    skip = (rng.uniform(0, 1, n) < 0.1).astype(np.uint8)
    return dict(kl=kl, pdesc=pdesc, bounds=np.array([0, 0, 640, 480], np.float32), Tcw=v["Tcw"], Ow=v["Ow"], K=np.array(K, np.float32),
                skip=skip, pos=v["pos"], normal=v["normal"], min_dist=v["min_dist"], max_dist=v["max_dist"], ml_desc=ml_desc, src=src,
                n_pdesc=n_pdesc)


def _args(f):
    return (f["kl"], f["pdesc"], f["bounds"], f["Tcw"], f["Ow"], f["K"], 1.2, float(np.float32(np.log(np.float32(1.2)))), f["skip"], f["pos"],
            f["normal"], f["min_dist"], f["max_dist"], f["ml_desc"], 3.0)


def test_oracle_lsd_fuse_search_properties():
    f = _line_fuse_problem(21)
    bi, bd, stop = oracle.lsd_fuse_search(*_args(f))
    n = len(f["pos"])
    assert stop == n
    assert (bi[f["skip"] > 0] == -1).all() and (bd[f["skip"] > 0] == 256).all()
    good = bd <= 50
    assert good.sum() > 20
    # a fused map line found the keyframe line generated from it, whose POINT-descriptor row carries its code
    gen = {int(i): j for j, i in enumerate(f["src"])}
    hit = [gen.get(int(i), -2) == int(bi[i]) for i in np.nonzero(good)[0]]
    assert np.mean(hit) > 0.95
    assert (bi[good] < f["n_pdesc"]).all()                       # line indices without a point-descriptor row are never returned
    ham = np.unpackbits(f["pdesc"][bi[good]] ^ f["ml_desc"][good], axis=1).sum(1)
    assert np.array_equal(ham, bd[good])
    # the early `return false`: a line behind the camera at index 37 ends the call; nothing from there on is looked at
    g = _line_fuse_problem(22, behind=37)
    bi2, bd2, stop2 = oracle.lsd_fuse_search(*_args(g))
    sk = g["skip"]
    first_unskipped_behind = 37 if not sk[37] else None
    if first_unskipped_behind is not None:
        assert stop2 == 37 and (bi2[37:] == -1).all() and (bd2[37:] == 256).all()
    # a tiny radius finds (almost) nothing
    a = list(_args(f)); a[-1] = 0.05
    assert (oracle.lsd_fuse_search(*a)[1] <= 50).sum() < good.sum() // 4


@pytest.mark.gpu
def test_gpu_distinctive_descriptors():
    import plslam_b200 as pl
    for seed in (3, 4):
        desc, off = _mp_lists(seed)
        best, out = pl.ComputeDistinctiveDescriptors(desc, off, return_desc=True)
        ob = oracle.distinctive_descriptors(desc, off)
        assert np.array_equal(best, ob)
        nz = ob >= 0
        assert np.array_equal(out[nz], desc[off[:-1][nz] + ob[nz]])
    empty = pl.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.zeros(1, np.int32))
    assert len(empty) == 0


@pytest.mark.gpu
def test_gpu_lsd_fuse_search():
    import plslam_b200 as pl
    for seed, behind in ((21, None), (22, 37), (23, 0), (24, None)):
        f = _line_fuse_problem(seed, behind)
        obi, obd, ostop = oracle.lsd_fuse_search(*_args(f))
        bi, bd, stop = pl.LSDmatcher().FuseSearch(*_args(f))
        assert stop == ostop and np.array_equal(bi, obi) and np.array_equal(bd, obd), (seed, behind)
    for th in (1.0, 8.0):
        a = list(_args(_line_fuse_problem(25))); a[-1] = th
        o = oracle.lsd_fuse_search(*a); g = pl.LSDmatcher().FuseSearch(*a)
        assert g[2] == o[2] and np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
