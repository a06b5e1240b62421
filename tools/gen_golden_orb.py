"""Generate tests/golden/orb_*.npz.

Pins (a) the third-party OpenCV arithmetic the reference calls (resize, GaussianBlur, FAST, fastAtan2) using
the cv2 4.13 wheel present in the build container, and (b) the oracle's end-to-end ORB output on the seeded
synthetic frames, so the GPU box (no /root/reference, cv2 optional) can check both.
Run from the repo root: python tools/gen_golden_orb.py
"""
import os, sys
import numpy as np
import cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200  # noqa  (synth only; no GPU needed)
from plslam_b200 import synth
import oracle

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)
rng = np.random.Generator(np.random.PCG64(7))
img = synth.synth_frame(640, 480, 1)
small = np.ascontiguousarray(img[100:196, 200:328])  # 96x128 crop

# (a) cv2 primitives on the crop
rs = cv2.resize(small, (107, 80), interpolation=cv2.INTER_LINEAR)
b7 = cv2.GaussianBlur(small, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
b5 = cv2.GaussianBlur(small, (5, 5), 1, 1, borderType=cv2.BORDER_REFLECT_101)
fast = {}
for th in (20, 7):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
    k = det.detect(small)
    fast[th] = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)
yx = rng.integers(-60000, 60000, (4000, 2)).astype(np.float32)
at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
np.savez_compressed(os.path.join(out, "orb_cv2_primitives.npz"), small=small, resize_107x80=rs, blur7=b7, blur5=b5,
                    fast20=fast[20], fast7=fast[7], atan_yx=yx, atan=at, cv2_version=cv2.__version__)

# (b) oracle end-to-end on seeded frames (inputs are regenerated from the seed by the tests)
for name, (w, h, seed, nf) in {"640x480_n1000": (640, 480, 1, 1000), "640x480_n2000": (640, 480, 1, 2000),
                               "752x480_n1000": (752, 480, 5, 1000), "1241x376_n2000": (1241, 376, 4, 2000)}.items():
    im = synth.synth_frame(w, h, seed)
    o = oracle.OrbOracle(nf, 1.2, 8, 20, 7)
    kps, desc = o.extract(im)
    np.savez_compressed(os.path.join(out, f"orb_oracle_{name}.npz"), kps=kps, desc=desc,
                        img_sum=np.int64(im.astype(np.int64).sum()), params=np.array([w, h, seed, nf]))
    print(name, len(kps))
