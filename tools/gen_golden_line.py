"""tests/golden/lsd_cv2_*.npz: cv2 4.13 createLineSegmentDetector() outputs on seeded synthetic frames (the pin for the
LSD restatement); tests/golden/line_oracle_*.npz: oracle LINEextractor outputs (regression pin, used on the GPU box)."""
import os, sys
import numpy as np, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200  # noqa
from plslam_b200 import synth
import oracle
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, (w, h, seed) in {"640x480_s1": (640, 480, 1), "640x480_s2": (640, 480, 2), "752x480_s5": (752, 480, 5), "1241x376_s4": (1241, 376, 4)}.items():
    img = synth.synth_frame(w, h, seed)
    segs = cv2.createLineSegmentDetector().detect(img)[0].reshape(-1, 4)
    np.savez_compressed(os.path.join(out, f"lsd_cv2_{name}.npz"), segments=segs, cv2_version=cv2.__version__)
    print(name, len(segs))
kl, desc, lf = oracle.line_extract(synth.synth_frame(640, 480, 1))
np.savez_compressed(os.path.join(out, "line_oracle_640x480_s1.npz"), kl=kl, desc=desc, lf=lf)
