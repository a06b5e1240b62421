"""tests/golden/frame_cv2.npz: cv2 outputs that pin the Frame-glue restatement (oracle/oracle_frame.cpp):
initUndistortRectifyMap / remap / undistortPoints for the TUM1 and EuRoC cameras, and cv2.gemm 3x3*3x1+3x1 samples that pin
the fp32 accumulation order used by every Rcw*x3Dw+tcw in the path.  Run here (cv2 is not on the GPU box)."""
import os, sys
import numpy as np, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200  # noqa
from plslam_b200 import synth
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frame_cv2.npz")
d = {"cv2_version": cv2.__version__}
for name, (K4, D5, w, h, seed) in {"tum1": (synth.TUM1_K, synth.TUM1_DIST, 640, 480, 1), "euroc": (synth.EUROC_K, synth.EUROC_DIST, 752, 480, 5)}.items():
    K = np.eye(3, dtype=np.float32); K[0, 0], K[1, 1], K[0, 2], K[1, 2] = K4
    D = np.array(D5, np.float32).reshape(5, 1)          # Tracking.cc:53-120 keeps both as CV_32F
    mx, my = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (w, h), cv2.CV_32F)
    img = synth.synth_frame(w, h, seed)
    und = cv2.remap(img, mx, my, cv2.INTER_LINEAR)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(0, w, 1500), rng.uniform(0, h, 1500)], 1).astype(np.float32)
    pts = np.concatenate([pts, np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32)])
    upts = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, np.eye(3), K).reshape(-1, 2)
    # sample of the maps (the full maps are 2.4 MB): every 7th row/col + the remap result checksum rows
    d[f"{name}_mx_s"] = mx[::7, ::7]; d[f"{name}_my_s"] = my[::7, ::7]
    d[f"{name}_und"] = und; d[f"{name}_pts"] = pts; d[f"{name}_upts"] = upts
rng = np.random.default_rng(11)
A = rng.normal(0, 1, (4000, 3, 3)).astype(np.float32); x = (rng.normal(0, 3, (4000, 3, 1))).astype(np.float32)
c = rng.normal(0, 1, (4000, 3, 1)).astype(np.float32)
d["gemm_A"], d["gemm_x"], d["gemm_c"] = A, x, c
d["gemm_out"] = np.stack([cv2.gemm(A[i], x[i], 1.0, c[i], 1.0) for i in range(4000)])
np.savez_compressed(out, **d)
print(out, os.path.getsize(out))
