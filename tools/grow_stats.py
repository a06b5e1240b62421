"""Counters of the ordered LSD region growing (library built with -DPL_GROW_STATS): python tools/grow_stats.py lib.so [B] [undist]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200 as pl
from plslam_b200 import synth
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8; W, H = 640, 480
frames = synth.synth_sequence(max(B, 2), W, H, seed=1)[:B]
if len(sys.argv) > 3:
    import bench
    frames = bench.make_inputs(B, 1234, W, H, bench.camera_of("tum")[0])[0] if hasattr(bench, "make_inputs") else frames
d = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
vp = C.c_void_p
L = C.CDLL(os.path.abspath(sys.argv[1]))
L.pl_last_error.restype = C.c_char_p
cfg = pl.binding.PLLineConfig(W, H, 200, 0.0, B, 0, 0)
h = vp()
L.pl_line_create.argtypes = [C.POINTER(pl.binding.PLLineConfig), C.POINTER(vp)]
L.pl_line_extract_batch_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, vp, vp, vp, vp, vp]
L.pl_line_capacity.argtypes = [vp]
assert L.pl_line_create(C.byref(cfg), C.byref(h)) == 0, L.pl_last_error()
cap = L.pl_line_capacity(h)
kl = torch.zeros((B, cap, 68), dtype=torch.uint8, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
lf = torch.zeros((B, cap, 3), dtype=torch.float64, device="cuda"); n = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
out = (C.c_ulonglong * 24)()
L.pl_line_grow_stats(None, 1)
rc = L.pl_line_extract_batch_dev(h, d.data_ptr(), W, W * H, B, None, kl.data_ptr(), desc.data_ptr(), lf.data_ptr(), n.data_ptr(), st.cuda_stream)
assert rc == 0, L.pl_last_error()
L.pl_line_grow_stats(out, 0)
names = ["regions grown", "grow steps (fast)", "steps with live", "decide iterations", "accepts (fast)", "exact arctangent iterations", "regions >= min size",
         "refine entered", "reduce rounds", "region2rect pixels", "grow steps (cold)", "accepts (cold)", "region2rect calls", "reduce pixels", "regions of 1 pixel", "regions of <= 4 pixels", "record loads in grow (fast)", "record loads in grow (cold)"]
for i, nm in enumerate(names): print(f"{nm:32s} {out[i] / B:12.1f} per frame")
