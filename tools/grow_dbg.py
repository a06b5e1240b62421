"""One batch through the line extractor; prints the grow kernel's counters (pl_line_debug_ctl)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plslam_b200 as pl
from plslam_b200 import synth
B = int(os.environ.get("B", 1))
seq = synth.synth_sequence(min(B, 16), 640, 480, seed=1)
frames = np.stack([seq[b % len(seq)] for b in range(B)])
ex = pl.LINEextractor(1, 1.2, 200, 0.0, max_batch=B)
L = pl.binding.lib()
L.pl_line_debug_ctl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
for rep in range(2):
    t = time.time()
    out = ex.extract_batch(frames) if B > 1 else ex(frames[0])
    dt = time.time() - t
ctl = np.zeros(16, np.int32)
L.pl_line_debug_ctl(ex._h, 0, ctl.ctypes.data, 16)
print(f"B={B} call {dt*1e3:.1f} ms; frame 0: seeds(NXT)={ctl[0]} F={ctl[1]} segs={ctl[3]} pool={ctl[5]} err={ctl[6]} selfaborts={ctl[8]} headredos={ctl[9]} "
      f"published_with_dep={ctl[10]} refined={ctl[12]} max_iters_per_warp={ctl[13]} lane_steps/32={ctl[14]} rq(h,t)=({ctl[7]},{ctl[15]})")
