import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle, plslam_b200 as pl
from plslam_b200 import synth
img = synth.synth_frame()
ex = pl.ORBextractor(1000, 1.2, 8, 20, 7)
kps, desc = ex(img)
o = oracle.OrbOracle(1000, 1.2, 8, 20, 7); okps, odesc = o.extract(img)
c = ex.debug_candidates(0); oc = o.candidates(0)
print("gpu", len(c), c[:12]); print("ora", len(oc), oc[:12])
gs = {(int(k['x']), int(k['y'])): int(k['response']) for k in c}
os_ = {(int(k['x']), int(k['y'])): int(k['response']) for k in oc}
both = set(gs) & set(os_)
print("common", len(both), "resp equal on common", sum(gs[k] == os_[k] for k in both))
only_g = sorted(set(gs) - set(os_))[:20]; print("only gpu", [(k, gs[k]) for k in only_g])
s = oracle.fast_score_map(img)
print("true score at gpu-only pts", [int(s[y + 16, x + 16]) for (x, y) in only_g])
