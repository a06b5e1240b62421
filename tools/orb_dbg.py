import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plslam_b200 as pl, oracle
from plslam_b200 import synth
img = synth.synth_frame(640, 480, 1)
kps, desc = pl.ORBextractor(1000, 1.2, 8, 20, 7)(img)
okps, odesc = oracle.OrbOracle(1000, 1.2, 8, 20, 7).extract(img)
print("TMA off" if os.environ.get("PLSLAM_NO_TMA") else "TMA on", len(kps), len(okps), kps.tobytes() == okps.tobytes(), np.array_equal(desc, odesc))
