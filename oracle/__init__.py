"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/oracle_orb.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.

Pinning (details in each source header and DESIGN.md §2): OpenCV primitives, LSD, undistortion / remap and the fp32 gemm
order are pinned to cv2 4.13 golden vectors; ORB orchestration, LBD, the matchers, the g2o LM and BA are "parity unpinned"
(the reference ships no vectors for them and cannot be built here).
"""
from .binding import *  # noqa
