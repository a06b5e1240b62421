"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/oracle_orb.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.

Pinning (details in each source header and DESIGN.md §2): OpenCV primitives, LSD, undistortion / remap and the fp32 gemm
order are pinned to cv2 4.13 golden vectors; the whole ORB extraction is pinned to the reference's OWN src/ORBextractor.cc, compiled
where it lies into oracle/_ref/libref_orb.so (RefOrb; byte-identical keypoints and descriptors, tests/test_oracle_orb_ref.py and
tests/golden/orb_ref_*.npz); KeyLine construction and the LBD descriptor are pinned the same way to the reference tree's
Thirdparty/line_descriptor sources (oracle/_ref/libref_line.so; tests/test_oracle_line_ref.py, tests/golden/line_ref_*.npz); the
point and line matchers (ORBmatcher.cc, LSDmatcher.cpp, MapPoint.cc -> libref_match.so) and the frame glue (the reference's Frame.cc
against its real Frame.h -> libref_frame.so: the whole monocular Frame constructor, grids, isInFrustum) are pinned the same way; the g2o LM
and BA are "parity unpinned" (the reference ships no vectors for them and Optimizer.cc cannot be built here).
"""
from .binding import *  # noqa
