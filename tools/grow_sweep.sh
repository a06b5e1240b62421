#!/bin/bash
# k_lsd_grow time of one launch at several batch sizes (CUDA events inside the library)
for B in 1 8 64 512 4736; do B=$B timeout 600 python tools/bench_grow.py pl-slam_b200/libplslam_b200.so; done
