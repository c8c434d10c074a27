#!/bin/bash
# Builds tmp_ab/libprobe.so: the product library with match_hamming_tc.cu compiled with -DMVS_TC_PROBE (per-CTA clock64 /
# %globaltimer counters of the matcher, read by tools/knn_probe.py through MVS_LIB_OVERRIDE).
# Run after `make -C mvslam_b200/csrc` (it links the other objects from mvslam_b200/csrc/build).
set -e
cd "$(dirname "$0")/../mvslam_b200/csrc"
mkdir -p ../../tmp_ab/objp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O2 -I../../include \
     -DMVS_TC_PROBE -c match_hamming_tc.cu -o ../../tmp_ab/objp/match_hamming_tc.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tmp_ab/libprobe.so \
     $(ls build/*.o | grep -v match_hamming_tc) ../../tmp_ab/objp/match_hamming_tc.o -lcuda -ldl
echo built tmp_ab/libprobe.so
