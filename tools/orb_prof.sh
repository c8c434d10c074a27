#!/bin/bash
# Run on the GPU box: ncu captures of the feature-extraction kernels (batch of 256 Tsukuba frames, nfeatures 2000).
set -u
mkdir -p gpurun_out
R=${1:-r1}
python tools/orb_bench.py 256 2000 3 > gpurun_out/orb_plain_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"orb_(fast|describe|blur|harris|select|resize)" -s 33 -c 11 -o gpurun_out/prof_orb_$R python tools/orb_bench.py 256 2000 1 > gpurun_out/ncu_orb_$R.log 2>&1
ls -la gpurun_out | tail -5
