#!/bin/bash
# Run on the GPU box (gpurun): round-2 benches + ncu captures for profiles/.  ncu runs only after the same command exited 0.
set -u
mkdir -p gpurun_out
R=r2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tsukuba_$R.json 2> gpurun_out/bench_tsukuba_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2>/dev/null
./tools/ubench > gpurun_out/ubench_$R.json
./tools/ubench_umma > gpurun_out/ubench_umma_$R.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launch_$R.log 2>&1
$CMD > gpurun_out/plain2_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"knn2|match_finalize|hypotheses|score|select|triangulate|finish" -s 21 -c 7 -o gpurun_out/prof_path_$R $CMD > gpurun_out/ncu_full_$R.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --solver fast --hypotheses 1024"
$CMD2 > gpurun_out/plain3_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hypotheses|score|triangulate" -s 9 -c 3 -o gpurun_out/prof_fast_$R $CMD2 > gpurun_out/ncu_fast_$R.log 2>&1
kill $SMI
ls -la gpurun_out | tail -20
# float L2 matcher (BASELINE config 4): launch list and one full capture of the tensor-core kernel
python tools/l2_bench.py 32768 4 > gpurun_out/l2_bench_$R.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/l2_launches_$R.csv python tools/l2_bench.py 32768 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:l2_gemm -s 2 -c 1 -o gpurun_out/prof_l2_$R python tools/l2_bench.py 32768 4 > gpurun_out/ncu_l2_$R.log 2>&1
# matcher clock / wait counters (needs tmp_ab/libprobe.so from tools/build_tc_probe.sh: match_hamming_tc.cu with -DMVS_TC_PROBE)
if [ -f tmp_ab/libprobe.so ]; then MVS_LIB_OVERRIDE=tmp_ab/libprobe.so TAG=probe H=1 SOLVER=reference REPS=10 python tools/knn_probe.py > gpurun_out/tc_probe_$R.json 2>/dev/null; fi
ls -la gpurun_out | tail -8
