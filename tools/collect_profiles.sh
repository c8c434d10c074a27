#!/bin/bash
# Run on the GPU box (gpurun): benches + ncu captures for profiles/.  One ncu "session" per gpurun call.
set -u
mkdir -p gpurun_out
R=${1:-r1}
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tsukuba_$R.json 2> gpurun_out/bench_tsukuba_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2>/dev/null
python bench.py --workload s8k --steps 5 --warmup 3 > gpurun_out/bench_s8k_$R.json 2>/dev/null
python bench.py --workload w512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w512_$R.json 2>/dev/null
python tools/l2_bench.py 32768 5 > gpurun_out/l2_bench_$R.json 2>/dev/null
./tools/ubench > gpurun_out/ubench_$R.json
python tools/orb_bench.py 256 2000 10 > gpurun_out/orb_bench_$R.json 2>/dev/null
python tools/orb_bench.py 1 2000 10 > gpurun_out/orb_bench1_$R.json 2>/dev/null
python tools/orb_bench.py 64 500 10 > gpurun_out/orb_bench500_$R.json 2>/dev/null
python tools/pnp_bench.py 1024 500 100 > gpurun_out/pnp_bench_$R.json 2>/dev/null
python tools/pnp_bench.py 64 5000 4096 > gpurun_out/pnp_bench_big_$R.json 2>/dev/null
python tools/ba_bench.py 1024 200 > gpurun_out/ba_bench_$R.json 2>/dev/null
python tools/ba_bench.py 256 1000 > gpurun_out/ba_bench_big_$R.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launch_$R.log 2>&1
$CMD > gpurun_out/plain2_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"knn2|hypotheses|score|select|triangulate" -s 28 -c 5 -o gpurun_out/prof_path_$R $CMD > gpurun_out/ncu_full_$R.log 2>&1
python tools/l2_bench.py 32768 2 > gpurun_out/plain3_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:l2_gemm -s 1 -c 1 -o gpurun_out/prof_l2_$R python tools/l2_bench.py 32768 2 > gpurun_out/ncu_l2_$R.log 2>&1
python bench.py --workload s8k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain4_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"score_kernel" -s 3 -c 1 -o gpurun_out/prof_score_s8k_$R python bench.py --workload s8k --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_score_$R.log 2>&1
python tools/orb_bench.py 256 2000 3 > gpurun_out/plain5_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"orb_" -s 18 -c 6 -o gpurun_out/prof_orb_$R python tools/orb_bench.py 256 2000 1 > gpurun_out/ncu_orb_$R.log 2>&1
ls -la gpurun_out
