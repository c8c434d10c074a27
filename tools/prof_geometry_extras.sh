set -u
mkdir -p gpurun_out
python tools/ba_bench.py 1024 200 > gpurun_out/ba_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ba_solve" -s 2 -c 1 -o gpurun_out/prof_ba_r1 python tools/ba_bench.py 1024 200 > gpurun_out/ncu_ba.log 2>&1
python tools/pnp_bench.py 1024 500 100 > gpurun_out/pnp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"pnp_" -s 9 -c 3 -o gpurun_out/prof_pnp_r1 python tools/pnp_bench.py 1024 500 100 > gpurun_out/ncu_pnp.log 2>&1
tail -n 2 gpurun_out/ncu_ba.log; tail -n 2 gpurun_out/ncu_pnp.log
