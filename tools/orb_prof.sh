set -u
mkdir -p gpurun_out
python tools/orb_bench.py 256 2000 3 > gpurun_out/orb_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:orb_ -c 60 --csv --log-file gpurun_out/orb_launches_256.csv python tools/orb_bench.py 256 2000 1 > gpurun_out/orb_ncu_256.log 2>&1
python tools/orb_bench.py 1 2000 3 > gpurun_out/orb_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum --clock-control none -k regex:orb_ -c 60 --csv --log-file gpurun_out/orb_launches_1.csv python tools/orb_bench.py 1 2000 1 > gpurun_out/orb_ncu_1.log 2>&1
cat gpurun_out/orb_plain1.log
