#!/bin/bash
# samples SM clock / power / throttle reasons every 20 ms while the given command runs
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 20 > /tmp/clk.csv &
SMI=$!
sleep 0.3
"$@"
sleep 0.1
kill $SMI
awk -F', ' '{c[$1]++; p+=$2; n++; if ($4 ~ /Active/ && $4 !~ /Not/) pc++} END {for (k in c) printf "%s MHz: %d samples\n", k, c[k]; printf "mean power %.0f W, sw_power_cap active in %d of %d samples\n", p/n, pc, n}' /tmp/clk.csv | sort -n
