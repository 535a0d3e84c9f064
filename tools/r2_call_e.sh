#!/usr/bin/env bash
# power / clock behaviour of the fused loop: nvidia-smi sampled every 100 ms while fused_lab runs 4000 pivots per config
set -u
out=gpurun_out/r2e
mkdir -p "$out"
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv -lms 100 > "$out/smi.csv" 2>&1 &
SMI=$!
sleep 1
timeout 600 python tools/fused_lab.py --pivots 4000 --depths 8 --minb 3 --items 2 --variants 0:128,1:0 > "$out/fused_lab_long.log" 2>&1
echo "exit $?"
sleep 1
kill $SMI
cat "$out/fused_lab_long.log"
awk -F, 'NR>1{print $2, $4, $6}' "$out/smi.csv" | sort | uniq -c | sort -rn | head -40
