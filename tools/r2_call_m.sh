#!/usr/bin/env bash
# Round-2 call M (1 GPU): persistent pricing engine (engine 2) and the look-ahead L2-resident kernel — parity, then cfg2 timing.
set -u
out=gpurun_out/r2m
mkdir -p "$out"
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 150 -k "engine or resident or single_rank or cfg2 or lookahead_state or dantzig" > "$out/gpu_parity_new.log" 2>&1
echo "new-path parity: exit $?" | tee -a "$out/summary.txt"
tail -n 15 "$out/gpu_parity_new.log"
timeout 200 python tools/cfg2_lab.py 1000 2000 resident,resident-r1,resident > "$out/cfg2_lab.log" 2>&1
echo "cfg2_lab: exit $?" | tee -a "$out/summary.txt"
cat "$out/cfg2_lab.log"
timeout 200 python tools/cfg2_lab.py 3000 1500 resident,resident-r1 >> "$out/cfg2_lab.log" 2>&1
timeout 200 python tools/cfg2_lab.py 200 20000 resident,resident-r1 >> "$out/cfg2_lab.log" 2>&1
tail -n 4 "$out/cfg2_lab.log"
