#!/usr/bin/env bash
# Round-2 call D (1 GPU): the lazy-guard fused update kernel — parity first, then the schedule sweep, then ncu.
set -u
out=gpurun_out/r2d
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > "$out/gpu_parity.log" 2>&1
echo "test_gpu_parity: exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/gpu_parity.log"
timeout 600 python tools/fused_lab.py --pivots 400 --depths 8 --minb 2,3 --items 1,2 \
    --variants 0:64,0:128,0:256 > "$out/fused_lab.log" 2>&1
echo "fused_lab sweep: exit $?" | tee -a "$out/summary.txt"
cat "$out/fused_lab.log"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:update_lazy_kernel -s 20 -c 1 \
    -o "$out/prof_lazy_r2d" -f python tools/fused_lab.py --pivots 400 --depths 8 --minb 3 --items 2 --variants 0:128 > "$out/ncu.log" 2>&1
echo "ncu: exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/ncu.log"
