#!/usr/bin/env bash
set -u
out=gpurun_out/r2g
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -q > "$out/gpu_tests.log" 2>&1
echo "pytest -m gpu: exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/gpu_tests.log"
for p in 1000 1008 2000; do timeout 300 python tools/fused_lab.py --pivots $p --depths 8 --minb 0 --items 0 --variants 0:0 2>&1 | grep fused; done | tee "$out/fixup_copy.log"
