#!/usr/bin/env bash
# Round-2 call T (1 GPU, what is left of the budget): the new whole-table cfg4 test alone.
set -u
out=gpurun_out/r2t
mkdir -p "$out"
timeout 50 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 45 -k "cfg4_whole_table" > "$out/cfg4_whole_table.log" 2>&1
echo "cfg4 whole-table test: exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/cfg4_whole_table.log"
