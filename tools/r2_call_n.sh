#!/usr/bin/env bash
# Round-2 call N (1 GPU): per-phase cycle breakdown of the two L2-resident kernels on cfg2.
set -u
out=gpurun_out/r2n
mkdir -p "$out"
SPX_RESIDENT_STAMPS=1 timeout 200 python tools/cfg2_lab.py 1000 2000 resident,resident-r1 > "$out/cfg2_stamps.log" 2>&1
echo "cfg2 stamps: exit $?" | tee -a "$out/summary.txt"
cat "$out/cfg2_stamps.log"
