#!/usr/bin/env bash
# Round-2 call B (1 GPU): the experimental fused update kernel — bit-exactness, then a schedule sweep on cfg4.
set -u
out=gpurun_out/r2b
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "late_ranks or fused_sharded_loop_single_rank" > "$out/owner_cases_1gpu.log" 2>&1
echo "owner cases (1 GPU): exit $?" | tee -a "$out/summary.txt"
SPX_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k experimental \
    > "$out/experimental_update.log" 2>&1
echo "experimental fused update kernel (bit-exactness): exit $?" | tee -a "$out/summary.txt"
timeout 600 python tools/fused_lab.py --pivots 400 --depths 8 --minb 2,3,4 \
    --variants 0,2:32,6:32,2:64,6:64,6:128,7:64,7:128,7:256,3:128 > "$out/fused_lab_variants.log" 2>&1
echo "fused_lab variant sweep: exit $?" | tee -a "$out/summary.txt"
tail -n 40 "$out"/*.log
