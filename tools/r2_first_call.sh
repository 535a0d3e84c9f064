#!/usr/bin/env bash
# First GPU call of round 2 (see DESIGN.md "Known open items"):
#   gpurun --gpus 2 --timeout 2400 -- bash tools/r2_first_call.sh
# 1. the single-GPU owner-rank cases added after round 1's GPU budget was spent,
# 2. the gated multi-GPU cases (entering column owned by a rank != 0, phase-1, degenerate), one pytest process per
#    exchange mode with a hard timeout each so that one hang cannot eat the call,
# 3. the experimental fused update kernel: bit-exactness, then a schedule sweep on cfg4 (tools/fused_lab.py),
# 4. a short N=2 bench (its preflight reports whether the fused exchange survived the triple-buffer change).
# Everything is logged under gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "late_ranks or fused_sharded_loop_single_rank" > "$out/owner_cases_1gpu.log" 2>&1
echo "owner cases (1 GPU): exit $?" | tee -a "$out/summary.txt"
for mode in fused p2p nccl nccl-ahead; do
    SPX_MULTIGPU_EXTENDED=1 timeout 480 python -m pytest tests/test_multigpu.py -m gpu -q \
        -k "$mode and not dense" -p no:cacheprovider > "$out/multigpu_$mode.log" 2>&1
    echo "multi-GPU extended [$mode]: exit $?" | tee -a "$out/summary.txt"
done
SPX_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k experimental \
    > "$out/experimental_update.log" 2>&1
echo "experimental fused update kernel (bit-exactness): exit $?" | tee -a "$out/summary.txt"
timeout 420 python tools/fused_lab.py --pivots 400 --depths 8 --minb 2,3,4 \
    --variants 0,2:32,6:32,6:64,6:128,7:64,7:128,7:256 > "$out/fused_lab_variants.log" 2>&1
echo "fused_lab variant sweep: exit $?" | tee -a "$out/summary.txt"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-batched > "$out/bench_n2.log" 2>&1
echo "bench N=2: exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out"/*.log
