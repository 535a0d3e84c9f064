#!/usr/bin/env bash
# Round-2 call A (2+ GPUs): the gated multi-GPU cases — entering column owned by a rank != 0, phase-1, degenerate —
# one pytest process per exchange mode with a hard timeout each, then a short N=2 bench.
#   gpurun --gpus 2 --timeout 1500 -- bash tools/r2_call_a.sh
set -u
out=gpurun_out/r2a
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt" 2>&1
for mode in fused p2p nccl nccl-ahead; do
    SPX_MULTIGPU_EXTENDED=1 timeout 420 python -m pytest tests/test_multigpu.py -m gpu -q \
        -k "$mode and not dense" -p no:cacheprovider > "$out/multigpu_$mode.log" 2>&1
    echo "multi-GPU extended [$mode]: exit $?" | tee -a "$out/summary.txt"
done
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-batched > "$out/bench_n2.log" 2>&1
echo "bench N=2: exit $?" | tee -a "$out/summary.txt"
tail -n 4 "$out"/*.log
