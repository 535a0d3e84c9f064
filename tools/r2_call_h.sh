#!/usr/bin/env bash
# Round-2 call H (4 GPUs): the multi-GPU parity cases on 4 ranks (entering column owned by every rank), then bench N=4, N=2.
set -u
out=gpurun_out/r2h
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt" 2>&1
for mode in fused p2p; do
    timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -k "$mode" -p no:cacheprovider > "$out/multigpu_$mode.log" 2>&1
    echo "multi-GPU [$mode] on $(nvidia-smi -L | wc -l) GPUs: exit $?" | tee -a "$out/summary.txt"
    tail -n 3 "$out/multigpu_$mode.log"
done
for n in 4 2; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > "$out/bench_n$n.json" 2> "$out/bench_n$n.err"
    echo "bench N=$n: exit $?" | tee -a "$out/summary.txt"
    tail -n 3 "$out/bench_n$n.err"
    cat "$out/bench_n$n.json"
done
