#!/usr/bin/env bash
# Round-2 call I (2 GPUs): the reworked sharded pricing kernel — single-rank cases, the multi-GPU fused cases, bench N=2.
set -u
out=gpurun_out/r2i
mkdir -p "$out"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coop or single_rank or late_ranks or dantzig" > "$out/single_rank.log" 2>&1
echo "single-rank sharded pricing cases: exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/single_rank.log"
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "fused" -p no:cacheprovider > "$out/multigpu_fused.log" 2>&1
echo "multi-GPU [fused]: exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/multigpu_fused.log"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > "$out/bench_n2.json" 2> "$out/bench_n2.err"
echo "bench N=2: exit $?" | tee -a "$out/summary.txt"
tail -n 4 "$out/bench_n2.err"
cat "$out/bench_n2.json"
