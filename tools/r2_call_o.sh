#!/usr/bin/env bash
# Round-2 call O (2 GPUs): the persistent pricing engine across ranks — parity cases (winner != rank 0 included), then bench N=2.
set -u
out=gpurun_out/r2o
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt" 2>&1
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 200 -k "fused-persistent" > "$out/multigpu_persistent.log" 2>&1
echo "multi-GPU [fused-persistent]: exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/multigpu_persistent.log"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-batched --price-engine persistent > "$out/bench_n2_persistent.json" 2> "$out/bench_n2_persistent.err"
echo "bench N=2 persistent: exit $?" | tee -a "$out/summary.txt"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2o/bench_n2_persistent.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, d["e2e"]["value"], d["implementation"], d["parity"][:80], d.get("pricing_level_breakdown"))
except Exception as e:
    print("no bench line:", e)
P
tail -n 5 "$out/bench_n2_persistent.err"
