#!/usr/bin/env bash
# Round-2 call R (8 GPUs): bench N=8 with the persistent pricing engine (the default from 8 ranks on).
set -u
out=gpurun_out/r2r
mkdir -p "$out"
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-batched > "$out/bench_n8_persistent.json" 2> "$out/bench_n8_persistent.err"
echo "bench N=8 (auto = persistent): exit $?" | tee -a "$out/summary.txt"
python - "$out/bench_n8_persistent.json" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, round(d["e2e"]["value"]), d["implementation"].get("price_engine", "")[:12],
          d["implementation"].get("exchange_fallback"), d["parity"][:60], d.get("pricing_level_breakdown"))
except Exception as e:
    print("no bench line:", e)
P
tail -n 3 "$out/bench_n8_persistent.err"
