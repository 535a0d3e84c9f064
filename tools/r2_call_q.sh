#!/usr/bin/env bash
# Round-2 call P (2 GPUs): persistent pricing engine after the submission-order fix — two parity cases, then bench N=2
# with the persistent engine and with the per-pass engine on 32 hardware work queues.
set -u
out=gpurun_out/r2q
mkdir -p "$out"
timeout 150 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 100 -x -k "fused-persistent and (24-1100 or 9-40)" > "$out/multigpu_persistent.log" 2>&1
rc=$?
echo "multi-GPU [fused-persistent, 2 cases]: exit $rc" | tee -a "$out/summary.txt"
tail -n 4 "$out/multigpu_persistent.log"
if [ $rc -ne 0 ]; then exit 0; fi
for cfg in "persistent 0"; do
  set -- $cfg
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 3 --warmup 3 --no-batched --price-engine $1 --max-connections $2 > "$out/bench_n2_$1_$2.json" 2> "$out/bench_n2_$1_$2.err"
  echo "bench N=2 $1 conn=$2: exit $?" | tee -a "$out/summary.txt"
  python - "$out/bench_n2_$1_$2.json" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, round(d["e2e"]["value"]), d["implementation"].get("price_engine", "")[:12],
          d["implementation"].get("exchange_fallback"), d["parity"][:60], (d.get("pricing_level_breakdown") or {}).get("us_per_level"))
except Exception as e:
    print("no bench line:", e)
P
done
