#!/usr/bin/env bash
# Round-2 call J (8 GPUs): bench N=8 with the sharded pricing kernel at 128 and 512 threads per CTA, then N=4.
set -u
out=gpurun_out/r2j
mkdir -p "$out"
run() {  # n tag extra-args...
    local n=$1 tag=$2; shift 2
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 "$@" > "$out/bench_$tag.json" 2> "$out/bench_$tag.err"
    echo "bench $tag: exit $?" | tee -a "$out/summary.txt"
    tail -n 2 "$out/bench_$tag.err"
    python - "$out/bench_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "parity_owner_changes")}, d["e2e"]["value"] if d.get("e2e") else None,
          d.get("pricing_level_breakdown"), (d.get("batched") or {}).get("lps_per_s"), (d.get("batched_1M") or {}).get("lps_per_s"))
    print(d["parity"][:120])
except Exception as e:
    print("no line:", e)
PY
}
run 8 n8_t128
run 8 n8_t512 --shard-threads 512 --no-batched
run 4 n4_t128
