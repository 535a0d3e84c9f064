#!/usr/bin/env bash
# Round-2 call L (8 GPUs): the sharded pricing kernel on a SUBSET of the SMs (the rest keep updating), N=8 and N=4.
set -u
out=gpurun_out/r2l
mkdir -p "$out"
run() {  # n tag extra-args...
    local n=$1 tag=$2; shift 2
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 2955$n bench.py --gpus $n --steps 3 --warmup 3 --no-batched "$@" > "$out/bench_$tag.json" 2> "$out/bench_$tag.err"
    echo "bench $tag: exit $?" | tee -a "$out/summary.txt"
    python - "$out/bench_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step")}, d["e2e"]["value"] if d.get("e2e") else None)
    print(d.get("pricing_level_breakdown"))
except Exception as e:
    print("no line:", e)
PY
}
run 8 n8_g64 --shard-ctas 64
run 8 n8_g96 --shard-ctas 96
run 4 n4_g64 --shard-ctas 64
run 4 n4_g148
