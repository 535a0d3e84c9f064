#!/usr/bin/env bash
# Round-2 call F (1 GPU): the whole -m gpu suite, smoke(), then bench.py both arms (driver settings).
set -u
out=gpurun_out/r2f
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -q -x > "$out/gpu_tests.log" 2>&1
echo "pytest -m gpu: exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/gpu_tests.log"
timeout 300 python __graft_entry__.py smoke > "$out/smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/smoke.log"
timeout 900 python bench.py --impl reference --steps ${STEPS:-5} --warmup ${WARMUP:-3} > "$out/bench_ref.json" 2> "$out/bench_ref.err"
echo "bench reference: exit $?" | tee -a "$out/summary.txt"
timeout 1200 python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} > "$out/bench.json" 2> "$out/bench.err"
echo "bench: exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/bench.err"
cat "$out/bench_ref.json" "$out/bench.json"
