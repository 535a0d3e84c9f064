#!/usr/bin/env bash
# Round-2 call S (1 GPU, last of the round's budget): the whole -m gpu suite on the shipped build, smoke(), then the ncu
# launch list of the bench command (bench.py itself exited 0 without ncu earlier in the round: gpurun_out/r2f, r2q, r2r).
set -u
out=gpurun_out/r2s
mkdir -p "$out"
timeout 80 python -m pytest tests -m gpu -q --timeout 60 > "$out/gpu_tests.log" 2>&1
echo "pytest -m gpu: exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/gpu_tests.log"
timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$out/smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/summary.txt"
tail -n 2 "$out/smoke.log"
timeout 55 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$out/launches_r2s.csv" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batched > "$out/ncu_bench.log" 2>&1
echo "ncu launch list: exit $?" | tee -a "$out/summary.txt"
tail -n 2 "$out/ncu_bench.log" | cut -c1-300
