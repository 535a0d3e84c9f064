#!/usr/bin/env python
"""Developer lab: run the cfg3 batch (65,536 LPs, 8x2) and Klee-Minty n=20 through the warp-resident
solver a few times and print kernel times (CUDA events).  Used under ncu for instruction counts."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplex_method_solver_b200 import workloads as W  # noqa: E402
from simplex_method_solver_b200.batched import DeviceBatch  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    km = "--km" in sys.argv
    T, C = W.gui_batch(65536, 0)
    tabs = torch.from_numpy(W.batch_flat(T, C)).cuda()
    for trace in (True, False):
        db = DeviceBatch(65536, 8, 2, max_pivots=64, trace=trace)
        ts = []
        for _ in range(reps):
            db.T.copy_(tabs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); db.run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        npiv = int(db.npiv.sum().item())
        print(f"cfg3 trace={trace}: best {min(ts):.4f} ms  {65536 / min(ts) / 1e3:.1f} M LPs/s  "
              f"{npiv / min(ts) / 1e6:.2f} G pivots/s", flush=True)
    if km:
        for n in (10, 16, 20):
            rows, c = W.klee_minty(n)
            flat = torch.from_numpy(np.concatenate([np.asarray(rows).reshape(-1), np.asarray(c)])[None, :]).cuda()
            db = DeviceBatch(1, n, n, max_pivots=1 << n, trace=True)
            db.T.copy_(flat)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); db.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            npiv = int(db.npiv[0].item())
            print(f"KM({n}): {npiv} pivots in {dt * 1e3:.1f} ms  {npiv / dt / 1e3:.1f} k pivots/s", flush=True)


if __name__ == "__main__":
    main()
