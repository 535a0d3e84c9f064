#!/usr/bin/env python
"""Developer lab: the C-side sharded look-ahead loop (spx_shard_*) with ONE rank on a shard-shaped
tableau (default 16384 x 4096 = the 8-GPU column block of cfg4): pivots/s and, under ncu, the
duration of every kernel of the side chain (ahead_candidate, peer_push, ahead_select) next to the
update they must hide behind."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplex_method_solver_b200 import _native as N  # noqa: E402
from simplex_method_solver_b200 import workloads as W  # noqa: E402
from simplex_method_solver_b200.parallel import PeerShardedTableau  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--pivots", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rows, c = W.dense_lp(a.n, a.m, 0)
    sh = PeerShardedTableau(a.n, a.m, 0, 1, torch.device("cuda", 0), trace_capacity=a.pivots * (a.reps + 1) + 64)
    sh.load(rows, c, max_pivots=a.pivots * (a.reps + 1) + 32)
    sh.run(a.pivots)
    st = sh.read_state()
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.run(a.pivots)
        e1.record()
        st = sh.read_state()
        ms = e0.elapsed_time(e1)
        cells = a.n * (a.m + 1) + a.m
        print(f"{a.n}x{a.m}: {a.pivots} pivots in {ms:.2f} ms = {ms / a.pivots * 1e3:.1f} us/pivot, "
              f"{a.pivots / ms * 1e3:.0f} pivots/s, {16.0 * cells * a.pivots / ms / 1e6:.0f} GB/s  status={st.status} npiv={st.npiv}",
              flush=True)
    sh.close()


if __name__ == "__main__":
    main()
