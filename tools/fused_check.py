#!/usr/bin/env python
"""Developer check: N pivots of cfg4 through the look-ahead loop and through the fused loop from
the same start; final tables, b, labels and traces must be bit-identical.  Also times one classic
K3 update after each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplex_method_solver_b200 import _native as N  # noqa: E402
from simplex_method_solver_b200 import workloads as W  # noqa: E402
from simplex_method_solver_b200.engine import DeviceTableau  # noqa: E402


def k3_time(tab, npiv):
    ts = []
    for q in range(6):
        tab.pick(npiv + q, sticky=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tab.update(npiv + q); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), tab.read_state()


def main():
    n, m = 16384, 32768
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    rows, c = W.dense_lp(n, m, 0)
    out = {}
    for mode in ("lookahead", "fused"):
        tab = DeviceTableau(n, m, trace_capacity=total + 64)
        tab.load(rows, c, max_pivots=total + 32)
        done = 0
        while done < total:
            k = min(1000, total - done)
            st, npiv = tab.solve(stop_after=k, chunk=k, lookahead=mode)
            done += k
        cur = npiv & 1
        out[mode] = (tab.A[cur].clone(), tab.b[cur].clone(), tab.trace[:npiv].clone(), tab.rowlab.clone(), tab.collab.clone())
        a = tab.A[cur][:, :m]
        print(mode, "status", st, "npiv", npiv, "finite", bool(torch.isfinite(a).all()), "absmax", float(a.abs().max()),
              "zeros", int((a == 0).sum()), "tiny(<1e-290)", int((a.abs() < 1e-290).sum()), flush=True)
        t, s = k3_time(tab, npiv)
        print(mode, f"classic K3 after {npiv} pivots: {t:.3f} ms  p={s.p!r} r={s.r} c={s.c}", flush=True)
        del tab
        torch.cuda.empty_cache()
    A0, b0, t0, r0, c0 = out["lookahead"]
    A1, b1, t1, r1, c1 = out["fused"]
    print("tables identical:", bool(torch.equal(A0[:, :m].view(torch.int64), A1[:, :m].view(torch.int64))),
          "b:", bool(torch.equal(b0[:n].view(torch.int64), b1[:n].view(torch.int64))),
          "trace:", bool(torch.equal(t0, t1)), "labels:", bool(torch.equal(r0, r1) and torch.equal(c0, c1)))


if __name__ == "__main__":
    main()
