#!/usr/bin/env python
"""Developer experiment: does the cooperative pricing kernel, launched on a high-priority side stream,
run CONCURRENTLY with a fused update kernel that fills the GPU?  (Timing only: the two kernels race on
the plan, the results are garbage.)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplex_method_solver_b200 import _native as N  # noqa: E402
from simplex_method_solver_b200 import workloads as W  # noqa: E402
from simplex_method_solver_b200.engine import DeviceTableau  # noqa: E402


def main():
    n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 32768)
    rows, c = W.dense_lp(n, m, 0)
    tab = DeviceTableau(n, m)
    tab.load(rows, c, max_pivots=1 << 40)
    st = tab.read_state(); st.reserved[0] = 0; tab.write_state(st)
    for _ in range(3):
        tab.fused_pass(8, 0)
    torch.cuda.synchronize()
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1)
    for trial in range(4):
        tab.fused_pass(8, 1)                       # a valid plan for the update
        torch.cuda.synchronize()
        e0, em, es = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(main_s)
        side.wait_event(e0)
        tab.fused_pass(8, 2)                       # update on the main stream (2.5 ms)
        em.record(main_s)
        with torch.cuda.stream(side):
            tab.fused_pass(8, 1)                   # pricing on the side stream, concurrently
            es.record(side)
        torch.cuda.synchronize()
        print(f"trial {trial}: update done at {e0.elapsed_time(em):.3f} ms, side pricing done at {e0.elapsed_time(es):.3f} ms",
              flush=True)
        st = tab.read_state(); st.status = N.PIVOT; tab.write_state(st)


if __name__ == "__main__":
    main()
