#!/usr/bin/env python
"""Developer lab: where does the upload of a 16384 x 32768 table from pinned host memory go?"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplex_method_solver_b200 import _native as N  # noqa: E402
from simplex_method_solver_b200.engine import DeviceTableau  # noqa: E402


def main():
    n, m = 16384, 32768
    pinned = torch.empty((n, m + 1), dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = 1.0
    c = np.ones(m)
    rows = pinned.numpy()
    tab = DeviceTableau(n, m)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tab.load(rows, c)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"load(): {dt * 1e3:.1f} ms  {rows.nbytes / dt / 1e9:.1f} GB/s", flush=True)
    # pieces
    dst = tab.A[0]
    src = pinned
    for name, fn in [
        ("body 2D copy (torch)", lambda: dst[:n, :m].copy_(src[:, :m], non_blocking=True)),
        ("b column (torch strided)", lambda: tab.b[0, :n].copy_(src[:, m], non_blocking=True)),
        ("contiguous 4.3 GB", lambda: tab.A[1].view(-1)[: n * (m + 1)].copy_(src.view(-1), non_blocking=True)),
    ]:
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"{name}: {dt * 1e3:.2f} ms", flush=True)


if __name__ == "__main__":
    main()
