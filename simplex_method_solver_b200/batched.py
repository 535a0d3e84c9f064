"""Batches of independent small LPs: one warp per LP (K4, csrc/spx_batched.cu).

Each LP is the reference's whole get_solution() loop
(/root/reference/src/simplex.py:179-199) run by one warp with its tableau in
shared memory; LPs never communicate, so a batch splits over GPUs by
contiguous ranges with no collective (see parallel.shard_range).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _native as N


class BatchResult(NamedTuple):
    status: np.ndarray            # [B] int32: N.OPTIMAL / N.INCORRECT / N.NOCONV / N.CAP
    npiv: np.ndarray              # [B] int32
    x: np.ndarray                 # [B, m]
    obj: np.ndarray               # [B]  function[0]*x1 + function[1]*x2  (simplex.py:49)
    tables: np.ndarray            # [B, cells] final reference-flat tables
    rowlab: np.ndarray            # [B, m] int32 label codes
    collab: np.ndarray            # [B, n] int32
    trace: Optional[np.ndarray]   # [B, max_pivots, 2] int32
    snaps: Optional[np.ndarray]   # [B, max_pivots+1, cells]


class DeviceBatch:
    """Device buffers for one batch shape; reusable across solves (bench / serving)."""

    def __init__(self, B: int, n: int, m: int, max_pivots: int = 64, trace: bool = True,
                 snapshots: bool = False, device=None):
        N.lib()
        self.device = torch.device(device if device is not None else "cuda")
        self.B, self.n, self.m, self.max_pivots = int(B), int(n), int(m), int(max_pivots)
        self.cells = n * (m + 1) + m
        if self.cells > N.load().spx_batched_max_cells():
            raise ValueError(f"{self.cells} cells per LP exceed the warp-resident limit; "
                             "use SimplexMethod / DeviceTableau for large tableaus")
        dev = self.device
        Bq = max(self.B, 1)
        self.T = torch.empty((Bq, self.cells), dtype=torch.float64, device=dev)
        self.x = torch.empty((Bq, m), dtype=torch.float64, device=dev)
        self.obj = torch.empty(Bq, dtype=torch.float64, device=dev)
        self.status = torch.empty(Bq, dtype=torch.int32, device=dev)
        self.npiv = torch.empty(Bq, dtype=torch.int32, device=dev)
        self.rowlab = torch.empty((Bq, m), dtype=torch.int32, device=dev)
        self.collab = torch.empty((Bq, n), dtype=torch.int32, device=dev)
        self.trace = (torch.zeros((Bq, max(self.max_pivots, 1), 2), dtype=torch.int32, device=dev)
                      if trace else None)
        self.snaps = (torch.empty((Bq, self.max_pivots + 1, self.cells), dtype=torch.float64, device=dev)
                      if snapshots else None)

    def upload(self, tables, non_blocking: bool = False):
        """tables: [B, cells] fp64 numpy array or (pinned) CPU tensor."""
        src = torch.from_numpy(tables) if isinstance(tables, np.ndarray) else tables
        self.T[: self.B].copy_(src, non_blocking=non_blocking)

    def run(self, rule: int = N.RULE_REFERENCE):
        """Launch the solver on the resident batch (asynchronous on the current stream)."""
        with torch.cuda.device(self.device):
            N.call("spx_solve_batched", self.T.data_ptr(), self.B, self.n, self.m, rule,
                   self.max_pivots, self.x.data_ptr(), self.obj.data_ptr(), self.status.data_ptr(),
                   self.npiv.data_ptr(), self.rowlab.data_ptr(), self.collab.data_ptr(),
                   N.ptr(self.trace), N.ptr(self.snaps),
                   torch.cuda.current_stream(self.device).cuda_stream)

    def result(self) -> BatchResult:
        B = self.B
        g = lambda t: None if t is None else t[:B].cpu().numpy()  # noqa: E731
        return BatchResult(g(self.status), g(self.npiv), g(self.x), g(self.obj), g(self.T),
                           g(self.rowlab), g(self.collab), g(self.trace), g(self.snaps))


def solve_batched(tables, n: int, m: int, max_pivots: int = 64, rule: str = "reference",
                  trace: bool = True, snapshots: bool = False, device=None) -> BatchResult:
    """Solve B independent LPs of one shape.

    tables: [B, cells] reference-flat fp64 (rows ``[a_1..a_m, b]`` x n, then the m
    function coefficients) — the flattened arguments of
    ``SimplexMethod(constraints, function)`` for each LP.
    """
    tables = np.ascontiguousarray(np.asarray(tables, dtype=np.float64))
    if tables.ndim != 2 or tables.shape[1] != n * (m + 1) + m:
        raise ValueError("tables must be [B, n*(m+1)+m]")
    db = DeviceBatch(tables.shape[0], n, m, max_pivots, trace, snapshots, device)
    if db.B == 0:
        z = np.zeros
        return BatchResult(z(0, np.int32), z(0, np.int32), z((0, m)), z(0), z((0, db.cells)),
                           z((0, m), np.int32), z((0, n), np.int32),
                           z((0, max_pivots, 2), np.int32) if trace else None,
                           z((0, max_pivots + 1, db.cells)) if snapshots else None)
    db.upload(tables)
    db.run(N.RULES[rule])
    return db.result()
