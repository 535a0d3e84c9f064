"""Device-resident tableau for the streaming pivot kernels (K1/K2/K3).

PyTorch is plumbing here: it allocates the fp64 device buffers and provides the
stream handle; every computation is a call through the C ABI
(include/spx_b200.h) into hand-written sm_100a kernels.

HBM layout (the "split" layout of the header): two ping-pong bodies
A[2][(n+1)][ld] (row n = the f row, ld = 16-double multiple so rows start on
128-byte lines), two b vectors, the gathered pivot column, int32 labels, the
128-byte device state and an optional pivot trace.  The current table is always
in buffer ``npiv & 1``.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _native as N


class Solution(NamedTuple):
    """Outcome of a pivot loop (the information get_solution() carries, without snapshots)."""
    status: int                # N.OPTIMAL / N.INCORRECT / N.NOCONV / N.CAP / N.PIVOT (stopped early)
    npiv: int
    trace: Optional[np.ndarray]   # [npiv, 2] int32 (r, c) or None
    x: np.ndarray              # [m] values of x1..xm (find_optimum generalised)
    obj2: float                # function[0]*x1 + function[1]*x2, simplex.py:49
    objective: float           # sum_j function[j]*x[j]
    rowlab: np.ndarray         # [m] int32 codes of the header labels
    collab: np.ndarray         # [n] int32 codes of the row labels


def as_rows_function(constraints, function):
    """Reference inputs (simplex.py:25) -> dense fp64 [n, m+1] rows and [m] function.

    Never mutates or aliases-for-write the caller's data; a C-contiguous fp64
    ndarray (e.g. a view of pinned memory) is used in place for the upload.
    """
    rows = np.asarray(constraints, dtype=np.float64)
    if rows.ndim != 2 or rows.shape[0] < 1 or rows.shape[1] < 2:
        raise ValueError("constraints must be n >= 1 rows of m+1 >= 2 numbers")
    rows = np.ascontiguousarray(rows)
    c = np.ascontiguousarray(np.asarray(function, dtype=np.float64).reshape(-1))
    if c.shape[0] != rows.shape[1] - 1:
        raise ValueError("function must have m = len(constraints[0]) - 1 entries")
    return rows, c


class DeviceTableau:
    def __init__(self, n: int, m: int, device=None, max_pivots: int = 1 << 62,
                 trace_capacity: int = 0):
        N.lib()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise N.NativeUnavailable("the pivot kernels only run on a CUDA device")
        self.n, self.m = int(n), int(m)
        L = N.load()
        self.ld = int(L.spx_ld(self.m))
        self.nb = (self.n + 15) // 16 * 16
        with torch.cuda.device(self.device):
            dev = self.device
            self.A = torch.empty((2, self.n + 1, self.ld), dtype=torch.float64, device=dev)
            self.b = torch.zeros((2, self.nb), dtype=torch.float64, device=dev)
            self.colbuf = torch.zeros(int(L.spx_colbuf_doubles(self.n)), dtype=torch.float64, device=dev)
            self.state = torch.zeros(ctypes.sizeof(N.SpxState) // 8, dtype=torch.int64, device=dev)
            self.rowlab = torch.empty(self.m, dtype=torch.int32, device=dev)
            self.collab = torch.empty(max(self.n, 1), dtype=torch.int32, device=dev)
            self.function = torch.empty(self.m, dtype=torch.float64, device=dev)
            self.x = torch.empty(self.m, dtype=torch.float64, device=dev)
            self.obj = torch.empty(2, dtype=torch.float64, device=dev)
            self.trace = (torch.empty((trace_capacity, 2), dtype=torch.int32, device=dev)
                          if trace_capacity > 0 else None)
            wbytes = max(int(L.spx_solve_workspace_bytes(self.n)), int(L.spx_fused_workspace_bytes(self.n, self.m)))
            self.work = torch.zeros(wbytes // 8, dtype=torch.float64, device=dev)
        self.trace_capacity = int(trace_capacity)
        self.max_pivots = int(max_pivots)
        self._keepalive = None

    # -- plumbing ---------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            N.call(name, *args, self._stream())

    def load(self, rows: np.ndarray, function: np.ndarray, max_pivots: Optional[int] = None):
        """SimplexMethod.__init__ (simplex.py:25-39): upload the table, reset labels and state."""
        assert rows.dtype == np.float64 and rows.flags.c_contiguous and rows.shape == (self.n, self.m + 1)
        assert function.dtype == np.float64 and function.shape == (self.m,)
        if max_pivots is not None:
            self.max_pivots = int(max_pivots)
        if self.trace is not None and self.max_pivots > self.trace_capacity:
            raise ValueError("max_pivots exceeds the trace capacity")
        self._call("spx_import_table", rows.ctypes.data, function.ctypes.data,
                   self.A[0].data_ptr(), self.b[0].data_ptr(), self.n, self.m, self.ld)
        self.function.copy_(torch.from_numpy(function), non_blocking=False)
        self._call("spx_init_state", self.state.data_ptr(), self.rowlab.data_ptr(),
                   self.collab.data_ptr(), self.n, self.m, self.max_pivots)
        # the sources of the async 2-D copies must outlive them
        torch.cuda.current_stream(self.device).synchronize()

    def load_device_flat(self, flat: torch.Tensor, buf: int):
        """Import a device reference-flat table into ping-pong buffer `buf` (state/labels untouched)."""
        rows_ptr = flat.data_ptr()
        fn_ptr = rows_ptr + 8 * self.n * (self.m + 1)
        self._call("spx_import_table", rows_ptr, fn_ptr, self.A[buf].data_ptr(),
                   self.b[buf].data_ptr(), self.n, self.m, self.ld)

    def read_state(self) -> N.SpxState:
        host = self.state.cpu().numpy()
        return N.SpxState.from_buffer_copy(host.tobytes())

    def write_state(self, st: N.SpxState):
        arr = np.frombuffer(bytes(st), dtype=np.int64).copy()
        self.state.copy_(torch.from_numpy(arr))

    def cur(self, npiv: int) -> int:
        return int(npiv) & 1

    # -- K1 + K2 ------------------------------------------------------------------
    def pick(self, npiv: int, rule: int = N.RULE_REFERENCE, sticky: bool = False):
        c = self.cur(npiv)
        self._call("spx_pick", self.A[c].data_ptr(), self.b[c].data_ptr(), self.n, self.m, self.ld,
                   rule, int(sticky), self.state.data_ptr(), self.colbuf.data_ptr())

    # -- K3 -----------------------------------------------------------------------
    def update(self, npiv: int):
        c = self.cur(npiv)
        self._call("spx_update", self.A[c].data_ptr(), self.A[c ^ 1].data_ptr(),
                   self.b[c].data_ptr(), self.b[c ^ 1].data_ptr(), self.n, self.m, self.ld,
                   self.state.data_ptr(), self.colbuf.data_ptr(), self.rowlab.data_ptr(),
                   self.collab.data_ptr(), N.ptr(self.trace) if npiv < self.trace_capacity else None)

    # -- the loop -------------------------------------------------------------------
    def solve(self, rule: int = N.RULE_REFERENCE, chunk: int = 64, stop_after: int = 0,
              lookahead=None):
        """Run the pivot loop on the device until a terminal status / cap / stop_after.

        lookahead selects the loop (all give identical results): None/"auto" — by size: the
        persistent L2-resident kernel for small tableaus, look-ahead streaming for big ones;
        False/"classic" — pick k, update k, ...; True/"lookahead" — price pivot k+1 on a side
        stream while update k streams; "resident" — the persistent cooperative kernel.
        """
        st, npiv = ctypes.c_int32(0), ctypes.c_int64(0)
        mode = N.LOOP_MODES[lookahead]
        with torch.cuda.device(self.device):
            N.call("spx_solve", self.A[0].data_ptr(), self.A[1].data_ptr(), self.b[0].data_ptr(),
                   self.b[1].data_ptr(), self.n, self.m, self.ld, rule, self.state.data_ptr(),
                   self.colbuf.data_ptr(), self.rowlab.data_ptr(), self.collab.data_ptr(),
                   N.ptr(self.trace), int(chunk), int(stop_after), mode, self.work.data_ptr(),
                   self.work.numel() * 8, ctypes.byref(st), ctypes.byref(npiv), self._stream())
        return st.value, npiv.value

    def fused_pass(self, depth: int, phase: int = 0, rule: int = N.RULE_REFERENCE):
        """One pass of the fused loop (phase 0), or its pricing (1) / streaming-update (2) kernel alone.
        The caller keeps state.reserved[0] = index of the buffer holding the current table."""
        self._call("spx_fused_pass", self.A[0].data_ptr(), self.A[1].data_ptr(), self.b[0].data_ptr(),
                   self.b[1].data_ptr(), self.n, self.m, self.ld, rule, int(depth), int(phase),
                   self.state.data_ptr(), self.work.data_ptr(), self.work.numel() * 8, self.rowlab.data_ptr(),
                   self.collab.data_ptr(), N.ptr(self.trace))

    # -- results ----------------------------------------------------------------------
    def export_flat(self, npiv: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Current table -> host reference-flat array (what .table / Info.table expose)."""
        cells = self.n * (self.m + 1) + self.m
        if out is None:
            out = np.empty(cells, dtype=np.float64)
        c = self.cur(npiv)
        base = out.ctypes.data
        self._call("spx_export_table", self.A[c].data_ptr(), self.b[c].data_ptr(), base,
                   base + 8 * self.n * (self.m + 1), self.n, self.m, self.ld)
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def export_device_flat(self, npiv: int, out: torch.Tensor):
        c = self.cur(npiv)
        base = out.data_ptr()
        self._call("spx_export_table", self.A[c].data_ptr(), self.b[c].data_ptr(), base,
                   base + 8 * self.n * (self.m + 1), self.n, self.m, self.ld)

    def b_host(self, npiv: int) -> np.ndarray:
        return self.b[self.cur(npiv), : self.n].cpu().numpy()

    def labels_host(self):
        return self.rowlab.cpu().numpy(), self.collab[: self.n].cpu().numpy()

    def extract(self, npiv: int):
        """find_optimum()/f() for all m variables on the device (simplex.py:48-68)."""
        c = self.cur(npiv)
        self._call("spx_extract", self.b[c].data_ptr(), self.n, self.m, self.collab.data_ptr(),
                   self.function.data_ptr(), self.x.data_ptr(), self.obj.data_ptr())
        x = self.x.cpu().numpy()
        obj = self.obj.cpu().numpy()
        return x, float(obj[0]), float(obj[1])

    def solution(self, status: int, npiv: int) -> Solution:
        x, obj2, objm = self.extract(npiv)
        rl, cl = self.labels_host()
        tr = self.trace[:npiv].cpu().numpy() if self.trace is not None else None
        return Solution(status, npiv, tr, x, obj2, objm, rl, cl)
