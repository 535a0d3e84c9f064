"""CPU oracle for the tableau pivot loop — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only
as the checker or the reported CPU baseline.  The product package
(``simplex_method_solver_b200``) never imports it and has no CPU fallback.

Parity status: PINNED against outputs of the reference itself
(``/root/reference/src/simplex.py`` executed in the build container; fixtures
in ``tests/golden/`` made by ``tests/golden/make_golden.py``).  See the header
of ``spx_oracle.c`` for the line-by-line citations.

The C restatement works on the reference's ragged table flattened row-major
("reference flat": n rows of m+1 cells, then the f row with m cells,
simplex.py:36-39).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import NamedTuple, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

PIVOT, OPTIMAL, INCORRECT, NOCONV, CAP = 1, 0, -1, -2, -3
STATUS_NAME = {PIVOT: "pivot", OPTIMAL: "optimal", INCORRECT: "incorrect system",
               NOCONV: "simplex method does not converge", CAP: "cap"}

_lib = None


def build(force: bool = False) -> str:
    """Compile liborc.so with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "spx_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liborc.so"] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = ctypes.CDLL(_LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    i64 = ctypes.c_int64
    L.orc_pick.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                           ctypes.POINTER(ctypes.c_int), dp]
    L.orc_pick.restype = ctypes.c_int
    L.orc_pick_rule.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int), dp]
    L.orc_pick_rule.restype = ctypes.c_int
    L.orc_solve_rule.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, ip, dp, ip, ip,
                                 ctypes.POINTER(i64)]
    L.orc_solve_rule.restype = ctypes.c_int
    L.orc_update.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.orc_update.restype = None
    L.orc_init_labels.argtypes = [ip, ip, ctypes.c_int, ctypes.c_int]
    L.orc_extract.argtypes = [dp, ctypes.c_int, ctypes.c_int, ip, dp, dp, dp, dp]
    L.orc_solve.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, i64, ip, dp, ip, ip,
                            ctypes.POINTER(i64)]
    L.orc_solve.restype = ctypes.c_int
    L.orc_solve_batched.argtypes = [dp, i64, ctypes.c_int, ctypes.c_int, i64, ip, ip, ip,
                                    dp, dp, dp, ip, ip]
    L.orc_solve_batched.restype = None
    L.orc_num_threads.restype = ctypes.c_int
    L.orc_set_num_threads.argtypes = [ctypes.c_int]
    _lib = L
    return L


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def n_cells(n: int, m: int) -> int:
    return n * (m + 1) + m


def flatten(constraints, function) -> tuple[np.ndarray, int, int]:
    """Reference inputs (simplex.py:25) -> reference-flat fp64 array."""
    rows = np.asarray(constraints, dtype=np.float64)
    if rows.ndim != 2:
        raise ValueError("constraints must be a rectangular list of rows")
    n, m1 = rows.shape
    m = m1 - 1
    c = np.asarray(function, dtype=np.float64).reshape(-1)
    if c.shape[0] != m:
        raise ValueError("function must have m entries")
    return np.concatenate([rows.reshape(-1), c]), n, m


def unflatten(T: np.ndarray, n: int, m: int) -> list[list[float]]:
    """Reference-flat -> the reference's ragged list of lists."""
    body = T[: n * (m + 1)].reshape(n, m + 1).tolist()
    body.append(T[n * (m + 1):].tolist())
    return body


RULES = {"reference": 0, "dantzig": 1}     # "dantzig" is an extension, see spx_oracle.c


def pick(T: np.ndarray, n: int, m: int, rule: str = "reference"):
    """simplex.py:70-141 -> (status, r, c, e)."""
    r, c, e = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
    st = lib().orc_pick_rule(_dp(T), n, m, RULES[rule], ctypes.byref(r), ctypes.byref(c), ctypes.byref(e))
    return st, r.value, c.value, e.value


def update(T: np.ndarray, n: int, m: int, r: int, c: int) -> np.ndarray:
    """simplex.py:149-177 arithmetic -> new reference-flat table."""
    N = np.empty_like(T)
    lib().orc_update(_dp(T), _dp(N), n, m, r, c)
    return N


def init_labels(n: int, m: int):
    return np.arange(m, dtype=np.int32), np.arange(m, m + n, dtype=np.int32)


def label_strings(rowlab, collab, m: int):
    """int codes -> the reference's header strings (simplex.py:30-33)."""
    def s(code):
        code = int(code)
        return f"x{code + 1}" if code < m else f"y{code - m + 1}"
    return [s(v) for v in rowlab] + ["-b"], [s(v) for v in collab] + ["f"]


class Solve(NamedTuple):
    status: int
    npiv: int
    trace: np.ndarray        # [npiv, 2] int32
    table: np.ndarray        # final reference-flat table
    rowlab: np.ndarray
    collab: np.ndarray
    x: np.ndarray            # [m]
    obj2: float              # function[0]*x1 + function[1]*x2 (simplex.py:49)
    objm: float              # full c.x
    snaps: Optional[np.ndarray]


def solve(constraints, function, max_pivots: int = 1_000_000, snapshots: bool = False,
          rule: str = "reference") -> Solve:
    T, n, m = flatten(constraints, function)
    return solve_flat(T, n, m, max_pivots, snapshots, rule=rule)


def solve_flat(T0: np.ndarray, n: int, m: int, max_pivots: int = 1_000_000,
               snapshots: bool = False, keep_trace: bool = True, rule: str = "reference") -> Solve:
    T = np.array(T0, dtype=np.float64, copy=True)
    cells = n_cells(n, m)
    function = T[n * (m + 1):].copy()
    scratch = np.empty(cells, dtype=np.float64)
    trace = np.zeros((max_pivots, 2), dtype=np.int32) if keep_trace else None
    snaps = np.empty((max_pivots + 1, cells), dtype=np.float64) if snapshots else None
    rowlab, collab = init_labels(n, m)
    npiv = ctypes.c_int64(0)
    st = lib().orc_solve_rule(_dp(T), _dp(scratch), n, m, RULES[rule], max_pivots, _ip(trace), _dp(snaps),
                              _ip(rowlab), _ip(collab), ctypes.byref(npiv))
    k = npiv.value
    x = np.zeros(m, dtype=np.float64)
    o2, om = ctypes.c_double(0.0), ctypes.c_double(0.0)
    lib().orc_extract(_dp(T), n, m, _ip(collab), _dp(function), _dp(x),
                      ctypes.byref(o2), ctypes.byref(om))
    return Solve(st, k, trace[:k].copy() if keep_trace else np.zeros((0, 2), np.int32), T,
                 rowlab, collab, x, o2.value, om.value,
                 snaps[: k + 1].copy() if snapshots else None)


class BatchSolve(NamedTuple):
    status: np.ndarray   # [B] int32
    npiv: np.ndarray     # [B] int32
    trace: np.ndarray    # [B, max_pivots, 2] int32
    tables: np.ndarray   # [B, cells]
    rowlab: np.ndarray   # [B, m]
    collab: np.ndarray   # [B, n]
    x: np.ndarray        # [B, m]
    obj2: np.ndarray     # [B]


def solve_batched(tables: np.ndarray, n: int, m: int, max_pivots: int = 64,
                  threads: int = 0) -> BatchSolve:
    """tables: [B, cells] reference-flat LPs of one shape."""
    T = np.array(tables, dtype=np.float64, copy=True, order="C")
    B, cells = T.shape
    assert cells == n_cells(n, m)
    function = T[:, n * (m + 1):].copy()
    trace = np.zeros((B, max_pivots, 2), dtype=np.int32)
    rowlab = np.zeros((B, m), dtype=np.int32)
    collab = np.zeros((B, n), dtype=np.int32)
    x = np.zeros((B, m), dtype=np.float64)
    obj2 = np.zeros(B, dtype=np.float64)
    status = np.zeros(B, dtype=np.int32)
    npiv = np.zeros(B, dtype=np.int32)
    if threads:
        lib().orc_set_num_threads(threads)
    lib().orc_solve_batched(_dp(T), B, n, m, max_pivots, _ip(trace), _ip(rowlab), _ip(collab),
                            _dp(function), _dp(x), _dp(obj2), _ip(status), _ip(npiv))
    return BatchSolve(status, npiv, trace, T, rowlab, collab, x, obj2)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(t: int) -> None:
    lib().orc_set_num_threads(t)
