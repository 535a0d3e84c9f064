"""Helpers shared by the CPU and GPU test modules."""
import hashlib

import numpy as np

END_TO_STATUS = {"optimal": 0, "incorrect system": -1, "simplex method does not converge": -2, "cap": -3}


def unhex(rows):
    return np.asarray([[float.fromhex(v) for v in r] for r in rows], dtype=np.float64)


def unhex1(vals):
    return np.asarray([float.fromhex(v) for v in vals], dtype=np.float64)


def case_inputs(case):
    """rows [n, m+1], c [m] of a golden case (stored as hex floats or as a generator spec)."""
    if "generator" in case:
        from simplex_method_solver_b200 import workloads as W
        g = case["generator"]
        assert g["kind"] == "dense_lp"
        return W.dense_lp(g["n"], g["m"], g["seed"])
    return unhex(case["rows"]), unhex1(case["c"])


def table_sha(flat) -> str:
    return hashlib.sha256(np.ascontiguousarray(flat, dtype="<f8").tobytes()).hexdigest()


def flat_of(rows, c):
    return np.concatenate([np.asarray(rows, dtype=np.float64).reshape(-1), np.asarray(c, dtype=np.float64)])


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def label_codes(names, m):
    """['x1','y2',...] -> int codes used on the device (x_j -> j-1, y_i -> m+i-1)."""
    return [int(s[1:]) - 1 if s[0] == "x" else m + int(s[1:]) - 1 for s in names]


def make_lp(n, m, seed, kind):
    """Test LPs of the sharded flows: rows [n, m+1], c [m].
    dense    : workloads.dense_lp — b > 0, the entering column stays among the FIRST columns (rank 0 owns it)
    late     : dense_lp with a positive objective on the first 95 % of the columns — the entering column starts
               on the LAST column block and then moves between blocks (owner rank != 0, owner changes)
    smallint : small integers — degenerate ties, phase-1 pivots (b < 0), 'incorrect' / 'does not converge' endings
    """
    from simplex_method_solver_b200 import workloads as W
    if kind == "dense":
        return W.dense_lp(n, m, seed)
    if kind == "late":
        rows, c = W.dense_lp(n, m, seed)
        c[: int(0.95 * m)] = np.abs(c[: int(0.95 * m)])
        return rows, c
    assert kind == "smallint", kind
    rng = np.random.default_rng(seed)
    A = rng.integers(-3, 4, (n, m)).astype(float)
    b = rng.integers(-2, 7, n).astype(float)
    c = rng.integers(-3, 4, m).astype(float)
    return np.hstack([A, b[:, None]]), c


def owners_of(trace, blocks):
    """owner rank of every pivot's entering column; blocks = [(col0, m_loc), ...]"""
    return [[k for k, (a, w) in enumerate(blocks) if a <= int(cc) < a + w][0] for _, cc in trace]
