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
