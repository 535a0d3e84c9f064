#!/usr/bin/env python
"""Long oracle prefix of cfg4 (16384 x 32768) — covers every TIMED pivot of bench.py.

    python tests/golden/make_cfg4_long.py            # K = 26000 pivots, marks every 1000 (hours on 8 cores)

Provenance "oracle": the pure-Python reference cannot hold this shape; the C restatement
(oracle/spx_oracle.c) is checked bit-for-bit against the live reference and every reference fixture by
tests/test_oracle.py, and the first 2000 pivots written here must equal tests/golden/cfg_digests.json["cfg4"].

Writes
  tests/golden/cfg4_long_trace.npy   int32 [K][2]  (r, c) per pivot
  tests/golden/cfg4_long.json        marks: cumulative pivot sha256, b / f sha256, body checksum
                                     (+ sha256 of the body at a few marks)

Body checksum (cheap on both sides, order-free, sensitive to every bit):
    sum over body cells (i < n, j < m) of  bits(T[i][j]) * (2*(i*m + j) + 1)   mod 2^64
(simplex_method_solver_b200.workloads.body_checksum_* compute the same thing).

The run checkpoints the table to $CFG4_CKPT (default /tmp/cfg4_ckpt) every 2000 pivots and resumes from it.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from simplex_method_solver_b200 import workloads as W  # noqa: E402

N_ROWS, N_COLS = 16384, 32768
K = int(os.environ.get("CFG4_PIVOTS", "26000"))
STEP = int(os.environ.get("CFG4_MARK_EVERY", "1000"))
SHA_BODY_AT = {2000, 6000, 25000, 26000}
CKPT = os.environ.get("CFG4_CKPT", "/tmp/cfg4_ckpt")
OUT_JSON = os.path.join(HERE, "cfg4_long.json")
OUT_TRACE = os.path.join(HERE, "cfg4_long_trace.npy")


def hx(v):
    return float(v).hex()


def mark_of(T, trace, k, n, m):
    body = T[: n * (m + 1)].reshape(n, m + 1)
    out = {
        "pivot_sha256": W.pivot_digest(trace[:k]),
        "b_sha256": hashlib.sha256(np.ascontiguousarray(body[:, m]).tobytes()).hexdigest(),
        "f_sha256": hashlib.sha256(T[n * (m + 1):].tobytes()).hexdigest(),
        "b_first4": [hx(v) for v in body[:4, m]],
        "f_first4": [hx(v) for v in T[n * (m + 1): n * (m + 1) + 4]],
        "body_checksum_u64": int(W.body_checksum_numpy(body[:, :m])),
    }
    if k in SHA_BODY_AT:
        h = hashlib.sha256()
        for i in range(0, n, 256):
            h.update(np.ascontiguousarray(body[i:i + 256, :m]).tobytes())
        out["body_sha256"] = h.hexdigest()
    return out


def main():
    n, m = N_ROWS, N_COLS
    cells = n * (m + 1) + m
    marks = {}
    trace = np.zeros((K, 2), np.int32)
    k0 = 0
    meta = {"generator": {"kind": "dense_lp", "n": n, "m": m, "seed": 0}, "provenance": "oracle",
            "made_by": "tests/golden/make_cfg4_long.py", "threads": oracle.num_threads()}
    if os.path.exists(CKPT + ".json") and os.path.exists(CKPT + ".f8"):
        st = json.load(open(CKPT + ".json"))
        k0 = st["k"]
        marks = st["marks"]
        meta["input_sha256"] = st["input_sha256"]
        trace[:k0] = np.asarray(st["trace"], np.int32).reshape(-1, 2)
        T = np.fromfile(CKPT + ".f8", dtype=np.float64)
        assert T.shape[0] == cells
        print("resumed at pivot", k0, flush=True)
    else:
        rows, c = W.dense_lp(n, m, 0)
        meta["input_sha256"] = W.input_digest(rows, c)
        T = np.concatenate([rows.reshape(-1), c])
        del rows
    N = np.empty_like(T)
    t0 = time.time()
    for k in range(k0, K):
        st, r, cc, _ = oracle.pick(T, n, m)
        assert st == oracle.PIVOT, (k, st)
        trace[k] = (r, cc)
        oracle.lib().orc_update(oracle._dp(T), oracle._dp(N), n, m, r, cc)
        T, N = N, T
        kk = k + 1
        if kk % STEP == 0 or kk in (16, 50, 100, 200, 400, 800, 1600) or kk == K:
            marks[str(kk)] = mark_of(T, trace, kk, n, m)
            print("cfg4 pivot", kk, round(time.time() - t0, 1), "s", marks[str(kk)]["pivot_sha256"][:16],
                  flush=True)
            np.save(OUT_TRACE, trace[:kk])
            with open(OUT_JSON, "w") as fh:
                json.dump(dict(meta, npiv=kk, marks=marks), fh, indent=1, sort_keys=True)
        if kk % 2000 == 0 and kk < K:
            T.tofile(CKPT + ".f8.tmp")
            os.replace(CKPT + ".f8.tmp", CKPT + ".f8")
            with open(CKPT + ".json", "w") as fh:
                json.dump({"k": kk, "marks": marks, "input_sha256": meta["input_sha256"],
                           "trace": trace[:kk].reshape(-1).tolist()}, fh)


if __name__ == "__main__":
    main()
