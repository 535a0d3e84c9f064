"""Real multi-GPU test of the column-sharded flow (needs >= 2 CUDA devices; skipped otherwise):
one process per GPU, the C-side look-ahead loop with NVLink peer mailboxes (PeerShardedTableau) and
the NCCL all-gather flow (ShardedTableau), each against the single-process oracle, bit for bit.
Run with:  gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu -q
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_lp(n, m, seed, kind):
    from util import make_lp
    return make_lp(n, m, seed, kind)


def _worker(rank, world, port, n, m, seed, cap, mode, kind, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=120))
    try:
        from simplex_method_solver_b200 import parallel as P
        rows, c = _make_lp(n, m, seed, kind)
        if mode == "simplexmethod":
            # the reference surface: every rank constructs the same SimplexMethod and calls solve()
            from simplex_method_solver_b200.simplex import SimplexMethod
            sm = SimplexMethod(rows, c, engine="sharded", max_pivots=cap)
            sol = sm.solve()
            sh = sm._sh
            out.put((rank, {"status": int(sol.status), "npiv": int(sol.npiv), "col0": sh.col0,
                            "body": sh.local_body().cpu().numpy().copy(), "b": sh.b_current().cpu().numpy().copy(),
                            "trace": sol.trace.copy(), "rowlab": sol.rowlab.copy(), "collab": sol.collab.copy(),
                            "x": sol.x.copy(), "objective": float(sol.objective), "labels": (list(sm.row), list(sm.column))}))
            dist.barrier()
            sm.close()
            return
        if mode in ("fused", "fused-persistent"):
            sh = P.FusedShardedTableau(n, m, rank, world, dev, trace_capacity=cap + 8, depth=5,
                                       lookahead="persistent" if mode == "fused-persistent" else "per-pass")
        elif mode == "p2p":
            sh = P.PeerShardedTableau(n, m, rank, world, dev, trace_capacity=cap + 8)
        else:
            sh = P.ShardedTableau(n, m, rank, world, dev, trace_capacity=cap + 8, lookahead=(mode == "nccl-ahead"))
        sh.load(rows, c, max_pivots=cap)
        status, npiv = sh.solve(cap, check_every=16)
        st = sh.sync()
        body = sh.local_body().cpu().numpy().copy() if mode not in ("p2p",) else \
            sh.A[sh._cur, :, : sh.m_loc].cpu().numpy().copy()
        b = (sh.b[sh._cur, :n] if mode == "p2p" else sh.b_current()).cpu().numpy().copy()
        out.put((rank, {"status": int(st.status), "npiv": int(st.npiv), "col0": sh.col0, "body": body, "b": b,
                        "trace": sh.trace[: int(st.npiv)].cpu().numpy().copy(),
                        "rowlab": sh.rowlab.cpu().numpy().copy(), "collab": sh.collab[:n].cpu().numpy().copy()}))
        dist.barrier()
        if mode in ("p2p", "fused", "fused-persistent"):
            sh.close()
    finally:
        dist.destroy_process_group()


# The "smallint" (degenerate ties, phase-1 pivots) and "late" (entering columns on the LAST ranks, ~30-60 owner
# changes) cases exercise the winner != rank 0 branch of every exchange.  First hardware run: round 2, 2 x B200,
# all four exchange modes green (profiles/r2/r2a_multigpu_extended_2gpu_summary.txt); they are no longer gated.


@pytest.mark.parametrize("mode", ["fused", "fused-persistent", "p2p", "nccl", "nccl-ahead", "simplexmethod"])
@pytest.mark.parametrize("n,m,cap,kind", [(300, 2600, 150, "dense"), (64, 1024, 400, "dense"),
                                          (9, 40, 60, "dense"),            # ranks >= 1 own no columns
                                          (12, 1300, 60, "smallint"), (20, 1100, 80, "smallint"),
                                          (24, 1100, 90, "late"), (40, 2100, 120, "late"),
                                          (64, 1600, 100, "late")])         # owners 19/3/41/37 at world 4
def test_sharded_flow_on_real_gpus(mode, n, m, cap, kind):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import oracle
    rows, c = _make_lp(n, m, 7, kind)
    o = oracle.solve(rows, c, max_pivots=cap)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, m, 7, cap, mode, kind, out), daemon=True)
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = dict(out.get(timeout=150) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:                      # a failed rank must not leave its peers spinning in a collective
            if p.is_alive():
                p.kill()
    body = np.zeros((n + 1, m))
    for r in range(world):
        g = got[r]
        assert (g["status"], g["npiv"]) == (o.status, o.npiv), (mode, r, g["status"], g["npiv"], o.status, o.npiv)
        assert g["trace"].tolist() == o.trace.tolist()
        assert g["rowlab"].tolist() == o.rowlab.tolist() and g["collab"].tolist() == o.collab.tolist()
        if mode == "simplexmethod":
            assert g["x"].tobytes() == o.x.tobytes() and float(g["objective"]).hex() == float(o.objm).hex()
        body[:, g["col0"]: g["col0"] + g["body"].shape[1]] = g["body"]
        assert np.array_equal(g["b"].view(np.uint64), o.table[: n * (m + 1)].reshape(n, m + 1)[:, m].copy().view(np.uint64))
    ob = np.zeros((n + 1, m))
    ob[:n] = o.table[: n * (m + 1)].reshape(n, m + 1)[:, :m]
    ob[n] = o.table[n * (m + 1):]
    assert np.array_equal(body.view(np.uint64), ob.view(np.uint64))
