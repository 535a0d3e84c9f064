"""CPU tests of the multi-GPU host logic (simplex_method_solver_b200/parallel.py) under gloo,
world_size 2 and 3: shard ranges, the all-gather message layout, ping-pong parity and
termination.  The three per-pivot kernels are replaced by the numpy stand-in in
tests/cpu_shard_ops.py; the collective path is the product's own.  The sharded pivot trace must
equal the single-process oracle trace exactly.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from simplex_method_solver_b200 import parallel as P  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [P.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def test_column_block_alignment():
    for m in (1, 2, 511, 512, 513, 2000, 32768, 40000):
        for world in (1, 2, 4, 8):
            blocks = [P.column_block(m, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == m
            for (c0, k0), (c1, _) in zip(blocks, blocks[1:]):
                assert c0 + k0 == c1 or k0 == 0 or c1 == m
            for c0, k in blocks:
                assert c0 % P.TILE_COLS == 0 or k == 0
    assert P.column_block(32768, 3, 8) == (3 * 4096, 4096)


def _sharded_worker(rank, world, port, n, m, seed, cap, kind, lookahead, out, local=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_shard_ops import CpuShardOps
        rows, c = _make_lp(n, m, seed, kind)
        sh = P.ShardedTableau(n, m, rank, world, device="cpu", trace_capacity=cap, ops=CpuShardOps(),
                              lookahead=lookahead)
        if local:
            # every rank packs ITS OWN block [n, m_loc + 1] (its columns, then b): the e2e path of bench.py at N > 1;
            # a rank without columns (m_loc == 0) packs the b column alone
            blk = np.ascontiguousarray(np.hstack([rows[:, sh.col0: sh.col0 + sh.m_loc], rows[:, m: m + 1]]))
            sh.load_local(blk, np.ascontiguousarray(c[sh.col0: sh.col0 + sh.m_loc]), max_pivots=cap)
        else:
            sh.load(rows, c, max_pivots=cap)
        status, npiv = sh.solve(cap, check_every=5)
        body = sh.local_body().numpy().copy()
        res = {"status": status, "npiv": npiv, "trace": sh.trace[:npiv].numpy().copy(),
               "b": sh.b_current().numpy().copy(), "col0": sh.col0, "body": body,
               "rowlab": sh.rowlab.numpy().copy(), "collab": sh.collab.numpy()[:n].copy()}
        out.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _make_lp(n, m, seed, kind):
    from util import make_lp
    return make_lp(n, m, seed, kind)


def _run_world(world, n, m, seed, cap, kind, lookahead, local=False):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n, m, seed, cap, kind, lookahead, out, local))
             for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("lookahead", [False, True])
@pytest.mark.parametrize("world,n,m,kind,seed", [
    (2, 24, 1100, "dense", 3),       # 3 column tiles over 2 ranks: uneven blocks
    (3, 12, 1300, "smallint", 5),    # degenerate ties, phase-1 pivots, error endings
    (2, 9, 40, "smallint", 8),       # fewer tiles than ranks: rank 1 owns no columns
    (3, 64, 1600, "late", 7),        # entering columns on ranks 1 and 2, ~30 owner changes in 60 pivots
])
def test_sharded_trace_equals_single_process_oracle(world, n, m, kind, seed, lookahead):
    _check_world(world, n, m, kind, seed, lookahead, local=False)


@pytest.mark.parametrize("world,n,m,kind,seed", [
    (2, 9, 40, "smallint", 8),       # fewer tiles than ranks: rank 1 packs a block WITHOUT columns (b alone)
    (3, 12, 600, "dense", 5),        # 2 tiles over 3 ranks: rank 2 owns nothing
    (2, 24, 1100, "late", 3),
])
def test_sharded_load_local_packed_blocks(world, n, m, kind, seed):
    _check_world(world, n, m, kind, seed, lookahead=True, local=True)


def _check_world(world, n, m, kind, seed, lookahead, local):
    import oracle
    cap = 60
    rows, c = _make_lp(n, m, seed, kind)
    o = oracle.solve(rows, c, max_pivots=cap)
    got = _run_world(world, n, m, seed, cap, kind, lookahead, local)
    body = np.zeros((n + 1, m))
    for r in range(world):
        g = got[r]
        assert g["status"] == o.status and g["npiv"] == o.npiv, (r, g["status"], o.status)
        assert g["trace"].tolist() == o.trace.tolist()
        assert g["rowlab"].tolist() == o.rowlab.tolist() and g["collab"].tolist() == o.collab.tolist()
        k = g["body"].shape[1]
        body[:, g["col0"]:g["col0"] + k] = g["body"]
        # b is replicated: bit-identical on every rank and equal to the oracle's
        assert np.array_equal(g["b"].view(np.uint64),
                              o.table[: n * (m + 1)].reshape(n, m + 1)[:, m].copy().view(np.uint64))
    ob = np.zeros((n + 1, m))
    ob[:n] = o.table[: n * (m + 1)].reshape(n, m + 1)[:, :m]
    ob[n] = o.table[n * (m + 1):]
    assert np.array_equal(body.view(np.uint64), ob.view(np.uint64))


def _batched_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from simplex_method_solver_b200 import workloads as W
        T, C = W.gui_batch(1001, seed=4)
        tabs = W.batch_flat(T, C)

        def cpu_solver(tables, n, m, max_pivots, rule, device):
            return oracle.solve_batched(tables, n, m, max_pivots=max_pivots)
        start, count, res = P.solve_batched_sharded(tabs, 8, 2, max_pivots=64, solver=cpu_solver)
        # gather the per-rank pivot counts the way bench.py aggregates them
        tot = torch.tensor([int(res.npiv.sum()), count], dtype=torch.int64)
        dist.all_reduce(tot)
        out.put((rank, start, count, res.npiv.copy(), tot.tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_batched_split_has_no_gaps_or_overlap():
    import oracle
    from simplex_method_solver_b200 import workloads as W
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batched_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    T, C = W.gui_batch(1001, seed=4)
    full = oracle.solve_batched(W.batch_flat(T, C), 8, 2, max_pivots=64)
    npiv = np.concatenate([g[3] for g in got])
    assert got[0][1] == 0 and got[0][2] + got[1][2] == 1001 and got[1][1] == got[0][2]
    assert np.array_equal(npiv, full.npiv)
    assert got[0][4] == [int(full.npiv.sum()), 1001]
