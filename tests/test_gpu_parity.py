"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the
golden fixtures generated from the live reference.  Bit-exact everywhere: pivot
sequences, labels and every fp64 cell; no tolerance is needed or used.
"""
import hashlib

import numpy as np
import pytest

import oracle
from simplex_method_solver_b200 import workloads as W
from util import END_TO_STATUS, bits, case_inputs, flat_of, label_codes, table_sha, unhex, unhex1

pytestmark = pytest.mark.gpu


class loop_mode:
    """'fused-engine' = the same with the persistent pricing engine; 'fused-coop' = the fused loop of spx_solve with look-ahead pricing (side stream, replay through the
    previous pass's pending levels) switched on; every other name passes through."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        from simplex_method_solver_b200 import _native as N
        if self.name == "fused-coop":                  # the fused loop of spx_solve with look-ahead pricing on
            assert N.lib().spx_set_option(9, 1) == 0
            return "fused"
        if self.name == "fused-engine":                # ... with the persistent pricing engine (one kernel per call)
            assert N.lib().spx_set_option(9, 2) == 0
            return "fused"
        if self.name == "fused-gpuwide":               # ... with the whole-GPU cooperative pricing kernel forced on
            assert N.lib().spx_set_option(8, 2) == 0
            return "fused"
        if self.name == "resident-ahead":              # the L2-resident kernel with look-ahead pricing inside the CTA
            assert N.lib().spx_set_option(15, 1) == 0
            return "resident"
        if isinstance(self.name, str) and self.name.startswith("fused-x"):   # "fused-x<variant>-<rows>[-<minb>[-<pairs>]]":
            parts = self.name[len("fused-x"):].split("-")                    # update kernel schedule
            assert N.lib().spx_set_option(10, int(parts[0])) == 0
            assert N.lib().spx_set_option(11, int(parts[1])) == 0
            assert N.lib().spx_set_option(7, int(parts[2]) if len(parts) > 2 else 0) == 0
            assert N.lib().spx_set_option(12, int(parts[3]) if len(parts) > 3 else 0) == 0
            return "fused"
        return self.name

    def __exit__(self, *a):
        from simplex_method_solver_b200 import _native as N
        if self.name in ("fused-coop", "fused-engine"):
            N.lib().spx_set_option(9, 0)
        if self.name == "fused-gpuwide":
            N.lib().spx_set_option(8, 0)
        if self.name == "resident-ahead":
            N.lib().spx_set_option(15, 0)
        if isinstance(self.name, str) and self.name.startswith("fused-x"):
            for k in (10, 11, 7, 12):
                N.lib().spx_set_option(k, 0)


@pytest.fixture(scope="module")
def spx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a CUDA device")
    from simplex_method_solver_b200 import _native, batched, engine, simplex
    _native.lib()

    class NS:
        pass
    ns = NS()
    ns.N, ns.batched, ns.engine, ns.simplex, ns.torch = _native, batched, engine, simplex, torch
    return ns


# --------------------------------------------------------------------------- get_solution()
def _check_infos(spx, case, engine):
    rows, c = case_inputs(case)
    sm = spx.simplex.SimplexMethod(rows.tolist(), c.tolist(), engine=engine)
    got = sm.get_solution()
    exp = case["get_solution"]
    assert len(got) == len(exp), case["name"]
    for g, e in zip(got, exp):
        if "error" in e:
            assert isinstance(g, spx.simplex.Error) and str(g) == e["error"], case["name"]
            continue
        assert isinstance(g, spx.simplex.Info)
        assert g.row == e["row"] and g.column == e["column"], case["name"]
        assert g.i == e["i"] and g.j == e["j"], case["name"]
        assert len(g.table) == len(e["table"])
        for gr, er in zip(g.table, e["table"]):
            assert [float(v).hex() for v in gr] == er, case["name"]
        assert float(g.x1).hex() == e["x1"] and float(g.x2).hex() == e["x2"], case["name"]
        assert float(g.optimum).hex() == e["optimum"], case["name"]
    # object state after the call matches the reference's (final labels / table)
    assert sm.row == case["row_labels"] and sm.column == case["column_labels"]
    assert [[float(v).hex() for v in r] for r in sm.table] == case["final_table"]


@pytest.mark.parametrize("engine", ["warp", "stream"])
def test_get_solution_matches_reference_snapshots(spx, ref_cases, engine):
    cases = [c for c in ref_cases if "get_solution" in c]
    assert len(cases) >= 15
    for case in cases:
        _check_infos(spx, case, engine)


def test_cfg1_snapshot_digest(spx, ref_cases):
    case = next(c for c in ref_cases if c["name"] == "ref_example_cfg1_205")
    sm = spx.simplex.SimplexMethod(W.CFG1_ROWS, W.CFG1_C)
    res = sm.get_solution()
    assert [(i.i, i.j) for i in res[:-1]] == [(2, 0), (3, 0), (1, 1), (0, 0)]
    assert W.snapshot_digest([i.table for i in res]) == case["snapshot_sha256"]
    assert case["snapshot_sha256"] == "0599fb0b6881268191a0d0490575cbe5891d0e5f5f7148bd4124e07a0a853183"
    assert (res[-1].x1, res[-1].x2) == (39.18192919380969, 26.096639697976617)
    assert res[-1].optimum == -65.2785688917863


def test_gui_adapter_compute_solution(spx, ref_cases):
    """N1: the headless compute_solution call site returns the reference's Info list and table views."""
    from simplex_method_solver_b200 import gui_adapter as G
    case = next(c for c in ref_cases if c["name"] == "ref_example_cfg1_205")
    sol = G.compute_solution(lines=W.CFG1_ROWS, grad=W.CFG1_C + [0])
    assert not sol.failed and len(sol.tables) == len(case["get_solution"]) == 5
    view = G.table_view(sol.tables[0])
    assert view["pivot"] == (2, 0) and view["optimum"] == 0 and view["f"] == [-1.0, -1.0]
    assert view["body"][0] == [-39.7, -96.0, 4060.8]
    last = G.table_view(sol.tables[-1])
    assert last["pivot"] is None and last["optimum"] == -65.28
    assert last["point"] == (39.18192919380969, 26.096639697976617)
    bad = G.compute_solution(lines=[[1.0, 1.0, -2.0], [-1.0, -1.0, 1.0]], grad=[1.0, 1.0, 0])
    assert bad.failed and str(bad.tables[-1]) == "incorrect system"


def test_problem_files_feed_the_batched_solver(spx):
    """N3: problems in the GUI's text format -> one batched solve == per-problem oracle solves."""
    from simplex_method_solver_b200 import problem_io as PIO
    T, C = W.gui_batch(64, seed=9)
    probs = [PIO.loads(PIO.dumps(T[k].tolist(), C[k].tolist() + [0], 10)) for k in range(64)]
    tabs, n, m = PIO.batch_tables(probs)
    res = spx.batched.solve_batched(tabs, n, m, max_pivots=64)
    orc = oracle.solve_batched(tabs, n, m, max_pivots=64)
    assert res.trace.tobytes() == orc.trace.tobytes() and res.x.tobytes() == orc.x.tobytes()


# --------------------------------------------------------------------------- all golden cases
@pytest.mark.parametrize("lookahead,chunk", [(False, 7), (True, 7), (True, 4), (True, 1), ("resident", 7), (None, 5),
                                             ("fused", 7), ("fused", 3), ("fused-coop", 7), ("fused-gpuwide", 5),
                                             ("fused-engine", 7), ("fused-engine", 40), ("resident-ahead", 7)])
def test_all_reference_cases_streaming_solver(spx, ref_cases, lookahead, chunk):
    """solve(): device-side loop, classic (pick k, update k, ...) and look-ahead (pivot k+1 priced
    from table k on a side stream while update k runs); trace, ending, labels, final table bits."""
    for case in ref_cases:
        rows, c = case_inputs(case)
        n, m = rows.shape[0], rows.shape[1] - 1
        sm = spx.simplex.SimplexMethod(rows, c, engine="stream")
        with loop_mode(lookahead) as mode_:
            sol = sm.solve(max_pivots=case["cap"], chunk=chunk, lookahead=mode_)
        assert sol.status == END_TO_STATUS[case["end"]], case["name"]
        assert sol.trace.tolist() == case["trace"], case["name"]
        flat = sm._dev.export_flat(sm._npiv)
        assert table_sha(flat) == case["final_table_sha256"], case["name"]
        assert sm.row == case["row_labels"] and sm.column == case["column_labels"], case["name"]
        if "x1" in case:
            x1, x2 = sm.find_optimum()
            assert float(x1).hex() == case["x1"] and float(x2).hex() == case["x2"], case["name"]
            assert float(sm.f(x1, x2)).hex() == case["f"], case["name"]
            assert float(sol.obj2).hex() == case["f"], case["name"]


def test_all_reference_cases_batched_solver(spx, ref_cases):
    """solve_batched(): same cases grouped by shape, one warp per LP."""
    groups = {}
    for case in ref_cases:
        rows, c = case_inputs(case)
        key = (rows.shape[0], rows.shape[1] - 1, case["cap"])
        groups.setdefault(key, []).append((case, flat_of(rows, c)))
    for (n, m, cap), items in groups.items():
        tables = np.stack([f for _, f in items])
        res = spx.batched.solve_batched(tables, n, m, max_pivots=cap)
        for k, (case, _) in enumerate(items):
            assert res.status[k] == END_TO_STATUS[case["end"]], case["name"]
            assert res.npiv[k] == len(case["trace"]), case["name"]
            assert res.trace[k, : res.npiv[k]].tolist() == case["trace"], case["name"]
            assert table_sha(res.tables[k]) == case["final_table_sha256"], case["name"]
            assert res.rowlab[k].tolist() == label_codes(case["row_labels"][:-1], m), case["name"]
            assert res.collab[k].tolist() == label_codes(case["column_labels"][:-1], m), case["name"]
            if "x1" in case:
                assert float(res.x[k, 0]).hex() == float(float.fromhex(case["x1"])).hex()
                assert float(res.x[k, 1]).hex() == float(float.fromhex(case["x2"])).hex()
                assert float(res.obj[k]).hex() == case["f"], case["name"]


# --------------------------------------------------------------------------- step API
def test_step_api_matches_oracle(spx):
    rng = np.random.default_rng(7)
    for t in range(40):
        n, m = int(rng.integers(1, 12)), int(rng.integers(2, 9))
        A = rng.integers(-3, 4, (n, m)).astype(float)
        b = rng.integers(-2, 7, n).astype(float)
        c = rng.integers(-3, 4, m).astype(float)
        rows = np.hstack([A, b[:, None]])
        sm = spx.simplex.SimplexMethod(rows.tolist(), c.tolist())
        T = flat_of(rows, c)
        for step in range(30):
            st, r, cc, e = oracle.pick(T, n, m)
            if st == oracle.PIVOT:
                assert sm.pick_element() == (True, r, cc, e)
                sm.recalculate_matrix()
                T = oracle.update(T, n, m, r, cc)
                got = np.asarray([v for row in sm.table for v in row])
                assert np.array_equal(bits(got), bits(T))
            elif st == oracle.OPTIMAL:
                ok, x1, x2, f = sm.pick_element()
                assert ok is False
                assert sm.recalculate_matrix() is None
                break
            else:
                with pytest.raises(ValueError, match=oracle.STATUS_NAME[st]):
                    sm.pick_element()
                with pytest.raises(ValueError, match=oracle.STATUS_NAME[st]):
                    sm.recalculate_matrix()
                break


@pytest.mark.parametrize("mode", [True, "resident", "fused"])
def test_lookahead_state_feeds_the_step_api_and_resumes(spx, mode):
    """A look-ahead / resident solve stopped by the cap leaves a state the step API and a resumed solve continue from."""
    rows, c = W.dense_lp(40, 70, 9)
    ref = oracle.solve(rows, c, max_pivots=10000)
    assert ref.status == oracle.OPTIMAL and ref.npiv > 25
    sm = spx.simplex.SimplexMethod(rows, c, engine="stream")
    sol = sm.solve(max_pivots=11, chunk=4, lookahead=mode)
    assert sol.status == spx.N.CAP and sol.trace.tolist() == ref.trace[:11].tolist()
    ok, r, cc, e = sm.pick_element()
    assert ok and [r, cc] == ref.trace[11].tolist()
    sm.recalculate_matrix()
    sol = sm.solve(max_pivots=10000, chunk=5, lookahead=mode)
    assert sol.status == spx.N.OPTIMAL and sol.npiv == ref.npiv
    # pivot 11 went through the step API after the traced capacity of the first solve: not traced
    assert sol.trace[:11].tolist() == ref.trace[:11].tolist()
    assert sol.trace[12:].tolist() == ref.trace[12:].tolist()
    assert np.array_equal(bits(sm._dev.export_flat(sm._npiv)), bits(ref.table))
    assert sol.x.tobytes() == ref.x.tobytes()


@pytest.mark.parametrize("lookahead", [False, True, "resident", "fused", "fused-coop", "fused-gpuwide", "fused-engine", "warp"])
def test_dantzig_rule_every_loop_against_the_restated_oracle(spx, dantzig_cases, lookahead):
    """rule='dantzig' (extension: most negative f cell, lowest index on ties) in every CUDA loop —
    classic, look-ahead, L2-resident, fused (one-CTA, cooperative, whole-GPU pricing), warp-resident batched —
    against fixtures the reference itself produced under that rule (tests/golden/make_golden.py::drive_dantzig) and,
    for a mid-size LP, against the oracle's restatement (oracle.solve(rule="dantzig"))."""
    if lookahead == "warp":
        for case in dantzig_cases:
            rows, c = case_inputs(case)
            n, m = rows.shape[0], rows.shape[1] - 1
            if n * (m + 1) + m > spx.N.load().spx_batched_max_cells():
                continue                                  # larger than a CTA-resident LP: the streaming loops cover it
            res = spx.batched.solve_batched(flat_of(rows, c)[None, :], n, m, max_pivots=case["cap"], rule="dantzig")
            assert int(res.status[0]) == END_TO_STATUS[case["end"]], case["name"]
            assert res.trace[0, : int(res.npiv[0])].tolist() == case["trace"], case["name"]
            assert table_sha(res.tables[0]) == case["final_table_sha256"], case["name"]
        return
    with loop_mode(lookahead) as mode_:
        for case in dantzig_cases:
            rows, c = case_inputs(case)
            sm = spx.simplex.SimplexMethod(rows, c, engine="stream", rule="dantzig")
            sol = sm.solve(max_pivots=case["cap"], chunk=7, lookahead=mode_)
            assert sol.status == END_TO_STATUS[case["end"]], case["name"]
            assert sol.trace.tolist() == case["trace"], case["name"]
            assert table_sha(sm._dev.export_flat(sm._npiv)) == case["final_table_sha256"], case["name"]
            assert sm.row == case["row_labels"] and sm.column == case["column_labels"], case["name"]
        rows, c = W.dense_lp(130, 1030, 9)
        o = oracle.solve(rows, c, max_pivots=300, rule="dantzig")
        sm = spx.simplex.SimplexMethod(rows, c, engine="stream", rule="dantzig")
        sol = sm.solve(max_pivots=300, chunk=11, lookahead=mode_)
        assert (sol.status, sol.npiv) == (o.status, o.npiv)
        assert sol.trace.tolist() == o.trace.tolist()
        assert np.array_equal(bits(sm._dev.export_flat(sm._npiv)), bits(o.table))


def test_print_table_matches_the_reference_format(spx, capsys):
    """print_table() (simplex.py:41-46): a tab-separated header of the row labels, then one line per table row led by
    its column label, cells rounded to 6 places — before and after a pivot."""
    rows, c = [[-1.0, -1.0, 10.0], [1.0, -2.0, 4.0], [0.5, 0.25, -1.0 / 3.0]], [-1.0, -5.0]
    sm = spx.simplex.SimplexMethod(rows, c)

    def expected(table, row, column):
        out = "\t" + "\t".join(row) + "\n"
        for k in range(len(table)):
            out += column[k] + "\t" + "\t".join(str(round(v, 6)) for v in table[k]) + "\n"
        return out

    sm.print_table()
    assert capsys.readouterr().out == expected(rows + [c], ['x1', 'x2', '-b'], ['y1', 'y2', 'y3', 'f'])
    sm.recalculate_matrix()
    o = oracle.solve(rows, c, max_pivots=1)
    tab = oracle.unflatten(o.table, 3, 2)
    rl, cl = oracle.label_strings(o.rowlab, o.collab, 2)
    sm.print_table()
    assert capsys.readouterr().out == expected(tab, rl, cl)


def test_inputs_not_mutated_and_attributes(spx):
    rows = [[-1.0, -1.0, 10.0], [1.0, -2.0, 4.0]]
    c = [-1.0, -5.0]
    keep = ([list(r) for r in rows], list(c))
    sm = spx.simplex.SimplexMethod(rows, c)
    assert (sm.n, sm.m, sm.invalid_index) == (2, 2, 3)
    assert sm.function is c
    assert sm.row == ["x1", "x2", "-b"] and sm.column == ["y1", "y2", "f"]
    assert len(sm.table) == 3 and len(sm.table[-1]) == 2 and sm.table[0] == rows[0]
    sm.get_solution()
    assert (rows, c) == keep


def test_cycling_input_hits_cap(spx):
    sm = spx.simplex.SimplexMethod([[1, 0, 3], [2, 0, 0]], [-1, 0], max_pivots=25)
    res = sm.get_solution()
    assert isinstance(res[-1], spx.simplex.Error) and str(res[-1]) == "pivot limit reached"
    assert [(i.i, i.j) for i in res[:-2]] == [(1, 0)] * 25


# --------------------------------------------------------------------------- unit: K1/K2/K3 on ragged shapes
@pytest.mark.parametrize("n,m", [(1, 2), (3, 1), (7, 15), (8, 16), (9, 17), (63, 511), (64, 512),
                                 (65, 513), (130, 1030), (257, 100), (40, 2049)])
def test_pick_update_bit_exact_ragged_shapes(spx, n, m):
    """One pick + one update per step vs the oracle, whole table compared bit for bit."""
    rng = np.random.default_rng(n * 1000 + m)
    rows, c = W.dense_lp(n, m, seed=n + m)
    # sprinkle exact zeros and sign changes so every branch of the ratio scan is reachable
    rows[rng.random(rows.shape) < 0.05] = 0.0
    if n > 2:
        rows[rng.integers(0, n), m] = 0.0
    dev = spx.engine.DeviceTableau(n, m, trace_capacity=64)
    dev.load(rows, c, max_pivots=64)
    T = flat_of(rows, c)
    npiv = 0
    for step in range(12):
        st, r, cc, e = oracle.pick(T, n, m)
        dev.pick(npiv)
        s = dev.read_state()
        assert s.status == st
        if st != oracle.PIVOT:
            break
        assert (s.r, s.c, s.p) == (r, cc, e)
        dev.update(npiv)
        npiv += 1
        T = oracle.update(T, n, m, r, cc)
        got = dev.export_flat(npiv)
        assert np.array_equal(bits(got), bits(T)), f"step {step}"
        assert dev.read_state().npiv == npiv


@pytest.mark.parametrize("mode", ["resident", "resident-ahead", "fused", "fused-coop", "fused-gpuwide", "fused-engine"])
@pytest.mark.parametrize("n,m", [(1, 2), (3, 1), (7, 15), (9, 17), (64, 512), (65, 513), (130, 1030), (257, 100), (40, 2049)])
def test_resident_and_fused_loops_bit_exact_ragged_shapes(spx, n, m, mode):
    with loop_mode(mode) as mode_:
        _ragged_loop(spx, n, m, mode_)


def _ragged_loop(spx, n, m, mode):
    """The persistent L2-resident loop and the F-pivots-per-pass fused loop in steps of 1..9 pivots vs the
    oracle, whole table each time (odd step counts leave the table in either ping-pong buffer)."""
    rng = np.random.default_rng(n * 31 + m)
    rows, c = W.dense_lp(n, m, seed=n + 2 * m)
    rows[rng.random(rows.shape) < 0.05] = 0.0
    dev = spx.engine.DeviceTableau(n, m, trace_capacity=64)
    dev.load(rows, c, max_pivots=64)
    T = flat_of(rows, c)
    npiv = 0
    for k in (1, 2, 3, 1, 9, 5):
        want = []
        for _ in range(k):
            st, r, cc, e = oracle.pick(T, n, m)
            if st != oracle.PIVOT:
                break
            want.append([r, cc])
            T = oracle.update(T, n, m, r, cc)
        status, got = dev.solve(stop_after=k, lookahead=mode)
        assert got == npiv + len(want)
        assert dev.trace[npiv:got].cpu().numpy().tolist() == want
        npiv = got
        assert np.array_equal(bits(dev.export_flat(npiv)), bits(T)), (n, m, npiv)
        st, r, cc, e = oracle.pick(T, n, m)
        if mode == "fused" and status == oracle.PIVOT and st != oracle.PIVOT and len(want) == k:
            continue                                     # fused stops unpriced: the next call reports the ending
        assert status == st
        if st != oracle.PIVOT:
            break
        if mode == "resident":
            s_ = dev.read_state()
            assert (s_.r, s_.c, s_.p) == (r, cc, e)      # priced, not yet applied


def test_ratio_scan_special_values(spx):
    """inf / NaN / signed zero in the ratio scan: same leaving row as the sequential reference scan."""
    inf, nan = float("inf"), float("nan")
    cases = [
        [[-1, 0, 5], [-2, 0, inf], [-1, 0, 5]],
        [[-1, 0, nan], [-2, 0, 4], [-1, 0, 5]],       # first eligible ratio NaN -> it wins
        [[-1, 0, 5], [-2, 0, nan], [-1, 0, 5]],       # later NaN ignored
        [[-0.0, 0, 5], [-2, 0, 0.0], [2, 0, -0.0]],
        [[-1, 0, 0.0], [1, 0, 0.0], [-1, 0, 5]],
        [[-inf, 0, 5], [-1, 0, 7]],
        [[1e-320, 0, 5], [-1e-320, 0, 5]],
    ]
    for rows in cases:
        rows = np.asarray(rows, dtype=np.float64)
        c = np.asarray([-1.0, 0.0])
        # keep b >= 0 so the phase-2 branch is taken where possible
        n, m = rows.shape[0], 2
        T = flat_of(rows, c)
        st, r, cc, e = oracle.pick(T, n, m)
        dev = spx.engine.DeviceTableau(n, m)
        dev.load(rows, c)
        dev.pick(0)
        s = dev.read_state()
        assert s.status == st, rows
        if st == oracle.PIVOT:
            assert (s.r, s.c) == (r, cc), rows
        res = spx.batched.solve_batched(T[None, :], n, m, max_pivots=1)
        o = oracle.solve_flat(T, n, m, max_pivots=1)
        assert res.status[0] == o.status and res.npiv[0] == o.npiv
        assert res.trace[0, : o.npiv].tolist() == o.trace.tolist()
        # NaN sign/payload is not defined by the reference (CPython) — NaN == NaN here
        both_nan = np.isnan(res.tables[0]) & np.isnan(o.table)
        assert (both_nan | (bits(res.tables[0]) == bits(o.table))).all()


# --------------------------------------------------------------------------- K3 internals
def test_hoisted_reciprocal_division_equals_div_rn(spx):
    """pivot_div (reciprocal hoisted out of the cell loop) == the compiler's div.rn.f64, bit for bit."""
    import ctypes
    torch = spx.torch
    L = spx.N.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    cnt = 1 << 24

    def bits(lo_exp, hi_exp, count):
        mant = torch.randint(0, 1 << 52, (count,), generator=g, device="cuda", dtype=torch.int64)
        exp = torch.randint(lo_exp, hi_exp + 1, (count,), generator=g, device="cuda", dtype=torch.int64)
        sign = torch.randint(0, 2, (count,), generator=g, device="cuda", dtype=torch.int64) << 63
        return (mant | (exp << 52) | sign).view(torch.float64)

    specials = torch.tensor([0.0, -0.0, 1.0, -1.0, float("inf"), -float("inf"), float("nan"), 5e-324, -5e-324,
                             2.2250738585072014e-308, 1.7976931348623157e308, 1e-300, 1e300, 3.0, 1.0 / 3.0,
                             0.9954454131449921, 2.0 ** -969, 2.0 ** -970, 2.0 ** 1017, 2.0 ** 1016],
                            dtype=torch.float64, device="cuda")
    suites = [
        (bits(1023 - 40, 1023 + 40, cnt), bits(1023 - 40, 1023 + 40, 4096)),      # ordinary tableau magnitudes
        (bits(0, 2046, cnt), bits(0, 2046, 4096)),                                # whole exponent range + denormals
        (bits(1, 120, cnt), bits(900, 1100, 1024)),                               # tiny numerators (guard 1)
        (bits(1000, 1046, cnt), bits(1, 60, 1024)),                               # overflowing quotients
        (bits(1, 80, cnt), bits(1980, 2046, 1024)),                               # underflowing quotients / huge p
        (specials.repeat_interleave(specials.numel()), specials),                 # specials x specials
    ]
    # near-midpoint quotients: a = q*p rounded, with q having a long run of 1s / 0s at the bottom
    q = (bits(1023, 1023, cnt).view(torch.int64) | 0x7FF).view(torch.float64)
    pp = bits(1023, 1023, 4096)
    suites.append((q * pp[torch.arange(cnt, device="cuda") % 4096], pp))
    for a, p in suites:
        a = a.contiguous(); p = p.contiguous()
        bad = ctypes.c_uint64(123)
        first = (ctypes.c_double * 2)()
        rc = L.spx_selftest_division(a.data_ptr(), p.data_ptr(), a.numel(), p.numel(), ctypes.byref(bad), first,
                                     torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        assert bad.value == 0, (bad.value, first[0].hex(), first[1].hex())


@pytest.mark.parametrize("opts", [{1: 1, 2: 2}, {1: 1, 2: 3}, {1: 1, 2: 4}, {1: 2, 3: 0}, {1: 2, 3: 1},
                                  {1: 2, 3: 0, 4: 7}, {1: 2, 3: 1, 4: 5}, {1: 1, 2: 3, 5: 64}, {1: 1, 2: 4, 5: 24},
                                  {1: 1, 2: 2, 5: 16}])
def test_every_update_kernel_variant_is_bit_exact(spx, opts):
    """Tiled (every register budget) and TMA-pipelined (both tile orders, odd grids) update kernels
    against the oracle on ragged shapes, including tiles with the pivot row/column and the f row."""
    L = spx.N.lib()
    try:
        for k, v in opts.items():
            assert L.spx_set_option(k, v) == 0
        for n, m in [(1, 2), (7, 15), (9, 17), (64, 512), (65, 513), (130, 1030), (40, 2049), (300, 700)]:
            rng = np.random.default_rng(n * 77 + m)
            rows, c = W.dense_lp(n, m, seed=n + m)
            rows[rng.random(rows.shape) < 0.05] = 0.0
            dev = spx.engine.DeviceTableau(n, m, trace_capacity=16)
            dev.load(rows, c, max_pivots=16)
            T = flat_of(rows, c)
            for npiv in range(6):
                st, r, cc, e = oracle.pick(T, n, m)
                dev.pick(npiv)
                s = dev.read_state()
                assert s.status == st
                if st != oracle.PIVOT:
                    break
                assert (s.r, s.c, s.p) == (r, cc, e)
                dev.update(npiv)
                T = oracle.update(T, n, m, r, cc)
                assert np.array_equal(bits(dev.export_flat(npiv + 1)), bits(T)), (opts, n, m, npiv)
    finally:
        for k, v in {1: 0, 2: 4, 3: 0, 4: 0, 5: 0}.items():
            L.spx_set_option(k, v)


# --------------------------------------------------------------------------- column-sharded flow
@pytest.mark.parametrize("exchange", ["copy", "mailbox"])
@pytest.mark.parametrize("lookahead", [False, True])
@pytest.mark.parametrize("world,n,m,kind", [(1, 20, 700, "dense"), (2, 24, 1100, "dense"), (4, 33, 2500, "dense"),
                                            (3, 12, 1300, "smallint"), (2, 9, 40, "smallint")])
def test_column_sharded_ranks_emulated_on_one_gpu(spx, world, n, m, kind, lookahead, exchange):
    """The sharded CUDA kernels (candidate / select / update with col0 > 0, and their look-ahead
    forms) with `world` ranks emulated in one process: the phases of every rank run in lockstep and
    the all-gather is a device copy.  Trace, labels, b and every body cell equal the oracle's."""
    if kind == "dense":
        rows, c = W.dense_lp(n, m, 3)
    else:
        rng = np.random.default_rng(5)
        rows = np.hstack([rng.integers(-3, 4, (n, m)).astype(float), rng.integers(-2, 7, (n, 1)).astype(float)])
        c = rng.integers(-3, 4, m).astype(float)
    _emulated_ranks(spx, world, n, m, rows, c, 40, lookahead, exchange)


def _emulated_ranks(spx, world, n, m, rows, c, cap, lookahead, exchange):
    from simplex_method_solver_b200 import parallel as P
    torch = spx.torch
    o = oracle.solve(rows, c, max_pivots=cap)
    boxes = None
    if exchange == "mailbox":                          # NVLink-style peer stores + flags, all boxes in one process
        shared = [None] * world
        boxes = [P.PeerMailboxes(n, r, world, "cuda", local_only_ptrs=shared) for r in range(world)]
        for bx in boxes:
            bx.finalize_shared()
    shards = [P.ShardedTableau(n, m, r, world, device="cuda", trace_capacity=cap + 8, lookahead=lookahead,
                               mailboxes=boxes[r] if boxes else None) for r in range(world)]
    for sh in shards:
        sh.load(rows, c, max_pivots=cap)

    def gather_all():
        torch.cuda.synchronize()                       # emulation only: every rank's send is complete
        if boxes:
            for sh in shards:                          # each rank pushes into every box, then everyone proceeds
                sh.phase_exchange()
        else:
            for sh in shards:
                for g, other in enumerate(shards):
                    sh.gathered[g].copy_(other.send)
        torch.cuda.synchronize()

    def lockstep(fn_local, fn_global):
        for sh in shards:
            fn_local(sh)
        gather_all()
        for sh in shards:
            fn_global(sh)

    if lookahead:
        def first_local(sh):
            cur, si = sh.npiv_enqueued & 1, sh.si
            sh.ops.candidate(sh.A[cur], sh.b[cur], sh.n, sh.m_loc, sh.ld, sh.col0, sh.rule, sh.states[si], sh.send)

        def first_global(sh):
            cur, si = sh.npiv_enqueued & 1, sh.si
            gathered, flags, seq = sh._gathered_and_flags()
            sh.ops.select(gathered, sh.world, sh.b[cur], sh.n, sh.rule, sh.states[si], sh.colbufs[si], flags, seq)
            sh.priced = True
        lockstep(first_local, first_global)
    for _ in range(cap + 3):                           # a few steps past the ending: terminal states propagate
        lockstep(lambda sh: sh.phase_local(), lambda sh: sh.phase_global())
    body = np.zeros((n + 1, m))
    for sh in shards:
        st = sh.sync()
        assert st.status == o.status and st.npiv == o.npiv, (sh.rank, st.status, st.npiv, o.status, o.npiv)
        assert sh.trace[: o.npiv].cpu().numpy().tolist() == o.trace.tolist()
        assert sh.rowlab.cpu().numpy().tolist() == o.rowlab.tolist()
        assert sh.collab[:n].cpu().numpy().tolist() == o.collab.tolist()
        body[:, sh.col0: sh.col0 + sh.m_loc] = sh.local_body().cpu().numpy()
        assert np.array_equal(bits(sh.b_current().cpu().numpy()),
                              bits(o.table[: n * (m + 1)].reshape(n, m + 1)[:, m].copy()))
    ob = np.zeros((n + 1, m))
    ob[:n] = o.table[: n * (m + 1)].reshape(n, m + 1)[:, :m]
    ob[n] = o.table[n * (m + 1):]
    assert np.array_equal(bits(body), bits(ob))
    if boxes:
        for bx in boxes:
            bx.close()


@pytest.mark.parametrize("pricing", [1, 2])
@pytest.mark.parametrize("n,m,depth", [(7, 15, 3), (64, 512, 8), (130, 1030, 5), (40, 2049, 8)])
def test_fused_pass_entry_point_both_pricing_kernels(spx, n, m, depth, pricing):
    """spx_fused_pass (the stand-alone pass bench.py times): one-CTA and whole-GPU cooperative pricing."""
    L = spx.N.lib()
    rows, c = W.dense_lp(n, m, seed=n + m)
    o = oracle.solve(rows, c, max_pivots=4 * depth)
    dev = spx.engine.DeviceTableau(n, m, trace_capacity=64)
    dev.load(rows, c, max_pivots=4 * depth)
    try:
        assert L.spx_set_option(8, pricing) == 0
        for _ in range(5):                                   # one pass more than the cap allows
            dev.fused_pass(depth, 0)
    finally:
        L.spx_set_option(8, 0)
    st = dev.read_state()
    assert (st.status, st.npiv) == (o.status, o.npiv)
    assert dev.trace[: st.npiv].cpu().numpy().tolist() == o.trace.tolist()
    cur = int(st.reserved[0]) & 1
    body = dev.A[cur, :, :m].cpu().numpy()
    ob = np.zeros((n + 1, m))
    ob[:n] = o.table[: n * (m + 1)].reshape(n, m + 1)[:, :m]
    ob[n] = o.table[n * (m + 1):]
    assert np.array_equal(bits(body), bits(ob))


@pytest.mark.parametrize("n,m,kind,depth", [(20, 700, "dense", 8), (33, 2500, "dense", 3), (12, 1300, "smallint", 8),
                                            (9, 40, "smallint", 5), (300, 700, "dense", 8)])
@pytest.mark.parametrize("ahead", ["per-pass", "persistent", False])
def test_fused_sharded_loop_single_rank(spx, n, m, kind, depth, ahead):
    """The column-sharded fused loop (cooperative pricing with the in-kernel exchange) with ONE rank:
    the whole code path except the cross-rank selection, against the oracle."""
    from simplex_method_solver_b200 import parallel as P
    if kind == "dense":
        rows, c = W.dense_lp(n, m, 3)
    else:
        rng = np.random.default_rng(5)
        rows = np.hstack([rng.integers(-3, 4, (n, m)).astype(float), rng.integers(-2, 7, (n, 1)).astype(float)])
        c = rng.integers(-3, 4, m).astype(float)
    cap = 45
    o = oracle.solve(rows, c, max_pivots=cap)
    sh = P.FusedShardedTableau(n, m, 0, 1, "cuda", trace_capacity=cap + 16, depth=depth, lookahead=ahead)
    try:
        sh.load(rows, c, max_pivots=cap)
        status, npiv = sh.solve(cap, check_every=7)
        assert (status, npiv) == (o.status, o.npiv)
        assert sh.trace[:npiv].cpu().numpy().tolist() == o.trace.tolist()
        assert sh.rowlab.cpu().numpy().tolist() == o.rowlab.tolist()
        assert sh.collab[:n].cpu().numpy().tolist() == o.collab.tolist()
        body = sh.local_body().cpu().numpy()
        ob = np.zeros((n + 1, m))
        ob[:n] = o.table[: n * (m + 1)].reshape(n, m + 1)[:, :m]
        ob[n] = o.table[n * (m + 1):]
        assert np.array_equal(bits(body), bits(ob))
        assert np.array_equal(bits(sh.b_current().cpu().numpy()), bits(o.table[: n * (m + 1)].reshape(n, m + 1)[:, m].copy()))
    finally:
        sh.close()


@pytest.mark.parametrize("n,m,kind", [(20, 700, "dense"), (33, 2500, "dense"), (12, 1300, "smallint"), (9, 40, "smallint")])
def test_simplexmethod_sharded_engine_reference_surface(spx, n, m, kind):
    """SimplexMethod(..., engine="sharded"): the column-sharded fused loop behind the reference's surface.  One rank here
    (the driver's box has one GPU; tests/test_multigpu.py runs the same call on 2-4 ranks): solve() in two instalments,
    x / objective / labels / trace against the oracle, find_optimum() and f(), and the step API refusing politely."""
    if kind == "dense":
        rows, c = W.dense_lp(n, m, 3)
    else:
        rng = np.random.default_rng(5)
        rows = np.hstack([rng.integers(-3, 4, (n, m)).astype(float), rng.integers(-2, 7, (n, 1)).astype(float)])
        c = rng.integers(-3, 4, m).astype(float)
    cap = 60
    sm = spx.simplex.SimplexMethod(rows, c, engine="sharded", max_pivots=cap)
    try:
        first = oracle.solve(rows, c, max_pivots=13)
        sol = sm.solve(max_pivots=13, chunk=5)
        assert (sol.status, sol.npiv) == (first.status, first.npiv)
        assert sol.trace.tolist() == first.trace.tolist()
        o = oracle.solve(rows, c, max_pivots=cap)
        if first.status == oracle.CAP:
            sol = sm.solve(max_pivots=cap - 13)
        assert (sol.status, sol.npiv) == (o.status, o.npiv)
        assert sol.trace.tolist() == o.trace.tolist()
        assert sol.rowlab.tolist() == o.rowlab.tolist() and sol.collab.tolist() == o.collab.tolist()
        assert sol.x.tobytes() == o.x.tobytes()
        assert float(sol.objective).hex() == float(o.objm).hex() and float(sol.obj2).hex() == float(o.obj2).hex()
        x1, x2 = sm.find_optimum()
        ref_sm_labels = [("x%d" % (v + 1)) if v < m else ("y%d" % (v - m + 1)) for v in o.collab.tolist()]
        assert sm.column[:-1] == ref_sm_labels
        want1 = float(o.x[0]) if "x1" in sm.column else 0
        want2 = float(o.x[1]) if "x2" in sm.column else 0
        assert (x1, x2) == (want1, want2)
        assert sm.f(x1, x2) == c[0] * x1 + c[1] * x2
        with pytest.raises(NotImplementedError):
            sm.get_solution()
        with pytest.raises(NotImplementedError):
            sm.pick_element()
    finally:
        sm.close()


# --------------------------------------------------------------------------- BASELINE configs
@pytest.mark.parametrize("lookahead", [False, True, "resident", "resident-ahead", "fused", "fused-coop", "fused-engine"])
def test_cfg2_dense_1000x2000_full_sequence(spx, cfg_digests, lookahead):
    with loop_mode(lookahead) as mode_:
        _cfg2_full(spx, cfg_digests, mode_)


def _cfg2_full(spx, cfg_digests, lookahead):
    g = cfg_digests["cfg2"]
    rows, c = W.dense_lp(1000, 2000, 0)
    assert W.input_digest(rows, c) == g["input_sha256"]
    sm = spx.simplex.SimplexMethod(rows, c, engine="stream")
    # the reference's own first 12 pivots and its table after them
    sol = sm.solve(max_pivots=12, chunk=12, lookahead=lookahead)
    assert sol.trace.tolist() == g["reference_first12"]["trace"]
    assert table_sha(sm._dev.export_flat(sm._npiv)) == g["reference_first12"]["table_sha256_after12"]
    # continue to optimality
    sol = sm.solve(max_pivots=200000 - 12, chunk=255, lookahead=lookahead)
    o = g["oracle_full"]
    assert sol.status == o["status"] == 0 and sol.npiv == o["npiv"] == 13579
    assert W.pivot_digest(sol.trace[:100]) == o["pivot_sha256_after100"]
    assert W.pivot_digest(sol.trace[:1000]) == o["pivot_sha256_after1000"]
    assert W.pivot_digest(sol.trace) == o["pivot_sha256_final"]
    assert table_sha(sm._dev.export_flat(sm._npiv)) == o["final_table_sha256"]
    assert hashlib.sha256(sol.x.astype("<f8").tobytes()).hexdigest() == o["x_sha256"]
    assert float(sol.objective).hex() == o["objm"] and float(sol.obj2).hex() == o["obj2"]
    assert hashlib.sha256(sol.collab.astype("<i4").tobytes()).hexdigest() == o["collab_sha256"]
    assert hashlib.sha256(sol.rowlab.astype("<i4").tobytes()).hexdigest() == o["rowlab_sha256"]


def test_cfg3_batched_65536(spx, cfg_digests):
    g = cfg_digests["cfg3"]
    T, C = W.gui_batch(65536, 0)
    assert W.input_digest(T, C) == g["input_sha256"]
    res = spx.batched.solve_batched(W.batch_flat(T, C), 8, 2, max_pivots=64)
    assert (res.status == 0).all()
    hist = {str(k): int(v) for k, v in zip(*np.unique(res.npiv, return_counts=True))}
    assert hist == g["pivot_histogram"]
    assert int(res.npiv.sum()) == g["total_pivots"] == 408212
    assert W.batch_pivot_digest(res.trace, res.npiv) == g["batch_pivot_sha256"]
    assert W.batch_solution_digest(res.x[:, 0], res.x[:, 1], res.obj) == g["batch_solution_sha256"]
    for k, e in enumerate(g["first256"]):
        assert res.trace[k, : res.npiv[k]].tolist() == e["trace"]
        assert table_sha(res.tables[k]) == e["final_table_sha256"]


@pytest.mark.parametrize("n", [10, 20])
def test_cfg5_klee_minty(spx, cfg_digests, n):
    g = cfg_digests[f"km{n}"]
    rows, c = W.klee_minty(n)
    cap = 1 << n
    res = spx.batched.solve_batched(flat_of(rows, c)[None, :], n, n, max_pivots=cap)
    assert res.status[0] == 0 and res.npiv[0] == g["npiv"] == (1 << n) - 1
    tr = res.trace[0, : res.npiv[0]]
    assert tr[:8].tolist() == g["first8"]
    assert W.pivot_digest(tr) == g["pivot_sha256"]
    assert table_sha(res.tables[0]) == g["final_table_sha256"]
    assert res.rowlab[0].tolist() == label_codes(g["row_labels"][:-1], n)
    assert res.collab[0].tolist() == label_codes(g["column_labels"][:-1], n)
    assert float(res.x[0, 0]).hex() == float(float.fromhex(g["x1"])).hex()
    assert float(res.obj[0]).hex() == g["f"]
    assert res.x[0, n - 1] == float(5 ** n) and (res.x[0, : n - 1] == 0).all()


@pytest.mark.parametrize("lookahead", [False, True, "resident", "fused"])
def test_cfg5_klee_minty10_streaming_equals_batched(spx, cfg_digests, lookahead):
    rows, c = W.klee_minty(10)
    sm = spx.simplex.SimplexMethod(rows, c, engine="stream")
    sol = sm.solve(max_pivots=2000, chunk=127, lookahead=lookahead)
    assert sol.status == 0 and sol.npiv == 1023
    assert W.pivot_digest(sol.trace) == cfg_digests["km10"]["pivot_sha256"]
    assert table_sha(sm._dev.export_flat(sm._npiv)) == cfg_digests["km10"]["final_table_sha256"]


def test_cfg4_16k_x_32k_prefix(spx, cfg_digests):
    """The 4.3 GB tableau: first 800 pivots against the oracle-generated golden prefix."""
    g = cfg_digests["cfg4"]
    n, m = 16384, 32768
    rows, c = W.dense_lp(n, m, 0)
    assert W.input_digest(rows, c) == g["input_sha256"]
    dev = spx.engine.DeviceTableau(n, m, trace_capacity=1024)
    dev.load(rows, c, max_pivots=1024)      # look-ahead prices pivot 801: keep the cap out of the way
    del rows
    dev.pick(0)
    s = dev.read_state()
    assert (s.r, s.c) == tuple(g["trace"][0]) and float(s.p).hex() == g["first_pivot_value"]
    done = 0
    for mark in (16, 50, 100, 200, 800):
        status, npiv = dev.solve(chunk=min(mark - done, 256), stop_after=mark - done)
        done = mark
        assert npiv == mark and status == spx.N.PIVOT
        tr = dev.trace[:mark].cpu().numpy()
        assert tr.tolist() == g["trace"][:mark]
        assert W.pivot_digest(tr) == g["marks"][str(mark)]["pivot_sha256"]
        b = dev.b_host(npiv)
        f = dev.A[dev.cur(npiv), n, :m].cpu().numpy()
        assert [float(v).hex() for v in b[:4]] == g["marks"][str(mark)]["b_first4"]
        assert [float(v).hex() for v in f[:4]] == g["marks"][str(mark)]["f_first4"]
        assert hashlib.sha256(b.tobytes()).hexdigest() == g["marks"][str(mark)]["b_sha256"]
        assert hashlib.sha256(f.tobytes()).hexdigest() == g["marks"][str(mark)]["f_sha256"]


def test_cfg4_whole_table_fused_2000_pivots_and_pivot_at_a_time_loop(spx):
    """Every cell of the 4.3 GB tableau, not only b and f: the fused loop (8 pivots per pass) after 2000 pivots and the
    pivot-at-a-time look-ahead loop (K3) after 100, each against the oracle's marks of tests/golden/cfg4_long.json —
    pivot sha256, b / f sha256 and the 64-bit checksum of all 536,870,912 body cells (so the two loops agree with each
    other through the oracle; tools/fused_check.py compares them directly)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg4_long.json")) as fh:
        marks = json.load(fh)["marks"]
    n, m = 16384, 32768
    rows, c = W.dense_lp(n, m, 0)

    def check(dev, npiv):
        mk = marks[str(npiv)]
        assert W.pivot_digest(dev.trace[:npiv].cpu().numpy()) == mk["pivot_sha256"]
        cur = dev.cur(npiv)
        assert hashlib.sha256(dev.b_host(npiv).tobytes()).hexdigest() == mk["b_sha256"]
        assert hashlib.sha256(dev.A[cur, n, :m].cpu().numpy().tobytes()).hexdigest() == mk["f_sha256"]
        assert W.body_checksum_torch(dev.A[cur, :n, :m]) == int(mk["body_checksum_u64"])

    dev = spx.engine.DeviceTableau(n, m, trace_capacity=2100)
    dev.load(rows, c, max_pivots=2100)
    status, npiv = dev.solve(chunk=1000, stop_after=2000, lookahead="fused")
    assert (status, npiv) == (spx.N.PIVOT, 2000)
    check(dev, 2000)
    dev.load(rows, c, max_pivots=200)
    status, npiv = dev.solve(chunk=100, stop_after=100, lookahead=True)
    assert (status, npiv) == (spx.N.PIVOT, 100)
    check(dev, 100)


# --------------------------------------------------------------------------- owner rank != 0 (kept last in this file)
@pytest.mark.parametrize("exchange", ["copy", "mailbox"])
@pytest.mark.parametrize("lookahead", [False, True])
@pytest.mark.parametrize("world,n,m,cap", [(2, 24, 1100, 60), (3, 64, 1600, 100), (4, 64, 1600, 100), (4, 40, 2100, 120)])
def test_column_sharded_ranks_emulated_entering_column_on_late_ranks(spx, world, n, m, cap, lookahead, exchange):
    """Same emulation as test_column_sharded_ranks_emulated_on_one_gpu on LPs whose entering column is owned by
    the LATER ranks and keeps changing owner (util.make_lp 'late'): in the dense and small-integer cases, and
    in the cfg4 prefix bench.py runs, rank 0 owns every entering column, so the winner != 0 branch of the
    select kernels and the pivot column outside rank 0's block are only exercised here."""
    from util import make_lp, owners_of
    from simplex_method_solver_b200.parallel import column_block
    rows, c = make_lp(n, m, 7, "late")
    o = oracle.solve(rows, c, max_pivots=cap)
    own = owners_of(o.trace.tolist(), [column_block(m, r, world) for r in range(world)])
    assert sum(g != 0 for g in own) >= 30                                 # the case really is what it claims
    _emulated_ranks(spx, world, n, m, rows, c, cap, lookahead, exchange)


# --------------------------------------------------------------------------- fused update kernel schedules
# "fused-x<variant>-<rows per warp strip>[-<min blocks>[-<column pairs per lane>]]": variant 0 = update_lazy_kernel (the
# default: ONE range test per cell per pass), 1 = round 1's update_fused_kernel (a range test per cell per level)
X_MODES = ["fused-x0-128", "fused-x0-64-2-1", "fused-x0-256-3-2", "fused-x0-8-2-2", "fused-x0-40-2-1", "fused-x0-4096-3-2",
           "fused-x0-24-3-1", "fused-x1-0", "fused-x1-0-3"]


@pytest.mark.parametrize("mode", X_MODES)
def test_fused_update_kernel_schedules(spx, ref_cases, cfg_digests, mode):
    """Every schedule of the fused update (kernel, tile height, occupancy target) must produce the oracle's bits:
    ragged shapes in steps of 1..9 pivots, the reference's golden cases (zeros, degenerate ties, phase 1 — the
    lazy guard's re-do path), and the first 300 pivots of cfg2."""
    with loop_mode(mode) as mode_:
        for n, m in [(1, 2), (3, 1), (7, 15), (65, 513), (130, 1030), (257, 100), (300, 700), (40, 2049)]:
            _ragged_loop(spx, n, m, mode_)
        for case in ref_cases:
            rows, c = case_inputs(case)
            sm = spx.simplex.SimplexMethod(rows, c, engine="stream")
            sol = sm.solve(max_pivots=case["cap"], chunk=5, lookahead=mode_)
            assert sol.status == END_TO_STATUS[case["end"]], case["name"]
            assert sol.trace.tolist() == case["trace"], case["name"]
            assert table_sha(sm._dev.export_flat(sm._npiv)) == case["final_table_sha256"], case["name"]
        rows, c = W.dense_lp(1000, 2000, 0)
        o = oracle.solve(rows, c, max_pivots=300)
        dev = spx.engine.DeviceTableau(1000, 2000, trace_capacity=400)
        dev.load(rows, c, max_pivots=300)
        dev.solve(stop_after=300, lookahead=mode_)
        assert dev.trace[:300].cpu().numpy().tolist() == o.trace.tolist()
        assert np.array_equal(bits(dev.export_flat(300)), bits(o.table))


def test_lazy_range_guard_adversarial_chains(spx):
    """The fused update divides WITHOUT a per-level range test and checks only the outputs of a pass
    (csrc/spx_fused.cu, update_lazy_kernel).  Adversarial 8-level chains — exact zeros, subnormals, cancellation
    down to 2^-1000 followed by growth, overflow, pivots at and beyond the guard's span — through the kernel's
    policy and through the per-level guarded division: every bit must agree."""
    import ctypes
    torch = spx.torch
    L = spx.N.lib()
    g = torch.Generator(device="cuda")
    g.manual_seed(20261018)
    F, cnt = 8, 1 << 20

    def rnd(lo_exp, hi_exp, shape):
        """random doubles with biased exponent in [lo_exp, hi_exp] (0 = zeros / subnormals), random sign"""
        numel = int(np.prod(shape))
        mant = torch.randint(0, 1 << 52, (numel,), generator=g, device="cuda", dtype=torch.int64)
        exp = torch.randint(lo_exp, hi_exp + 1, (numel,), generator=g, device="cuda", dtype=torch.int64)
        sign = torch.randint(0, 2, (numel,), generator=g, device="cuda", dtype=torch.int64) << 63
        return (mant | (exp << 52) | sign).view(torch.float64).reshape(shape)

    def sprinkle(x, frac, value):
        m = torch.rand(x.shape, generator=g, device="cuda") < frac
        return torch.where(m, torch.full_like(x, value), x)

    def growth(e0, step):
        """level l multipliers with exponent e0 + step * l (+- 3): what a tiny cell meets on its way up"""
        base = torch.arange(F, device="cuda", dtype=torch.int64)[None, :] * step + e0
        jit = torch.randint(-3, 4, (cnt, F), generator=g, device="cuda", dtype=torch.int64)
        e = (base + jit).clamp(1, 2046)
        mant = torch.randint(0, 1 << 52, (cnt, F), generator=g, device="cuda", dtype=torch.int64)
        sign = torch.randint(0, 2, (cnt, F), generator=g, device="cuda", dtype=torch.int64) << 63
        return (mant | (e << 52) | sign).view(torch.float64)

    one = torch.ones((cnt, F), dtype=torch.float64, device="cuda")
    pivots = [rnd(1023 - 2, 1023 + 2, (F,)), rnd(1023 - 100, 1023 + 100, (F,)),
              torch.tensor([2.0 ** -100, -2.0 ** 100, 1.5 * 2.0 ** 100, 2.0 ** -100 * 1.999, 1.0, -3.0, 0.1, 7e29],
                           dtype=torch.float64, device="cuda"),
              torch.tensor([1.0, 2.0 ** -101, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0], dtype=torch.float64, device="cuda"),   # guarded
              torch.tensor([1.0, 1.0, 2.0 ** 500, 1.0, 5e-324, 1.0, 1.0, 1.0], dtype=torch.float64, device="cuda")]  # guarded
    suites = []
    # ordinary magnitudes, with exact zeros sprinkled over cells, rows and columns (sparse tableaus)
    suites.append((sprinkle(rnd(1023 - 30, 1023 + 30, (cnt,)), 0.1, 0.0), sprinkle(rnd(1023 - 30, 1023 + 30, (cnt, F)), 0.2, 0.0),
                   sprinkle(rnd(1023 - 30, 1023 + 30, (cnt, F)), 0.2, -0.0)))
    # whole exponent range
    suites.append((rnd(0, 2046, (cnt,)), rnd(0, 2046, (cnt, F)), rnd(0, 2046, (cnt, F))))
    # tiny cells (subnormal .. 2^-900) meeting multipliers that grow by 2^step per level, products rj * ci with ci = 1
    for e0, step in [(1, 40), (10, 54), (30, 55), (60, 56), (1, 60), (1, 100), (100, 0), (1, 128)]:
        suites.append((sprinkle(rnd(0, 123, (cnt,)), 0.05, 0.0), growth(e0, step), one))
    # cancellation: rj * ci equals t * p to the last bits at level 0 (ci = p0), then ordinary / tiny levels
    for lo, hi in [(1023 - 20, 1023 + 20), (1, 200)]:
        t0 = rnd(1023 - 940, 1023 - 900, (cnt,)) if lo == 1 else rnd(1023 - 20, 1023 + 20, (cnt,))
        rj = rnd(lo, hi, (cnt, F))
        rj[:, 0] = t0 * (1.0 + (torch.randint(-2, 3, (cnt,), generator=g, device="cuda").double() * 2.0 ** -52))
        suites.append((t0, rj, None))                      # ci[:, 0] = p[0] is filled in per pivot set
    # overflow on the way
    suites.append((rnd(2000, 2046, (cnt,)), rnd(1023 - 5, 1023 + 600, (cnt, F)), rnd(1023 - 5, 1023 + 400, (cnt, F))))
    redo_total = 0
    for p in pivots:
        p = p.contiguous()
        for t0, rj, ci in suites:
            if ci is None:
                ci = rnd(1023 - 20, 1023 + 20, (cnt, F))
                ci[:, 0] = p[0]
            t0 = t0.contiguous(); rj = rj.contiguous(); ci = ci.contiguous()
            lz, ref = torch.empty_like(t0), torch.empty_like(t0)
            redo = ctypes.c_uint64(0)
            rc = L.spx_selftest_lazy_guard(t0.data_ptr(), p.data_ptr(), rj.data_ptr(), ci.data_ptr(), F, cnt,
                                           lz.data_ptr(), ref.data_ptr(), ctypes.byref(redo),
                                           torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            bad = (lz.view(torch.int64) != ref.view(torch.int64)).nonzero()
            assert bad.numel() == 0, (int(bad.numel()), int(bad[0]), float(t0[bad[0]]).hex(),
                                      float(lz[bad[0]]).hex(), float(ref[bad[0]]).hex())
            redo_total += redo.value
    assert redo_total > 0                                   # the suites do reach the re-do path
