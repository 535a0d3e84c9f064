"""CPU tests of the next-row adapters: the problem-file format of the
reference GUI and the headless compute_solution call site (the solver call itself needs a GPU and
is covered in tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from simplex_method_solver_b200 import problem_io as PIO
from simplex_method_solver_b200 import workloads as W


def test_problem_file_round_trip(tmp_path):
    rows = [[-39.7, -96.0, 4060.8], [-45.5, 45.3, 600.6], [45.5, -7.4, -54.6], [24.2, 45.1, -1091.42]]
    grad = [-1.0, -1.0, 0]
    text = PIO.dumps(rows, grad, 100)
    # exactly what save_state writes: str() of every number, lim last without a newline (main.py:387-395)
    assert text.splitlines()[0] == "-39.7,-96.0,4060.8" and text.splitlines()[-2] == "-1.0,-1.0,0"
    assert text.endswith("\n100") and not text.endswith("\n")
    p = PIO.loads(text)
    assert p.rows == rows and p.grad == [-1.0, -1.0, 0.0] and p.lim == 100
    path = tmp_path / "lp.txt"
    PIO.save(str(path), rows, grad, 7)
    q = PIO.load(str(path))
    assert q.rows == rows and q.lim == 7
    y, c = PIO.solver_inputs(q)
    assert y == W.CFG1_ROWS and c == W.CFG1_C


def test_problem_file_errors():
    with pytest.raises(ValueError):
        PIO.loads("1,2,3")                       # a single line: main.py:417-420
    with pytest.raises(ValueError):
        PIO.loads("1,2\n1,2,0\n5")               # a row without three numbers: main.py:437-438
    with pytest.raises(ValueError):
        PIO.loads("1,2,3\n1,2\n5")               # a bad gradient: main.py:443-444
    with pytest.raises(ValueError):
        PIO.loads("1,2,3\n1,2,0\nx")             # lim must be an int: main.py:472


def test_batch_tables_layout():
    T, C = W.gui_batch(5, seed=2)
    probs = [PIO.Problem(T[k].tolist(), C[k].tolist() + [0.0], 10) for k in range(5)]
    tabs, n, m = PIO.batch_tables(probs)
    assert (n, m) == (8, 2) and tabs.shape == (5, 26)
    assert np.array_equal(tabs, W.batch_flat(T, C))
    with pytest.raises(ValueError):
        PIO.batch_tables(probs + [PIO.Problem([[1.0, 2.0, 3.0]], [1.0, 1.0, 0.0], 1)])
    with pytest.raises(ValueError):
        PIO.batch_tables([])


def test_gui_adapter_inputs_from_atom_like():
    from simplex_method_solver_b200 import gui_adapter as G

    class Line:
        def __init__(self, coeffs):
            self.coeffs = coeffs

    class AtomLike:
        lines = [Line([1, 2, 3]), Line(["4.5", 5, 6])]
        grad = [7, 8, 0]
    y, c = G.inputs_from_atom(AtomLike())
    assert y == [[1.0, 2.0, 3.0], [4.5, 5.0, 6.0]] and c == [7, 8]
    assert AtomLike.grad == [7, 8, 0]                     # deep-copied, never mutated (main.py:312)
    with pytest.raises(ValueError):
        G.compute_solution(lines=[], grad=[1, 1, 0])
