"""Headless adapter for the reference GUI's solver call site.

``MainWindow.compute_solution`` (/root/reference/src/main.py:308-317) builds the solver inputs
from the observable model — ``y = [list(map(float, row.coeffs)) for row in atom.lines]`` and
``c = deepcopy(atom.grad)[:-1]`` — runs ``SimplexMethod(y, c).get_solution()`` and keeps the list
of ``Info`` snapshots that ``on_combo_box_changed`` (main.py:319-326) feeds to
``TableWidget.update_from_info`` (table_widget.py:85-134) and ``PlotWidget.draw_point``
(plot_widget.py:432).  This module is that call site without PyQt5: it accepts any object with the
``Atom`` shape (``lines[i].coeffs``, ``grad``) or plain sequences, and returns what the GUI stores.
"""
from __future__ import annotations

import copy
from typing import NamedTuple, Sequence

from .simplex import Error, Info, SimplexMethod


class Solution(NamedTuple):
    tables: list            # list[Info | Error] — MainWindow.tables (main.py:313)
    rows: list              # y, the [a1, a2, b] rows as floats (main.py:309-311)
    grad: list              # atom.grad[:-1] (main.py:317)
    failed: bool            # any(isinstance(t, Error)) — the check of main.py:342-346


def inputs_from_atom(atom):
    """(y, c) exactly as main.py:309-312 derives them from an ``Atom``-like object."""
    y = [list(map(float, row.coeffs)) for row in atom.lines]
    c = copy.deepcopy(list(atom.grad))[:-1]
    return y, c


def compute_solution(atom=None, *, lines: Sequence[Sequence[float]] = None, grad: Sequence[float] = None,
                     **solver_kwargs) -> Solution:
    """The body of ``MainWindow.compute_solution`` on the B200 solver.

    Pass an ``Atom``-like object, or ``lines`` (rows ``[a1, a2, b]``) and ``grad`` (``[g1, g2, 0]``,
    the trailing cell is dropped as in main.py:312).  Raises ``ValueError`` when there is no
    constraint line, where the GUI shows its "incorrect data" box (main.py:329-331).
    """
    if atom is not None:
        y, c = inputs_from_atom(atom)
    else:
        y = [list(map(float, r)) for r in lines]
        c = list(grad)[:-1]
    if len(y) == 0:
        raise ValueError("no constraint lines: the simplex method needs at least one")
    tables = SimplexMethod(y, c, **solver_kwargs).get_solution()
    return Solution(tables, y, list(c), any(isinstance(t, Error) for t in tables))


def table_view(info: Info, ndigits: int = 2):
    """What ``TableWidget.update_from_info`` renders (table_widget.py:92-131): the first three
    columns of every constraint row and the first two of the f row, rounded to ``ndigits``, plus
    the pivot cell to highlight (``None`` on the last snapshot) and the rounded optimum."""
    body = [[round(float(v), ndigits) for v in row[:3]] for row in info.table[:-1]]
    f_row = [round(float(v), ndigits) for v in info.table[-1][:2]]
    pivot = None if info.i is None else (info.i, info.j)
    return {"row_labels": list(info.column), "column_labels": list(info.row), "body": body, "f": f_row,
            "pivot": pivot, "optimum": round(float(info.optimum), ndigits), "point": (info.x1, info.x2)}
