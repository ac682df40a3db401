"""Pins the oracle against test/test_saddle_point.jl:32-74 (exact equality) and
the host mirror's select_initial_primal_weight against the same answers."""
import numpy as np

from folp_b200 import _marshal
from oracle import oracle
from shared_problems import example_lp, example_qp


def _norm2(v):
    return float(np.sqrt(np.cumsum(np.square(v))[-1]))


def test_select_initial_primal_weight():  # :32-65
    lp1 = example_lp()
    lp2 = example_lp()
    lp2.objective_vector = np.zeros(4)
    lp3 = example_lp()
    lp3.right_hand_side = np.zeros(3)
    imp = 1.3
    expect = imp * _norm2([5.0, 2.0, 1.0, 1.0]) / _norm2([12.0, 7.0, 1.0])
    for fn in (lambda p: oracle.select_initial_primal_weight(p, np.ones(4), np.ones(3), imp),
               lambda p: _marshal.select_initial_primal_weight(p, imp)):
        assert fn(lp1) == expect
        assert fn(lp2) == imp
        assert fn(lp3) == imp


def test_compute_lagrangian_value():  # :67-74
    assert oracle.lagrangian_value(example_lp(), np.zeros(4), np.zeros(3)) == -14.0
    qp = example_qp()
    assert oracle.lagrangian_value(qp, [1.0, 1.0], [0.0]) == 0.5
    assert oracle.lagrangian_value(qp, [1.0, 1.0], [1.0]) == 1.5
    assert oracle.lagrangian_value(qp, [0.25, 0.0], [0.0]) == -0.125
