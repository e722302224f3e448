"""Explicit Runge-Kutta tableaus for the propagate-and-compress integrators with a time-dependent
Hamiltonian -- mirror of renormalizer/utils/rk.py:37-190 (RungeKutta): same method names, and the
(a, b, c) arrays in the shapes the reference's integrator reads (b has one row per embedded order)."""
from fractions import Fraction as F

import numpy as np

# method -> (a rows (strictly lower triangular part), b rows, c, orders of the b rows)
_TABLEAUS = {
    "Forward_Euler": ([[]], [[1]], [0], (1,)),
    "midpoint_RK2": ([[], [1]], [[0, 1]], [0, 1], (2,)),
    "Heun_RK2": ([[], [F(1, 2)]], [[0, 1]], [0, F(1, 2)], (2,)),
    "Ralston_RK2": ([[], [F(2, 3)]], [[F(1, 4), F(3, 4)]], [0, F(2, 3)], (2,)),
    "Kutta_RK3": ([[], [F(1, 2)], [-1, 2]], [[F(1, 6), F(2, 3), F(1, 6)]], [0, F(1, 2), 1], (3,)),
    "C_RK4": ([[], [F(1, 2)], [0, F(1, 2)], [0, 0, 1]],
              [[F(1, 6), F(1, 3), F(1, 3), F(1, 6)]], [0, F(1, 2), F(1, 2), 1], (4,)),
    "38rule_RK4": ([[], [F(1, 3)], [F(-1, 3), 1], [1, -1, 1]],
                   [[F(1, 8), F(3, 8), F(3, 8), F(1, 8)]], [0, F(1, 3), F(2, 3), 1], (4,)),
}
_FEHLBERG_A = [[], [F(1, 4)], [F(3, 32), F(9, 32)], [F(1932, 2197), F(-7200, 2197), F(7296, 2197)],
               [F(439, 216), -8, F(3680, 513), F(-845, 4104)],
               [F(-8, 27), 2, F(-3544, 2565), F(1859, 4104), F(-11, 40)]]
_FEHLBERG_C = [0, F(1, 4), F(3, 8), F(12, 13), 1, F(1, 2)]
_FEHLBERG_B5 = [F(16, 135), 0, F(6656, 12825), F(28561, 56430), F(-9, 50), F(2, 55)]
_FEHLBERG_B4 = [F(25, 216), 0, F(1408, 2565), F(2197, 4104), F(-1, 5), 0]
_TABLEAUS["Fehlberg5"] = (_FEHLBERG_A, [_FEHLBERG_B5], _FEHLBERG_C, (5,))
_TABLEAUS["RKF45"] = (_FEHLBERG_A, [_FEHLBERG_B5, _FEHLBERG_B4], _FEHLBERG_C, (5, 4))
_TABLEAUS["Cash-Karp45"] = (
    [[], [F(1, 5)], [F(3, 40), F(9, 40)], [F(3, 10), F(-9, 10), F(6, 5)],
     [F(-11, 54), F(5, 2), F(-70, 27), F(35, 27)],
     [F(1631, 55296), F(175, 512), F(575, 13824), F(44275, 110592), F(253, 4096)]],
    [[F(37, 378), 0, F(250, 621), F(125, 594), 0, F(512, 1771)],
     [F(2825, 27648), 0, F(18575, 48384), F(13525, 55296), F(277, 14336), F(1, 4)]],
    [0, F(1, 5), F(3, 10), F(3, 5), 1, F(7, 8)], (5, 4))
# the reference's "midpoint_RK2" / "Heun_RK2" / "Ralston_RK2" are the one-parameter family
# a21 = alpha, b = (1 - 1/(2 alpha), 1/(2 alpha)), c2 = alpha with alpha = 1, 1/2, 2/3 (rk.py:73-88)
for _name, _alpha in (("midpoint_RK2", F(1)), ("Heun_RK2", F(1, 2)), ("Ralston_RK2", F(2, 3))):
    _TABLEAUS[_name] = ([[], [_alpha]], [[1 - 1 / (2 * _alpha), 1 / (2 * _alpha)]], [0, _alpha], (2,))

method_list = list(_TABLEAUS)


class RungeKutta:
    def __init__(self, method="C_RK4"):
        if method not in _TABLEAUS:
            raise AssertionError(f"unknown Runge-Kutta method {method}")
        self.method = method
        rows, b, c, order = _TABLEAUS[method]
        n = len(c)
        a = np.zeros((n, n))
        for i, row in enumerate(rows):
            for j, v in enumerate(row):
                a[i, j] = float(v)
        self.stage = n
        self.order = tuple(order)
        self.tableau = [a, np.array([[float(v) for v in row] for row in b]).reshape(-1, n),
                        np.array([float(v) for v in c])]
