"""Offset-free linear MPC of the linearised CSTR with a TERMINAL EQUALITY constraint (TermCons = True).

Same plant, model, estimator and weights as examples/lmpc_cstr.py, but the horizon is shortened to 15 steps and the
terminal state is pinned to the target, x_N = x_s (Control_Calc.py:197-198), instead of being weighted by the DARE
cost alone.  The bounds are wide enough for x_s to be reachable in 15 steps.
"""
import numpy as np
import scipy.linalg as scla
from casadi import SX

Nsim, N, h = 40, 15, 1

xp = SX.sym("xp", 3); x = SX.sym("x", 3); u = SX.sym("u", 2); y = SX.sym("y", 3); d = SX.sym("d", 3)

A = np.array([[0.2511, -3.368e-03, -7.056e-04], [11.06, 0.3296, -2.545], [0.0, 0.0, 1.0]])
B = np.array([[-5.426e-03, 1.53e-05], [1.297, 0.1218], [0.0, -6.592e-02]])
C = np.eye(3)
Ap, Bp, Cp = A.copy(), B.copy(), C.copy()


def def_pxp(t):
    return [np.array([0.05, 0.0, 0.0]) if t <= 10 else np.zeros(3)]


offree = "lin"
Bd = np.eye(3); Cd = np.zeros((3, 3))
x0_p = np.zeros((3, 1)); x0_m = np.zeros((3, 1)); u0 = np.zeros((2, 1))

kal = True
Q_kf = scla.block_diag(1.0e-7 * np.eye(3), np.eye(3)); R_kf = 1.0e-7 * np.eye(3); P0 = 1.0e-8 * np.eye(6)


def defSP(t):
    ysp = np.array([0.05, 0.0, 0.0]) if t <= 20 else np.array([0.0, 0.0, 0.05])
    return [ysp, np.zeros(2), np.zeros(3)]


umin = -10.0 * np.ones((2, 1)); umax = 10.0 * np.ones((2, 1))
xmin = -10.0 * np.ones(3); xmax = 10.0 * np.ones((3, 1))

Qss = np.diag([20.0, 0.0, 1.0]); Rss = np.zeros((2, 2))
Q = np.diag([1.0, 0.0, 1.0]); R = 0.1 * np.eye(2)
TermCons = True
