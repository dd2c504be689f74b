"""Offset-free linear MPC of a linearised CSTR (3 states, 2 inputs, 3 outputs) with a Kalman filter.

Problem file in the CPCLAB-UNIPI/MPC-code user format; states the configuration BASELINE.json calls
"Ex_LMPC_CSTR": discrete linear model = plant, additive state/output disturbances on the plant, linear
disturbance model (Bd = I, Cd = 0), horizon 50, h = 1, set-point change at t = 15.
"""
import numpy as np
import scipy.linalg as scla
from casadi import SX

Nsim, N, h = 100, 50, 1

xp = SX.sym("xp", 3); x = SX.sym("x", 3); u = SX.sym("u", 2); y = SX.sym("y", 3); d = SX.sym("d", 3)

A = np.array([[0.2511, -3.368e-03, -7.056e-04], [11.06, 0.3296, -2.545], [0.0, 0.0, 1.0]])
B = np.array([[-5.426e-03, 1.53e-05], [1.297, 0.1218], [0.0, -6.592e-02]])
C = np.eye(3)
Ap, Bp, Cp = A.copy(), B.copy(), C.copy()


def def_pxp(t):                     # unmeasured state disturbance acting on the plant until t = 20
    return [np.array([0.1, 0.0, 0.0]) if t <= 20 else np.zeros(3)]


def def_pyp(t):                     # constant output disturbance on the plant
    return [np.array([0.1, 0.1, 0.0])]


offree = "lin"
Bd = np.eye(3); Cd = np.zeros((3, 3))
x0_p = 3 * np.ones((3, 1)); x0_m = 3 * np.ones((3, 1)); u0 = np.zeros((2, 1))

kal = True
Q_kf = scla.block_diag(1.0e-7 * np.eye(3), np.eye(3)); R_kf = 1.0e-7 * np.eye(3); P0 = 1.0e-8 * np.eye(6)


def defSP(t):
    ysp = np.array([0.2, 0.0, 0.0]) if t <= 15 else np.array([0.0, 0.0, 0.1])
    return [ysp, np.zeros(2), np.zeros(3)]     # ysp, usp, xsp


umin = -10.0 * np.ones((2, 1)); umax = 10.0 * np.ones((2, 1))
xmin = np.array([-10.0, -8.0, -10.0]); xmax = 10.0 * np.ones((3, 1))
ymin = np.array([-10.0, -8.0, -10.0]); ymax = 10.0 * np.ones(3)

Qss = np.diag([20.0, 0.0, 1.0]); Rss = np.zeros((2, 2))
Q = np.diag([1.0, 0.0, 1.0]); R = 0.1 * np.eye(2)
