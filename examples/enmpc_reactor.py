"""Economic nonlinear MPC of an isothermal CSTR with consecutive reactions A -> B -> C.

Problem file in the CPCLAB-UNIPI/MPC-code user format; states the configuration BASELINE.json calls
"Ex_ENMPC" with its extended-Kalman-filter estimator branch (the file's `mhe_mod = 'off'` alternative; moving-horizon
estimation is outside the accelerated path): states (cA, cB) both measured, input = dilution rate, economic stage
cost integrated over each interval (ContForm), steady-state economic target, terminal penalty, horizon 25, h = 2.
"""
import numpy as np
import scipy.linalg as scla
from casadi import SX, mtimes, vertcat

Nsim, N, h, Mx = 21, 25, 2.0, 10

xp = SX.sym("xp", 2); x = SX.sym("x", 2); u = SX.sym("u", 1); y = SX.sym("y", 2); d = SX.sym("d", 2)
StateFeedback = True

cA0, V, k1, k2 = 1.0, 1.0, 1.0, 0.05        # feed concentration, volume, rate constants
alfa, beta = 1.0, 4.0                        # reactant price, product price


def _reactor(xv, uv):
    return vertcat(uv[0] * (cA0 - xv[0]) / V - k1 * xv[0], -uv[0] * xv[1] / V + k1 * xv[0] - k2 * xv[1])


def User_fxp_Cont(xp, t, u, pxp, pxmp):
    return _reactor(xp, u)


def User_fxm_Cont(x, u, d, t, px):
    return _reactor(x, u)


offree = "lin"
Bd = np.zeros((2, 2)); Cd = np.eye(2)
x0_p = np.array([0.9, 0.1]); x0_m = np.array([1.2, 0.5]); u0 = np.array([0.0])

ekf = True
Q_kf = scla.block_diag(1.0e-8 * np.eye(2), np.eye(2)); R_kf = 1.0e-8 * np.eye(2); P0 = 1.0e-8 * np.eye(4)

umin = [0.00]; umax = [2.0]
xmin = np.array([0.00, 0.00]); xmax = np.array([1.00, 1.00])


def User_fssobj(x, u, y, xsp, usp, ysp):          # economic steady-state objective
    return u[0] * (alfa * cA0 - beta * y[1])


def User_fobj_Cont(x, u, y, xs, us, ys):          # economic stage cost, integrated over the interval
    return u[0] * (alfa * cA0 - beta * y[1])


def User_vfin(x, xs):
    diffx = x - xs
    return mtimes(diffx.T, mtimes(2000, diffx))


Sol_itmax = 200
ContForm = True
