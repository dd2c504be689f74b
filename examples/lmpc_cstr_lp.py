"""The linear CSTR problem of `lmpc_cstr.py` with LINEAR-PROGRAMMING costs: stage cost r_x |x - xs| + r_u |u - us|, target
cost rss_y y + rss_u |u| (the reference's `r_x` / `r_u` / `rss_y` / `rss_u` branch, Utilities.py:289-298,341-352,
MPC_code.py:202-232).  No reference example ships this form; it exercises `fabs` / `sign` through the code generator."""
import numpy as np
import scipy.linalg as scla
from casadi import SX, mtimes

Nsim, N, h = 20, 20, 1

xp = SX.sym("xp", 3); x = SX.sym("x", 3); u = SX.sym("u", 2); y = SX.sym("y", 3); d = SX.sym("d", 3)

A = np.array([[0.2511, -3.368e-03, -7.056e-04], [11.06, 0.3296, -2.545], [0.0, 0.0, 1.0]])
B = np.array([[-5.426e-03, 1.53e-05], [1.297, 0.1218], [0.0, -6.592e-02]])
C = np.eye(3)
Ap, Bp, Cp = A.copy(), B.copy(), C.copy()

offree = "lin"
Bd = np.eye(3); Cd = np.zeros((3, 3))
x0_p = 3 * np.ones((3, 1)); x0_m = 3 * np.ones((3, 1)); u0 = np.zeros((2, 1))

kal = True
Q_kf = scla.block_diag(1.0e-7 * np.eye(3), np.eye(3)); R_kf = 1.0e-7 * np.eye(3); P0 = 1.0e-8 * np.eye(6)


def defSP(t):
    return [np.array([0.2, 0.0, 0.0]), np.zeros(2), np.zeros(3)]     # ysp, usp, xsp


umin = -10.0 * np.ones((2, 1)); umax = 10.0 * np.ones((2, 1))
xmin = np.array([-10.0, -20.0, -10.0]); xmax = 10.0 * np.ones((3, 1))
ymin = np.array([-10.0, -20.0, -10.0]); ymax = 10.0 * np.ones(3)

rss_y = np.array([[1.0, 0.1, 1.0]]); rss_u = np.array([[0.1, 0.1]])
r_x = np.array([[1.0, 0.1, 1.0]]); r_u = np.array([[0.1, 0.1]])


def User_vfin(x, xs):              # a linear model without Q has no DARE terminal weight (MPC_code.py:250-255): the user's
    return mtimes(x.T, x)
