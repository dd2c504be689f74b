"""Nonlinear MPC of a quadruple-tank process stated in DISCRETE time, Luenberger observer, Delta-u bounds.

Problem file in the CPCLAB-UNIPI/MPC-code user format; same configuration as the reference's "Ex_NMPC_dis":
states (u1_prev, u2_prev, h1, h2, h3, h4) - the two valve commands are carried as states - inputs = the two valve
commands, measured the two lower tank levels.  The user supplies the one-step map directly (`User_fxm_Dis`): five
hand-unrolled classic RK4 sub-steps of the tank dynamics per h = 5 s, with the levels clamped to [0, 20] by `if_else`
before every evaluation of the right-hand side.  Output-disturbance model (Bd = 0, Cd = I), observer gain K = [0; I],
Q / S (Delta-u) stage cost, user terminal cost 100 |x|^2, bounds on u, x, y and Delta-u.  The plant is the same map
with an additive step disturbance on the tank levels (`def_pxp`).
"""
import numpy as np
from casadi import SX, if_else, mtimes, vertcat

Nsim, N, h = 1000, 50, 5.0

xp = SX.sym("xp", 6); x = SX.sym("x", 6); u = SX.sym("u", 2); y = SX.sym("y", 2); d = SX.sym("d", 2)

_g = 981.0                                   # cm/s^2
_a = (0.071, 0.057, 0.071, 0.057)            # outlet cross-sections, cm^2
_A = (28.0, 32.0, 28.0, 32.0)                # tank cross-sections, cm^2
_gm1, _gm2 = 0.7, 0.6                        # flow splitting factors
_K1 = (_a[0] + _a[3]) * (2.0 * _g * 20.0) ** 0.5 / 100.0
_K2 = (_a[1] + _a[2]) * (2.0 * _g * 20.0) ** 0.5 / 100.0


def _tanks(lv, uv):
    """Level dynamics of the four tanks; `lv` is clamped IN PLACE to [0, 20] first (as the reference does)."""
    for i in range(lv.shape[0]):
        lv[i] = if_else(lv[i] < 0, 0., lv[i])
        lv[i] = if_else(lv[i] > 20, 20., lv[i])
    f = SX(4, 1)
    f[0] = -(_a[0] / _A[0]) * (2.0 * _g * lv[0]) ** 0.5 + (_a[2] / _A[0]) * (2.0 * _g * lv[2]) ** 0.5 + (_gm1 / _A[0]) * _K1 * uv[0]
    f[1] = -(_a[1] / _A[1]) * (2.0 * _g * lv[1]) ** 0.5 + (_a[3] / _A[1]) * (2.0 * _g * lv[3]) ** 0.5 + (_gm2 / _A[1]) * _K2 * uv[1]
    f[2] = -(_a[2] / _A[2]) * (2.0 * _g * lv[2]) ** 0.5 + ((1.0 - _gm2) / _A[2]) * _K2 * uv[1]
    f[3] = -(_a[3] / _A[3]) * (2.0 * _g * lv[3]) ** 0.5 + ((1.0 - _gm1) / _A[3]) * _K1 * uv[0]
    return f


def _one_step(xv, uv):
    """[u; RK4 of the tank levels over h with 5 sub-steps]."""
    out = SX(xv.size1(), 1)
    dt = h / 5
    lv = xv[2:6]
    out[0:2] = uv
    for _ in range(5):
        k1 = _tanks(lv, uv)
        k2 = _tanks(lv + dt / 2.0 * k1, uv)
        k3 = _tanks(lv + dt / 2.0 * k2, uv)
        k4 = _tanks(lv + dt * k3, uv)
        lv = lv + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    out[2:6] = lv
    return out


def User_fxp_Dis(x, t, u, pxp, pxmp):
    return _one_step(x, u)


def User_fyp(x, u, t, pyp, pymp):
    return vertcat(x[2], x[3])


def def_pxp(t):                              # additive plant disturbance on the tank levels
    if t <= 2250:
        return [np.array([0., 0., 0.5, 0., 0., 0.])]
    if t <= 4000:
        return [np.array([0., 0., 0., 0.5, 0., 0.])]
    return [np.zeros(6)]


def User_fxm_Dis(x, u, d, t, px):
    return _one_step(x, u)


def User_fym(x, u, d, t, px):
    return vertcat(x[2], x[3])


offree = "lin"
Bd = np.zeros((6, 2)); Cd = np.eye(2)
x0_p = np.array([39.5794, 38.1492, 11.9996, 12.1883, 1.51364, 1.42194]); x0_m = x0_p.copy()
u0 = np.array([39.5794, 38.1492])

lue = True
K = np.vstack([np.zeros((6, 2)), np.eye(2)])

_USP = np.array([39.5185, 38.1743])
_SP = [(50, [11.9996, 12.1883], [50.0, 50.0, 10.0, 10.0, 2.0, 2.0]),
       (1000, [11.9996, 6.0], [60.0, 50.0, 12.0, 8.0, 2.0, 2.0]),
       (2000, [6.0, 6.0], [60.0, 40.0, 12.0, 8.0, 2.0, 2.0]),
       (3000, [12.0, 12.0], [40.0, 40.0, 8.0, 8.0, 2.0, 2.0]),
       (4000, [8.0, 12.0], [40.0, 60.0, 8.0, 12.0, 2.0, 2.0]),
       (5000, [10.0, 10.0], [50.0, 50.0, 10.0, 10.0, 2.0, 2.0]),
       (float("inf"), [8.0, 12.0], [40.0, 40.0, 8.0, 12.0, 2.0, 2.0])]


def defSP(t):
    for t_end, ysp, xsp in _SP:
        if t <= t_end:
            return [np.array(ysp), _USP.copy(), np.array(xsp)]


umin = np.array([0.0, 0.0]); umax = np.array([100.0, 100.0])
xmin = np.zeros((6, 1)); xmax = np.array([100.0, 100.0, 20.0, 20.0, 20.0, 20.0])
ymin = np.array([0.0, 0.0]); ymax = np.array([20.0, 20.0])
Dumin = np.array([-50.0, -50.0]); Dumax = np.array([50.0, 50.0])

Qss = np.eye(2); Sss = np.zeros((2, 2))
Q = np.diag([1e3, 1e3, 1.0, 1.0, 1e-6, 1e-6]); S = 10.0 * np.eye(2)


def User_vfin(x, xs):
    return mtimes(x.T, mtimes(100.0, x))
