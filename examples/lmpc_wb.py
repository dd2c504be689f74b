"""Offset-free linear MPC of the Wood-Berry distillation column (2 x 2, four first-order lags) with
a Luenberger observer and a Delta-u penalty.

Problem file in the CPCLAB-UNIPI/MPC-code user format; states the configuration BASELINE.json calls
"Ex_LMPC_WB": plant and model differ (the model's poles are perturbed), output disturbance model
(Bd = 0, Cd = I), observer gain K = [0; I], cost on C'QyC and S (Delta-u form), horizon 50, h = 1.
"""
import numpy as np
from casadi import SX

Nsim, N, h = 100, 50, 1

xp = SX.sym("xp", 4); x = SX.sym("x", 4); u = SX.sym("u", 2); y = SX.sym("y", 2); d = SX.sym("d", 2)

_poles = np.array([0.8871, 0.8324, 0.9092, 0.8703])
Ap = np.diag(_poles)
Bp = np.array([[1, 0], [1, 0], [0.0, 1.0], [0, 2.0]])
Cp = np.array([[1.4447, 0.0, -1.7169, 0.0], [0.0, 1.1064, 0.0, -1.2579]])

A = np.diag(_poles + 2 * np.array([0.01, -0.01, -0.01, 0.01]))      # model/plant mismatch
B = Bp.copy()
C = Cp.copy()

offree = "lin"
Bd = np.zeros((4, 2)); Cd = np.eye(2)
x0_p = np.zeros((4, 1)); x0_m = np.zeros((4, 1)); u0 = np.zeros((2, 1))

lue = True
K = np.vstack((np.zeros((4, 2)), np.eye(2)))


def defSP(t):
    ysp = np.array([0.0, 0.0]) if t <= 10 else np.array([1.0, -1.0])
    return [ysp, np.zeros(2), np.zeros(4)]      # ysp, usp, xsp


umin = -0.5 * np.ones((2, 1)); umax = 0.5 * np.ones((2, 1))

Qss = np.diag([1, 1]); Rss = np.zeros((2, 2))
Q = C.T @ np.diag([1, 1]) @ C
S = np.diag([10, 20])
