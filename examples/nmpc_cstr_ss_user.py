"""The CSTR NMPC problem of examples/nmpc_cstr.py with user constraints on the steady-state TARGET problem
(`User_g_ineq_SS`, `User_h_eq_SS`; Target_Calc.py:87-109,146-153): the jacket duty at the target is limited (active at
the nominal operating point, so the target moves away from the set-point) and the target level is tied to the outlet
flow by a nonlinear relation."""
import os

from casadi import vertcat

exec(compile(open(os.path.join(os.path.dirname(__file__), "nmpc_cstr.py")).read(), "nmpc_cstr.py", "exec"))


def User_g_ineq_SS(x, u, y, d, t, px, py):
    return vertcat((x[1] - u[0]) * x[2] - 15.9)                       # jacket duty at the target (K m)


def User_h_eq_SS(x, u, y, d, t, px, py):
    return vertcat(y[1] - 0.6528 - 5.0 * (u[1] ** 2 - 0.01))           # level target follows the outlet flow


Nsim = 30
