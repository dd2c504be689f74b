"""The CSTR NMPC problem of examples/nmpc_cstr.py with a user-defined NONLINEAR stage inequality
(`User_g_ineq`, Control_Calc.py:94-96,132-137,244-245): the heat removed through the jacket,
proportional to (T - Tc) * level, is limited.  The limit is chosen so that it is active during the
transient that follows a perturbed start, which changes the optimal input trajectory."""
import os

from casadi import vertcat

exec(compile(open(os.path.join(os.path.dirname(__file__), "nmpc_cstr.py")).read(), "nmpc_cstr.py", "exec"))


def User_g_ineq(x, u, y, d, t, px, py):
    return vertcat((x[1] - u[0]) * x[2] - 16.8,          # jacket duty (K m)
                   0.55 - y[1] - 0.05 * (u[1] - 0.1))    # level margin tied to the outlet flow


Nsim = 30
