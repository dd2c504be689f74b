"""The CSTR NMPC problem of examples/nmpc_cstr.py with a terminal equality constraint, x_N = x_s (TermCons = True,
Control_Calc.py:197-198): nonlinear model, exact Hessian, N = 50 - the case that exercises the terminal-multiplier
extension of the Riccati sweep together with the RK4 derivative kernels."""
import os

exec(compile(open(os.path.join(os.path.dirname(__file__), "nmpc_cstr.py")).read(), "nmpc_cstr.py", "exec"))
TermCons = True
Nsim = 30
