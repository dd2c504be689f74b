"""`ssjacid = True` (SS_JAC_ID.py:14-201, MPC_code.py:84-91): the nonlinear model is linearised at a steady state found near
(x0_m, u0).  Ex_LMPC_nlplant.py prints exactly such a linearisation of its CSTR (KAT2): giving the tool that CSTR as the
MODEL must reproduce the printed A, B to their printed digits, and the loop built from the result must behave like the
hand-linearised example."""
import os

import numpy as np

from conftest import ROOT
from mpc_code_b200.loader import load_example
from mpc_code_b200.problem import build_problem, make_specs


def _namespace():
    ns = load_example(os.path.join(ROOT, "examples", "lmpc_nlplant.py"))
    printed = dict(A=np.array(ns["A"]), B=np.array(ns["B"]))
    for name in ("A", "B", "xlin", "ulin"):
        del ns[name]
    fxp = ns["User_fxp_Cont"]
    ns["User_fxm_Cont"] = lambda x, u, d, t, px: fxp(x, t, u, None, None)
    ns["ssjacid"] = True
    ns["Bd"] = np.zeros((3, 2)); ns["Cd"] = np.eye(2)            # Bd = B is not known before the linearisation
    return ns, printed


def test_ssjacid_reproduces_the_printed_linearisation():
    ns, printed = _namespace()
    prob = build_problem(ns)
    A, B, xlin, ulin = np.asarray(ns["A"]), np.asarray(ns["B"]), ns["xlin"], ns["ulin"]
    # a steady state of the RK4 map next to the guess (the printed point is one only to its 3-4 printed digits)
    print("xlin", xlin, "ulin", ulin, "max |A - printed|", np.abs(A - printed["A"]).max(), "max |B - printed|", np.abs(B - printed["B"]).max())
    assert np.abs(xlin - np.array([0.5, 350.0, 0.659])).max() < 0.5 and np.abs(ulin - np.array([300.0, 0.1])).max() < 0.5
    assert np.abs(A - printed["A"]).max() < 0.05 * np.abs(printed["A"]).max() and np.abs(B - printed["B"]).max() < 0.05
    assert abs(np.linalg.eigvals(A).real.max() - np.linalg.eigvals(printed["A"]).real.max()) < 0.05      # the unstable pole (2.15)
    # the linear model is exact at the linearisation point and first-order accurate around it
    x1 = np.asarray(prob.Fx_model(xlin, ulin, prob.h, np.zeros(2), 0.0, np.zeros(3))).ravel()
    assert np.abs(x1 - xlin).max() < 1e-8
    ss, ocp = make_specs(prob)
    assert prob.flags["Sol_Hess_constdyn"] == "yes" and prob.estimator["type"] == "kal"
