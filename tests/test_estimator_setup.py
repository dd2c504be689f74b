"""Kkalss (steady-state Kalman gain, Estimator.py:103-229 / MPC_code.py:339-363): host-side set-up for `kalss = True`."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, **over):
    import mpc_code_b200  # noqa: F401
    from mpc_code_b200.loader import load_example
    return load_example(os.path.join(ROOT, "examples", name + ".py"), overrides=over)


def test_gain_is_the_limit_of_the_time_varying_filter():
    from mpc_code_b200.problem import build_problem
    ns = _load("lmpc_cstr", kal=False, kalss=True)
    p = build_problem(ns)
    assert p.estimator["type"] == "kalss"
    K = p.estimator["K"]
    A = np.block([[ns["A"], ns["Bd"]], [np.zeros((3, 3)), np.eye(3)]]); C = np.hstack([ns["C"], ns["Cd"]])
    P = np.array(ns["P0"], dtype=float)
    for _ in range(3000):                                              # kalman(), Estimator.py:288-309
        Kt = P @ C.T @ np.linalg.inv(C @ P @ C.T + ns["R_kf"]); P = A @ ((np.eye(6) - Kt @ C) @ P) @ A.T + ns["Q_kf"]
    assert np.abs(Kt - K).max() < 1e-10
    assert np.abs(np.linalg.eigvals(A - A @ K @ C)).max() < 1.0        # the observer is stable (Estimator.py:224-226)


def test_linearised_model_path_matches_given_matrices():
    """linmod = 'no': A and C are taken as Jacobians of Fx_model / Fy_model at (x_ss, u_ss) - for a linear model they
    must reproduce the gain computed from the matrices themselves."""
    from mpc_code_b200.estimator_setup import Kkalss
    from mpc_code_b200.problem import build_problem
    ns = _load("lmpc_cstr", kal=False, kalss=True)
    p = build_problem(ns)
    s = p.sym
    K2 = Kkalss(p.ny, p.nd, p.nx, ns["Q_kf"], ns["R_kf"], "lin", "no", s["x"], s["u"], s["k"], s["d"], s["t"], p.h, s["px"], s["py"],
                np.zeros(3), np.zeros(2), None, None, Bd=ns["Bd"], Cd=ns["Cd"], Fx=p.Fx_model, Fy=p.Fy_model)
    assert np.abs(K2 - p.estimator["K"]).max() < 1e-12


def test_nonlinear_disturbance_model_has_no_steady_state_gain():
    """offree = 'nl': the reference augments A with an identity block that is not coupled to the states
    (Estimator.py:183-190), so the disturbance states are undetectable and SciPy's DARE solver raises - here as there."""
    from mpc_code_b200.problem import build_problem
    ns = _load("nmpc_cstr", ekf=False, kalss=True)
    ns["x_ss"], ns["u_ss"] = ns["x0_m"], ns["u0"]
    with pytest.raises(np.linalg.LinAlgError):
        build_problem(ns)
