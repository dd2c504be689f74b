"""Estimator updates in NumPy (TEST INFRASTRUCTURE - see ``oracle/__init__.py``).

Restates ``Estimator.py:231-261`` (`kalss`, also used for the Luenberger observer),
``:263-311`` (`kalman`) and ``:313-386`` (`ekf`) on the augmented state ``xi = [x; d]`` built at
``MPC_code.py:546-575``.  ``cmod`` supplies the augmented maps with their Jacobians
(`oracle.cmodel.model_functions`).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as scla


def _split(prob, xi):
    return (xi[:prob.nx], xi[prob.nx:]) if prob.flags["offree"] != "no" else (xi, np.zeros(prob.nd))


def kalss(prob, cmod, y_act, u_k, K, xi_min, t_k, p_y):
    """``xi = xi- + K (y - Fy(xi-))`` (``Estimator.py:253-259``)."""
    x, d = _split(prob, xi_min)
    yhat, _ = cmod.orc_fyes_d(x, d, u_k, t_k, p_y)
    return xi_min + K @ (y_act - yhat.ravel())


def kalman(prob, cmod, y_act, u_k, Q, R, P_min, xi_min, t_k, p_y, p_x, ts):
    """Linear Kalman filter (``Estimator.py:288-309``): the Jacobians are those of the (linear) maps."""
    x, d = _split(prob, xi_min)
    _, A = cmod.orc_fxes_d(x, d, u_k, ts, t_k, p_x)
    yhat, C = cmod.orc_fyes_d(x, d, u_k, t_k, p_y)
    K = np.linalg.solve((C @ P_min @ C.T + R).T, (P_min @ C.T).T).T          # :297
    P_corr = (np.eye(A.shape[0]) - K @ C) @ P_min                             # :300
    xi = xi_min + K @ (y_act - yhat.ravel())                                  # :303-306
    P_plus = A @ P_corr @ A.T + Q                                             # :309
    return P_plus, P_corr, xi


def ekf(prob, cmod, y_act, u_k, Q, R, P_min, xi_min, ts, t_k, p_y, p_x):
    """Extended Kalman filter (``Estimator.py:340-381``); A is taken at the *corrected* state, old input."""
    x, d = _split(prob, xi_min)
    yhat, C = cmod.orc_fyes_d(x, d, u_k, t_k, p_y)                            # :340-348
    inbrackets = scla.inv(np.linalg.multi_dot([C, P_min, C.T]) + R)           # :354
    K = np.linalg.multi_dot([P_min, C.T, inbrackets])                         # :355
    P_corr = P_min - np.linalg.multi_dot([K, C, P_min])                       # :358
    xi = xi_min + K @ (y_act - yhat.ravel())                                  # :361-367
    xc, dc = _split(prob, xi)
    _, A = cmod.orc_fxes_d(xc, dc, u_k, ts, t_k, p_x)                         # :370-376
    P_plus = np.linalg.multi_dot([A, P_corr, A.T]) + Q                        # :381
    return P_plus, P_corr, xi
