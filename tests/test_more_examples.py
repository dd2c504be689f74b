"""The reference examples outside the four BASELINE configurations, from the repo's own restatements (examples/) against
oracle fixtures generated from the UNMODIFIED reference files (tests/golden/make_golden.py --only-examples;
tests/test_examples_equivalence.py proves the restatements define the same problems):

* nmpc_dis        discrete-time user map with `if_else` clamps and hand-unrolled RK4, Delta-u bounds, user terminal
                  cost, time-varying plant parameter `def_pxp`, Luenberger observer   (Ex_NMPC_dis.py:75-77,120-125)
* lmpcxp_nlplant  model with one state more than the plant (nx = 4, nxp = 3), Kalman filter (Ex_LMPCxp_nlplant.py:98)

CPU: device code through the harness build.  GPU: the CUDA library through the C ABI, python loop and fused step."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

R = np.load(os.path.join(GOLDEN, "ref_examples_oracle.npz"))
KEYS = ("U", "X_HAT", "D_HAT", "XS", "US", "Xp", "Yp")
CASES = [("nmpc_dis", 16), ("lmpcxp_nlplant", 110)]


def _check(rec, tag, ns, keys=KEYS):
    assert np.array_equal(np.asarray(rec["STATUS_DYN"])[:ns, 0], R[tag + "_STATUS_DYN"][:ns])
    assert np.array_equal(np.asarray(rec["ITER_DYN"])[:ns, 0], R[tag + "_ITER_DYN"][:ns])
    assert np.array_equal(np.asarray(rec["STATUS_SS"])[:ns, 0], R[tag + "_STATUS_SS"][:ns])
    for key in keys:
        ref = R["%s_%s" % (tag, key)][:ns]
        diff = np.abs(np.asarray(rec[key])[:ns, 0, :] - ref).max()
        assert diff < 1e-6, (key, diff)
    f = np.asarray(rec["F_DYN"])[:ns, 0]; fr = R[tag + "_F_DYN"][:ns]
    assert np.all(np.abs(f - fr) <= 1e-8 * np.maximum(1.0, np.abs(fr)))


@pytest.mark.parametrize("name,ns", [("nmpc_dis", 12), ("lmpcxp_nlplant", 12)])
def test_device_code_on_cpu_matches_reference_fixture(name, ns, request):
    from harness_loop import HarnessLoop
    b = request.getfixturevalue(name)
    _check(HarnessLoop(b, 1).run(ns), name, ns)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ns", CASES)
@pytest.mark.parametrize("fused", [False, True])
def test_gpu_closed_loop_matches_reference_fixture(name, ns, fused, request):
    from mpc_code_b200.mpc_loop import CompiledProblem
    b = request.getfixturevalue(name)
    B = 3
    ctl = CompiledProblem(b.prob, name).controller(B)
    rec = {k: v.cpu().numpy() for k, v in ctl.run(ns, fused=fused).items()}
    for i in range(1, B):                                   # identical instances: identical trajectories
        assert np.array_equal(rec["U"][:, i], rec["U"][:, 0])
    _check(rec, name, ns)
