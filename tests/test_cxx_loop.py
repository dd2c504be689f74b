"""The C++/OpenMP CPU arm of bench.py (oracle/cxx_loop.cpp: host build of the solver sources, closed loop) against the
independent NumPy oracle's closed-loop fixture - the baseline must compute the same thing it is timed on."""
import os

import numpy as np

from conftest import GOLDEN

G = np.load(os.path.join(GOLDEN, "nmpc_oracle.npz"))


def test_cxx_closed_loop_matches_oracle_fixture(nmpc):
    from oracle.cxx_baseline import CxxLoop, range_bounds_of
    loop = CxxLoop("nmpc_cstr", nmpc.prob, nmpc.ss, nmpc.ocp, range_bounds_of(nmpc.ocp))
    x0, noise = G["cl_x0"], G["cl_noise"]
    rec = loop.run(noise.shape[0], x0, x0, noise)
    assert np.array_equal(rec["STATUS_DYN"], G["cl_STATUS_DYN"]) and np.array_equal(rec["ITER_DYN"], G["cl_ITER_DYN"])
    assert np.abs(rec["U"] - G["cl_U"]).max() < 1e-6
    assert rec["threads"] >= 1
