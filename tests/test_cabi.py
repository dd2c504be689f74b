"""The C-ABI boundary: every symbol include/mpcb.h declares is exported, sizes are reported, no compute without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mpc_code_b200.solvers import C_SYMBOLS, MpcbLibrary, MpcbHandle


def _declared():
    text = open(os.path.join(ROOT, "include", "mpcb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcb_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(nmpc):
    names = _declared()
    assert len(names) >= 15 and set(names) == set(C_SYMBOLS)
    lib = ctypes.CDLL(nmpc.lib["so"])
    for n in names:
        assert hasattr(lib, n), n


def test_dims_and_flop_table_are_reported(nmpc):
    L = MpcbLibrary(nmpc.lib["so"])
    d, p = L.dims, nmpc.prob
    assert L.lib.mpcb_abi_version() == 2
    assert (d.nx, d.nu, d.ny, d.nd, d.N, d.Mx) == (p.nx, p.nu, p.ny, p.nd, p.N, 10)
    assert (d.nw, d.npar, d.ng, d.nwss, d.nparss, d.nxi) == (p.nw, p.npar, p.ny, p.nx + p.nu + p.ny, p.npar_ss, p.nxi)
    assert d.has_ocp == 1 and d.has_target == 1
    assert L.model_flops("mdl_f") == nmpc.lib["gen"]["flops"]["mdl_f"] > 0 and L.model_flops("nope") == -1
    o = L.default_opts()
    assert (o.max_iter, o.tol, o.mu_init, o.bound_relax_factor, o.honor_original_bounds) == (100, 1e-8, 0.1, 1e-8, 0)


def test_no_cpu_fallback_without_a_device(nmpc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU"):
        MpcbHandle(MpcbLibrary(nmpc.lib["so"]), 4)
    from mpc_code_b200.control_calc import opt_dyn
    from mpc_code_b200.solvers import BatchedNlpSolver
    s = BatchedNlpSolver("ocp", nmpc.ocp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s(x0=None, p=None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc-code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
