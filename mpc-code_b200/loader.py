"""Load an unmodified reference-style problem file (``Ex_*.py``) into a namespace.

The reference reads a problem by star-importing ``Default_Values`` and then the user's file
into the driver's globals and probing for names (``MPC_code.py:23-28,55,94-167``).  The same
is done here: defaults first, the file executed on top, with stand-in modules for the
packages those files import (``casadi``, ``casadi.tools``, ``matplotlib.pylab``,
``past.utils``, ``Utilities``) so that no edit of the user's file is needed.
"""
from __future__ import annotations

import contextlib
import math
import os
import sys
import types

import numpy as np

from . import sx as _sx
from .defaults import default_namespace

_SHIMMED = ("casadi", "casadi.tools", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot",
            "past", "past.utils", "Utilities")


def _public(mod):
    return {k: v for k, v in vars(mod).items() if not k.startswith("_")}


def _make_shims():
    casadi = types.ModuleType("casadi")
    casadi.__dict__.update({k: v for k, v in _public(_sx).items()
                            if k not in ("annotations", "math", "np", "S", "Expr", "List", "Sequence")})
    tools = types.ModuleType("casadi.tools")
    tools.simpleRK = _sx.simpleRK
    casadi.tools = tools

    mpl = types.ModuleType("matplotlib")
    pylab = types.ModuleType("matplotlib.pylab")
    pylab.linspace = np.linspace

    def _noplot(*a, **k):
        return None
    for name in ("figure", "plot", "step", "xlabel", "ylabel", "legend", "grid", "savefig", "close", "show",
                 "subplot", "title", "ion", "ioff"):
        setattr(pylab, name, _noplot)
    mpl.pylab = pylab
    mpl.pyplot = pylab

    past = types.ModuleType("past")
    utils = types.ModuleType("past.utils")
    utils.old_div = lambda a, b: a / b  # the files use true division (`from __future__ import division`)
    past.utils = utils

    from . import model_factory
    util = types.ModuleType("Utilities")
    util.__dict__.update(_public(casadi))
    util.__dict__.update(np=np, math=math, old_div=utils.old_div)
    import scipy.linalg as scla
    util.scla = scla
    for name in model_factory.__all__:
        setattr(util, name, getattr(model_factory, name))
    return {"casadi": casadi, "casadi.tools": tools, "matplotlib": mpl, "matplotlib.pylab": pylab,
            "matplotlib.pyplot": pylab, "past": past, "past.utils": utils, "Utilities": util}


@contextlib.contextmanager
def shim_modules():
    """Temporarily install the stand-in modules in ``sys.modules``."""
    saved = {name: sys.modules.get(name) for name in _SHIMMED}
    sys.modules.update(_make_shims())
    try:
        yield
    finally:
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod


def load_example(path: str, overrides: dict | None = None, source_edits=()) -> dict:
    """Execute ``path`` on top of the default flags and return the resulting namespace.

    ``overrides`` are applied after execution (e.g. ``{"Nsim": 20}``); they cannot change
    branches taken inside the file.  ``source_edits`` is a sequence of ``(old, new)`` text
    replacements applied to the source before execution, for switches a file hard-codes
    (e.g. ``("mhe_mod = 'on'", "mhe_mod = 'off'")`` in the reference's Ex_ENMPC.py:109).
    """
    path = os.path.abspath(path)
    with open(path, "r") as fh:
        source = fh.read()
    for old, new in source_edits:
        if old not in source:
            raise ValueError("source edit %r does not match anything in %s" % (old, path))
        source = source.replace(old, new)
    ns = default_namespace()
    ns["__name__"] = os.path.splitext(os.path.basename(path))[0]
    ns["__file__"] = path
    with shim_modules():
        code = compile(source, path, "exec")
        exec(code, ns)
    if overrides:
        ns.update(overrides)
    ns["_example_path"] = path
    return ns
