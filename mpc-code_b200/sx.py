"""CasADi-compatible matrix layer (``SX``, ``DM``, ``Function`` and friends).

Only the slice of the CasADi API the reference's user files and builders touch
is provided (reference usage: ``Ex_NMPC.py:27-31,57,68-76``, ``Ex_NMPC_dis.py:75-77``
in-place element assignment, ``Utilities.py:155-183`` ``Function``/``simpleRK``,
``Estimator.py:343-345`` ``jacobian``).  Matrices are dense, column-major for
``reshape``/``vec`` exactly like CasADi, and hold `symbolic.Expr` scalars.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np

from . import symbolic as S
from .symbolic import Expr

inf = math.inf
pi = math.pi


def _is_num(a):
    return isinstance(a, (int, float, bool, np.integer, np.floating, np.bool_))


class SX:
    """Dense matrix of scalar expressions."""

    __array_priority__ = 2000.0
    __array_ufunc__ = None

    def __init__(self, data=None, m=None):
        if data is None:
            arr = np.empty((0, 1), dtype=object)
        elif isinstance(data, SX):
            arr = data._a.copy()
        elif _is_num(data) and m is not None and isinstance(data, (int, np.integer)):
            arr = np.empty((int(data), int(m)), dtype=object)  # SX(n, m): zeros
            arr[...] = S.ZERO
        elif isinstance(data, Expr):
            arr = np.empty((1, 1), dtype=object)
            arr[0, 0] = data
        elif _is_num(data):
            arr = np.empty((1, 1), dtype=object)
            arr[0, 0] = S.const(data)
        else:
            arr = _to_obj(data)
        self._a = arr

    # -- constructors --------------------------------------------------------
    @staticmethod
    def sym(name, n=1, m=1):
        if isinstance(n, (tuple, list)):
            n, m = n
        arr = np.empty((n, m), dtype=object)
        for j in range(m):
            for i in range(n):
                arr[i, j] = S.sym(name if n * m == 1 else "%s_%d" % (name, i + j * n))
        out = SX.__new__(SX)
        out._a = arr
        return out

    @staticmethod
    def zeros(n=1, m=1):
        if isinstance(n, (tuple, list)):
            n, m = n
        return _wrap(_filled(n, m, S.ZERO))

    @staticmethod
    def ones(n=1, m=1):
        if isinstance(n, (tuple, list)):
            n, m = n
        return _wrap(_filled(n, m, S.ONE))

    @staticmethod
    def eye(n):
        a = _filled(n, n, S.ZERO)
        for i in range(n):
            a[i, i] = S.ONE
        return _wrap(a)

    @staticmethod
    def inf(n=1, m=1):
        return _wrap(_filled(n, m, S.const(math.inf)))

    # -- shape ---------------------------------------------------------------
    @property
    def shape(self): return self._a.shape
    def size1(self): return self._a.shape[0]
    def size2(self): return self._a.shape[1]
    def size(self, axis=None): return self._a.shape if axis is None else self._a.shape[axis - 1]
    def numel(self): return self._a.size
    def rows(self): return self._a.shape[0]
    def columns(self): return self._a.shape[1]
    def is_empty(self): return self._a.size == 0
    def is_scalar(self): return self._a.size == 1
    def is_symbolic(self): return all(e.op == "sym" for e in self._a.ravel(order="F"))
    def is_constant(self): return all(e.op == "const" for e in self._a.ravel(order="F"))
    @property
    def T(self): return _wrap(self._a.T.copy())
    def __len__(self): return self._a.shape[0]

    def elements(self) -> List[Expr]:
        return list(self._a.ravel(order="F"))

    nz = property(lambda self: self.elements())

    def reshape(self, shape, m=None):
        if m is not None:
            shape = (shape, m)
        return _wrap(self._a.reshape(tuple(shape), order="F").copy())

    def _as_scalar_expr(self) -> Expr:
        if self._a.size != 1:
            raise TypeError("expected a 1x1 SX, got %s" % (self._a.shape,))
        return self._a.ravel()[0]

    def __float__(self):
        return float(self._as_scalar_expr())

    def full(self):
        return np.array([[float(e) for e in row] for row in self._a], dtype=float).reshape(self._a.shape)

    def __array__(self, dtype=None, copy=None):
        return self.full() if dtype is None else self.full().astype(dtype)

    # -- indexing ------------------------------------------------------------
    def _index(self, key):
        a = self._a
        if isinstance(key, tuple):
            r, c = key
            r = _norm_idx(r, a.shape[0])
            c = _norm_idx(c, a.shape[1])
            return ("rc", r, c)
        # single index: linear, column-major (vectors: along the vector)
        return ("lin", _norm_idx(key, a.size), None)

    def __getitem__(self, key):
        kind, r, c = self._index(key)
        if kind == "rc":
            return _wrap(self._a[np.ix_(r, c)].copy())
        flat = self._a.ravel(order="F")
        sel = flat[r]
        if self._a.shape[0] == 1 and self._a.shape[1] > 1:
            return _wrap(sel.reshape(1, -1).copy())
        return _wrap(sel.reshape(-1, 1).copy())

    def __setitem__(self, key, value):
        kind, r, c = self._index(key)
        v = _to_obj(value)
        if kind == "rc":
            tgt = (len(r), len(c))
            self._a[np.ix_(r, c)] = _bcast(v, tgt)
            return
        vals = _bcast(v, (len(r), 1)).ravel(order="F") if v.size != len(r) else v.ravel(order="F")
        nrow = self._a.shape[0]
        for k, lin in enumerate(r):
            self._a[lin % nrow, lin // nrow] = vals[k]

    # -- arithmetic ----------------------------------------------------------
    def _ew(self, other, fn, swap=False):
        b = _to_obj(other)
        a = self._a
        if a.shape != b.shape:
            if b.size == 1:
                b = _bcast(b, a.shape)
            elif a.size == 1:
                a = _bcast(a, b.shape)
            else:
                raise ValueError("dimension mismatch %s vs %s" % (a.shape, b.shape))
        out = np.empty(a.shape, dtype=object)
        fa, fb, fo = a.ravel(), b.ravel(), out.ravel()
        for i in range(fa.size):
            fo[i] = fn(fb[i], fa[i]) if swap else fn(fa[i], fb[i])
        return _wrap(fo.reshape(a.shape))

    def __add__(self, o): return self._ew(o, S.add)
    def __radd__(self, o): return self._ew(o, S.add, True)
    def __sub__(self, o): return self._ew(o, S.sub)
    def __rsub__(self, o): return self._ew(o, S.sub, True)
    def __mul__(self, o): return self._ew(o, S.mul)
    def __rmul__(self, o): return self._ew(o, S.mul, True)
    def __truediv__(self, o): return self._ew(o, S.div)
    def __rtruediv__(self, o): return self._ew(o, S.div, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __pow__(self, o): return self._ew(o, S.power)
    def __rpow__(self, o): return self._ew(o, S.power, True)
    def __neg__(self): return _map1(self, S.neg)
    def __pos__(self): return self
    def __abs__(self): return _map1(self, lambda e: S.unary("fabs", e))
    def __lt__(self, o): return self._ew(o, lambda a, b: S.binary("lt", a, b))
    def __le__(self, o): return self._ew(o, lambda a, b: S.binary("le", a, b))
    def __gt__(self, o): return self._ew(o, lambda a, b: S.binary("lt", b, a))
    def __ge__(self, o): return self._ew(o, lambda a, b: S.binary("le", b, a))
    def __eq__(self, o): return self._ew(o, lambda a, b: S.binary("eq", a, b))
    def __ne__(self, o): return self._ew(o, lambda a, b: S.binary("ne", a, b))
    __hash__ = object.__hash__
    def __matmul__(self, o): return mtimes(self, o)
    def __rmatmul__(self, o): return mtimes(o, self)

    def __repr__(self):
        return "SX(%s)" % (self._a.tolist(),)


MX = SX  # the reference mixes MX (NLP level) and SX (model level); one graph type serves both here


class DM(np.ndarray):
    """Numeric matrix with the few CasADi conveniences the reference's loop uses."""

    def __new__(cls, data=0.0, m=None):
        if m is not None and isinstance(data, (int, np.integer)):
            arr = np.zeros((int(data), int(m)))
        else:
            if isinstance(data, SX):
                data = data.full()
            arr = np.array(data, dtype=float)
            if arr.ndim == 0:
                arr = arr.reshape(1, 1)
            elif arr.ndim == 1:
                arr = arr.reshape(-1, 1)
        return arr.view(cls)

    @staticmethod
    def zeros(n=1, m=1):
        if isinstance(n, (tuple, list)): n, m = n
        return np.zeros((n, m)).view(DM)
    @staticmethod
    def ones(n=1, m=1):
        if isinstance(n, (tuple, list)): n, m = n
        return np.ones((n, m)).view(DM)
    @staticmethod
    def eye(n): return np.eye(n).view(DM)
    @staticmethod
    def inf(n=1, m=1): return np.full((n, m), math.inf).view(DM)
    def size1(self): return self.shape[0]
    def size2(self): return self.shape[1] if self.ndim > 1 else 1
    def numel(self): return self.size
    def full(self): return np.asarray(self)
    def __array_finalize__(self, obj): pass


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------

def _wrap(arr) -> SX:
    out = SX.__new__(SX)
    out._a = arr
    return out


def _filled(n, m, e):
    a = np.empty((int(n), int(m)), dtype=object)
    a[...] = e
    return a


def _to_obj(v) -> np.ndarray:
    """Anything matrix-like -> 2-D object array of Expr (1-D -> column)."""
    if isinstance(v, SX):
        return v._a
    if isinstance(v, Expr):
        return _filled(1, 1, v)
    if _is_num(v):
        return _filled(1, 1, S.const(v))
    if isinstance(v, (list, tuple)):
        if len(v) and all(isinstance(e, (SX, Expr)) or _is_num(e) for e in v):
            flat = [S._lift(e) for e in v]
            out = np.empty((len(flat), 1), dtype=object)
            out[:, 0] = flat
            return out
        v = np.array(v, dtype=float)
    arr = np.asarray(v)
    if arr.dtype == object:
        arr2 = arr if arr.ndim == 2 else arr.reshape(-1, 1)
        out = np.empty(arr2.shape, dtype=object)
        for idx, e in np.ndenumerate(arr2):
            out[idx] = S._lift(e)
        return out
    arr = arr.astype(float)
    if arr.ndim == 0:
        arr = arr.reshape(1, 1)
    elif arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    out = np.empty(arr.shape, dtype=object)
    for idx, e in np.ndenumerate(arr):
        out[idx] = S.const(e)
    return out


def _bcast(v: np.ndarray, shape) -> np.ndarray:
    if v.shape == tuple(shape):
        return v
    if v.size == 1:
        return _filled(shape[0], shape[1], v.ravel()[0])
    if v.size == shape[0] * shape[1]:
        return v.reshape(shape, order="F")
    raise ValueError("cannot fit %s into %s" % (v.shape, shape))


def _norm_idx(k, n):
    if isinstance(k, slice):
        return list(range(*k.indices(n)))
    if isinstance(k, (int, np.integer)):
        k = int(k)
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError("index %d out of range %d" % (k, n))
        return [k]
    if isinstance(k, SX):
        return [int(float(e)) for e in k.elements()]
    return [int(i) for i in np.asarray(k).ravel()]


def _map1(x, fn):
    if _is_num(x):
        return None
    a = _to_obj(x)
    out = np.empty(a.shape, dtype=object)
    fa, fo = a.ravel(), out.ravel()
    for i in range(fa.size):
        fo[i] = fn(fa[i])
    return _wrap(fo.reshape(a.shape))


def _symbolic(x):
    return isinstance(x, (SX, Expr))


def _unary_factory(name, pyfun):
    def f(x):
        if _symbolic(x):
            return _map1(x, lambda e: S.unary(name, e))
        if _is_num(x):
            return pyfun(x)
        return getattr(np, {"fabs": "abs", "asin": "arcsin", "acos": "arccos", "atan": "arctan"}.get(name, name))(np.asarray(x, dtype=float))
    f.__name__ = name
    return f


exp = _unary_factory("exp", math.exp)
log = _unary_factory("log", math.log)
sqrt = _unary_factory("sqrt", math.sqrt)
sin = _unary_factory("sin", math.sin)
cos = _unary_factory("cos", math.cos)
tan = _unary_factory("tan", math.tan)
tanh = _unary_factory("tanh", math.tanh)
sinh = _unary_factory("sinh", math.sinh)
cosh = _unary_factory("cosh", math.cosh)
asin = _unary_factory("asin", math.asin)
acos = _unary_factory("acos", math.acos)
atan = _unary_factory("atan", math.atan)
fabs = _unary_factory("fabs", abs)
sign = _unary_factory("sign", lambda v: float((v > 0) - (v < 0)))


def _binary_factory(name, pyfun):
    def f(a, b):
        if _symbolic(a) or _symbolic(b):
            a = a if isinstance(a, SX) else SX(a)
            return a._ew(b, lambda p, q: S.binary(name, p, q))
        return pyfun(a, b)
    f.__name__ = name
    return f


fmin = _binary_factory("fmin", lambda a, b: np.minimum(a, b))
fmax = _binary_factory("fmax", lambda a, b: np.maximum(a, b))
power = _binary_factory("pow", lambda a, b: np.power(a, b))
atan2 = _binary_factory("atan2", lambda a, b: np.arctan2(a, b))
logic_and = _binary_factory("and", lambda a, b: float(bool(a) and bool(b)))
logic_or = _binary_factory("or", lambda a, b: float(bool(a) or bool(b)))


def sq(x):
    return x * x


def sumsqr(x):
    x = SX(x) if not isinstance(x, SX) else x
    acc = S.ZERO
    for e in x.elements():
        acc = S.add(acc, S.unary("sq", e))
    return SX(acc)


def if_else(c, a, b, short_circuit=False):
    if not (_symbolic(c) or _symbolic(a) or _symbolic(b)):
        return a if c else b
    C, A, B = _to_obj(c), _to_obj(a), _to_obj(b)
    shape = max((A.shape, B.shape, C.shape), key=lambda s: s[0] * s[1])
    C, A, B = _bcast(C, shape), _bcast(A, shape), _bcast(B, shape)
    out = np.empty(shape, dtype=object)
    for idx in np.ndindex(*shape):
        out[idx] = S.if_else_s(C[idx], A[idx], B[idx])
    return _wrap(out)


def vertcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    if not any(isinstance(a, SX) for a in args):
        parts = [np.asarray(DM(a)) for a in args if np.size(a)]
        return DM(np.vstack(parts)) if parts else DM.zeros(0, 1)
    parts = [_to_obj(a) for a in args]
    parts = [p for p in parts if p.size]
    if not parts:
        return SX()
    return _wrap(np.vstack(parts))


def horzcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    if not any(isinstance(a, SX) for a in args):
        parts = [np.asarray(DM(a)) for a in args if np.size(a)]
        return DM(np.hstack(parts)) if parts else DM.zeros(1, 0)
    parts = [_to_obj(a) for a in args]
    parts = [p for p in parts if p.size]
    return _wrap(np.hstack(parts))


def veccat(*args):
    return vertcat(*[vec(a) for a in args])


def vec(a):
    if isinstance(a, SX):
        return a.reshape((a.numel(), 1))
    return DM(np.asarray(DM(a)).reshape(-1, 1, order="F"))


def reshape(a, n, m=None):
    if m is None:
        n, m = n
    if isinstance(a, SX):
        return a.reshape((n, m))
    return DM(np.asarray(DM(a)).reshape((n, m), order="F"))


def transpose(a):
    return a.T


def mtimes(a, b=None, *rest):
    if b is None:  # mtimes([a, b, c])
        seq = list(a)
        out = seq[0]
        for nxt in seq[1:]:
            out = mtimes(out, nxt)
        return out
    if rest:
        out = mtimes(a, b)
        for nxt in rest:
            out = mtimes(out, nxt)
        return out
    if not (_symbolic(a) or _symbolic(b)):
        A, B = np.asarray(DM(a)), np.asarray(DM(b))
        if A.size == 1 or B.size == 1:
            return DM(A * B)
        return DM(A @ B)
    A, B = _to_obj(a), _to_obj(b)
    if A.size == 1 or B.size == 1:
        return SX(a) * b if isinstance(a, SX) or not isinstance(b, SX) else b * a
    if A.shape[1] != B.shape[0]:
        raise ValueError("mtimes: %s x %s" % (A.shape, B.shape))
    out = np.empty((A.shape[0], B.shape[1]), dtype=object)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            acc = S.ZERO
            for k in range(A.shape[1]):
                acc = S.add(acc, S.mul(A[i, k], B[k, j]))
            out[i, j] = acc
    return _wrap(out)


def dot(a, b):
    return mtimes(vec(SX(a)).T, vec(SX(b)))


def sum1(a):
    A = _to_obj(a)
    return mtimes(SX.ones(1, A.shape[0]), _wrap(A))


def sum2(a):
    A = _to_obj(a)
    return mtimes(_wrap(A), SX.ones(A.shape[1], 1))


def diag(a):
    if isinstance(a, SX):
        A = a._a
        if 1 in A.shape:
            v = A.ravel(order="F")
            out = _filled(v.size, v.size, S.ZERO)
            for i in range(v.size):
                out[i, i] = v[i]
            return _wrap(out)
        return _wrap(np.array([[A[i, i]] for i in range(min(A.shape))], dtype=object))
    return DM(np.diag(np.asarray(a, dtype=float).ravel())) if np.ndim(a) < 2 or 1 in np.shape(a) else DM(np.diag(np.asarray(a)))


def inv(a):
    if isinstance(a, SX):
        if not a.is_constant():
            raise NotImplementedError("symbolic matrix inverse")
        a = a.full()
    return DM(np.linalg.inv(np.asarray(a, dtype=float)))


def solve(a, b):
    return DM(np.linalg.solve(np.asarray(DM(a)), np.asarray(DM(b))))


def norm_2(a):
    return sqrt(sumsqr(a))


def jacobian(f, x):
    f = SX(f) if not isinstance(f, SX) else f
    rows = S.jacobian_entries(f.elements(), x.elements())
    out = _filled(f.numel(), x.numel(), S.ZERO)
    for i, r in enumerate(rows):
        for j, e in enumerate(r):
            out[i, j] = e
    return _wrap(out)


def gradient(f, x):
    f = SX(f) if not isinstance(f, SX) else f
    g = S.reverse_gradient(f._as_scalar_expr(), x.elements())
    return SX(g)


def hessian(f, x):
    f = SX(f) if not isinstance(f, SX) else f
    H = S.hessian_entries(f._as_scalar_expr(), x.elements())
    n = x.numel()
    out = _filled(n, n, S.ZERO)
    for i in range(n):
        for j in range(n):
            out[i, j] = H[i][j]
    return _wrap(out), gradient(f, x)


def jtimes(f, x, v, tr=False):
    J = jacobian(f, x)
    return mtimes(J.T, v) if tr else mtimes(J, v)


def substitute(f, x, v):
    f = SX(f) if not isinstance(f, SX) else f
    xs, vs = x.elements(), _to_obj(v).ravel(order="F")
    out = S.substitute(f.elements(), {a.uid: b for a, b in zip(xs, vs)})
    return SX(out).reshape(f.shape)


class Function:
    """Named map from symbolic inputs to symbolic outputs (CasADi ``Function`` look-alike).

    Calling it with symbolic arguments substitutes; with numeric arguments it evaluates
    and returns ``DM``.  ``meta`` carries builder annotations (e.g. that the map is an RK4
    of a continuous right-hand side) so the GPU path can use its hand-written integrator.
    """

    def __init__(self, name, ins, outs, in_names=None, out_names=None, opts=None):
        self.name_ = name
        self.ins = [i if isinstance(i, SX) else SX(i) for i in ins]
        for i in self.ins:
            if not i.is_symbolic():
                raise ValueError("Function %s: inputs must be purely symbolic" % name)
        self.outs = [o if isinstance(o, SX) else SX(o) for o in outs]
        self.in_names = in_names
        self.out_names = out_names
        self.meta = {}
        known = set()
        for i in self.ins:
            known.update(e.uid for e in i.elements())
        free = [s for s in S.symbols_of([e for o in self.outs for e in o.elements()]) if s.uid not in known]
        if free:
            raise ValueError("Function %s has free variables: %s" % (name, [f.val for f in free]))

    def name(self): return self.name_
    def n_in(self): return len(self.ins)
    def n_out(self): return len(self.outs)
    def size1_in(self, i): return self.ins[i].size1()
    def size1_out(self, i): return self.outs[i].size1()
    def sx_in(self, i=None): return list(self.ins) if i is None else self.ins[i]

    def _bind(self, args):
        if len(args) != len(self.ins):
            raise TypeError("%s expects %d arguments, got %d" % (self.name_, len(self.ins), len(args)))
        symbolic = any(isinstance(a, SX) and not a.is_constant() for a in args)
        return symbolic

    def call(self, args):
        symbolic = self._bind(args)
        flat_out = [e for o in self.outs for e in o.elements()]
        if symbolic:
            mapping = {}
            for decl, a in zip(self.ins, args):
                vals = _to_obj(a).ravel(order="F")
                if vals.size != decl.numel():
                    raise ValueError("%s: argument of size %d for input of size %d" % (self.name_, vals.size, decl.numel()))
                for s_, v_ in zip(decl.elements(), vals):
                    mapping[s_.uid] = v_
            res = S.substitute(flat_out, mapping)
        else:
            values = {}
            for decl, a in zip(self.ins, args):
                vals = (a.full() if isinstance(a, SX) else np.asarray(a, dtype=float)).ravel(order="F")
                if vals.size != decl.numel():
                    raise ValueError("%s: argument of size %d for input of size %d" % (self.name_, vals.size, decl.numel()))
                for s_, v_ in zip(decl.elements(), vals):
                    values[s_.uid] = float(v_)
            res = S.evaluate(flat_out, values)
        outs, pos = [], 0
        for o in self.outs:
            n = o.numel()
            chunk = res[pos:pos + n]
            pos += n
            if symbolic:
                outs.append(SX(chunk).reshape(o.shape) if n else SX())
            else:
                outs.append(DM(np.array(chunk, dtype=float).reshape(o.shape, order="F")))
        return outs

    def __call__(self, *args, **kwargs):
        if kwargs:
            if self.in_names is None:
                raise TypeError("%s has no named inputs" % self.name_)
            args = [kwargs[n] for n in self.in_names]
            outs = self.call(args)
            return dict(zip(self.out_names, outs))
        outs = self.call(list(args))
        return outs[0] if len(outs) == 1 else tuple(outs)

    def eval_batch(self, *args):
        """Vectorised numeric evaluation: each argument is ``[B, n_i]``; returns ``[B, n_out]`` arrays."""
        values = {}
        B = None
        for decl, a in zip(self.ins, args):
            a = np.asarray(a, dtype=float)
            if a.ndim == 1:
                a = a.reshape(1, -1) if decl.numel() == a.size else a.reshape(-1, 1)
            B = a.shape[0] if B is None else max(B, a.shape[0])
            for j, s_ in enumerate(decl.elements()):
                values[s_.uid] = a[:, j]
        outs = []
        for o in self.outs:
            res = S.evaluate(o.elements(), values)
            outs.append(np.stack([np.broadcast_to(np.asarray(r, dtype=float), (B,)) for r in res], axis=1)
                        if res else np.zeros((B, 0)))
        return outs[0] if len(outs) == 1 else tuple(outs)

    def jacobian_wrt(self, i_in, i_out=0):
        return jacobian(self.outs[i_out], self.ins[i_in])

    def __repr__(self):
        return "Function(%s: %s -> %s)" % (self.name_, [i.numel() for i in self.ins], [o.numel() for o in self.outs])


def simpleRK(f, N=10, order=4):
    """Fixed-step classic Runge-Kutta over one interval (stands in for ``casadi.tools.simpleRK``).

    ``f(x, p)`` is the right-hand side; the returned ``Function(x0, p, h)`` performs ``N`` RK4
    sub-steps of ``h / N``.  Reference call sites: ``Utilities.py:70,168``; the classic-RK4
    semantics are pinned by the Jacobians printed in ``Ex_LMPC_nlplant.py:85-91`` (KAT2).
    """
    if order != 4:
        raise NotImplementedError("only the classic 4th-order scheme is used by the reference")
    nx, npar = f.ins[0].numel(), f.ins[1].numel()
    x0 = SX.sym("x0", nx)
    p = SX.sym("p", npar)
    h = SX.sym("h", 1)
    dt = h / N
    x = x0
    for _ in range(N):
        k1 = f(x, p)
        k2 = f(x + dt / 2 * k1, p)
        k3 = f(x + dt / 2 * k2, p)
        k4 = f(x + dt * k3, p)
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    F = Function("F_RK", [x0, p, h], [x])
    F.meta.update(kind="rk4", rhs=f, substeps=N)
    return F
