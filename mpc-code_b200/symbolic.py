"""Scalar-expression DAG with hash-consing, symbolic AD and evaluation.

This is the host-side tracer that stands in for the CasADi ``SX`` layer the
reference is written against (reference: every ``Ex_*.py`` builds its model
from ``SX.sym`` symbols, e.g. ``Ex_NMPC.py:27-31,114-150``; the driver turns
them into ``Function`` objects in ``Utilities.py:102-245``).  It records what
the user wrote as a DAG of scalar nodes; `codegen.py` turns DAGs into C/CUDA.

Nothing here runs on the hot path: it executes once per problem, at build time.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import numpy as np

# ----------------------------------------------------------------------------
# scalar nodes
# ----------------------------------------------------------------------------

_UNARY = ("neg", "exp", "log", "sqrt", "sin", "cos", "tan", "tanh", "fabs", "sign",
          "not", "sq", "asin", "acos", "atan", "sinh", "cosh")
_BINARY = ("add", "sub", "mul", "div", "pow", "lt", "le", "eq", "ne", "and", "or",
           "fmin", "fmax", "atan2")
_TERNARY = ("if_else",)


class Expr:
    """One scalar node.  Instances are unique per (op, operands) - compare with ``is``."""

    __slots__ = ("op", "args", "val", "uid")
    __array_priority__ = 1000.0
    __array_ufunc__ = None

    def __init__(self, op, args=(), val=None, uid=0):
        self.op = op
        self.args = args
        self.val = val
        self.uid = uid

    # -- arithmetic ----------------------------------------------------------
    def __add__(self, o): return add(self, _lift(o))
    def __radd__(self, o): return add(_lift(o), self)
    def __sub__(self, o): return sub(self, _lift(o))
    def __rsub__(self, o): return sub(_lift(o), self)
    def __mul__(self, o): return mul(self, _lift(o))
    def __rmul__(self, o): return mul(_lift(o), self)
    def __truediv__(self, o): return div(self, _lift(o))
    def __rtruediv__(self, o): return div(_lift(o), self)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __neg__(self): return neg(self)
    def __pos__(self): return self
    def __pow__(self, o): return power(self, _lift(o))
    def __rpow__(self, o): return power(_lift(o), self)
    def __abs__(self): return unary("fabs", self)
    def __lt__(self, o): return binary("lt", self, _lift(o))
    def __le__(self, o): return binary("le", self, _lift(o))
    def __gt__(self, o): return binary("lt", _lift(o), self)
    def __ge__(self, o): return binary("le", _lift(o), self)
    __hash__ = object.__hash__

    def is_const(self): return self.op == "const"

    def __float__(self):
        if self.op != "const":
            raise TypeError("symbolic expression has no numeric value")
        return float(self.val)

    def __repr__(self):
        if self.op == "const":
            return repr(self.val)
        if self.op == "sym":
            return str(self.val)
        return "%s(%s)" % (self.op, ",".join(repr(a) for a in self.args))


_TABLE: Dict[tuple, Expr] = {}
_COUNTER = [0]


def _new(op, args=(), val=None) -> Expr:
    _COUNTER[0] += 1
    return Expr(op, args, val, _COUNTER[0])


def const(v) -> Expr:
    v = float(v)
    if v == 0.0:
        v = 0.0  # fold -0.0
    key = ("const", v) if not math.isnan(v) else ("const", "nan")
    e = _TABLE.get(key)
    if e is None:
        e = _new("const", (), v)
        _TABLE[key] = e
    return e


ZERO = const(0.0)
ONE = const(1.0)
TWO = const(2.0)
HALF = const(0.5)


def sym(name: str) -> Expr:
    return _new("sym", (), name)  # never shared: two symbols of one name are distinct


def _lift(o) -> Expr:
    if isinstance(o, Expr):
        return o
    if isinstance(o, (bool, np.bool_)):
        return ONE if o else ZERO
    if isinstance(o, (int, float, np.integer, np.floating)):
        return const(o)
    if isinstance(o, np.ndarray) and o.size == 1:
        return _lift(o.reshape(-1)[0])
    if hasattr(o, "_as_scalar_expr"):
        return o._as_scalar_expr()
    raise TypeError("cannot use %r in a scalar expression" % (type(o),))


def _mk(op, *args) -> Expr:
    key = (op,) + tuple(a.uid for a in args)
    e = _TABLE.get(key)
    if e is None:
        e = _new(op, tuple(args))
        _TABLE[key] = e
    return e


_PYFUN = {
    "neg": lambda a: -a, "exp": math.exp, "log": math.log, "sqrt": math.sqrt, "sin": math.sin,
    "cos": math.cos, "tan": math.tan, "tanh": math.tanh, "fabs": abs,
    "sign": lambda a: (a > 0) - (a < 0), "not": lambda a: float(a == 0.0), "sq": lambda a: a * a,
    "asin": math.asin, "acos": math.acos, "atan": math.atan, "sinh": math.sinh, "cosh": math.cosh,
    "add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
    "div": lambda a, b: a / b, "pow": lambda a, b: a ** b,
    "lt": lambda a, b: float(a < b), "le": lambda a, b: float(a <= b), "eq": lambda a, b: float(a == b),
    "ne": lambda a, b: float(a != b), "and": lambda a, b: float(a != 0 and b != 0),
    "or": lambda a, b: float(a != 0 or b != 0), "fmin": min, "fmax": max, "atan2": math.atan2,
    "if_else": lambda c, a, b: a if c != 0 else b,
}


def _after(a: Expr, b: Expr) -> bool:
    """Canonical operand order of the commutative operations: non-constants by creation order, constants last.
    Constants are interned for the life of the process, so their uid says when some EARLIER trace first used the value;
    ordering by it made the generated text (and with it the library digest) depend on what had been traced before."""
    ac, bc = a.op == "const", b.op == "const"
    if ac != bc:
        return ac
    return a.uid > b.uid


def add(a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        return const(a.val + b.val)
    if a is ZERO:
        return b
    if b is ZERO:
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    if a.op == "neg":
        return sub(b, a.args[0])
    if _after(a, b):  # canonical order (commutative)
        a, b = b, a
    return _mk("add", a, b)


def sub(a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        return const(a.val - b.val)
    if b is ZERO:
        return a
    if a is ZERO:
        return neg(b)
    if a is b:
        return ZERO
    if b.op == "neg":
        return add(a, b.args[0])
    return _mk("sub", a, b)


def neg(a: Expr) -> Expr:
    if a.op == "const":
        return const(-a.val)
    if a.op == "neg":
        return a.args[0]
    if a.op == "sub":
        return sub(a.args[1], a.args[0])
    return _mk("neg", a)


def mul(a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        return const(a.val * b.val)
    if a is ZERO or b is ZERO:
        return ZERO
    if a is ONE:
        return b
    if b is ONE:
        return a
    if a.op == "const" and a.val == -1.0:
        return neg(b)
    if b.op == "const" and b.val == -1.0:
        return neg(a)
    if a.op == "neg" and b.op == "neg":
        return mul(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(mul(a.args[0], b))
    if b.op == "neg":
        return neg(mul(a, b.args[0]))
    if a is b:
        return unary("sq", a)
    if _after(a, b):
        a, b = b, a
    return _mk("mul", a, b)


def div(a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        return const(a.val / b.val) if b.val != 0.0 else const(math.copysign(math.inf, a.val) if a.val else math.nan)
    if a is ZERO:
        return ZERO
    if b is ONE:
        return a
    if b.op == "const" and b.val == -1.0:
        return neg(a)
    if a.op == "neg" and b.op == "neg":
        return div(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(div(a.args[0], b))
    if b.op == "neg":
        return neg(div(a, b.args[0]))
    return _mk("div", a, b)


def power(a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        return const(a.val ** b.val)
    if b.op == "const":
        if b.val == 0.0:
            return ONE
        if b.val == 1.0:
            return a
        if b.val == 2.0:
            return unary("sq", a)
        if b.val == 0.5:
            return unary("sqrt", a)
        if b.val == -1.0:
            return div(ONE, a)
    return _mk("pow", a, b)


def unary(op: str, a: Expr) -> Expr:
    if op == "neg":
        return neg(a)
    if a.op == "const":
        return const(_PYFUN[op](a.val))
    if op == "fabs" and a.op in ("fabs", "sq", "exp"):
        return a
    if op in ("sq", "fabs", "cos", "cosh") and a.op == "neg":
        return unary(op, a.args[0])
    return _mk(op, a)


def binary(op: str, a: Expr, b: Expr) -> Expr:
    if op == "add": return add(a, b)
    if op == "sub": return sub(a, b)
    if op == "mul": return mul(a, b)
    if op == "div": return div(a, b)
    if op == "pow": return power(a, b)
    if a.op == "const" and b.op == "const":
        return const(_PYFUN[op](a.val, b.val))
    return _mk(op, a, b)


def if_else_s(c: Expr, a: Expr, b: Expr) -> Expr:
    if c.op == "const":
        return a if c.val != 0.0 else b
    if a is b:
        return a
    return _mk("if_else", c, a, b)


# ----------------------------------------------------------------------------
# graph utilities
# ----------------------------------------------------------------------------

def topo_order(outputs: Iterable[Expr]) -> List[Expr]:
    """Nodes reachable from ``outputs`` in dependency order (iterative DFS)."""
    seen = set()
    order: List[Expr] = []
    for root in outputs:
        if root.uid in seen:
            continue
        stack = [(root, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0:
                if node.uid in seen:
                    continue
                seen.add(node.uid)
            if i < len(node.args):
                stack.append((node, i + 1))
                child = node.args[i]
                if child.uid not in seen:
                    stack.append((child, 0))
            else:
                order.append(node)
    return order


def symbols_of(outputs: Iterable[Expr]) -> List[Expr]:
    return [n for n in topo_order(outputs) if n.op == "sym"]


def substitute(outputs: Sequence[Expr], mapping: Dict[int, Expr]) -> List[Expr]:
    """Rebuild ``outputs`` with symbols (by uid) replaced through ``mapping``."""
    memo: Dict[int, Expr] = dict(mapping)
    for n in topo_order(outputs):
        if n.uid in memo:
            continue
        if not n.args:
            memo[n.uid] = n
            continue
        a = [memo[c.uid] for c in n.args]
        if all(x is y for x, y in zip(a, n.args)):
            memo[n.uid] = n
        elif len(a) == 1:
            memo[n.uid] = unary(n.op, a[0])
        elif len(a) == 2:
            memo[n.uid] = binary(n.op, a[0], a[1])
        else:
            memo[n.uid] = if_else_s(a[0], a[1], a[2])
    return [memo[o.uid] for o in outputs]


_NPFUN = {
    "neg": np.negative, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos,
    "tan": np.tan, "tanh": np.tanh, "fabs": np.abs, "sign": np.sign,
    "not": lambda a: (np.asarray(a) == 0).astype(float), "sq": np.square,
    "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "sinh": np.sinh, "cosh": np.cosh,
    "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "pow": np.power,
    "lt": lambda a, b: np.less(a, b).astype(float), "le": lambda a, b: np.less_equal(a, b).astype(float),
    "eq": lambda a, b: np.equal(a, b).astype(float), "ne": lambda a, b: np.not_equal(a, b).astype(float),
    "and": lambda a, b: np.logical_and(a, b).astype(float), "or": lambda a, b: np.logical_or(a, b).astype(float),
    "fmin": np.minimum, "fmax": np.maximum, "atan2": np.arctan2,
    "if_else": lambda c, a, b: np.where(np.asarray(c) != 0, a, b),
}


def evaluate(outputs: Sequence[Expr], values: Dict[int, object]) -> list:
    """Numerically evaluate; ``values`` maps symbol uid -> float or ndarray (broadcast)."""
    memo = dict(values)
    for n in topo_order(outputs):
        if n.uid in memo:
            continue
        if n.op == "const":
            memo[n.uid] = n.val
        elif n.op == "sym":
            raise KeyError("free symbol %r has no value" % (n.val,))
        else:
            with np.errstate(all="ignore"):
                memo[n.uid] = _NPFUN[n.op](*[memo[c.uid] for c in n.args])
    return [memo[o.uid] for o in outputs]


# ----------------------------------------------------------------------------
# symbolic differentiation
# ----------------------------------------------------------------------------

def _partials(n: Expr) -> List[Expr]:
    """d n / d arg_i for each operand, as expressions."""
    op, a = n.op, n.args
    if op == "add": return [ONE, ONE]
    if op == "sub": return [ONE, const(-1.0)]
    if op == "mul": return [a[1], a[0]]
    if op == "div":  # a/b
        return [div(ONE, a[1]), neg(div(n, a[1]))]
    if op == "neg": return [const(-1.0)]
    if op == "sq": return [mul(TWO, a[0])]
    if op == "exp": return [n]
    if op == "log": return [div(ONE, a[0])]
    if op == "sqrt": return [div(HALF, n)]
    if op == "sin": return [unary("cos", a[0])]
    if op == "cos": return [neg(unary("sin", a[0]))]
    if op == "tan": return [add(ONE, unary("sq", n))]
    if op == "tanh": return [sub(ONE, unary("sq", n))]
    if op == "sinh": return [unary("cosh", a[0])]
    if op == "cosh": return [unary("sinh", a[0])]
    if op == "asin": return [div(ONE, unary("sqrt", sub(ONE, unary("sq", a[0]))))]
    if op == "acos": return [neg(div(ONE, unary("sqrt", sub(ONE, unary("sq", a[0])))))]
    if op == "atan": return [div(ONE, add(ONE, unary("sq", a[0])))]
    if op == "fabs": return [unary("sign", a[0])]
    if op == "pow":
        if a[1].op == "const":
            return [mul(a[1], power(a[0], const(a[1].val - 1.0))), ZERO]
        return [mul(a[1], power(a[0], sub(a[1], ONE))), mul(n, unary("log", a[0]))]
    if op == "fmin":
        c = binary("le", a[0], a[1])
        return [c, unary("not", c)]
    if op == "fmax":
        c = binary("le", a[1], a[0])
        return [c, unary("not", c)]
    if op == "atan2":
        den = add(unary("sq", a[0]), unary("sq", a[1]))
        return [div(a[1], den), neg(div(a[0], den))]
    if op == "if_else":
        return [ZERO, if_else_s(a[0], ONE, ZERO), if_else_s(a[0], ZERO, ONE)]
    if op in ("sign", "not", "lt", "le", "eq", "ne", "and", "or"):
        return [ZERO] * len(a)
    raise NotImplementedError(op)


def forward_derivative(outputs: Sequence[Expr], seeds: Dict[int, Expr]) -> List[Expr]:
    """Directional derivative: ``seeds`` maps symbol uid -> tangent expression."""
    dot: Dict[int, Expr] = {}
    for n in topo_order(outputs):
        if n.op == "sym":
            dot[n.uid] = seeds.get(n.uid, ZERO)
        elif n.op == "const":
            dot[n.uid] = ZERO
        else:
            tangents = [dot[c.uid] for c in n.args]
            if all(t is ZERO for t in tangents):
                dot[n.uid] = ZERO
                continue
            acc = ZERO
            for p, t in zip(_partials(n), tangents):
                if t is not ZERO and p is not ZERO:
                    acc = add(acc, mul(p, t))
            dot[n.uid] = acc
    return [dot[o.uid] for o in outputs]


def reverse_gradient(output: Expr, wrt: Sequence[Expr], seed: Expr = ONE) -> List[Expr]:
    """Gradient of one scalar with respect to the symbols in ``wrt`` (reverse sweep)."""
    order = topo_order([output])
    bar: Dict[int, Expr] = {output.uid: seed}
    for n in reversed(order):
        b = bar.get(n.uid)
        if b is None or b is ZERO or not n.args:
            continue
        for p, c in zip(_partials(n), n.args):
            if p is ZERO or c.op == "const":
                continue
            contrib = mul(b, p)
            prev = bar.get(c.uid)
            bar[c.uid] = contrib if prev is None else add(prev, contrib)
    return [bar.get(w.uid, ZERO) for w in wrt]


def jacobian_entries(outputs: Sequence[Expr], wrt: Sequence[Expr]) -> List[List[Expr]]:
    """Dense Jacobian as a list of rows.  Picks forward or reverse by shape."""
    m, n = len(outputs), len(wrt)
    if m == 0 or n == 0:
        return [[] for _ in range(m)]
    if n <= m:
        cols = [forward_derivative(outputs, {w.uid: ONE}) for w in wrt]
        return [[cols[j][i] for j in range(n)] for i in range(m)]
    return [reverse_gradient(o, wrt) for o in outputs]


def hessian_entries(output: Expr, wrt: Sequence[Expr]) -> List[List[Expr]]:
    """Symmetric Hessian of a scalar (forward over reverse)."""
    g = reverse_gradient(output, wrt)
    n = len(wrt)
    H = [[ZERO] * n for _ in range(n)]
    for j, w in enumerate(wrt):
        col = forward_derivative(g, {w.uid: ONE})
        for i in range(n):
            H[i][j] = col[i]
    for i in range(n):  # one representative per symmetric pair (same value analytically)
        for j in range(i + 1, n):
            H[i][j] = H[j][i]
    return H


def op_counts(outputs: Sequence[Expr]) -> Dict[str, int]:
    """Operation histogram of the DAG that produces ``outputs`` (after CSE)."""
    counts: Dict[str, int] = {}
    for n in topo_order(outputs):
        if n.op in ("const", "sym"):
            continue
        counts[n.op] = counts.get(n.op, 0) + 1
    return counts


def flop_count(outputs: Sequence[Expr]) -> int:
    """Every arithmetic / transcendental / select node counts as one FLOP (SURVEY 8d rule)."""
    return int(sum(op_counts(outputs).values()))
