"""Number-type-polymorphic math, the Python stand-in for Haskell's `RealFloat a` class.

A system's coordinate map is written once against these functions and then runs on
  * `Tr` tracers (to record the tape that crosses the C ABI — hamilton_b200.mkSystem),
  * plain floats / numpy scalars (ordinary evaluation),
  * sympy expressions (the independent cross-check in oracle/crosscheck.py),
exactly as mkSystem's `forall a. RealFloat a => Vector n a -> Vector m a` argument runs on
`ad`'s number types in the reference (src/Numeric/Hamilton.hs:212-215)."""
import math

# hb_opcode (include/hamilton_b200.h)
(OP_INPUT, OP_CONST, OP_PARAM, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_RECIP, OP_ABS, OP_SIGNUM, OP_SQRT, OP_EXP,
 OP_LOG, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH, OP_ASINH, OP_ACOSH, OP_ATANH,
 OP_POW, OP_POWI, OP_ATAN2) = range(29)


class Tape:
    def __init__(self, n_in):
        self.n_in = n_in
        self.ops = []          # (op, a, b, c)

    def push(self, op, a=0, b=0, c=0.0):
        self.ops.append((op, a, b, float(c)))
        return len(self.ops) - 1

    def inputs(self):
        return [Tr(self, self.push(OP_INPUT, j)) for j in range(self.n_in)]


class Tr:
    """Tracing number: every operation appends a node to its tape."""
    __slots__ = ("t", "id")

    def __init__(self, tape, node):
        self.t, self.id = tape, node

    def _lift(self, x):
        if isinstance(x, Tr):
            if x.t is not self.t:
                raise ValueError("mixing values of two different tapes")
            return x
        return Tr(self.t, self.t.push(OP_CONST, 0, 0, float(x)))

    def _bin(self, op, o, swap=False):
        o = self._lift(o)
        a, b = (o, self) if swap else (self, o)
        return Tr(self.t, self.t.push(op, a.id, b.id))

    def __add__(self, o): return self._bin(OP_ADD, o)
    def __radd__(self, o): return self._bin(OP_ADD, o, True)
    def __sub__(self, o): return self._bin(OP_SUB, o)
    def __rsub__(self, o): return self._bin(OP_SUB, o, True)
    def __mul__(self, o): return self._bin(OP_MUL, o)
    def __rmul__(self, o): return self._bin(OP_MUL, o, True)
    def __truediv__(self, o): return self._bin(OP_DIV, o)
    def __rtruediv__(self, o): return self._bin(OP_DIV, o, True)
    def __neg__(self): return Tr(self.t, self.t.push(OP_NEG, self.id))
    def __pos__(self): return self

    def __pow__(self, e):
        # Haskell: (^) for integer literals, (**) otherwise
        if isinstance(e, int) and not isinstance(e, bool):
            return Tr(self.t, self.t.push(OP_POWI, self.id, 0, float(e)))
        return self._bin(OP_POW, e)

    def __rpow__(self, base): return self._bin(OP_POW, base, True)
    def __abs__(self): return Tr(self.t, self.t.push(OP_ABS, self.id))


def _un(op, name, x):
    if isinstance(x, Tr):
        return Tr(x.t, x.t.push(op, x.id))
    if hasattr(x, "free_symbols") or type(x).__module__.startswith("sympy"):
        import sympy
        return getattr(sympy, {"arcsin": "asin"}.get(name, name))(x)
    return getattr(math, name)(x)


def sin(x): return _un(OP_SIN, "sin", x)
def cos(x): return _un(OP_COS, "cos", x)
def tan(x): return _un(OP_TAN, "tan", x)
def asin(x): return _un(OP_ASIN, "asin", x)
def acos(x): return _un(OP_ACOS, "acos", x)
def atan(x): return _un(OP_ATAN, "atan", x)
def sinh(x): return _un(OP_SINH, "sinh", x)
def cosh(x): return _un(OP_COSH, "cosh", x)
def tanh(x): return _un(OP_TANH, "tanh", x)
def asinh(x): return _un(OP_ASINH, "asinh", x)
def acosh(x): return _un(OP_ACOSH, "acosh", x)
def atanh(x): return _un(OP_ATANH, "atanh", x)
def exp(x): return _un(OP_EXP, "exp", x)
def log(x): return _un(OP_LOG, "log", x)
def sqrt(x): return _un(OP_SQRT, "sqrt", x)


def recip(x):
    if isinstance(x, Tr):
        return Tr(x.t, x.t.push(OP_RECIP, x.id))
    return 1 / x


def atan2(y, x):
    if isinstance(y, Tr):
        return y._bin(OP_ATAN2, x)
    if isinstance(x, Tr):
        return x._bin(OP_ATAN2, y, True)
    if hasattr(y, "free_symbols") or hasattr(x, "free_symbols"):
        import sympy
        return sympy.atan2(y, x)
    return math.atan2(y, x)


def trace(fn, n_in):
    """Runs `fn` on tracers; returns (tape, [output node ids]).  `fn` returns a scalar or a sequence."""
    t = Tape(n_in)
    out = fn(t.inputs())
    scalar = not isinstance(out, (list, tuple))
    outs = [out] if scalar else list(out)
    ids = []
    for o in outs:
        if not isinstance(o, Tr):       # a constant output (e.g. f = const)
            o = Tr(t, t.push(OP_CONST, 0, 0, float(o)))
        ids.append(o.id)
    return t, ids
