"""The reference's example systems (app/Examples.hs:61-183) and the BASELINE.json benchmark
systems, written — like the Haskell originals — as number-type-polymorphic functions, so the same
definitions feed mkSystem (tracers), sympy (oracle/crosscheck.py) and plain floats.

Each `*_def` returns (inertia, f, u, n, u_on_cartesian); each constructor without the suffix
builds the System through mkSystem / mkSystem'.  Builtin ids (hb_builtin) select the ahead-of-time
compiled versions of the same systems."""
import math

from . import num
from .api import System, mkSystem, mkSystem_

PENDULUM, DOUBLE_PENDULUM, ROOM, TWO_BODY, SPRING, BEZIER, TRIPLE_PENDULUM, CHAIN12, SPRING1D = range(9)
BUILTIN_NAMES = ["pendulum", "double_pendulum", "room", "two_body", "spring", "bezier", "triple_pendulum", "chain12", "spring1d"]


def logistic(pos, ht, width):
    """app/Examples.hs:601-605"""
    beta = math.log(0.9 / (1 - 0.9)) / width
    return lambda x: ht / (1 + num.exp(-(beta * (x - pos))))


def pendulum_def():                                   # app/Examples.hs:64-69
    return ([1, 1], lambda q: [num.sin(q[0]), 0.5 - num.cos(q[0])], lambda x: x[1], 1, True)


def double_pendulum_def(m1=1.0, m2=1.0):              # app/Examples.hs:78-89
    def f(q):
        t1, t2 = q
        return [num.sin(t1), 1 - num.cos(t1), num.sin(t1) + num.sin(t2) / 2, 1 - num.cos(t1) - num.cos(t2) / 2]
    return ([m1, m1, m2, m2], f, lambda x: 5 * (m1 * x[1] + m2 * x[3]), 2, True)


def room_def():                                       # app/Examples.hs:99-112
    def u(q):
        x, y = q
        return (2 * y + (1 - logistic(-1, 10, 0.1)(y)) + logistic(1, 10, 0.1)(y)
                + (1 - logistic(-2, 10, 0.1)(x)) + logistic(2, 10, 0.1)(x))
    return ([1, 1], lambda q: [q[0], q[1]], u, 2, False)


def two_body_def(m1=5.0, m2=0.5):                     # app/Examples.hs:123-138
    mT = m1 + m2

    def f(q):
        r, th = q
        r1, r2 = r * (-(m2 / mT)), r * (m1 / mT)
        return [r1 * num.cos(th), r1 * num.sin(th), r2 * num.cos(th), r2 * num.sin(th)]
    return ([m1, m1, m2, m2], f, lambda q: -((m1 * m2) / q[0]), 2, False)


def spring_def(mB=2.0, mW=1.0, k=10.0):               # app/Examples.hs:148-158
    def f(q):
        r, x, th = q
        return [r, r + (1 + x) * num.sin(th), (1 + x) * (-num.cos(th))]

    def u(q):
        r, x, th = q
        return (k * x ** 2.0 / 2 + (1 - logistic(-1.5, 25, 0.1)(r)) + logistic(1.5, 25, 0.1)(r)
                + mB * ((1 + x) * (-num.cos(th))))
    return ([mB, mW, mW], f, u, 3, False)


BEZIER_DEFAULT = [(-1, -1), (-2, 1), (0, 1), (1, -1), (2, 1)]    # app/Examples.hs:350


def bezier_def(points=BEZIER_DEFAULT):                # app/Examples.hs:171-179, bezierCurve :607-627
    npts = len(points) - 1

    def f(q):
        t = q[0]
        bx, by = 0.0, 0.0
        for i, (px, py) in enumerate(points):
            coef = math.comb(npts, i) * (1 - t) ** (npts - i) * t ** i
            bx, by = bx + px * coef, by + py * coef
        return [bx, by]
    return ([1, 1], f, lambda q: (1 - logistic(0, 5, 0.05)(q[0])) + logistic(1, 5, 0.05)(q[0]), 1, False)


def pendulum_chain_def(masses, lengths, g=5.0):
    n = len(masses)

    def f(q):
        out, sx, sy = [], 0.0, 1.0
        for k in range(n):
            sx = sx + lengths[k] * num.sin(q[k])
            sy = sy - lengths[k] * num.cos(q[k])
            out += [sx, sy]
        return out

    def u(x):
        acc = 0.0
        for k in range(n):
            acc = acc + masses[k] * x[2 * k + 1]
        return g * acc
    w = [m for mk in masses for m in (mk, mk)]
    return (w, f, u, n, True)


def triple_pendulum_def(m=(1.0, 1.0, 1.0), l=(1.0, 0.5, 0.5)):
    return pendulum_chain_def(list(m), list(l))


def chain12_def():
    return pendulum_chain_def([1.0] * 12, [1.0] * 12)


def spring1d_def(k=10.0, alpha=0.3):                  # synthetic
    ca, sa = math.cos(alpha), math.sin(alpha)
    return ([1, 1], lambda q: [q[0] * ca, q[0] * sa], lambda q: k * (q[0] * q[0]) / 2, 1, False)


DEFS = [pendulum_def, double_pendulum_def, room_def, two_body_def, spring_def, bezier_def, triple_pendulum_def,
        chain12_def, spring1d_def]


def from_def(d):
    inertia, f, u, n, cart = d
    return (mkSystem_ if cart else mkSystem)(inertia, f, u, n)


def builtin(sid, params=None):
    """Ahead-of-time compiled fixture (hb_system_builtin)."""
    return System.builtin(sid, params)
