"""Shared test data: per-system sampling boxes for random Phases (away from coordinate singularities)."""
import numpy as np

PI = np.pi
# name -> (builtin id, lo[2n], hi[2n])
BOXES = {
    "pendulum": (0, [-PI, -1], [PI, 1]),
    "double_pendulum": (1, [-PI, -PI, -1, -1], [PI, PI, 1, 1]),
    "room": (2, [-1.5, -0.5, -1, -1], [1.5, 0.5, 1, 1]),
    "two_body": (3, [1, -PI, -1, 1], [3, PI, 1, 5]),
    "spring": (4, [-1, -0.3, -1, -1, -1, -1], [1, 0.3, 1, 1, 1, 1]),
    "bezier": (5, [0.1, -0.5], [0.9, 0.5]),
    "triple_pendulum": (6, [-PI] * 3 + [-1] * 3, [PI] * 3 + [1] * 3),
    "chain12": (7, [-PI] * 12 + [-1] * 12, [PI] * 12 + [1] * 12),
    "spring1d": (8, [-1, -1], [1, 1]),
}
SEED = 0x48414D49


def random_phases(name, N, seed=1234):
    sid, lo, hi = BOXES[name]
    rng = np.random.default_rng(seed + sid)
    lo, hi = np.array(lo, float), np.array(hi, float)
    return lo + (hi - lo) * rng.random((N, lo.size))


def tape_args(system):
    """(m, n, inertia, f_ops, f_outs, u_ops, u_out, u_on_cartesian) of a hamilton_b200 tape System, for the oracle."""
    inertia, f_ops, f_outs, u_ops, u_out, cart = system.tapes
    return system.m, system.n, inertia, f_ops, f_outs, u_ops, u_out, cart


def maxerr(a, b):
    """Largest ABSOLUTE component-wise difference — the tolerance BASELINE.json's north_star states (|dq|, |dp| < 1e-10)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


def relerr(a, b):
    """abs / (1 + |ref|): only for values too large for an absolute 1e-10 to be representable (|q| ~ 1e8: one ulp is 1e-8)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (1.0 + np.abs(b)))) if a.size else 0.0
