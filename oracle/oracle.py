"""ctypes binding of the CPU oracle (oracle/hamilton_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under hamilton_b200/ does.  PARITY UNPINNED (see the C file's header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# hb_builtin ids (include/hamilton_b200.h)
PENDULUM, DOUBLE_PENDULUM, ROOM, TWO_BODY, SPRING, BEZIER, TRIPLE_PENDULUM, CHAIN12, SPRING1D = range(9)


class HbOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double)]


class HbTape(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_ops", C.c_int32), ("ops", C.POINTER(HbOp)),
                ("n_out", C.c_int32), ("outs", C.POINTER(C.c_int32))]


class HoStats(C.Structure):
    _fields_ = [("rhs_evals", C.c_long), ("steps", C.c_long), ("rejects", C.c_long)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "hamilton_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        vp = C.c_void_p
        L.ho_builtin.restype = vp
        L.ho_builtin.argtypes = [C.c_int, dp, C.c_int]
        L.ho_from_tape.restype = vp
        L.ho_from_tape.argtypes = [C.c_int, C.c_int, dp, C.POINTER(HbTape), C.POINTER(HbTape), C.c_int, dp, C.c_int]
        L.ho_free.argtypes = [vp]
        L.ho_dims.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ho_underlying_pos.argtypes = [vp, dp, dp]
        L.ho_pe.restype = C.c_double
        L.ho_pe.argtypes = [vp, dp]
        for nm in ("ho_jacobian", "ho_hessian", "ho_potential_grad"):
            getattr(L, nm).argtypes = [vp, dp, dp]
        L.ho_momenta.argtypes = [vp, dp, dp, dp]
        L.ho_velocities.argtypes = [vp, dp, dp, dp]
        for nm in ("ho_keC", "ho_keP", "ho_lagrangian", "ho_hamiltonian"):
            getattr(L, nm).restype = C.c_double
            getattr(L, nm).argtypes = [vp, dp, dp]
        L.ho_ham_eqs.argtypes = [vp, dp, dp, dp, dp]
        L.ho_rk4.argtypes = [vp, C.c_double, C.c_int, dp]
        L.ho_evolve_ham.argtypes = [vp, dp, dp, dp, C.c_int, dp, C.POINTER(HoStats)]
        L.ho_step_ham.argtypes = [vp, C.c_double, dp, dp, dp, dp, C.POINTER(HoStats)]
        L.ho_batch_step.restype = C.c_long
        L.ho_batch_step.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_long, dp, C.c_int]
        L.ho_batch_ham_eqs.argtypes = [vp, C.c_long, dp, dp, C.c_int]
        L.ho_init_random.argtypes = [vp, C.c_uint64, C.c_long, C.c_long, dp, dp, dp]
        L.ho_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def make_tape(ops, outs, n_in):
    """ops: list of (op, a, b, c); returns (HbTape, keepalive)."""
    arr = (HbOp * max(1, len(ops)))()
    for k, (op, a, b, c) in enumerate(ops):
        arr[k].op, arr[k].a, arr[k].b, arr[k].c = int(op), int(a), int(b), float(c)
    o = (C.c_int32 * len(outs))(*[int(x) for x in outs])
    t = HbTape(int(n_in), len(ops), arr, len(outs), o)
    return t, (arr, o)


class OracleSystem:
    """The oracle's `System m n` (src/Numeric/Hamilton.hs:160-169)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle: could not create system")
        self._h = handle
        m, n = C.c_int(), C.c_int()
        lib().ho_dims(self._h, C.byref(m), C.byref(n))
        self.m, self.n = m.value, n.value

    @classmethod
    def builtin(cls, sid, params=None):
        if params is None:
            return cls(lib().ho_builtin(int(sid), None, 0))
        p, pp = _d(params)
        return cls(lib().ho_builtin(int(sid), pp, len(p)))

    @classmethod
    def from_tape(cls, m, n, inertia, f_ops, f_outs, u_ops, u_out, u_on_cartesian, params=()):
        ft, k1 = make_tape(f_ops, f_outs, n)
        ut, k2 = make_tape(u_ops, [u_out], m if u_on_cartesian else n)
        w, wp = _d(inertia)
        p, pp = _d(list(params) if len(params) else [0.0])
        h = lib().ho_from_tape(m, n, wp, C.byref(ft), C.byref(ut), int(bool(u_on_cartesian)), pp, len(params))
        return cls(h)

    def __del__(self):
        try:
            lib().ho_free(self._h)
        except Exception:
            pass

    # --- single-trajectory API -------------------------------------------------------------
    def underlying_pos(self, q):
        q, qp = _d(q); x = np.empty(self.m); lib().ho_underlying_pos(self._h, qp, x.ctypes.data_as(C.POINTER(C.c_double))); return x

    def pe(self, q):
        q, qp = _d(q); return lib().ho_pe(self._h, qp)

    def jacobian(self, q):
        q, qp = _d(q); J = np.empty((self.m, self.n)); lib().ho_jacobian(self._h, qp, J.ctypes.data_as(C.POINTER(C.c_double))); return J

    def hessian(self, q):
        q, qp = _d(q); H = np.empty((self.n, self.m, self.n)); lib().ho_hessian(self._h, qp, H.ctypes.data_as(C.POINTER(C.c_double))); return H

    def potential_grad(self, q):
        q, qp = _d(q); g = np.empty(self.n); lib().ho_potential_grad(self._h, qp, g.ctypes.data_as(C.POINTER(C.c_double))); return g

    def momenta(self, q, v):
        q, qp = _d(q); v, vp = _d(v); p = np.empty(self.n); lib().ho_momenta(self._h, qp, vp, p.ctypes.data_as(C.POINTER(C.c_double))); return p

    def velocities(self, q, p):
        q, qp = _d(q); p, pp = _d(p); v = np.empty(self.n)
        if lib().ho_velocities(self._h, qp, pp, v.ctypes.data_as(C.POINTER(C.c_double))):
            raise ArithmeticError("oracle: singular mass matrix")
        return v

    def _scalar(self, name, a, b):
        a, ap = _d(a); b, bp = _d(b); return getattr(lib(), name)(self._h, ap, bp)

    def keC(self, q, v): return self._scalar("ho_keC", q, v)
    def keP(self, q, p): return self._scalar("ho_keP", q, p)
    def lagrangian(self, q, v): return self._scalar("ho_lagrangian", q, v)
    def hamiltonian(self, q, p): return self._scalar("ho_hamiltonian", q, p)

    def ham_eqs(self, q, p):
        q, qp = _d(q); p, pp = _d(p); dq = np.empty(self.n); dp = np.empty(self.n)
        rc = lib().ho_ham_eqs(self._h, qp, pp, dq.ctypes.data_as(C.POINTER(C.c_double)), dp.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise ArithmeticError("oracle: hamEqs failed rc=%d" % rc)
        return dq, dp

    def rk4(self, y, dt, nsteps=1):
        y, yp = _d(np.array(y, dtype=np.float64)); lib().ho_rk4(self._h, dt, nsteps, yp); return y

    def step_ham(self, r, q, p, stats=False):
        q, qp = _d(q); p, pp = _d(p); qo = np.empty(self.n); po = np.empty(self.n); st = HoStats()
        rc = lib().ho_step_ham(self._h, r, qp, pp, qo.ctypes.data_as(C.POINTER(C.c_double)), po.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st))
        if rc:
            raise ArithmeticError("oracle: stepHam failed rc=%d" % rc)
        return (qo, po, st) if stats else (qo, po)

    def evolve_ham(self, q0, p0, ts, stats=False):
        q0, qp = _d(q0); p0, pp = _d(p0); ts, tp = _d(ts); out = np.empty((len(ts), 2 * self.n)); st = HoStats()
        rc = lib().ho_evolve_ham(self._h, qp, pp, tp, len(ts), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st))
        if rc:
            raise ArithmeticError("oracle: evolveHam failed rc=%d" % rc)
        return (out, st) if stats else out

    # --- batch (AOS N x 2n) ------------------------------------------------------------------
    def batch_step(self, y, integ, dt, nsteps=1, threads=1):
        y = np.array(y, dtype=np.float64, order="C"); N = y.shape[0]
        bad = lib().ho_batch_step(self._h, int(integ), dt, nsteps, N, y.ctypes.data_as(C.POINTER(C.c_double)), threads)
        return y, bad

    def batch_ham_eqs(self, y, threads=1):
        y, yp = _d(y); dy = np.empty_like(y)
        lib().ho_batch_ham_eqs(self._h, y.shape[0], yp, dy.ctypes.data_as(C.POINTER(C.c_double)), threads); return dy

    def init_random(self, seed, first, N, lo, hi):
        lo, lp = _d(lo); hi, hp = _d(hi); y = np.empty((N, 2 * self.n))
        lib().ho_init_random(self._h, C.c_uint64(seed), first, N, lp, hp, y.ctypes.data_as(C.POINTER(C.c_double))); return y


def max_threads():
    return lib().ho_max_threads()
