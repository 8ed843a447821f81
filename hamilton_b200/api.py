"""Python mirror of Numeric.Hamilton's API surface over the C ABI (include/hamilton_b200.h).

Same names, argument meaning and error behaviour as the reference module
(src/Numeric/Hamilton.hs:28-70): System, mkSystem, mkSystem' (spelled mkSystem_), Config, Phase,
underlyingPos, pe, momenta, toPhase, keC, lagrangian, velocities, fromPhase, keP, hamiltonian, hamEqs,
stepHam, evolveHam, evolveHam' (evolveHam_), stepHamC, evolveHamC, evolveHamC' (evolveHamC_) —
plus the batched entry points the GPU engine adds.  Everything computes on the GPU through the
C ABI; this module holds no numerics of its own and raises if the library or a device is missing.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import num

_dp = C.POINTER(C.c_double)


def _vec(x, n=None, name="vector"):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError("%s must have %d components, got %d" % (name, n, a.size))
    return a


def _p(a):
    return a.ctypes.data_as(_dp)


@dataclass
class Config:
    """Cfg { cfgPositions, cfgVelocities }  (src/Numeric/Hamilton.hs:103-113)"""
    cfgPositions: np.ndarray
    cfgVelocities: np.ndarray


@dataclass
class Phase:
    """Phs { phsPositions, phsMomenta }  (src/Numeric/Hamilton.hs:133-143)"""
    phsPositions: np.ndarray
    phsMomenta: np.ndarray


Cfg, Phs = Config, Phase


def _make_tape(tape, outs):
    arr = (L.HbOp * max(1, len(tape.ops)))()
    for k, (op, a, b, c) in enumerate(tape.ops):
        arr[k].op, arr[k].a, arr[k].b, arr[k].c = op, a, b, c
    o = (C.c_int32 * len(outs))(*outs)
    return L.HbTape(tape.n_in, len(tape.ops), arr, len(outs), o), (arr, o)


class System:
    """Opaque `System m n` (src/Numeric/Hamilton.hs:160-169): a handle to compiled device code."""

    def __init__(self, handle, tapes=None):
        self._h = C.c_void_p(handle)
        m, n = C.c_int32(), C.c_int32()
        L.check(L.lib().hb_system_dims(self._h, C.byref(m), C.byref(n)))
        self.m, self.n = m.value, n.value
        self.tapes = tapes      # (inertia, f_ops, f_outs, u_ops, u_out, u_on_cartesian) for tape systems

    def __del__(self):
        try:
            if self._h:
                L.lib().hb_system_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def builtin(cls, sid, params=None):
        h = C.c_void_p()
        if params is None:
            L.check(L.lib().hb_system_builtin(int(sid), None, 0, C.byref(h)))
        else:
            p = _vec(params)
            L.check(L.lib().hb_system_builtin(int(sid), _p(p), p.size, C.byref(h)))
        return cls(h.value)

    def source(self):
        need = L.lib().hb_system_source(self._h, None, 0)
        buf = C.create_string_buffer(need)
        L.lib().hb_system_source(self._h, buf, need)
        return buf.value.decode()

    def params(self):
        n = L.lib().hb_system_params(self._h, None, 0)
        buf = np.zeros(max(n, 1))
        L.lib().hb_system_params(self._h, _p(buf), n)
        return buf[:n]

    # ---- batched entry points ---------------------------------------------------------------
    def _io(self, x, name, flags=False):
        """Returns (pointer, memspace, device) for numpy or torch input; data arrays must be float64, flags int32 — the C
        library reads and writes exactly N x d doubles (N int32 flags), so anything else would run past the buffer."""
        if isinstance(x, np.ndarray):
            if x.dtype != (np.int32 if flags else np.float64) or not x.flags["C_CONTIGUOUS"]:
                raise ValueError("%s must be a C-contiguous %s array" % (name, "int32" if flags else "float64"))
            return x.ctypes.data, L.HOST, None
        import torch
        if isinstance(x, torch.Tensor):
            if x.dtype != (torch.int32 if flags else torch.float64) or not x.is_contiguous():
                raise ValueError("%s must be a contiguous %s tensor" % (name, "int32" if flags else "float64"))
            if x.is_cuda:
                return x.data_ptr(), L.DEVICE, x.device
            return x.data_ptr(), L.HOST, None     # CPU tensor (possibly pinned): host memspace
        raise TypeError("%s must be a numpy array or a torch tensor" % name)

    def _shape(self, y, d, layout, name):
        if y.ndim != 2:
            raise ValueError("%s must be 2-D" % name)
        N = y.shape[0] if layout == L.AOS else y.shape[1]
        if (y.shape[1] if layout == L.AOS else y.shape[0]) != d:
            raise ValueError("%s has the wrong number of components for layout" % name)
        return N

    def _alloc_like(self, y, shape, dtype=None):
        if isinstance(y, np.ndarray):
            return np.empty(shape, dtype=dtype or np.float64)
        import torch
        dt = {None: torch.float64, np.int32: torch.int32}[dtype]
        return torch.empty(shape, dtype=dt, device=y.device, pin_memory=(not y.is_cuda and y.is_pinned()))

    def _run(self, y, outs, call, out_shape=None):
        """call(mem, stream) -> status, with the device of `y` current and its torch stream passed through.
        outs = [out, flags?]: `out` must have exactly `out_shape` (default: the shape of y), everything must live in the same
        memory space and, on the GPU, on the same device."""
        ptr, mem, dev = self._io(y, "input")
        want = tuple(y.shape) if out_shape is None else tuple(out_shape)
        for k, o in enumerate(outs):
            if o is None:
                continue
            _, omem, odev = self._io(o, "flags" if k else "output", flags=bool(k))
            if omem != mem or odev != dev:
                raise ValueError("input, output and flags must live in the same memory space (and on the same device)")
            if k == 0 and tuple(o.shape) != want:
                raise ValueError("output has shape %s, expected %s" % (tuple(o.shape), want))
        if mem == L.DEVICE:
            import torch
            with torch.cuda.device(dev):
                L.check(L.lib().hb_set_device(dev.index))
                L.check(call(mem, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        else:
            L.check(call(mem, None))

    @staticmethod
    def _vp(x):
        if x is None:
            return None
        return C.c_void_p(x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr())

    def _flags_ok(self, flags, N):
        if flags is None:
            return
        ok = (flags.dtype == np.int32) if isinstance(flags, np.ndarray) else (str(flags.dtype) == "torch.int32")
        size = flags.size if isinstance(flags, np.ndarray) else flags.numel()
        if not ok or size != N:
            raise ValueError("flags must be int32 with one entry per trajectory")

    def batch_step(self, y, dt, nsteps=1, integ=L.RK4, layout=L.AOS, out=None, flags=None):
        """stepHam over a batch: `nsteps` steps of size dt per trajectory (hb_batch_step)."""
        N = self._shape(y, 2 * self.n, layout, "y")
        if out is None:
            out = self._alloc_like(y, y.shape)
        self._flags_ok(flags, N)
        self._run(y, [out, flags], lambda mem, st: L.lib().hb_batch_step(
            self._h, int(integ), float(dt), int(nsteps), N, int(layout), mem, self._vp(y), self._vp(out), self._vp(flags), st))
        return out

    def batch_ham_eqs(self, y, layout=L.AOS, out=None, flags=None):
        N = self._shape(y, 2 * self.n, layout, "y")
        if out is None:
            out = self._alloc_like(y, y.shape)
        self._flags_ok(flags, N)
        self._run(y, [out, flags], lambda mem, st: L.lib().hb_batch_ham_eqs(
            self._h, N, int(layout), mem, self._vp(y), self._vp(out), self._vp(flags), st))
        return out

    def batch_evolve(self, y0, ts, integ=L.RKF45_GSL, rk4_substeps=1, layout=L.AOS, out=None, flags=None):
        """evolveHam over a batch sharing the grid ts; returns s stacked batches (s, *y0.shape)."""
        N = self._shape(y0, 2 * self.n, layout, "y0")
        ts = _vec(ts)
        if out is None:
            out = self._alloc_like(y0, (ts.size,) + tuple(y0.shape))
        self._flags_ok(flags, N)
        self._run(y0, [out, flags], lambda mem, st: L.lib().hb_batch_evolve(
            self._h, int(integ), int(rk4_substeps), N, int(layout), mem, self._vp(y0), _p(ts), ts.size, self._vp(out),
            self._vp(flags), st), out_shape=(ts.size,) + tuple(y0.shape))
        return out

    def batch_to_phase(self, c, layout=L.AOS, out=None):
        N = self._shape(c, 2 * self.n, layout, "c")
        if out is None:
            out = self._alloc_like(c, c.shape)
        self._run(c, [out], lambda mem, st: L.lib().hb_batch_to_phase(self._h, N, int(layout), mem, self._vp(c), self._vp(out), st))
        return out

    def batch_from_phase(self, y, layout=L.AOS, out=None, flags=None):
        N = self._shape(y, 2 * self.n, layout, "y")
        if out is None:
            out = self._alloc_like(y, y.shape)
        self._flags_ok(flags, N)
        self._run(y, [out, flags], lambda mem, st: L.lib().hb_batch_from_phase(
            self._h, N, int(layout), mem, self._vp(y), self._vp(out), self._vp(flags), st))
        return out

    def batch_energies(self, y, layout=L.AOS, out=None, flags=None):
        """Columns: keP, pe, hamiltonian, lagrangian."""
        N = self._shape(y, 2 * self.n, layout, "y")
        if out is None:
            out = self._alloc_like(y, (N, 4))
        self._flags_ok(flags, N)
        self._run(y, [out, flags], lambda mem, st: L.lib().hb_batch_energies(
            self._h, N, int(layout), mem, self._vp(y), self._vp(out), self._vp(flags), st), out_shape=(N, 4))
        return out

    def batch_underlying_pos(self, q, layout=L.AOS, out=None):
        N = self._shape(q, self.n, layout, "q")
        if out is None:
            out = self._alloc_like(q, (N, self.m) if layout == L.AOS else (self.m, N))
        self._run(q, [out], lambda mem, st: L.lib().hb_batch_underlying_pos(self._h, N, int(layout), mem, self._vp(q), self._vp(out), st),
                  out_shape=(N, self.m) if layout == L.AOS else (self.m, N))
        return out

    def batch_init_random(self, seed, first, N, lo, hi, layout=L.AOS, device=None):
        """Counter-based synthetic initial Phases generated on the device (torch tensor)."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = 2 * self.n
        y = torch.empty((N, d) if layout == L.AOS else (d, N), dtype=torch.float64, device=dev)
        lo, hi = _vec(lo, d, "lo"), _vec(hi, d, "hi")
        with torch.cuda.device(dev):
            L.check(L.lib().hb_set_device(dev.index))
            L.check(L.lib().hb_batch_init_random(self._h, C.c_uint64(seed), int(first), int(N), int(layout), _p(lo), _p(hi),
                                                 C.c_void_p(y.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return y


# ------------------------------------------------------------------------ construction ---------
def _mk(inertia, f, u, n, u_on_cartesian, params=()):
    inertia = _vec(inertia)
    m = inertia.size
    if n is None:
        raise ValueError("n (number of generalized coordinates) is required: Python has no type-level Nat")
    ft, fouts = num.trace(f, n)
    if len(fouts) != m:
        raise ValueError("f returned %d coordinates but inertia has %d" % (len(fouts), m))
    ut, uouts = num.trace(u, m if u_on_cartesian else n)
    if len(uouts) != 1:
        raise ValueError("the potential must return a scalar")
    tf, k1 = _make_tape(ft, fouts)
    tu, k2 = _make_tape(ut, uouts)
    p = _vec(list(params)) if len(params) else None
    h = C.c_void_p()
    L.check(L.lib().hb_system_from_tape(m, n, _p(inertia), C.byref(tf), C.byref(tu), int(bool(u_on_cartesian)),
                                        _p(p) if p is not None else None, 0 if p is None else p.size, C.byref(h)))
    return System(h.value, tapes=(inertia, ft.ops, fouts, ut.ops, uouts[0], bool(u_on_cartesian)))


def mkSystem(inertia, f, u, n=None):
    """mkSystem (src/Numeric/Hamilton.hs:201-225): inertia R m; f: generalized -> Cartesian; u: potential on the
    generalized coordinates.  f and u must be written against hamilton_b200.num (number-type polymorphic)."""
    return _mk(inertia, f, u, n, False)


def mkSystem_(inertia, f, u, n=None):
    """mkSystem' (src/Numeric/Hamilton.hs:238-254): u is a function of the Cartesian coordinates."""
    return _mk(inertia, f, u, n, True)


# ------------------------------------------------------------------ single-trajectory API ------
def underlyingPos(s, q):
    q = _vec(q, s.n, "q"); x = np.empty(s.m)
    L.check(L.lib().hb_underlying_pos(s._h, _p(q), _p(x))); return x


def pe(s, q):
    q = _vec(q, s.n, "q"); u = C.c_double()
    L.check(L.lib().hb_pe(s._h, _p(q), C.byref(u))); return u.value


def momenta(s, c):
    q, v = _vec(c.cfgPositions, s.n), _vec(c.cfgVelocities, s.n); p = np.empty(s.n)
    L.check(L.lib().hb_momenta(s._h, _p(q), _p(v), _p(p))); return p


def toPhase(s, c):
    return Phase(_vec(c.cfgPositions, s.n).copy(), momenta(s, c))


def velocities(s, ph):
    q, p = _vec(ph.phsPositions, s.n), _vec(ph.phsMomenta, s.n); v = np.empty(s.n)
    L.check(L.lib().hb_velocities(s._h, _p(q), _p(p), _p(v))); return v


def fromPhase(s, ph):
    return Config(_vec(ph.phsPositions, s.n).copy(), velocities(s, ph))


def _scalar(fn, s, a, b):
    a, b = _vec(a, s.n), _vec(b, s.n); r = C.c_double()
    L.check(fn(s._h, _p(a), _p(b), C.byref(r))); return r.value


def keC(s, c): return _scalar(L.lib().hb_ke_c, s, c.cfgPositions, c.cfgVelocities)
def lagrangian(s, c): return _scalar(L.lib().hb_lagrangian, s, c.cfgPositions, c.cfgVelocities)
def keP(s, ph): return _scalar(L.lib().hb_ke_p, s, ph.phsPositions, ph.phsMomenta)
def hamiltonian(s, ph): return _scalar(L.lib().hb_hamiltonian, s, ph.phsPositions, ph.phsMomenta)


def hamEqs(s, ph):
    """(dq/dt, dp/dt)  (src/Numeric/Hamilton.hs:370-387)"""
    q, p = _vec(ph.phsPositions, s.n), _vec(ph.phsMomenta, s.n); dq = np.empty(s.n); dp = np.empty(s.n)
    L.check(L.lib().hb_ham_eqs(s._h, _p(q), _p(p), _p(dq), _p(dp))); return dq, dp


def stepHam(r, s, ph):
    """stepHam r s p (src/Numeric/Hamilton.hs:390-402): the reference's adaptive RKF45 solve over (0, r)."""
    q, p = _vec(ph.phsPositions, s.n), _vec(ph.phsMomenta, s.n); qo = np.empty(s.n); po = np.empty(s.n)
    L.check(L.lib().hb_step_ham(s._h, float(r), _p(q), _p(p), _p(qo), _p(po))); return Phase(qo, po)


def evolveHam(s, p0, ts):
    """evolveHam (src/Numeric/Hamilton.hs:433-462); len(ts) >= 2 (the type-level `2 <= s`)."""
    ts = _vec(ts)
    if ts.size < 2:
        raise ValueError("evolveHam needs at least two times (2 <= s)")
    q, p = _vec(p0.phsPositions, s.n), _vec(p0.phsMomenta, s.n); out = np.empty((ts.size, 2 * s.n))
    L.check(L.lib().hb_evolve_ham(s._h, _p(q), _p(p), _p(ts), ts.size, _p(out)))
    return [Phase(r[:s.n].copy(), r[s.n:].copy()) for r in out]


def evolveHam_(s, p0, ts):
    """evolveHam' (src/Numeric/Hamilton.hs:409-429): [] -> []; [x] -> grid [0, x] with the first point dropped."""
    ts = list(ts)
    if not ts:
        return []
    if len(ts) == 1:
        return evolveHam(s, p0, [0.0, ts[0]])[1:]
    return evolveHam(s, p0, ts)


def stepHamC(r, s, c):
    """stepHamC (src/Numeric/Hamilton.hs:505-515)"""
    q, v = _vec(c.cfgPositions, s.n), _vec(c.cfgVelocities, s.n); qo = np.empty(s.n); vo = np.empty(s.n)
    L.check(L.lib().hb_step_ham_c(s._h, float(r), _p(q), _p(v), _p(qo), _p(vo))); return Config(qo, vo)


def evolveHamC(s, c0, ts):
    """evolveHamC (src/Numeric/Hamilton.hs:488-498)"""
    ts = _vec(ts)
    if ts.size < 2:
        raise ValueError("evolveHamC needs at least two times (2 <= s)")
    q, v = _vec(c0.cfgPositions, s.n), _vec(c0.cfgVelocities, s.n); out = np.empty((ts.size, 2 * s.n))
    L.check(L.lib().hb_evolve_ham_c(s._h, _p(q), _p(v), _p(ts), ts.size, _p(out)))
    return [Config(r[:s.n].copy(), r[s.n:].copy()) for r in out]


def evolveHamC_(s, c0, ts):
    """evolveHamC' (src/Numeric/Hamilton.hs:470-480)"""
    ts = list(ts)
    if not ts:
        return []
    if len(ts) == 1:
        return evolveHamC(s, c0, [0.0, ts[0]])[1:]
    return evolveHamC(s, c0, ts)
