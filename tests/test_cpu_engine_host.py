"""The DEVICE engine compiled for the host (tests/host_engine_harness.cpp, HB_HOST_EMU) against the CPU oracle.

The GPU parity suite (-m gpu) is the real gate; this file makes the same code paths — generated hamEqs in both forms,
the closed-form / LDL^T solves, RK4, the GSL-RKF45 stepper, controller and evolve loop, the table-driven sincos and the
Newton reciprocal — checkable on a machine without a GPU.  Tolerance as in the GPU suite: 1e-10 per step; the observed
differences are rounding-level."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import hamilton_b200 as hb
from tests.common import BOXES, maxerr, random_phases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENGINE = os.path.join(ROOT, "hamilton_b200", "csrc", "engine", "hb_engine.cuh")
_dp = C.POINTER(C.c_double)
NAMES = [n for n in BOXES if n != "chain12"]          # chain12 has its own (slower to compile) test below
_cache = {}
_tmpdirs = []


def _p(a):
    return a.ctypes.data_as(_dp)


def harness(name, kind="aot", defines=(), definition=None):
    """Builds (once per session) the host image of the engine + the generated Sys of a built-in system (or of `definition`,
    a hamilton_b200.systems *_def tuple traced like a user system)."""
    key = (name, kind, tuple(defines))
    if key in _cache:
        return _cache[key]
    if kind == "aot" and definition is None:
        s = hb.systems.builtin(BOXES[name][0])
    else:
        os.environ["HB_JIT_SKIP_COMPILE"] = "1"        # symbolic stage only: the host harness compiles the source itself
        try:
            s = hb.systems.from_def(definition if definition is not None else hb.systems.DEFS[BOXES[name][0]]())
        finally:
            del os.environ["HB_JIT_SKIP_COMPILE"]
    holder = tempfile.TemporaryDirectory(prefix="hb_hostemu_")   # removed when the test session's cache is dropped
    _tmpdirs.append(holder)
    tmp = holder.name
    src = os.path.join(tmp, "sys.inc")
    with open(src, "w") as f:
        f.write(s.source())
    sname = re.search(r"struct (\w+) \{", s.source()).group(1)
    so = os.path.join(tmp, "engine_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off"] + ["-D" + d for d in defines] + ['-DENGINE_HEADER="%s"' % ENGINE,
                           '-DSYS_SOURCE="%s"' % src, "-DSYS_NAME=" + sname, os.path.join(ROOT, "tests", "host_engine_harness.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.rcp_fast.restype = C.c_double
    lib.rcp_fast.argtypes = [C.c_double]
    lib.sincos_fast.argtypes = [C.c_double, _dp, _dp]
    prm = np.zeros(64)
    pv = s.params()
    prm[:len(pv)] = pv
    _cache[key] = (lib, prm, s)
    return _cache[key]


@pytest.mark.parametrize("name", NAMES)
def test_engine_ham_eqs_and_maps_on_host(name, oracle_mod):
    lib, prm, s = harness(name)
    o = oracle_mod.OracleSystem.builtin(BOXES[name][0])
    n = o.n
    for y in random_phases(name, 20):
        dy = np.empty(2 * n)
        assert lib.ham_eqs(_p(prm), _p(y), _p(dy)) == 0
        dq, dp = o.ham_eqs(y[:n], y[n:])
        assert maxerr(dy, np.r_[dq, dp]) < 1e-12
        v = y[n:].copy()                                   # use the momenta slot as a velocity vector
        p = np.empty(n); vb = np.empty(n); U = C.c_double()
        assert lib.config_maps(_p(prm), _p(np.ascontiguousarray(y[:n])), _p(v), _p(p), _p(vb), C.byref(U)) == 0
        assert maxerr(p, o.momenta(y[:n], v)) < 1e-12 and maxerr(vb, v) < 1e-11 and abs(U.value - o.pe(y[:n])) < 1e-12 * (1 + abs(U.value))


@pytest.mark.parametrize("name", NAMES)
def test_engine_rk4_and_rkf45_on_host(name, oracle_mod):
    lib, prm, s = harness(name)
    o = oracle_mod.OracleSystem.builtin(BOXES[name][0])
    ys = random_phases(name, 12)
    want, bad = o.batch_step(ys, 0, 0.01, 3)
    assert bad == 0
    for y, w in zip(ys, want):
        got = y.copy()
        assert lib.rk4_steps(_p(prm), _p(got), C.c_double(0.01), 3) == 0
        assert maxerr(got, w) < 1e-10
    for dt, k in ((0.01, 2), (1.0 / 12, 3)):              # the second makes the controller work (and reject, for some)
        want, bad = o.batch_step(ys, 1, dt, k)
        assert bad == 0
        for y, w in zip(ys, want):
            got = y.copy()
            assert lib.rkf45_steps(_p(prm), _p(got), C.c_double(dt), k) == 0
            assert maxerr(got, w) < 1e-9


def test_engine_rkf45_rejections_and_evolve_on_host(oracle_mod):
    """The room at the demo's frame step: rejected sub-steps; and an evolveHam grid with h / FSAL carried across rows."""
    lib, prm, s = harness("room")
    o = oracle_mod.OracleSystem.builtin(2)
    q, p = np.array([-1.0, 0.25]), np.array([np.cos(np.pi / 4), np.sin(np.pi / 4)])
    y = np.r_[q, p]
    rej = 0
    for _ in range(60):
        q, p, st = o.step_ham(1.0 / 12, q, p, stats=True)
        rej += st.rejects
        assert lib.rkf45_steps(_p(prm), _p(y), C.c_double(1.0 / 12), 1) == 0
        assert maxerr(y, np.r_[q, p]) < 1e-9
        y = np.r_[q, p]                                    # teacher forcing
    assert rej > 0
    lib, prm, s = harness("double_pendulum")
    o = oracle_mod.OracleSystem.builtin(1)
    y0 = random_phases("double_pendulum", 1)[0]
    ts = np.linspace(0.0, 1.0, 11)
    out = np.empty((len(ts), 4))
    assert lib.evolve_rkf45(_p(prm), _p(y0), _p(ts), len(ts), _p(out)) == 0
    assert maxerr(out, o.evolve_ham(y0[:2], y0[2:], ts)) < 1e-9


def test_engine_jit_source_on_host(oracle_mod):
    """The tape path (literal parameters, as NVRTC would compile it) through the same harness."""
    lib, prm, s = harness("triple_pendulum", "jit")
    o = oracle_mod.OracleSystem.builtin(BOXES["triple_pendulum"][0])
    for y in random_phases("triple_pendulum", 8):
        dy = np.empty(6)
        assert lib.ham_eqs(_p(prm), _p(y), _p(dy)) == 0
        dq, dp = o.ham_eqs(y[:3], y[3:])
        assert maxerr(dy, np.r_[dq, dp]) < 1e-12


def test_engine_large_system_shared_memory_path_on_host(oracle_mod):
    """chain12 (n = 12): hpre / parked values / LDL^T / hpost and the shared-memory RK4 the kernels use for n >= 8."""
    lib, prm, s = harness("chain12")
    o = oracle_mod.OracleSystem.builtin(BOXES["chain12"][0])
    ys = random_phases("chain12", 4)
    for y in ys:
        dy = np.empty(24)
        assert lib.ham_eqs(_p(prm), _p(y), _p(dy)) == 0
        dq, dp = o.ham_eqs(y[:12], y[12:])
        assert maxerr(dy, np.r_[dq, dp]) < 1e-11
    want, bad = o.batch_step(ys, 0, 0.01, 2)
    for y, w in zip(ys, want):
        got = y.copy()
        assert lib.rk4_steps(_p(prm), _p(got), C.c_double(0.01), 2) == 0
        assert maxerr(got, w) < 1e-10


@pytest.mark.parametrize("defines", [("HB_RK4_STAGE_LOOP=0",), ("HB_SOLVE_COLS=0",)])
def test_large_system_switches_keep_results_on_host(defines, oracle_mod):
    """chain12 with the RK4 stages unrolled instead of looped / the LDL^T and substitutions in row instead of column order:
    the same results (the stage loop bit for bit, the solve order to rounding) against the oracle, RK4 and adaptive."""
    lib, prm, s = harness("chain12", defines=defines)
    ref, _, _ = harness("chain12")
    o = oracle_mod.OracleSystem.builtin(BOXES["chain12"][0])
    ys = random_phases("chain12", 3)
    want, bad = o.batch_step(ys, 0, 0.01, 2)
    want45, bad45 = o.batch_step(ys, 1, 0.01, 1)
    assert bad == 0 and bad45 == 0
    for y, w, w45 in zip(ys, want, want45):
        got, base = y.copy(), y.copy()
        assert lib.rk4_steps(_p(prm), _p(got), C.c_double(0.01), 2) == 0
        assert ref.rk4_steps(_p(prm), _p(base), C.c_double(0.01), 2) == 0
        assert maxerr(got, w) < 1e-10
        if defines == ("HB_RK4_STAGE_LOOP=0",):
            assert np.array_equal(got, base)
        else:
            assert maxerr(got, base) < 1e-13
        got = y.copy()
        assert lib.rkf45_steps(_p(prm), _p(got), C.c_double(0.01), 1) == 0
        assert maxerr(got, w45) < 1e-10


def test_engine_at_the_largest_supported_size_on_host(oracle_mod):
    """n = HB_MAX_N = 16: a 16-link chain (System 32 16) traced as a user system — symbolic stage, generated hpre / hpost, the
    16 x 16 LDL^T and the shared-memory RK4, against the oracle's tape interpreter (dense jets + explicit inverse)."""
    from tests.common import tape_args
    d = hb.systems.pendulum_chain_def([1.0] * 16, [1.0] * 16)
    lib, prm, s = harness("chain16", kind="jit", definition=d)
    assert (s.m, s.n) == (32, 16)
    o = oracle_mod.OracleSystem.from_tape(*tape_args(s))
    rng = np.random.default_rng(16)
    ys = np.c_[rng.uniform(-np.pi, np.pi, size=(3, 16)), rng.uniform(-1, 1, size=(3, 16))]
    for y in ys:
        dy = np.empty(32)
        assert lib.ham_eqs(_p(prm), _p(y), _p(dy)) == 0
        dq, dp = o.ham_eqs(y[:16], y[16:])
        assert maxerr(dy, np.r_[dq, dp]) < 1e-10
    want, bad = o.batch_step(ys, 0, 0.01, 2)
    assert bad == 0
    for y, w in zip(ys, want):
        got = y.copy()
        assert lib.rk4_steps(_p(prm), _p(got), C.c_double(0.01), 2) == 0
        assert maxerr(got, w) < 1e-10


def test_fast_sincos_on_host():
    """hb_sincos<FAST>: 2048-entry table, one-FMA reduction, 1 + 2 polynomial terms.  Error bound 2.5e-16 + 3.9e-17 |x|
    (the reduction constant's rounding = a 0.36-ulp perturbation of the argument); domain |x| < 2^31 pi / 1024 = 6.59e6,
    everything else (huge, inf, nan) flags `oob`."""
    lib, _, _ = harness("pendulum")
    rng = np.random.default_rng(7)
    edge = 2.0 ** 31 * np.pi / 1024
    xs = np.r_[rng.uniform(-np.pi, np.pi, 20000), rng.uniform(-100, 100, 20000), rng.uniform(-1e5, 1e5, 5000),
               rng.uniform(-edge, edge, 5000), rng.uniform(-1e-3, 1e-3, 2000),
               np.arange(-4096, 4097) * (np.pi / 1024), (np.arange(-4096, 4097) + 0.5) * (np.pi / 1024),
               [0.0, -0.0, 99999.9, -99999.9, 1e-300, 0.999 * edge, -0.999 * edge]]
    s, c = C.c_double(), C.c_double()
    worst = 0.0
    for x in xs:
        assert lib.sincos_fast(float(x), C.byref(s), C.byref(c)) == 0, x
        err = max(abs(s.value - np.sin(x)), abs(c.value - np.cos(x)))
        worst = max(worst, err / (2.5e-16 + 4.0e-17 * abs(x)))
    assert worst < 1.0
    near = rng.uniform(-np.pi, np.pi, 20000)     # where every mechanical system of the fixtures lives: < 3.5e-16 absolute
    for x in near:
        lib.sincos_fast(float(x), C.byref(s), C.byref(c))
        assert max(abs(s.value - np.sin(x)), abs(c.value - np.cos(x))) < 3.5e-16
    for x in (1.001 * edge, -1.001 * edge, 1e9, -1e9, 1e300, float("inf"), float("-inf"), float("nan")):
        assert lib.sincos_fast(x, C.byref(s), C.byref(c)) != 0, x


def test_fast_exp_on_host():
    """hb_exp<FAST>: 64-entry 2^(j/64) table, two-term Cody-Waite reduction, degree-5 polynomial: relative error < 2.5e-16
    against a 60-digit reference over the whole domain |x| < 512; everything else (huge, inf, nan) flags `oob`."""
    import mpmath as mp
    mp.mp.dps = 40
    lib, _, _ = harness("pendulum")
    lib.exp_fast.argtypes = [C.c_double, _dp]
    rng = np.random.default_rng(13)
    ln2_64 = np.log(2.0) / 64
    xs = np.r_[rng.uniform(-1, 1, 6000), rng.uniform(-40, 40, 6000), rng.uniform(-511.9, 511.9, 6000), rng.uniform(-1e-8, 1e-8, 500),
               np.arange(-3000, 3001) * ln2_64, (np.arange(-3000, 3001) + 0.5) * ln2_64,   # table nodes and the reduction's break points
               [0.0, -0.0, 1e-300, -1e-300, 511.999, -511.999, np.log(2.0), -np.log(2.0)]]
    e = C.c_double()
    worst = 0.0
    for x in xs:
        assert lib.exp_fast(float(x), C.byref(e)) == 0, x
        ref = mp.exp(mp.mpf(float(x)))
        worst = max(worst, float(abs(mp.mpf(e.value) - ref) / ref))
    assert worst < 2.5e-16, worst
    for x in (512.0, -512.0, 600.0, -1e9, 1e300, float("inf"), float("-inf"), float("nan")):
        assert lib.exp_fast(x, C.byref(e)) != 0, x


def test_fast_reciprocal_on_host():
    """hb_rcp: 20-bit seed (emulated MUFU.RCP64H) + one cubic step: within 1 ulp over the exponent range the solves see."""
    lib, _, _ = harness("pendulum")
    rng = np.random.default_rng(11)
    d = np.exp(rng.uniform(np.log(1e-150), np.log(1e150), 20000)) * rng.choice([-1.0, 1.0], 20000)
    got = np.array([lib.rcp_fast(float(x)) for x in d])
    assert np.max(np.abs(got * d - 1.0)) < 2.5e-16


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 12])
def test_spd_solve_on_host(n):
    """hb_spd_solve<N>: closed forms for N <= 3, unrolled LDL^T beyond; non-SPD input raises HB_FLAG_NOT_SPD."""
    lib, _, _ = harness("pendulum")
    rng = np.random.default_rng(100 + n)
    for _ in range(50):
        G = rng.normal(size=(n, n))
        A = G @ G.T + 0.5 * np.eye(n)
        b = rng.normal(size=n)
        packed = np.array([A[j, k] for j in range(n) for k in range(j + 1)])
        x = np.empty(n)
        assert lib.spd_solve(n, _p(packed), _p(b), _p(x)) == 0
        assert np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-11 * np.linalg.cond(A)
    A = -np.eye(n)
    packed = np.array([A[j, k] for j in range(n) for k in range(j + 1)])
    assert lib.spd_solve(n, _p(packed), _p(np.ones(n)), _p(np.empty(n))) & 1


# ---- whole kernels on the host: the grid-stride bodies, layouts, flags, the slow retry, init_random ---------------------
K_STEP_RK4, K_STEP_RKF45, K_EVOLVE_RK4, K_EVOLVE_RKF45, K_HAM_EQS, K_TO_PHASE, K_FROM_PHASE, K_ENERGIES, K_UPOS = range(9)
AOS, SOA = 0, 1


def run_kernel(lib, prm, kid, inp, out, N, dt=0.0, nsteps=1, layout=AOS, flags=None, ts=None, substeps=1, grid=3, block=128):
    tsp = _p(ts) if ts is not None else None
    fl = flags.ctypes.data_as(C.POINTER(C.c_int32)) if flags is not None else None
    rc = lib.run_kernel(kid, _p(inp), _p(out), fl, tsp, C.c_longlong(N), C.c_double(dt), nsteps, layout, 0 if ts is None else len(ts),
                        substeps, _p(prm), grid, block)
    assert rc == 0


@pytest.mark.parametrize("name", ["double_pendulum", "pendulum", "spring", "chain12"])
@pytest.mark.parametrize("layout", [AOS, SOA])
def test_kernel_bodies_step_on_host(name, layout, oracle_mod):
    """step_rk4 / step_rkf45 kernels, 3 CTAs x 128 emulated threads over a ragged batch (every thread walks several
    trajectories with the prefetch of the next Phase), both layouts, out of place and in place."""
    lib, prm, s = harness(name)
    o = oracle_mod.OracleSystem.builtin(BOXES[name][0])
    N = 70 if name == "chain12" else 1000
    y = random_phases(name, N)
    d = y.shape[1]
    for kid, integ, steps, tol in ((K_STEP_RK4, 0, 2, 1e-10), (K_STEP_RKF45, 1, 1, 1e-9)):
        if name == "chain12" and kid == K_STEP_RKF45:
            continue
        want, bad = o.batch_step(y, integ, 0.01, steps)
        assert bad == 0
        yin = np.ascontiguousarray(y.T) if layout == SOA else y.copy()
        out = np.full_like(yin, np.nan)
        fl = np.zeros(N, np.int32)
        run_kernel(lib, prm, kid, yin, out, N, dt=0.01, nsteps=steps, layout=layout, flags=fl, grid=1 if name == "chain12" else 3)
        got = out.T if layout == SOA else out
        assert maxerr(got, want) < tol * steps and not fl.any()
        run_kernel(lib, prm, kid, yin, yin, N, dt=0.01, nsteps=steps, layout=layout, grid=2)      # in place
        assert np.array_equal(yin, out)


def test_kernel_bodies_other_entry_points_on_host(oracle_mod):
    lib, prm, s = harness("double_pendulum")
    o = oracle_mod.OracleSystem.builtin(1)
    N, n, m = 300, 2, 4
    y = random_phases("double_pendulum", N)
    out = np.empty_like(y)
    run_kernel(lib, prm, K_HAM_EQS, y, out, N)
    for i in (0, 17, 299):
        dq, dp = o.ham_eqs(y[i, :n], y[i, n:])
        assert maxerr(out[i], np.r_[dq, dp]) < 1e-12
    cfg = y.copy()                                          # read as Config [q, v]
    ph = np.empty_like(y)
    run_kernel(lib, prm, K_TO_PHASE, cfg, ph, N)
    back = np.empty_like(y)
    run_kernel(lib, prm, K_FROM_PHASE, ph, back, N)
    assert maxerr(back, cfg) < 1e-11
    for i in (3, 150):
        assert maxerr(ph[i, n:], o.momenta(cfg[i, :n], cfg[i, n:])) < 1e-12
    en = np.empty((N, 4))
    run_kernel(lib, prm, K_ENERGIES, y, en, N)
    for i in (5, 250):
        T, U = o.keP(y[i, :n], y[i, n:]), o.pe(y[i, :n])
        assert maxerr(en[i], np.array([T, U, T + U, T - U])) < 1e-12
    q = np.ascontiguousarray(y[:, :n])
    x = np.empty((N, m))
    run_kernel(lib, prm, K_UPOS, q, x, N)
    assert maxerr(x[7], o.underlying_pos(q[7])) < 1e-13
    # evolve: s rows of the batch, first row = initial state, RK4 sub-steps or GSL-RKF45 carried across rows
    ts = np.linspace(0.0, 0.3, 4)
    rows = np.empty((len(ts),) + y.shape)
    run_kernel(lib, prm, K_EVOLVE_RKF45, y, rows, N, ts=ts)
    assert np.array_equal(rows[0], y)
    for i in (0, 123):
        assert maxerr(rows[:, i, :], o.evolve_ham(y[i, :n], y[i, n:], ts)) < 1e-9
    run_kernel(lib, prm, K_EVOLVE_RK4, y, rows, N, ts=ts, substeps=5)
    want, _ = o.batch_step(y, 0, 0.02, 5)                  # 0.1 per row = 5 sub-steps of 0.02
    assert maxerr(rows[1], want) < 1e-10


def test_kernel_flags_and_slow_retry_on_host(oracle_mod):
    """Singular mass matrices raise HB_FLAG_NOT_SPD for exactly those trajectories (OR-ed into the caller's flags);
    angles outside the fast sincos domain are redone out of line with libm and still match the oracle, in place too."""
    lib, prm, s = harness("two_body")
    y = random_phases("two_body", 200)
    y[5, 0] = 0.0
    y[-1, 0] = 0.0                                          # r = 0
    fl = np.zeros(200, np.int32)
    fl[7] = 64
    out = np.empty_like(y)
    run_kernel(lib, prm, K_STEP_RK4, y, out, 200, dt=0.01, flags=fl)
    assert sorted(np.nonzero(fl)[0].tolist()) == [5, 7, 199] and fl[7] == 64 and fl[5] != 0
    lib, prm, s = harness("double_pendulum")
    o = oracle_mod.OracleSystem.builtin(1)
    y = random_phases("double_pendulum", 64)
    y[3, 0] += 2e5
    y[40, 1] -= 7e7
    want, bad = o.batch_step(y, 0, 0.01, 2)
    assert bad == 0
    buf = y.copy()
    run_kernel(lib, prm, K_STEP_RK4, buf, buf, 64, dt=0.01, nsteps=2, grid=1)
    assert maxerr(buf, want) < 1e-10 * 2


def test_kernel_slow_retry_outside_the_exp_domain_on_host(oracle_mod):
    """room: 25 units outside a wall the logistic's exp argument (beta = 22) leaves hb_exp's domain |x| < 512; the kernel body
    redoes that trajectory out of line with libm's exp and still matches the oracle (in place)."""
    lib, prm, s = harness("room")
    o = oracle_mod.OracleSystem.builtin(2)
    y = random_phases("room", 64)
    y[2, 0] = 25.0
    y[33, 1] = -25.0
    want, bad = o.batch_step(y, 0, 0.01, 2)
    assert bad == 0 and np.all(np.isfinite(want))
    buf = y.copy()
    run_kernel(lib, prm, K_STEP_RK4, buf, buf, 64, dt=0.01, nsteps=2, grid=1)
    assert maxerr(buf, want) < 1e-10 * 2


@pytest.mark.parametrize("layout", [AOS, SOA])
def test_init_random_kernel_on_host_is_bit_identical(layout, oracle_mod):
    lib, prm, s = harness("double_pendulum")
    o = oracle_mod.OracleSystem.builtin(1)
    lo, hi = [np.array(v, dtype=float) for v in BOXES["double_pendulum"][1:]]
    N, first = 777, 1 << 20
    out = np.empty((N, 4) if layout == AOS else (4, N))
    lib.run_init_random(_p(out), C.c_longlong(N), 4, layout, C.c_ulonglong(0x48414D49), C.c_longlong(first), _p(lo), _p(hi), 2, 128)
    want = o.init_random(0x48414D49, first, N, lo, hi)
    assert np.array_equal(out if layout == AOS else out.T, want)


@pytest.mark.parametrize("defines", [("HB_LAYSPEC=0",), ("HB_SOLVE_COLS=0",), ("HB_REG_PREFETCH=0", "HB_ASYNC_STAGE=0"), ("HB_SC_CW2=1",)])
def test_engine_experiment_switches_keep_results_on_host(defines, oracle_mod):
    """The compile-time experiment switches of the engine (profiles/ A/Bs, round-2 candidates) must not change results:
    same kernels, same ragged batch, against the oracle."""
    lib, prm, s = harness("double_pendulum", defines=defines)
    o = oracle_mod.OracleSystem.builtin(1)
    N = 901
    y = random_phases("double_pendulum", N)
    y[11, 0] += 3e5                                        # one trajectory through the slow retry
    for layout in (AOS, SOA):
        want, bad = o.batch_step(y, 0, 0.01, 2)
        assert bad == 0
        yin = np.ascontiguousarray(y.T) if layout == SOA else y.copy()
        out = np.full_like(yin, np.nan)
        run_kernel(lib, prm, K_STEP_RK4, yin, out, N, dt=0.01, nsteps=2, layout=layout, grid=2)
        assert maxerr(out.T if layout == SOA else out, want) < 2e-10
        run_kernel(lib, prm, K_STEP_RK4, yin, yin, N, dt=0.01, nsteps=2, layout=layout, grid=3)
        assert np.array_equal(yin, out)
