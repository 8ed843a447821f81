"""CPU suite, part 1: the oracle itself (no GPU).  The reference ships no tests or golden vectors and cannot be
built here, so the oracle is pinned by (i) an independent sympy/scipy derivation, (ii) physics invariants and closed
forms, (iii) committed golden vectors of its own output (regression), (iv) its two internal routes (native fixtures
vs tape interpreter) agreeing."""
import json
import os

import numpy as np
import pytest

import hamilton_b200 as hb
from tests.common import BOXES, maxerr, random_phases, tape_args

NAMES = list(BOXES)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")


def unhx(v, shape=None):
    a = np.array([float.fromhex(s) for s in v])
    return a.reshape(shape) if shape else a


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


@pytest.mark.parametrize("name", NAMES)
def test_ham_eqs_matches_independent_sympy_derivation(name, oracle_mod):
    from oracle.crosscheck import SympySystem
    sid = BOXES[name][0]
    S, o = SympySystem(hb.systems.DEFS[sid]()), oracle_mod.OracleSystem.builtin(sid)
    n = o.n
    for r in random_phases(name, 6):
        dq, dp = o.ham_eqs(r[:n], r[n:])
        dq2, dp2 = S.ham_eqs(r[:n], r[n:])
        assert maxerr(np.r_[dq, dp], np.r_[dq2, dp2]) < 1e-12
        assert abs(o.hamiltonian(r[:n], r[n:]) - S.hamiltonian(r[:n], r[n:])) < 1e-12 * (1 + abs(S.hamiltonian(r[:n], r[n:])))


@pytest.mark.parametrize("name", ["pendulum", "double_pendulum", "two_body", "spring", "triple_pendulum"])
def test_step_ham_matches_high_order_integration(name, oracle_mod):
    """GSL-semantics RKF45 (eps 1.49e-8) vs DOP853 at 1e-12 on the sympy-derived RHS: agreement at the solver's tolerance."""
    from oracle.crosscheck import SympySystem
    sid = BOXES[name][0]
    S, o = SympySystem(hb.systems.DEFS[sid]()), oracle_mod.OracleSystem.builtin(sid)
    n = o.n
    for r in random_phases(name, 3):
        q, p = o.step_ham(0.1, r[:n], r[n:])
        assert maxerr(np.r_[q, p], S.integrate(r, 0.1)) < 2e-7
        y = o.rk4(r, 0.001, 100)
        assert maxerr(y, S.integrate(r, 0.1)) < 1e-9     # classical RK4, h = 1e-3


def test_double_pendulum_closed_form_mass_matrix(oracle_mod):
    """M = [[m1+m2, m2 cos(t1-t2)/2], [., m2/4]] for app/Examples.hs:82-88 (hand derivation)."""
    m1, m2 = 1.3, 0.7
    o = oracle_mod.OracleSystem.builtin(1, [m1, m2])
    for t1, t2 in [(0.3, 1.1), (-2.0, 0.4), (3.0, -3.0)]:
        J = o.jacobian([t1, t2])
        M = J.T @ np.diag([m1, m1, m2, m2]) @ J
        assert np.allclose(M, [[m1 + m2, 0.5 * m2 * np.cos(t1 - t2)], [0.5 * m2 * np.cos(t1 - t2), 0.25 * m2]], atol=1e-14)
        H = o.hessian([t1, t2])                       # H[j][i][k] = d2 f_i / dq_j dq_k  (tr2, src/Numeric/Hamilton.hs:227-233)
        assert np.allclose(H[0, :, 0], [-np.sin(t1), np.cos(t1), -np.sin(t1), np.cos(t1)], atol=1e-15)
        assert np.allclose(H[1, :, 1], [0, 0, -0.5 * np.sin(t2), 0.5 * np.cos(t2)], atol=1e-15)
        assert np.allclose(H[0, :, 1], 0) and np.allclose(H[1, :, 0], 0)


def test_survey_provisional_markers(oracle_mod):
    """values an independent numpy restatement produced at survey time (sanity markers)."""
    o = oracle_mod.OracleSystem.builtin(1)
    q0 = np.array([np.pi / 2, 0.0]); p0 = o.momenta(q0, [0, 0])
    assert np.allclose(p0, 0) and abs(o.hamiltonian(q0, p0) - 7.5) < 1e-14
    q, p, st = o.step_ham(0.01, q0, p0, stats=True)
    assert abs(q[0] - 1.570546326793870) < 1e-13 and abs(q[1] - 6.250000671335e-08) < 1e-18
    assert abs(p[0] + 9.999999812500e-02) < 1e-13 and abs(p[1] + 1.562499198870e-09) < 1e-19
    assert (st.steps, st.rejects) == (4, 0) and st.rhs_evals == 25       # h = 1e-4, 5e-4, 2.5e-3, remainder; FSAL
    for _ in range(9):
        q, p = o.step_ham(0.01, q, p)
    assert abs(q[0] - 1.545795236257720) < 1e-12 and abs(p[0] + 0.9998125075312316) < 1e-12


def test_invariants(oracle_mod):
    # energy drift of 1000 iterated stepHam 0.01 on the default double pendulum stays at the solver's tolerance
    o = oracle_mod.OracleSystem.builtin(1)
    q, p = np.array([np.pi / 2, 0.0]), np.zeros(2)
    for _ in range(1000):
        q, p = o.step_ham(0.01, q, p)
    assert abs(o.hamiltonian(q, p) - 7.5) < 1e-6
    # two-body: U depends on r only (app/Examples.hs:138) => p_theta conserved exactly by the equations
    tb = oracle_mod.OracleSystem.builtin(3)
    y = random_phases("two_body", 5)
    for r in y:
        dq, dp = tb.ham_eqs(r[:2], r[2:])
        assert abs(dp[1]) < 1e-13
        out = tb.evolve_ham(r[:2], r[2:], np.linspace(0, 2, 5))
        assert np.max(np.abs(out[:, 3] - r[3])) < 1e-6
    # fromPhase . toPhase = id (src/Numeric/Hamilton.hs:284,337)
    for name in NAMES:
        s = oracle_mod.OracleSystem.builtin(BOXES[name][0]); n = s.n
        for r in random_phases(name, 3):
            assert maxerr(s.velocities(r[:n], s.momenta(r[:n], r[n:])), r[n:]) < 1e-10
    # small-angle pendulum: period 2 pi sqrt(l/g) with l = 1, g = 1 (W = (1,1), U = y)
    pe = oracle_mod.OracleSystem.builtin(0)
    ts = np.linspace(0, 2 * np.pi, 3)
    out = pe.evolve_ham([1e-2], [0.0], ts)
    assert abs(out[-1, 0] - 1e-2) < 1e-6 and abs(out[1, 0] + 1e-2) < 1e-6
    # time reversal: step forward, flip momenta, step forward again, flip back
    dpn = oracle_mod.OracleSystem.builtin(1)
    r = random_phases("double_pendulum", 1)[0]
    q, p = dpn.step_ham(0.05, r[:2], r[2:])
    q2, p2 = dpn.step_ham(0.05, q, -p)
    assert maxerr(np.r_[q2, -p2], r) < 1e-6


def test_controller_rejects_at_demo_rate(oracle_mod):
    """dt = 1/12 (app/Examples.hs:415,429): the room's logistic walls force rejected sub-steps."""
    o = oracle_mod.OracleSystem.builtin(2)
    q, p = np.array([-1.0, 0.25]), np.array([np.cos(np.pi / 4), np.sin(np.pi / 4)])
    rej = steps = 0
    for _ in range(120):
        q, p, st = o.step_ham(1.0 / 12, q, p, stats=True)
        rej += st.rejects; steps = max(steps, st.steps)
    assert rej > 0 and steps >= 6
    assert abs(o.hamiltonian(q, p) - o.hamiltonian([-1.0, 0.25], [np.cos(np.pi / 4), np.sin(np.pi / 4)])) < 1e-4


def test_evolve_grid_semantics(oracle_mod):
    """evolveHam carries h and the FSAL derivative across grid points; stepHam restarts at dt/100 (so they differ slightly)."""
    o = oracle_mod.OracleSystem.builtin(1)
    r = random_phases("double_pendulum", 1)[0]
    ts = np.linspace(0, 0.5, 6)
    out, st = o.evolve_ham(r[:2], r[2:], ts, stats=True)
    assert np.array_equal(out[0], r)
    q, p, n_it = r[:2], r[2:], 0
    for _ in range(5):
        q, p, s1 = o.step_ham(0.1, q, p, stats=True); n_it += s1.rhs_evals
    assert st.rhs_evals < n_it                       # fewer RHS evaluations than 5 fresh solves
    assert 0 < maxerr(out[-1], np.r_[q, p]) < 1e-6


@pytest.mark.parametrize("sid,dt,nsteps", [(1, 0.01, 5), (1, 1.0 / 12, 12), (2, 1.0 / 12, 60), (3, 1.0 / 12, 12), (4, 0.05, 10)])
def test_rkf45_two_restatements_agree(sid, dt, nsteps, oracle_mod):
    """The C oracle's GSL-RKF45 (stepper + standard controller + evolve loop) against tests/ref_rkf45.py, a second
    restatement written independently in Python, on the oracle's own hamEqs: same accepted/rejected step sequence, same
    RHS count, states equal to rounding — iterated stepHam (fresh solve per step) and one evolveHam grid (h carried)."""
    from tests import ref_rkf45 as R
    o = oracle_mod.OracleSystem.builtin(sid)
    n = o.n
    name = [k for k, v in BOXES.items() if v[0] == sid][0]
    y = random_phases(name, 3)[2]
    if sid == 2:
        y = np.array([-1.0, 0.25, np.cos(np.pi / 4), np.sin(np.pi / 4)])       # the room: walls force rejections

    def f(v):
        dq, dp = o.ham_eqs(v[:n], v[n:])
        return np.r_[dq, dp]

    q, p = y[:n].copy(), y[n:].copy()
    yy = y.copy()
    rejects = 0
    for _ in range(nsteps):
        q, p, st = o.step_ham(dt, q, p, stats=True)
        rows, st2 = R.ode_solve(f, yy, [0.0, dt])
        yy = rows[-1]
        assert (st.steps, st.rejects, st.rhs_evals) == (st2.steps, st2.rejects, st2.rhs_evals)
        assert maxerr(np.r_[q, p], yy) < 1e-12
        rejects += st.rejects
    if sid == 2:
        assert rejects > 0
    ts = np.linspace(0.0, 6 * dt, 7)
    out, st = o.evolve_ham(y[:n], y[n:], ts, stats=True)
    rows, st2 = R.ode_solve(f, y, ts)
    assert (st.steps, st.rejects, st.rhs_evals) == (st2.steps, st2.rejects, st2.rhs_evals)
    assert maxerr(out, rows) < 1e-11


@pytest.mark.parametrize("name", NAMES)
def test_golden_vectors_regression(name, oracle_mod, gold):
    g = gold[name]; n = g["n"]
    o = oracle_mod.OracleSystem.builtin(BOXES[name][0])
    y = unhx(g["y"], (4, 2 * n))
    assert np.array_equal(y, random_phases(name, 4, seed=777))
    tol = 1e-13
    assert maxerr(o.batch_ham_eqs(y), unhx(g["ham_eqs"], (4, 2 * n))) < tol
    assert maxerr(o.batch_step(y, 0, 0.01, 1)[0], unhx(g["rk4_1"], (4, 2 * n))) < tol
    assert maxerr(o.batch_step(y, 0, 0.01, 10)[0], unhx(g["rk4_10"], (4, 2 * n))) < tol
    assert maxerr(o.batch_step(y, 1, 0.01, 1)[0], unhx(g["step_ham"], (4, 2 * n))) < tol
    assert maxerr(o.batch_step(y, 1, 1.0 / 12, 1)[0], unhx(g["step_ham_demo"], (4, 2 * n))) < 1e-11
    ts = unhx(g["ts"])
    ev = np.array([o.evolve_ham(r[:n], r[n:], ts) for r in y])
    assert maxerr(ev, unhx(g["evolve"], ev.shape)) < 1e-11


@pytest.mark.parametrize("name", [n for n in NAMES if n != "chain12"])
def test_tape_interpreter_agrees_with_native_fixture(name, oracle_mod):
    """The oracle's two routes: native C fixtures (restating app/Examples.hs) vs its tape interpreter fed with the
    tapes recorded from the Python definitions (hamilton_b200.num tracers)."""
    from hamilton_b200 import num
    sid = BOXES[name][0]
    inertia, f, u, n, cart = hb.systems.DEFS[sid]()
    ft, fouts = num.trace(f, n)
    ut, uouts = num.trace(u, len(inertia) if cart else n)
    ot = oracle_mod.OracleSystem.from_tape(len(inertia), n, inertia, ft.ops, fouts, ut.ops, uouts[0], cart)
    on = oracle_mod.OracleSystem.builtin(sid)
    y = random_phases(name, 8)
    assert maxerr(ot.batch_ham_eqs(y), on.batch_ham_eqs(y)) < 1e-13
    assert maxerr(ot.batch_step(y, 1, 0.01, 1)[0], on.batch_step(y, 1, 0.01, 1)[0]) < 1e-13


def test_oracle_batch_threads_deterministic(oracle_mod):
    o = oracle_mod.OracleSystem.builtin(1)
    y = random_phases("double_pendulum", 999)
    a, _ = o.batch_step(y, 0, 0.01, 3, threads=1)
    b, _ = o.batch_step(y, 0, 0.01, 3, threads=4)
    assert np.array_equal(a, b)
    lo, hi = BOXES["double_pendulum"][1:]
    r = o.init_random(0x48414D49, 7, 100, lo, hi)
    assert np.all(r >= np.array(lo)) and np.all(r < np.array(hi))
    assert np.array_equal(r[3:], o.init_random(0x48414D49, 10, 97, lo, hi))      # counter-based: shards line up
