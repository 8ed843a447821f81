"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerance: |dq|, |dp| < 1e-10 per step (BASELINE.json north_star), ABSOLUTE component-wise error (tests.common.maxerr); the
single exception is the out-of-domain test, whose angles of 1e5..1e9 cannot even be represented to 1e-10 (relerr there)."""
import numpy as np
import pytest

import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES, SEED, maxerr, random_phases, relerr, tape_args

pytestmark = pytest.mark.gpu
TOL = 1e-10
NAMES = list(BOXES)
SMALL = [n for n in NAMES if n != "chain12"]


def systems_for(name, O):
    sid = BOXES[name][0]
    aot = hb.systems.builtin(sid)
    return aot, O.OracleSystem.builtin(sid)


@pytest.fixture(scope="module")
def jit_cache():
    return {}


def jit_system(name, cache):
    if name not in cache:
        cache[name] = hb.systems.from_def(hb.systems.DEFS[BOXES[name][0]]())
    return cache[name]


@pytest.mark.parametrize("name", NAMES)
def test_ham_eqs_aot_vs_oracle(name, oracle_mod):
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 257)
    assert maxerr(g.batch_ham_eqs(y), o.batch_ham_eqs(y)) < TOL


@pytest.mark.parametrize("name", SMALL)
def test_ham_eqs_jit_vs_oracle_tape(name, oracle_mod, jit_cache):
    """Tape systems: NVRTC-compiled product vs the oracle's tape interpreter AND its native fixture."""
    g = jit_system(name, jit_cache)
    m, n, w, fo, fouts, uo, uout, cart = tape_args(g)
    ot = oracle_mod.OracleSystem.from_tape(m, n, w, fo, fouts, uo, uout, cart)
    on = oracle_mod.OracleSystem.builtin(BOXES[name][0])
    y = random_phases(name, 129)
    dg = g.batch_ham_eqs(y)
    assert maxerr(dg, ot.batch_ham_eqs(y)) < TOL
    assert maxerr(dg, on.batch_ham_eqs(y)) < TOL


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("nsteps", [1, 10])
def test_rk4_step_vs_oracle(name, nsteps, oracle_mod):
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 64 if name == "chain12" else 300)
    yo, bad = o.batch_step(y, 0, 0.01, nsteps)
    assert bad == 0
    fl = np.zeros(len(y), np.int32)
    yg = g.batch_step(y, 0.01, nsteps, integ=L.RK4, flags=fl)
    assert not fl.any()
    assert maxerr(yg, yo) < TOL * nsteps


@pytest.mark.parametrize("name", NAMES)
def test_step_ham_rkf45_vs_oracle(name, oracle_mod):
    """Reference semantics: one `stepHam 0.01` = fresh GSL-RKF45 solve over (0, 0.01)."""
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 32 if name == "chain12" else 200)
    yo, bad = o.batch_step(y, 1, 0.01, 1)
    assert bad == 0
    yg = g.batch_step(y, 0.01, 1, integ=L.RKF45_GSL)
    assert maxerr(yg, yo) < TOL


@pytest.mark.parametrize("name", ["double_pendulum", "two_body", "room"])
def test_step_ham_demo_rate_with_rejections(name, oracle_mod):
    """dt = 1/12 (the demo's frame step, app/Examples.hs:415,429) makes the controller reject steps."""
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 100)
    yo, bad = o.batch_step(y, 1, 1.0 / 12, 3)
    yg = g.batch_step(y, 1.0 / 12, 3, integ=L.RKF45_GSL)
    assert bad == 0
    assert maxerr(yg, yo) < 1e-9   # 3 chained adaptive solves; step sequences identical, rounding differs


@pytest.mark.parametrize("layout", [L.AOS, L.SOA])
@pytest.mark.parametrize("mem", ["host", "device"])
def test_layouts_and_memspaces(layout, mem, oracle_mod):
    import torch
    g, o = systems_for("double_pendulum", oracle_mod)
    y = random_phases("double_pendulum", 1000)
    yo, _ = o.batch_step(y, 0, 0.01, 2)
    yin = np.ascontiguousarray(y.T) if layout == L.SOA else y
    if mem == "device":
        t = torch.from_numpy(yin).cuda()
        out = g.batch_step(t, 0.01, 2, layout=layout)
        torch.cuda.synchronize()
        out = out.cpu().numpy()
    else:
        out = g.batch_step(yin, 0.01, 2, layout=layout)
    if layout == L.SOA:
        out = out.T
    assert maxerr(out, yo) < TOL


def test_config1_anchor_teacher_forced(oracle_mod):
    """BASELINE config 1: double pendulum, Cfg (pi/2, 0) (0, 0), 1000 x stepHam 0.01, compared every step
    from the oracle's state (teacher forcing), for both integrators."""
    g, o = systems_for("double_pendulum", oracle_mod)
    q, p = np.array([np.pi / 2, 0.0]), o.momenta([np.pi / 2, 0.0], [0.0, 0.0])
    states = [np.r_[q, p]]
    for _ in range(1000):
        q, p = o.step_ham(0.01, q, p)
        states.append(np.r_[q, p])
    S = np.array(states)
    got = g.batch_step(S[:-1].copy(), 0.01, 1, integ=L.RKF45_GSL)
    assert np.max(np.abs(got - S[1:])) < TOL
    rk4_o, _ = o.batch_step(S[:-1].copy(), 0, 0.01, 1)
    rk4_g = g.batch_step(S[:-1].copy(), 0.01, 1, integ=L.RK4)
    assert np.max(np.abs(rk4_g - rk4_o)) < TOL


@pytest.mark.parametrize("name", ["double_pendulum", "two_body", "spring"])
def test_evolve_grid_vs_oracle(name, oracle_mod):
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 16)
    ts = np.linspace(0.0, 1.0, 11)
    got = g.batch_evolve(y, ts, integ=L.RKF45_GSL)     # (s, N, 2n)
    n = o.n
    for i in range(len(y)):
        ref = o.evolve_ham(y[i, :n], y[i, n:], ts)
        assert maxerr(got[:, i, :], ref) < 1e-9
    assert np.array_equal(got[0], y)                    # first row is the initial state


@pytest.mark.parametrize("name", NAMES)
def test_config_phase_maps_and_energies(name, oracle_mod):
    g, o = systems_for(name, oracle_mod)
    y = random_phases(name, 50)
    n = o.n
    c = g.batch_from_phase(y)
    vo = np.array([o.velocities(r[:n], r[n:]) for r in y])
    assert maxerr(c[:, n:], vo) < TOL
    back = g.batch_to_phase(c)
    assert maxerr(back, y) < 1e-9                       # fromPhase . toPhase = id
    po = np.array([o.momenta(r[:n], r[n:]) for r in c])
    assert maxerr(back[:, n:], po) < 1e-9
    e = g.batch_energies(y)
    eo = np.array([[o.keP(r[:n], r[n:]), o.pe(r[:n]), o.hamiltonian(r[:n], r[n:]), 0.0] for r in y])
    assert maxerr(e[:, :3], eo[:, :3]) < TOL
    assert maxerr(e[:, 3], eo[:, 0] - eo[:, 1]) < TOL
    x = g.batch_underlying_pos(np.ascontiguousarray(y[:, :n]))
    xo = np.array([o.underlying_pos(r[:n]) for r in y])
    assert maxerr(x, xo) < TOL


def test_single_trajectory_api_matches_oracle(oracle_mod):
    """The Haskell-shaped calls (README.md:124-165 usage) on one Phase."""
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM, [1.0, 2.0])
    o = oracle_mod.OracleSystem.builtin(1, [1.0, 2.0])
    c0 = hb.Cfg([1.0, 0.0], [0.0, 0.5])
    ph = hb.toPhase(s, c0)
    assert maxerr(ph.phsMomenta, o.momenta(c0.cfgPositions, c0.cfgVelocities)) < TOL
    assert abs(hb.keC(s, c0) - o.keC(c0.cfgPositions, c0.cfgVelocities)) < TOL
    assert abs(hb.lagrangian(s, c0) - o.lagrangian(c0.cfgPositions, c0.cfgVelocities)) < TOL
    assert abs(hb.hamiltonian(s, ph) - o.hamiltonian(ph.phsPositions, ph.phsMomenta)) < TOL
    assert abs(hb.pe(s, ph.phsPositions) - o.pe(ph.phsPositions)) < TOL
    dq, dp = hb.hamEqs(s, ph)
    dqo, dpo = o.ham_eqs(ph.phsPositions, ph.phsMomenta)
    assert maxerr(dq, dqo) < TOL and maxerr(dp, dpo) < TOL
    p1 = hb.stepHam(0.1, s, ph)
    qo, po = o.step_ham(0.1, ph.phsPositions, ph.phsMomenta)
    assert maxerr(np.r_[p1.phsPositions, p1.phsMomenta], np.r_[qo, po]) < TOL
    ts = np.arange(0, 1.05, 0.1)
    ev = hb.evolveHam(s, ph, ts)
    ref = o.evolve_ham(ph.phsPositions, ph.phsMomenta, ts)
    assert maxerr(np.array([np.r_[e.phsPositions, e.phsMomenta] for e in ev]), ref) < 1e-9
    assert hb.evolveHam_(s, ph, []) == []
    one = hb.evolveHam_(s, ph, [0.1])
    assert len(one) == 1 and maxerr(one[0].phsPositions, p1.phsPositions) < TOL
    c1 = hb.stepHamC(0.1, s, c0)
    assert maxerr(c1.cfgVelocities, o.velocities(qo, po)) < 1e-9


def test_flags_on_singular_mass_matrix():
    """two-body at r = 0 has J^T W J singular: reference `inv` would throw; we flag, never abort."""
    g = hb.systems.builtin(hb.systems.TWO_BODY)
    y = np.array([[0.0, 0.3, 0.1, 1.0], [2.0, 0.3, 0.1, 1.0]])
    fl = np.zeros(2, np.int32)
    g.batch_ham_eqs(y, flags=fl)
    assert fl[0] != 0 and fl[1] == 0
    with pytest.raises(hb.NumericError):
        hb.hamEqs(g, hb.Phase([0.0, 0.3], [0.1, 1.0]))


def test_init_random_matches_oracle(oracle_mod):
    import torch
    g, o = systems_for("double_pendulum", oracle_mod)
    lo, hi = BOXES["double_pendulum"][1:]
    y = g.batch_init_random(SEED, 5, 1000, lo, hi)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), o.init_random(SEED, 5, 1000, lo, hi))
    ys = g.batch_init_random(SEED, 5, 1000, lo, hi, layout=L.SOA)
    assert np.array_equal(ys.cpu().numpy().T, y.cpu().numpy())


@pytest.mark.parametrize("name", NAMES)
def test_gpu_vs_committed_golden_vectors(name):
    """The CUDA path against tests/golden/oracle_vectors.json (oracle output frozen by oracle/make_golden.py)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")) as f:
        g = json.load(f)[name]
    n = g["n"]
    un = lambda v, shape: np.array([float.fromhex(s) for s in v]).reshape(shape)   # noqa: E731
    y = un(g["y"], (4, 2 * n))
    s = hb.systems.builtin(BOXES[name][0])
    assert maxerr(s.batch_ham_eqs(y), un(g["ham_eqs"], (4, 2 * n))) < TOL
    assert maxerr(s.batch_step(y, 0.01, 1, integ=L.RK4), un(g["rk4_1"], (4, 2 * n))) < TOL
    assert maxerr(s.batch_step(y, 0.01, 10, integ=L.RK4), un(g["rk4_10"], (4, 2 * n))) < 10 * TOL
    assert maxerr(s.batch_step(y, 0.01, 1, integ=L.RKF45_GSL), un(g["step_ham"], (4, 2 * n))) < TOL
    assert maxerr(s.batch_step(y, 1.0 / 12, 1, integ=L.RKF45_GSL), un(g["step_ham_demo"], (4, 2 * n))) < 1e-9
    ts = un(g["ts"], (-1,))
    ev = s.batch_evolve(y, ts, integ=L.RKF45_GSL)                   # (s, 4, 2n)
    assert maxerr(np.transpose(ev, (1, 0, 2)), un(g["evolve"], (4, len(ts), 2 * n))) < 1e-9
    e = s.batch_energies(y)
    assert maxerr(e[:, :3], un(g["energies"], (4, 3))) < TOL
    assert maxerr(s.batch_underlying_pos(np.ascontiguousarray(y[:, :n])), un(g["upos"], (4, g["m"]))) < TOL


def test_large_batch_properties_at_full_size():
    """BASELINE configs[1] at full size (1,048,576 trajectories): size-independent properties instead of an oracle run —
    energy conservation of RK4 (O(dt^4) per unit time), time reversal, and agreement of the AOS and SOA paths."""
    import torch
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    lo, hi = BOXES["double_pendulum"][1:]
    N = 1 << 20
    y0 = s.batch_init_random(SEED, 0, N, lo, hi)
    e0 = s.batch_energies(y0)[:, 2]
    y1 = s.batch_step(y0, 0.001, 100, integ=L.RK4)
    e1 = s.batch_energies(y1)[:, 2]
    assert float((e1 - e0).abs().max()) < 1e-7
    back = y1.clone(); back[:, 2:] *= -1
    y2 = s.batch_step(back, 0.001, 100, integ=L.RK4); y2[:, 2:] *= -1
    assert float((y2 - y0).abs().max()) < 1e-7
    soa = s.batch_step(y0.t().contiguous(), 0.001, 100, integ=L.RK4, layout=L.SOA)
    assert torch.equal(soa.t().contiguous(), y1)
    fl = torch.zeros(N, dtype=torch.int32, device=y0.device)
    s.batch_step(y0, 0.01, 1, integ=L.RKF45_GSL, flags=fl)
    assert int(fl.sum()) == 0


@pytest.mark.parametrize("name,log2n", [("pendulum", 21), ("two_body", 21), ("spring1d", 21), ("triple_pendulum", 20), ("chain12", 18)])
def test_other_baseline_configs_at_full_size(name, log2n):
    """BASELINE configs[2..4] at their full per-GPU sizes (2 x 2,097,152 mixed (2,1)/(4,2); 1,048,576 triple pendulums per
    GPU; 262,144 twelve-link chains): energy drift of RK4, time reversal, AOS/SOA agreement, no failure flags, and
    bit-identity with the same trajectories stepped as a small batch (results must not depend on the batch size,
    grid shape or on which resident CTA a trajectory lands in)."""
    import torch
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid)
    N = 1 << log2n
    steps, dt = (20, 0.001) if name == "chain12" else (50, 0.001)
    y0 = s.batch_init_random(SEED + 11, 0, N, lo, hi)
    fl = torch.zeros(N, dtype=torch.int32, device=y0.device)
    e0 = s.batch_energies(y0)[:, 2]
    y1 = s.batch_step(y0, dt, steps, integ=L.RK4, flags=fl)
    assert int(fl.sum()) == 0
    e1 = s.batch_energies(y1)[:, 2]
    scale = 1.0 + e0.abs()
    assert float(((e1 - e0).abs() / scale).max()) < (1e-6 if name == "chain12" else 1e-7)
    n = s.n
    back = y1.clone(); back[:, n:] *= -1
    y2 = s.batch_step(back, dt, steps, integ=L.RK4); y2[:, n:] *= -1
    assert float((y2 - y0).abs().max()) < 1e-6
    soa = s.batch_step(y0.t().contiguous(), dt, steps, integ=L.RK4, layout=L.SOA)
    assert torch.equal(soa.t().contiguous(), y1)
    idx = torch.tensor([0, 1, 31, 32, 12345, N // 2 + 7, N - 129, N - 1], device=y0.device)
    small = s.batch_step(y0[idx].contiguous(), dt, steps, integ=L.RK4)
    assert torch.equal(small, y1[idx])


@pytest.mark.parametrize("env", [{}, {"HB_HOST_DIRECT": "2"}, {"HB_HOST_DIRECT": "0"}, {"HB_HOST_DIRECT": "0", "HB_HOST_GRAPH": "0"}],
                         ids=["zero_copy", "hybrid", "staged_graph", "staged_streams"])
def test_host_paths_match_device_path(env):
    """The HB_MEM_HOST data paths (picked per process by environment, so each runs in its own interpreter): zero-copy
    kernel on page-locked buffers, copy-engine upload + kernels storing straight to the host, chunked staging replayed
    from a cached CUDA graph, chunked staging submitted to three streams.  tests/host_paths_check.py compares each bit for bit with the device-pointer path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "host_paths_check.py")], cwd=root, env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "host paths ok" in r.stdout


def test_out_of_domain_angles_take_the_slow_path(oracle_mod):
    """|q| >= 6.6e6 leaves the fast sincos domain (2e5 and 1e5 stay inside it): those trajectories are redone out of line with libdevice math.
    Mixed in one warp with ordinary trajectories; also in-place (y_out == y_in) so the retry must re-read intact input."""
    g, o = systems_for("double_pendulum", oracle_mod)
    y = random_phases("double_pendulum", 64)
    y[3, 0] += 2.0e5; y[17, 1] -= 7.5e8; y[40, 0] = 1.0e5
    yo, bad = o.batch_step(y, 0, 0.01, 3)
    assert bad == 0
    got = g.batch_step(y, 0.01, 3, integ=L.RK4)
    assert relerr(got, yo) < 1e-9            # q ~ 7.5e8: one ulp of q is 1.2e-7, an absolute 1e-10 is not representable
    buf = y.copy()
    g.batch_step(buf, 0.01, 3, integ=L.RK4, out=buf)
    assert np.array_equal(buf, got)
    assert relerr(g.batch_step(y, 0.01, 1, integ=L.RKF45_GSL), o.batch_step(y, 1, 0.01, 1)[0]) < 1e-9
    assert relerr(g.batch_ham_eqs(y), o.batch_ham_eqs(y)) < 1e-9
    e = g.batch_energies(y)
    assert maxerr(e[:, 2], [o.hamiltonian(r[:2], r[2:]) for r in y]) < 1e-9


def test_out_of_domain_exp_takes_the_slow_path(oracle_mod):
    """The room's logistic walls (app/Examples.hs:601-605) evaluate exp(-beta (x - pos)) with beta = 22: 25 units outside the
    walls the argument leaves the table-driven hb_exp's domain (|x| < 512) and the trajectory is redone out of line with
    libdevice's exp.  Mixed into warps of ordinary trajectories; RK4, the adaptive stepper and hamEqs; in place too."""
    g, o = systems_for("room", oracle_mod)
    y = random_phases("room", 96)
    y[2, 0] = 25.0; y[33, 1] = -25.0; y[64, 0] = -24.5; y[64, 1] = 26.0
    yo, bad = o.batch_step(y, 0, 0.01, 3)
    assert bad == 0 and np.all(np.isfinite(yo))
    got = g.batch_step(y, 0.01, 3, integ=L.RK4)
    assert maxerr(got, yo) < 3e-10
    buf = y.copy()
    g.batch_step(buf, 0.01, 3, integ=L.RK4, out=buf)
    assert np.array_equal(buf, got)
    assert maxerr(g.batch_step(y, 0.01, 1, integ=L.RKF45_GSL), o.batch_step(y, 1, 0.01, 1)[0]) < 1e-10
    assert maxerr(g.batch_ham_eqs(y), o.batch_ham_eqs(y)) < 1e-10


def _random_system(rng, m, n):
    """A random smooth coordinate map f: R^n -> R^m with full column rank near the origin and a random potential,
    written against hamilton_b200.num like a user would write mkSystem's arguments."""
    from hamilton_b200 import num
    A = rng.normal(size=(m, n)) * 0.4
    for j in range(n):
        A[j, j] += 1.5                                   # f_j = 1.5 q_j + ... : J^T W J is SPD in the sampling box
    B, Cc = rng.normal(size=(m, n)) * 0.3, rng.normal(size=(m, n))
    E = rng.normal(size=(m,)) * 0.2
    w = rng.uniform(0.5, 2.0, size=m)
    ku = rng.normal(size=(n,))
    pick = rng.integers(0, 4, size=m)

    def f(q):
        out = []
        for i in range(m):
            acc = 0.0
            for j in range(n):
                acc = acc + A[i, j] * q[j] + B[i, j] * num.sin(q[j] + Cc[i, j])
            if pick[i] == 0:
                acc = acc + E[i] * q[0] * q[n - 1]
            elif pick[i] == 1:
                acc = acc + E[i] * num.exp(0.3 * q[i % n])
            elif pick[i] == 2:
                acc = acc + E[i] * num.sqrt(2.0 + q[i % n] ** 2)
            else:
                acc = acc + E[i] / (2.0 + num.cos(q[i % n]))
            out.append(acc)
        return out

    def u(q):
        acc = 0.0
        for j in range(n):
            acc = acc + ku[j] * num.cos(q[j]) + 0.1 * q[j] ** 2 + 0.05 * num.tanh(q[j]) * q[(j + 1) % n]
        return acc + num.log(3.0 + q[0] * q[0]) + num.atan(q[n - 1]) * 0.2
    return list(w), f, u, n


@pytest.mark.parametrize("m,n,seed", [(2, 1, 1), (3, 2, 2), (4, 3, 3), (6, 4, 4), (5, 5, 5)])
def test_random_tape_systems_match_oracle(m, n, seed, oracle_mod):
    """mkSystem on arbitrary user maps: the whole chain tracer -> symbolic AD -> NVRTC -> engine against the oracle's
    tape interpreter (dense jets + explicit inverse), for hamEqs, RK4 and reference-semantics stepHam."""
    rng = np.random.default_rng(100 + seed)
    w, f, u, _ = _random_system(rng, m, n)
    g = hb.mkSystem(w, f, u, n=n)
    mm, nn, ww, fo, fouts, uo, uout, cart = tape_args(g)
    o = oracle_mod.OracleSystem.from_tape(mm, nn, ww, fo, fouts, uo, uout, cart)
    y = np.c_[rng.uniform(-0.8, 0.8, size=(97, n)), rng.uniform(-1, 1, size=(97, n))]
    fl = np.zeros(97, np.int32)
    assert maxerr(g.batch_ham_eqs(y, flags=fl), o.batch_ham_eqs(y)) < TOL and not fl.any()
    yo, bad = o.batch_step(y, 0, 0.01, 3)
    assert bad == 0 and maxerr(g.batch_step(y, 0.01, 3, integ=L.RK4), yo) < 3 * TOL
    yo, bad = o.batch_step(y, 1, 0.02, 1)
    assert bad == 0 and maxerr(g.batch_step(y, 0.02, 1, integ=L.RKF45_GSL), yo) < TOL
    e = g.batch_energies(y)
    assert maxerr(e[:, 2], [o.hamiltonian(r[:n], r[n:]) for r in y]) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU ensembles behind the C ABI (hb_ensemble_*): every GPU the box has, one process, no torch.distributed
def test_ensemble_python_binding_matches_oracle(oracle_mod):
    """Ensemble over all visible GPUs: device-side initial Phases (global indices), RK4 steps, one all-gather; compared with
    the oracle on the same splitmix64 stream (absolute error on q, p)."""
    import torch
    ndev = torch.cuda.device_count()
    sid, lo, hi = BOXES["triple_pendulum"]
    s = hb.systems.builtin(sid)
    N = 4099                                            # ragged shards whenever ndev > 1
    ens = hb.ensemble.Ensemble(s, N, ndev)
    assert ens.shards()[0] == 0 and ens.shards()[-1] == N
    ens.init_random(SEED, lo, hi)
    ens.step(0.01, nsteps=1, launches=5)
    got, _ms = ens.gather()
    o = oracle_mod.OracleSystem.builtin(sid)
    y0 = o.init_random(SEED, 0, N, lo, hi)
    want, bad = o.batch_step(y0, 0, 0.01, 5)
    assert bad == 0 and int(ens.flags().sum()) == 0
    assert float(np.max(np.abs(got - want))) < TOL
    # upload path + RKF45 semantics
    ens.upload(y0)
    ens.step(0.01, nsteps=1, launches=1, integ=L.RKF45_GSL)
    got2, _ = ens.gather()
    want2, _ = o.batch_step(y0, 1, 0.01, 1)
    assert float(np.max(np.abs(got2 - want2))) < TOL
    ens.close()


def test_ensemble_cpp_host_all_gpus(tmp_path):
    """tests/ensemble_main.cpp: a C++ host drives every GPU of the box through hb_ensemble_* with no Python in the loop;
    the gathered Phases equal a single-GPU recomputation bit for bit on every device."""
    import json
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "hamilton_b200", "lib")
    exe = str(tmp_path / "ensemble_main")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "ensemble_main.cpp"), "-o", exe, "-I/usr/local/cuda/include",
                           "-L" + lib, "-lhamilton_b200", "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-lcudart"])
    ndev = torch.cuda.device_count()
    for n_traj in (1 << 16, (1 << 16) + 5):            # equal shards (ncclAllGather) and ragged ones (grouped ncclBroadcast)
        r = subprocess.run([exe, str(ndev), "6", str(n_traj), "7", "check"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        out = json.loads(r.stdout.strip().splitlines()[-1])
        assert out["ndev"] == ndev and out["mismatches"] == 0 and out["flagged"] == 0


# ---------------------------------------------------------------------------------------------------------------------
# parity at FULL batch sizes: the GPU steps the whole BASELINE batch, the oracle steps thousands of randomly chosen
# trajectories of it (same splitmix64 stream, so the initial Phases are bit-identical), absolute error on q and p.
FULL = [("double_pendulum", 20, 4096), ("pendulum", 21, 4096), ("two_body", 21, 4096), ("spring1d", 21, 4096),
        ("triple_pendulum", 20, 4096), ("chain12", 18, 4096)]


@pytest.mark.parametrize("name,log2n,nsample", FULL)
def test_full_size_batches_against_the_oracle(name, log2n, nsample, oracle_mod):
    """Every BASELINE config at its full per-GPU size: classical RK4 (1 step and 3 steps) and one reference-semantics
    `stepHam 0.01` (adaptive GSL RKF45) of the WHOLE batch on the GPU; 4096 random rows (512 for the adaptive chain) are
    compared with the oracle."""
    import torch
    sid, lo, hi = BOXES[name]
    s, o = hb.systems.builtin(sid), oracle_mod.OracleSystem.builtin(sid)
    N = 1 << log2n
    y0 = s.batch_init_random(SEED + 5, 0, N, lo, hi)
    y0h = o.init_random(SEED + 5, 0, N, lo, hi)
    rng = np.random.default_rng(log2n * 1000 + sid)
    idx = np.unique(np.r_[0, 1, 31, 32, 33, N - 1, N - 32, N - 33, rng.integers(0, N, size=nsample)])
    idt = torch.from_numpy(idx).to(y0.device)
    assert np.array_equal(y0[idt].cpu().numpy(), y0h[idx])              # identical inputs
    threads = oracle_mod.max_threads()
    fl = torch.zeros(N, dtype=torch.int32, device=y0.device)
    for nsteps in (1, 3):
        got = s.batch_step(y0, 0.01, nsteps, integ=L.RK4, flags=fl)[idt].cpu().numpy()
        want, bad = o.batch_step(y0h[idx], 0, 0.01, nsteps, threads=threads)
        assert bad == 0 and maxerr(got, want) < TOL * nsteps
    k = idx[:512] if name == "chain12" else idx
    kt = torch.from_numpy(k).to(y0.device)
    got = s.batch_step(y0, 0.01, 1, integ=L.RKF45_GSL, flags=fl)[kt].cpu().numpy()
    want, bad = o.batch_step(y0h[k], 1, 0.01, 1, threads=threads)
    assert bad == 0 and maxerr(got, want) < TOL
    assert int(fl.sum()) == 0


def test_evolve_ham_c_matches_oracle(oracle_mod):
    """evolveHamC / evolveHamC' (src/Numeric/Hamilton.hs:470-498): toPhase -> evolveHam -> fmap fromPhase, through
    hb_evolve_ham_c; including the list variant's edge cases ([] and a single time)."""
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM, [1.0, 2.0])
    o = oracle_mod.OracleSystem.builtin(1, [1.0, 2.0])
    c0 = hb.Cfg([1.0, 0.3], [0.2, -0.5])
    p0 = o.momenta(c0.cfgPositions, c0.cfgVelocities)
    ts = np.linspace(0.0, 1.0, 11)
    ref = o.evolve_ham(c0.cfgPositions, p0, ts)                          # rows [q, p]
    want = np.array([np.r_[r[:2], o.velocities(r[:2], r[2:])] for r in ref])
    got = hb.evolveHamC(s, c0, ts)
    assert len(got) == len(ts)
    assert maxerr(np.array([np.r_[c.cfgPositions, c.cfgVelocities] for c in got]), want) < 1e-9    # 10 chained adaptive intervals
    assert maxerr(np.r_[got[0].cfgPositions, got[0].cfgVelocities], np.r_[c0.cfgPositions, c0.cfgVelocities]) < TOL   # row 0 = the initial Config
    assert hb.evolveHamC_(s, c0, []) == []
    one = hb.evolveHamC_(s, c0, [0.1])
    st = hb.stepHamC(0.1, s, c0)
    assert len(one) == 1 and maxerr(one[0].cfgPositions, st.cfgPositions) < TOL and maxerr(one[0].cfgVelocities, st.cfgVelocities) < TOL
    # a triple pendulum as well (3 x 3 solve in fromPhase)
    s3, o3 = hb.systems.builtin(hb.systems.TRIPLE_PENDULUM), oracle_mod.OracleSystem.builtin(6)
    c3 = hb.Cfg([0.5, -0.4, 1.1], [0.1, 0.2, -0.3])
    p3 = o3.momenta(c3.cfgPositions, c3.cfgVelocities)
    ts3 = [0.0, 0.05, 0.1, 0.2]
    ref3 = o3.evolve_ham(c3.cfgPositions, p3, ts3)
    want3 = np.array([np.r_[r[:3], o3.velocities(r[:3], r[3:])] for r in ref3])
    got3 = hb.evolveHamC(s3, c3, ts3)
    assert maxerr(np.array([np.r_[c.cfgPositions, c.cfgVelocities] for c in got3]), want3) < 1e-9


def test_largest_supported_system_n16(oracle_mod):
    """n = HB_MAX_N = 16 (n = 13..16 untested): a 16-link chain (System 32 16) as a tape system — NVRTC build of
    hamEqs, RK4 and the adaptive stepper against the oracle's tape interpreter.  (A 16 x 16 packed mass matrix is 136 doubles:
    it cannot live in registers, the kernel spills to local memory — correct, slow)"""
    s = hb.systems.from_def(hb.systems.pendulum_chain_def([1.0] * 16, [1.0] * 16))
    o = oracle_mod.OracleSystem.from_tape(*tape_args(s))
    rng = np.random.default_rng(1616)
    y = np.c_[rng.uniform(-np.pi, np.pi, size=(96, 16)), rng.uniform(-1, 1, size=(96, 16))]
    fl = np.zeros(96, np.int32)
    assert maxerr(s.batch_ham_eqs(y, flags=fl), o.batch_ham_eqs(y)) < TOL and not fl.any()
    yo, bad = o.batch_step(y, 0, 0.01, 2, threads=oracle_mod.max_threads())
    assert bad == 0 and maxerr(s.batch_step(y, 0.01, 2, integ=L.RK4), yo) < 2 * TOL
    yo, bad = o.batch_step(y[:16], 1, 0.01, 1, threads=oracle_mod.max_threads())
    assert bad == 0 and maxerr(s.batch_step(y[:16], 0.01, 1, integ=L.RKF45_GSL), yo) < TOL


def test_advice_r1_edge_cases(oracle_mod):
    """(1) `stepHam r` with r <= 0 returns the Phase unchanged (hmatrix-gsl's `while (t < t1)` never runs);
    (2) inertias around 1e-170 make the closed forms' determinant underflow although the mass matrix is perfectly
    invertible (LAPACK's `inv` in the reference has no problem): the fast path's failed pivot test only sends the
    trajectory to the scale-safe LDL^T of the slow path — finite velocities, no HB_FLAG_NOT_SPD."""
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    y = random_phases("double_pendulum", 40)
    for dt in (0.0, -0.01):
        assert np.array_equal(s.batch_step(y, dt, 3, integ=L.RKF45_GSL), y)
    tiny = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM, [1e-170, 2e-170])
    o = oracle_mod.OracleSystem.builtin(1, [1e-170, 2e-170])
    yt = y.copy(); yt[:, 2:] *= 1e-170                      # momenta on the scale of the inertias: velocities are O(1)
    fl = np.zeros(len(yt), np.int32)
    c = tiny.batch_from_phase(yt, flags=fl)
    assert not fl.any() and np.isfinite(c).all()
    want = np.array([np.r_[r[:2], o.velocities(r[:2], r[2:])] for r in yt])
    assert float(np.max(np.abs(c - want) / (1e-300 + np.abs(want)).clip(1e-3))) < 1e-9
    dy = tiny.batch_ham_eqs(yt, flags=fl)
    assert not fl.any() and np.isfinite(dy).all()
    assert maxerr(dy[:, :2], o.batch_ham_eqs(yt)[:, :2]) < 1e-9
