"""cpp/hamilton.hpp — the compiled-language host mirror of Numeric.Hamilton — built against the C-ABI library and run
the way the reference's README uses the Haskell API (README.md:88-165)."""
import os
import subprocess

import numpy as np
import pytest

import hamilton_b200 as hb
from hamilton_b200 import num

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "cpp_mirror")
    lib = os.path.join(ROOT, "hamilton_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp_mirror_test.cpp"), "-o", out,
                           "-L" + lib, "-lhamilton_b200", "-Wl,-rpath," + lib])
    return out


def test_cpp_mirror_builds_and_traces_a_system(exe):
    """mkSystem' with generic lambdas -> tape -> symbolic AD -> NVRTC, no GPU needed."""
    assert subprocess.check_output([exe, "create"], text=True).strip() == "created"


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(exe, oracle_mod):
    out = subprocess.check_output([exe], text=True)
    vals = {}
    for line in out.splitlines():
        k, *v = line.split()
        vals[k] = v
    m1, m2 = 1.0, 2.0
    f = lambda q: [num.sin(q[0]), -num.cos(q[0]), num.sin(q[0]) + num.sin(q[1]) / 2.0, -num.cos(q[0]) - num.cos(q[1]) / 2.0]   # noqa: E731
    u = lambda x: 5.0 * (m1 * x[1] + m2 * x[3])   # noqa: E731
    ft, fo = num.trace(f, 2)
    ut, uo = num.trace(u, 4)
    o = oracle_mod.OracleSystem.from_tape(4, 2, [m1, m1, m2, m2], ft.ops, fo, ut.ops, uo[0], True)
    q0, v0 = [1.0, 0.0], [0.0, 0.5]
    p0 = o.momenta(q0, v0)
    fl = lambda k: np.array([float(x) for x in vals[k]])   # noqa: E731
    assert np.allclose(fl("momenta"), p0, rtol=0, atol=1e-12)
    assert abs(fl("hamiltonian")[0] - o.hamiltonian(q0, p0)) < 1e-12
    dq, dp = o.ham_eqs(q0, p0)
    assert np.allclose(fl("hamEqs"), np.r_[dq, dp], rtol=0, atol=1e-12)
    ts = [0.1 * k for k in range(11)]
    assert np.allclose(fl("evolve_last"), o.evolve_ham(q0, p0, ts)[-1], rtol=0, atol=1e-9)
    qs, ps = o.step_ham(0.1, q0, p0)
    assert np.allclose(fl("step"), np.r_[qs, ps], rtol=0, atol=1e-10)
    assert np.allclose(fl("stepC"), np.r_[qs, o.velocities(qs, ps)], rtol=0, atol=1e-9)
    assert vals["evolve_prime_sizes"] == ["0", "1"] and "throws_on_short_grid" in vals
