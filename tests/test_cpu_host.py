"""CPU suite, part 2: host logic of the product — the C-ABI library loads and exports every symbol the header
declares, the symbolic system compiler's output is checked against the oracle by compiling the generated `Sys` struct
as host code, tape validation errors, the Python mirror's API edge cases, and the N>1 sharding/gather plumbing over gloo.
No compute entry point is exercised without a GPU (they must refuse: there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from hamilton_b200 import num
from tests.common import BOXES, maxerr, random_phases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = list(BOXES)


def test_abi_library_exports_every_declared_symbol():
    lib = L.lib()
    assert lib.hb_abi_version() == 2
    hdr = open(os.path.join(ROOT, "include", "hamilton_b200.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"hb_op", "hb_tape", "hb_status"}
    assert declared == set(L.ABI_SYMBOLS), declared ^ set(L.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.hb_last_error() is not None


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    with pytest.raises(hb.NoDeviceError):
        hb.hamEqs(s, hb.Phase([1.0, 0.0], [0.0, 0.5]))
    with pytest.raises(hb.NoDeviceError):
        s.batch_step(np.zeros((4, 4)), 0.01)
    n = C.c_int32(-1)
    assert L.lib().hb_device_count(C.byref(n)) == L.ERR_NO_DEVICE and n.value == 0
    with pytest.raises(hb.NoDeviceError):
        hb.ensemble.Ensemble(s, 1024, 1)


def _host_eval(system, name, tmp):
    """Compiles the generated Sys struct as host C++ and returns eval(q) -> (J, H, gU, x, U, w)."""
    src = os.path.join(tmp, name + "_sys.inc")
    with open(src, "w") as f:
        f.write(system.source())
    sname = re.search(r"struct (\w+) \{", system.source()).group(1)
    so = os.path.join(tmp, name + ".so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", '-DSYS_SOURCE="%s"' % src, "-DSYS_NAME=" + sname,
                           os.path.join(ROOT, "tests", "host_eval_harness.cpp"), "-o", so])
    lib = C.CDLL(so)
    d = (C.c_int * 4)()
    lib.dims(d)
    m, n = d[0], d[1]
    assert (m, n) == (system.m, system.n)
    prm = np.r_[system.params(), 0.0]
    dp = C.POINTER(C.c_double)

    def ev(q):
        q = np.ascontiguousarray(q, float)
        J, H, gU, x, U, w = np.zeros((m, n)), np.zeros((n, m, n)), np.zeros(n), np.zeros(m), C.c_double(), np.zeros(m)
        lib.eval(prm.ctypes.data_as(dp), q.ctypes.data_as(dp), J.ctypes.data_as(dp), H.ctypes.data_as(dp), gU.ctypes.data_as(dp),
                 x.ctypes.data_as(dp), C.byref(U), w.ctypes.data_as(dp))
        return J, H, gU, x, U.value, w

    def ev_sym(q, pp):
        """symbolic hamEqs path: (dq, dp, A, U) or None when the system uses the direct contraction"""
        q, pp = np.ascontiguousarray(q, float), np.ascontiguousarray(pp, float)
        dq, dpp, A, U = np.zeros(n), np.zeros(n), np.zeros((n, n)), C.c_double()
        lib.eval_symham.restype = C.c_int
        ok = lib.eval_symham(prm.ctypes.data_as(dp), q.ctypes.data_as(dp), pp.ctypes.data_as(dp), dq.ctypes.data_as(dp),
                             dpp.ctypes.data_as(dp), A.ctypes.data_as(dp), C.byref(U))
        return (dq, dpp, A, U.value) if ok else None
    ev.sym = ev_sym
    return ev, (d[2], d[3])


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kind", ["aot", "jit"])
def test_system_compiler_output_matches_oracle_derivatives(name, kind, oracle_mod):
    """Symbolic 2nd-order forward AD (csrc/symbolic.cpp + sysgen.cpp) vs the oracle's dense jets: J, every Hessian slice,
    grad U, f, U and the inertia vector, for the AOT built-ins (runtime PARAM leaves) and tape systems (literals)."""
    sid = BOXES[name][0]
    if kind == "jit":
        os.environ["HB_JIT_SKIP_COMPILE"] = "1"      # symbolic stage only: NVRTC is exercised by test_nvrtc_*
        try:
            g = hb.systems.from_def(hb.systems.DEFS[sid]())
        finally:
            del os.environ["HB_JIT_SKIP_COMPILE"]
    else:
        g = hb.systems.builtin(sid)
    o = oracle_mod.OracleSystem.builtin(sid)
    with tempfile.TemporaryDirectory() as tmp:
        ev, (nj, nh) = _host_eval(g, name + kind, tmp)
        for r in random_phases(name, 5):
            q = r[: o.n]
            J, H, gU, x, U, w = ev(q)
            assert maxerr(J, o.jacobian(q)) < 1e-13
            assert maxerr(H, o.hessian(q)) < 1e-13
            assert maxerr(gU, o.potential_grad(q)) < 1e-13
            assert maxerr(x, o.underlying_pos(q)) < 1e-13 and abs(U - o.pe(q)) < 1e-13 * (1 + abs(U))
            # symbolic hamEqs (csrc/polyform.cpp): mass matrix, potential and (dq, dp) against the oracle's literal
            # restatement of src/Numeric/Hamilton.hs:370-387
            sym = ev.sym(q, r[o.n:])
            if name != "bezier" or kind == "jit":
                assert sym is not None, "expected the symbolic form to be selected"
            if sym is not None:
                dq, dpp, A, Us = sym
                Jo, wo = o.jacobian(q), w
                Mo = Jo.T @ (wo[:, None] * Jo)
                wdq, wdp = o.ham_eqs(q, r[o.n:])
                assert maxerr(A, Mo) < 1e-13 * (1 + np.abs(Mo).max())
                assert abs(Us - o.pe(q)) < 1e-13 * (1 + abs(Us))
                scale = 1 + max(np.abs(wdq).max(), np.abs(wdp).max())
                assert maxerr(dq, wdq) < 1e-12 * scale and maxerr(dpp, wdp) < 1e-12 * scale
        # sparsity is structural: chain12's Hessian tensor is "diagonal" (156 of 3456 entries), its J lower-triangular
        if name == "chain12":
            assert (nj, nh) == (156, 156)
        if name == "room":
            assert (nj, nh) == (2, 0)


def test_nvrtc_compiles_a_tape_system_without_a_gpu():
    """hb_system_from_tape = symbolic AD + CUDA source + NVRTC (sm_100a cubin); needs no device until the first launch."""
    s = hb.mkSystem_([1.0, 1.0], lambda q: [num.sin(q[0]), 0.5 - num.cos(q[0])], lambda x: x[1], n=1)
    assert (s.m, s.n) == (2, 1)
    assert "hb_sincos" in s.source() and "struct HbSysJit" in s.source()


def test_jit_disk_cache_hit_and_miss(tmp_path):
    """The cubin of a tape system is kept on disk keyed by (engine header, generated source, arch, options, NVRTC version):
    the second construction of the same System reads it back instead of compiling; a different System misses;
    HB_JIT_CACHE=0 bypasses the cache; a corrupt file is ignored and replaced."""
    import time
    old = {k: os.environ.get(k) for k in ("HB_JIT_CACHE_DIR", "HB_JIT_CACHE")}
    os.environ["HB_JIT_CACHE_DIR"] = str(tmp_path / "cache")
    os.environ.pop("HB_JIT_CACHE", None)

    def build(k):
        t0 = time.perf_counter()
        s = hb.mkSystem_([1.0, 1.0], lambda q: [num.sin(k * q[0]), 0.5 - num.cos(q[0])], lambda x: x[1], n=1)
        return s, time.perf_counter() - t0

    try:
        _, t_cold = build(1.25)
        files = sorted(os.listdir(tmp_path / "cache"))
        assert len(files) == 1 and files[0].endswith(".cubin") and not any(".tmp." in f for f in files)
        with open(tmp_path / "cache" / files[0], "rb") as f:
            blob = f.read()
        assert blob[:4] == b"\x7fELF"
        _, t_warm = build(1.25)
        assert sorted(os.listdir(tmp_path / "cache")) == files and t_warm < 0.5 * t_cold
        build(1.75)                                                  # another System: another entry
        assert len(os.listdir(tmp_path / "cache")) == 2
        os.environ["HB_JIT_CACHE"] = "0"
        _, t_off = build(1.25)
        assert t_off > 2 * t_warm and len(os.listdir(tmp_path / "cache")) == 2
        del os.environ["HB_JIT_CACHE"]
        with open(tmp_path / "cache" / files[0], "wb") as f:         # a truncated / foreign file must not be loaded
            f.write(b"garbage")
        build(1.25)
        with open(tmp_path / "cache" / files[0], "rb") as f:
            assert f.read() == blob                                   # recompiled (deterministically) and rewritten
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_tape_validation_errors():
    lib = L.lib()

    def mk(ops, outs, n_in):
        arr = (L.HbOp * len(ops))()
        for k, (op, a, b, c) in enumerate(ops):
            arr[k].op, arr[k].a, arr[k].b, arr[k].c = op, a, b, c
        o = (C.c_int32 * len(outs))(*outs)
        return L.HbTape(n_in, len(ops), arr, len(outs), o), (arr, o)
    w = (C.c_double * 1)(1.0)
    good_u, k0 = mk([(num.OP_INPUT, 0, 0, 0.0)], [0], 1)
    h = C.c_void_p()
    # forward reference
    bad, k1 = mk([(num.OP_INPUT, 0, 0, 0.0), (num.OP_ADD, 0, 5, 0.0)], [1], 1)
    assert lib.hb_system_from_tape(1, 1, w, C.byref(bad), C.byref(good_u), 0, None, 0, C.byref(h)) == L.ERR_TAPE
    assert b"earlier node" in lib.hb_last_error()
    # unknown opcode
    bad, k2 = mk([(99, 0, 0, 0.0)], [0], 1)
    assert lib.hb_system_from_tape(1, 1, w, C.byref(bad), C.byref(good_u), 0, None, 0, C.byref(h)) == L.ERR_TAPE
    # wrong arity of f
    assert lib.hb_system_from_tape(2, 1, w, C.byref(good_u), C.byref(good_u), 0, None, 0, C.byref(h)) == L.ERR_TAPE
    # dimensions out of range, null pointers
    assert lib.hb_system_from_tape(1, 17, w, C.byref(good_u), C.byref(good_u), 0, None, 0, C.byref(h)) == L.ERR_INVALID
    assert lib.hb_system_builtin(99, None, 0, C.byref(h)) == L.ERR_INVALID
    assert lib.hb_batch_step(None, 0, 0.01, 1, 1, 0, 0, None, None, None, None) == L.ERR_INVALID


def test_python_mirror_argument_checks():
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    assert (s.m, s.n) == (4, 2)
    assert np.allclose(s.params(), [1.0, 1.0])
    assert np.allclose(hb.systems.builtin(hb.systems.TWO_BODY).params(), [5.0, 0.5, -(0.5 / 5.5), 5.0 / 5.5, 2.5])
    with pytest.raises(ValueError):
        hb.evolveHam(s, hb.Phase([0, 0], [0, 0]), [0.0])          # the type-level `2 <= s`
    assert hb.evolveHam_(s, hb.Phase([0, 0], [0, 0]), []) == []      # evolveHam' [] = []
    assert hb.evolveHamC_(s, hb.Config([0, 0], [0, 0]), []) == []
    with pytest.raises(ValueError):
        hb.hamEqs(s, hb.Phase([0.0], [0.0]))                        # wrong vector length
    with pytest.raises(ValueError):
        s.batch_step(np.zeros((4, 3)), 0.01)
    with pytest.raises(ValueError):
        hb.mkSystem([1, 1], lambda q: [q[0]], lambda q: q[0], n=1)   # f returns 1 coordinate, inertia has 2


def test_tracer_records_haskell_operator_split():
    """x ** 2.0 -> POW (Floating (**)), x ** 2 -> POWI (Num (^)); constants lift like fromInteger/realToFrac."""
    t, outs = num.trace(lambda q: [q[0] ** 2.0, q[0] ** 3, 2 * q[0] - 1, num.atan2(q[0], 1.0)], 1)
    ops = [o[0] for o in t.ops]
    assert num.OP_POW in ops and num.OP_POWI in ops and num.OP_ATAN2 in ops and len(outs) == 4
    assert num.sin(0.5) == np.sin(0.5)                               # same functions on plain floats
    import sympy
    assert num.cos(sympy.Symbol("x")).diff(sympy.Symbol("x")) == -sympy.sin(sympy.Symbol("x"))


def test_shard_partition():
    for n, w in [(10, 3), (8388608, 8), (5, 8), (0, 2)]:
        parts = [hb.ensemble.shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, _) in zip(parts, parts[1:]):
            assert f0 + c0 == f1


def _gloo_worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import hamilton_b200 as hbm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = O.OracleSystem.builtin(O.DOUBLE_PENDULUM)
    lo, hi = BOXES["double_pendulum"][1:]
    first, count = hbm.ensemble.shard(n_total, rank, world)
    y = o.init_random(0x48414D49, first, count, lo, hi)            # what batch_init_random generates on each GPU
    out, _ = o.batch_step(y, 0, 0.01, 2)                             # stand-in for the GPU step (tests may use the oracle)
    full = hbm.ensemble.gather_final(torch.from_numpy(out))
    # the same gather with the block split known up front (no count exchange) and, for even shards, a preallocated output
    full2 = hbm.ensemble.gather_final(torch.from_numpy(out), n_total=n_total, out=torch.empty_like(full))
    assert torch.equal(full, full2)
    try:
        hbm.ensemble.gather_final(torch.from_numpy(out), n_total=n_total + world)   # shard sizes no longer match
        bad = False
    except ValueError:
        bad = True
    assert bad
    q.put((rank, full.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])
def test_ensemble_gather_world_size_2_gloo(n_total, oracle_mod):
    """N>1 path on CPU: block sharding + counter-based init + all-gather reassemble exactly the single-process result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle_mod.OracleSystem.builtin(1)
    lo, hi = BOXES["double_pendulum"][1:]
    want, _ = o.batch_step(o.init_random(0x48414D49, 0, n_total, lo, hi), 0, 0.01, 2)
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                                  cwd=ROOT, text=True)
    import json
    j = json.loads(out.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["unit"] == "steps/s"


@pytest.mark.parametrize("m,n,seed", [(2, 1, 1), (3, 2, 2), (4, 3, 3), (6, 4, 4), (5, 5, 5)])
def test_system_compiler_on_random_user_maps(m, n, seed, oracle_mod):
    """Arbitrary user maps (sin/exp/sqrt/recip/tanh/log/atan/powers): symbolic derivatives vs the oracle's dense jets."""
    from tests.common import tape_args
    from tests.test_gpu_parity import _random_system
    rng = np.random.default_rng(100 + seed)
    w, f, u, _ = _random_system(rng, m, n)
    os.environ["HB_JIT_SKIP_COMPILE"] = "1"
    try:
        g = hb.mkSystem(w, f, u, n=n)
    finally:
        del os.environ["HB_JIT_SKIP_COMPILE"]
    mm, nn, ww, fo, fouts, uo, uout, cart = tape_args(g)
    o = oracle_mod.OracleSystem.from_tape(mm, nn, ww, fo, fouts, uo, uout, cart)
    with tempfile.TemporaryDirectory() as tmp:
        ev, _ = _host_eval(g, "rand%d" % seed, tmp)
        for _ in range(6):
            q = rng.uniform(-0.8, 0.8, size=n)
            J, H, gU, x, U, wv = ev(q)
            assert maxerr(J, o.jacobian(q)) < 1e-13 and maxerr(H, o.hessian(q)) < 1e-13
            assert maxerr(gU, o.potential_grad(q)) < 1e-13 and abs(U - o.pe(q)) < 1e-13 and maxerr(wv, w) == 0


@pytest.mark.parametrize("m,n,seed", [(2, 1, 1), (3, 2, 2), (4, 3, 3), (6, 4, 4), (5, 5, 5)])
@pytest.mark.parametrize("force", ["1", None])
def test_symbolic_ham_eqs_on_random_user_maps(m, n, seed, force, oracle_mod):
    """The symbolic mass-matrix / force-polynomial form of hamEqs (csrc/polyform.cpp), forced on (HB_SYMH=1) and under the
    compiler's own cost model, against the oracle's literal restatement for arbitrary user maps."""
    from tests.common import tape_args
    from tests.test_gpu_parity import _random_system
    rng = np.random.default_rng(100 + seed)
    w, f, u, _ = _random_system(rng, m, n)
    os.environ["HB_JIT_SKIP_COMPILE"] = "1"
    if force:
        os.environ["HB_SYMH"] = force
    try:
        g = hb.mkSystem(w, f, u, n=n)
    finally:
        del os.environ["HB_JIT_SKIP_COMPILE"]
        os.environ.pop("HB_SYMH", None)
    mm, nn, ww, fo, fouts, uo, uout, cart = tape_args(g)
    o = oracle_mod.OracleSystem.from_tape(mm, nn, ww, fo, fouts, uo, uout, cart)
    with tempfile.TemporaryDirectory() as tmp:
        ev, _ = _host_eval(g, "symrand%d" % seed, tmp)
        checked = 0
        for _ in range(6):
            q, pp = rng.uniform(-0.8, 0.8, size=n), rng.uniform(-1, 1, size=n)
            sym = ev.sym(q, pp)
            if sym is None:
                assert not force or "self-check" in g.source() or "too large" in g.source()
                continue
            dq, dpp, A, Us = sym
            Jo = o.jacobian(q)
            Mo = Jo.T @ (np.asarray(w)[:, None] * Jo)
            wdq, wdp = o.ham_eqs(q, pp)
            cond = np.linalg.cond(Mo)
            assert maxerr(A, Mo) < 1e-13 * (1 + np.abs(Mo).max()) and abs(Us - o.pe(q)) < 1e-13 * (1 + abs(Us))
            scale = (1 + max(np.abs(wdq).max(), np.abs(wdp).max())) * max(1.0, cond)
            assert maxerr(dq, wdq) < 1e-13 * scale and maxerr(dpp, wdp) < 1e-13 * scale
            checked += 1
        if force:
            assert checked > 0 or "hamEqs form: direct" in g.source()


def test_haskell_shim_matches_header():
    """haskell/Numeric/Hamilton/B200.hs cannot be compiled here (no GHC): check what can be checked statically — every
    `foreign import` names a function of include/hamilton_b200.h with the same arity and argument kinds, nothing is left
    `undefined`, and every name of the reference's export list (src/Numeric/Hamilton.hs:28-70) is defined."""
    hs = open(os.path.join(ROOT, "haskell", "Numeric", "Hamilton", "B200.hs")).read()
    hdr = open(os.path.join(ROOT, "include", "hamilton_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_ *]*?[ *])(hb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        kinds = []
        for a in ([] if args in ("", "void") else args.split(",")):
            a = a.strip()
            if "*" in a:
                kinds.append("ptr")
            elif re.match(r"(const\s+)?double\b", a):
                kinds.append("double")
            elif re.match(r"(const\s+)?(int64_t)\b", a):
                kinds.append("i64")
            elif re.match(r"(const\s+)?(uint64_t)\b", a):
                kinds.append("u64")
            elif re.match(r"(const\s+)?(int32_t|hb_integrator|hb_layout|hb_memspace|hb_builtin|hb_status)\b", a):
                kinds.append("i32")
            elif re.match(r"size_t\b", a):
                kinds.append("size")
            else:
                raise AssertionError("unparsed C parameter %r of %s" % (a, name))
        protos[name] = (ret, kinds)
    hs_kind = {"Int32": "i32", "Int64": "i64", "Word64": "u64", "Double": "double"}
    imports = re.findall(r'foreign import ccall safe "(&?)(hb_[a-z0-9_]+)"\s*\n?\s*\w+ ::\s*([^\n]*)', hs)
    assert len(imports) >= 24
    for amp, name, sig in imports:
        assert name in protos, name
        sig = sig.split("--")[0].strip()
        if amp:                                   # address of a finaliser: FunPtr (Ptr X -> IO ())
            assert sig.startswith("FunPtr") and protos[name][1] == ["ptr"], name
            continue
        parts = [x.strip() for x in re.split(r"->(?![^()]*\))", sig)]
        res, params = parts[-1], parts[:-1]
        got = ["ptr" if x.startswith("Ptr") or x.startswith("(Ptr") else hs_kind[x] for x in params]
        assert got == protos[name][1], (name, got, protos[name][1])
        assert res in ("IO Int32", "IO CString"), (name, res)
    assert not re.search(r"=\s*undefined\b", hs) and "undefined;" not in hs
    for name in ("mkSystem", "mkSystem'", "underlyingPos", "toPhase", "fromPhase", "momenta", "velocities", "keC", "keP", "pe", "lagrangian",
                 "hamiltonian", "hamEqs", "stepHam", "evolveHam", "evolveHam'", "stepHamC", "evolveHamC", "evolveHamC'", "batchStep", "batchEvolve"):
        assert re.search(r"^" + re.escape(name) + r"\s*(::|\n\s+::)", hs, flags=re.M), name
        assert re.search(r"^" + re.escape(name) + r"\s[^:\n]*=", hs, flags=re.M), name
    for inst in ("Num", "Fractional", "Floating", "Eq", "Ord", "Real", "RealFrac", "RealFloat"):
        assert "instance %s Tr where" % inst in hs, inst


def test_python_mirror_rejects_buffers_the_c_library_would_overrun():
    """dtype and shape of every caller-supplied buffer are validated before the pointer reaches the C ABI."""
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    y = np.zeros((8, 4))
    with pytest.raises(ValueError):
        s.batch_step(y.astype(np.int32), 0.01)                      # int32 data read as doubles: twice its size
    with pytest.raises(ValueError):
        s.batch_step(y, 0.01, out=np.zeros((4, 4)))                 # too small an output
    with pytest.raises(ValueError):
        s.batch_step(y, 0.01, out=np.zeros((8, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        s.batch_step(y, 0.01, flags=np.zeros(8, dtype=np.int64))
    with pytest.raises(ValueError):
        s.batch_step(y, 0.01, flags=np.zeros(4, dtype=np.int32))
    with pytest.raises(ValueError):
        s.batch_energies(y, out=np.zeros((8, 3)))
    with pytest.raises(ValueError):
        s.batch_evolve(y, [0.0, 0.1, 0.2], out=np.zeros((2, 8, 4)))


def test_evolve_rejects_a_zero_length_first_interval_for_the_adaptive_integrator():
    """the reference's initial step is (ts[1] - ts[0]) / 100; zero makes GSL (and a GPU kernel) spin forever, so the
    grid is rejected before anything is launched (argument validation needs no device)."""
    s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
    y = np.zeros((4, 4))
    with pytest.raises(hb.HamiltonError) as ei:
        s.batch_evolve(y, [0.0, 0.0, 0.1], integ=L.RKF45_GSL)
    assert ei.value.status == L.ERR_INVALID
    with pytest.raises(hb.HamiltonError) as ei:
        s.batch_evolve(y, [0.0, 0.1, 0.05])
    assert ei.value.status == L.ERR_INVALID
    # misaligned array-of-Phases pointers are refused, not dereferenced (they would be a sticky device fault)
    buf = np.zeros(4 * 4 + 1)
    with pytest.raises(hb.HamiltonError) as ei:
        s.batch_step(buf[1:].reshape(4, 4), 0.01)
    assert ei.value.status == L.ERR_INVALID


def test_host_and_engine_agree_on_the_shared_memory_constants():
    """csrc/runtime.cpp sizes the dynamic shared memory of every launch from constants it cannot include from the engine header
    (the header is CUDA): the two definitions must say the same, and HB_TAB_BYTES must hold the staged table image
    (2048 sin/cos pairs + 64 powers of two + the mbarrier)."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hamilton_b200", "csrc")
    eng = open(os.path.join(root, "engine", "hb_engine.cuh")).read()
    run = open(os.path.join(root, "runtime.cpp")).read()
    for name in ("HB_TAB_BYTES", "HB_BIG_N", "HB_WSTORE_MAXD"):
        a = re.search(r"#define\s+%s\s+(\d+)" % name, eng)
        b = re.search(r"#define\s+%s\s+(\d+)" % name, run)
        assert a and b and a.group(1) == b.group(1), (name, a and a.group(1), b and b.group(1))
    tab = int(re.search(r"#define\s+HB_TAB_BYTES\s+(\d+)", eng).group(1))
    assert tab >= 2048 * 16 + 64 * 8 + 8 and tab % 128 == 0


@pytest.mark.parametrize("sid,bound", [(0, 14), (1, 70), (6, 124), (7, 1130)])
def test_system_compiler_operation_counts_do_not_regress(sid, bound):
    """Arithmetic operations per RHS of the symbolic hamEqs form, as the system compiler's cost model counts them (a
    transcendental = 14).  The kernels are instruction-issue-bound, so every operation the compiler leaves in the DAG is
    paid for: like-term collection (c1 x + c2 x) took the 12-link chain from 1195 to 1130 and the triple pendulum's gravity
    gradient from 5 (s + 2 s) to 15 s.  Pendulum, double pendulum, triple pendulum, chain12."""
    src = hb.systems.builtin(sid).source()
    m = re.search(r"cost model: symbolic (\d+), direct (\d+)", src)
    assert m, src[:300]
    assert int(m.group(1)) <= bound, (int(m.group(1)), bound)
    assert "hamEqs form: symbolic" in src
