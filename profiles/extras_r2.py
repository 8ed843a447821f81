"""Round-2 secondary measurements: the one-evaluation kernels (hamEqs, toPhase, fromPhase, energies, underlyingPos) of the
BASELINE systems against the HBM roofline, and the evolveHam output path.  Writes gpurun_out/<dir>/extras_r2.json.
Timing: one CUDA graph of K launches over a ring of buffers larger than L2, replayed once warm and once timed (as bench.py).
usage: python profiles/extras_r2.py <outdir>"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
import torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES

PEAK = 6532.2
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass

def graph_time(body, K):
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(K): body(i)
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best / K   # ms per launch

res = []
def one_eval(name, log2n, K=200):
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid)
    N, n, m = 1 << log2n, s.n, s.m
    bytes_pair = 2 * N * 2 * n * 8
    ring = max(2, int(600e6 // bytes_pair) + 1)
    ins = [s.batch_init_random(11 + r, 0, N, lo, hi) for r in range(ring)]
    outs = [torch.empty_like(b) for b in ins]
    e_out = [torch.empty((N, 4), dtype=torch.float64, device="cuda") for _ in range(ring)]
    q_in = [b[:, :n].contiguous() for b in ins]
    u_out = [torch.empty((N, m), dtype=torch.float64, device="cuda") for _ in range(ring)]
    for r in range(min(3, ring)):   # warm every kernel (module load, shape cache) outside any capture
        s.batch_ham_eqs(ins[r], out=outs[r]); s.batch_to_phase(ins[r], out=outs[r]); s.batch_from_phase(ins[r], out=outs[r])
        s.batch_energies(ins[r], out=e_out[r]); s.batch_underlying_pos(q_in[r], out=u_out[r])
    torch.cuda.synchronize()
    cases = [
        ("hamEqs", lambda i: s.batch_ham_eqs(ins[i % ring], out=outs[i % ring]), 32 * n),
        ("toPhase", lambda i: s.batch_to_phase(ins[i % ring], out=outs[i % ring]), 32 * n),
        ("fromPhase", lambda i: s.batch_from_phase(ins[i % ring], out=outs[i % ring]), 32 * n),
        ("energies", lambda i: s.batch_energies(ins[i % ring], out=e_out[i % ring]), 16 * n + 32),
        ("underlyingPos", lambda i: s.batch_underlying_pos(q_in[i % ring], out=u_out[i % ring]), 8 * (n + m)),
    ]
    for kname, fn, bpt in cases:
        ms = graph_time(fn, K)
        gbs = N * bpt / (ms * 1e-3) / 1e9
        res.append({"system": name, "N": N, "kernel": kname, "us_per_launch": ms * 1e3, "evals_per_s": N / (ms * 1e-3),
                    "algorithmic_bytes_per_trajectory": bpt, "hbm_GBps_algorithmic": gbs, "frac_of_hbm_peak": gbs / PEAK})
        print(res[-1], flush=True)

one_eval("double_pendulum", 20)
one_eval("triple_pendulum", 20)
one_eval("pendulum", 21)
one_eval("two_body", 21)
one_eval("chain12", 18, K=50)

# evolveHam output path (f)2: 64 grid points, 262,144 trajectories -> 64 batches written
def evolve(name, log2n, integ, sub, npts=64):
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid)
    N = 1 << log2n
    y0 = s.batch_init_random(3, 0, N, lo, hi)
    ts = np.arange(npts) * 0.01
    out = torch.empty((npts, N, 2 * s.n), dtype=torch.float64, device="cuda")
    for _ in range(2): s.batch_evolve(y0, ts, integ=integ, rk4_substeps=sub, out=out)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.batch_evolve(y0, ts, integ=integ, rk4_substeps=sub, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    r = {"system": name, "N": N, "kernel": "evolve_rk4 x%d substeps" % sub if integ == L.RK4 else "evolve_rkf45 (evolveHam)",
         "grid_points": npts, "ms": best, "intervals_per_s": N * (npts - 1) / (best * 1e-3),
         "output_GBps": npts * N * 2 * s.n * 8 / (best * 1e-3) / 1e9}
    if integ == L.RK4: r["rk4_steps_per_s"] = N * (npts - 1) * sub / (best * 1e-3)
    res.append(r); print(r, flush=True)

evolve("double_pendulum", 18, L.RK4, 4)
evolve("double_pendulum", 18, L.RKF45_GSL, 1)
evolve("chain12", 16, L.RK4, 1, npts=16)
evolve("chain12", 16, L.RKF45_GSL, 1, npts=16)

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
os.makedirs(out, exist_ok=True)
json.dump(res, open(os.path.join(out, "extras_r2.json"), "w"), indent=1)
