"""Secondary measurements (not the bench line): the other BASELINE configs and the reference-semantics integrator.
Writes gpurun_out/extras.json.  Timing: CUDA events, 3 warm-ups, ring of buffers larger than L2."""
import json, sys, os
sys.path.insert(0, ".")
import numpy as np
import torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES

def timed(fn, reps):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def run(name, N, integ, nsteps, reps, layout=L.AOS, ring=5):
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid)
    ins = [s.batch_init_random(7 + r, 0, N, lo, hi, layout=layout) for r in range(ring)]
    outs = [torch.empty_like(b) for b in ins]
    ms = timed(lambda i: s.batch_step(ins[i % ring], 0.01, nsteps, integ=integ, out=outs[i % ring], layout=layout), reps)
    sps = N * nsteps / (ms * 1e-3)
    return {"system": name, "N": N, "integrator": "rk4" if integ == L.RK4 else "rkf45_gsl", "steps_per_launch": nsteps,
            "ms_per_launch": ms, "steps_per_s": sps, "hbm_GBps_algorithmic": sps * 32 * s.n / 1e9 / nsteps if nsteps == 1 else None}

res = []
res.append(run("double_pendulum", 1 << 20, L.RK4, 1, 50, ring=9))
res.append(run("double_pendulum", 1 << 20, L.RK4, 16, 10, ring=9))
res.append(run("double_pendulum", 1 << 20, L.RKF45_GSL, 1, 10, ring=9))      # reference semantics: stepHam 0.01
res.append(run("pendulum", 1 << 21, L.RK4, 1, 30, ring=9))                  # config 3 (the reference's System 2 1)
res.append(run("spring1d", 1 << 21, L.RK4, 1, 30, ring=9))                  # config 3 (synthetic 1-D spring)
res.append(run("two_body", 1 << 21, L.RK4, 1, 30, ring=5))                  # config 3
res.append(run("triple_pendulum", 1 << 20, L.RK4, 1, 30, ring=5))           # config 4 per-GPU shard
res.append(run("triple_pendulum", 1 << 20, L.RKF45_GSL, 1, 5, ring=5))
res.append(run("chain12", 1 << 18, L.RK4, 1, 10, ring=5))                   # config 5
res.append(run("chain12", 1 << 18, L.RK4, 1, 10, layout=L.SOA, ring=5))
res.append(run("chain12", 1 << 18, L.RK4, 8, 3, ring=5))
# config 3 mixed on two streams: 2M pendulum + 2M two-body concurrently
sp, st = hb.systems.builtin(0), hb.systems.builtin(3)
Np = 1 << 21
yp = sp.batch_init_random(1, 0, Np, *BOXES["pendulum"][1:]); op = torch.empty_like(yp)
yt = st.batch_init_random(1, 0, Np, *BOXES["two_body"][1:]); ot = torch.empty_like(yt)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def mixed(i):
    with torch.cuda.stream(s1): sp.batch_step(yp, 0.01, 1, out=op)
    with torch.cuda.stream(s2): st.batch_step(yt, 0.01, 1, out=ot)
for _ in range(3): mixed(0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): mixed(i)
s1.synchronize(); s2.synchronize(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
res.append({"system": "config3 mixed: 2M pendulum + 2M two_body on two streams", "N": 2 * Np, "ms_per_launch_pair": ms, "steps_per_s": 2 * Np / (ms * 1e-3)})
# evolve output path: 64 grid points, RK4 x4 substeps, 256K trajectories -> 64 x 8 MiB written
s = hb.systems.builtin(1)
y0 = s.batch_init_random(3, 0, 1 << 18, *BOXES["double_pendulum"][1:])
ts = np.linspace(0, 0.63, 64)
out = torch.empty((64,) + tuple(y0.shape), dtype=torch.float64, device=y0.device)
ms = timed(lambda i: s.batch_evolve(y0, ts, integ=L.RK4, rk4_substeps=4, out=out), 5)
res.append({"system": "double_pendulum evolve (64 outputs x 4 RK4 substeps)", "N": 1 << 18, "ms": ms, "steps_per_s": (1 << 18) * 63 * 4 / (ms * 1e-3),
            "output_GBps": out.numel() * 8 / (ms * 1e-3) / 1e9})
ms = timed(lambda i: s.batch_evolve(y0, ts, integ=L.RKF45_GSL, out=out), 3)
res.append({"system": "double_pendulum evolveHam RKF45_GSL (64 outputs, dt 0.01)", "N": 1 << 18, "ms": ms, "intervals_per_s": (1 << 18) * 63 / (ms * 1e-3)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/extras.json", "w"), indent=1)
for r in res: print(json.dumps(r))
