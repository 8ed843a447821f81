"""A/B of an engine compile-time switch on evolveHam (GSL RKF45 carried across a 64-point grid, dt 0.01) and on stepHam.
usage: HB_JIT_DEFINES=... python profiles/exp/exp_jit_evolve.py"""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES
sid, lo, hi = BOXES["double_pendulum"]
s = hb.systems.from_def(hb.systems.DEFS[sid]())
N = 1 << 18
y0 = s.batch_init_random(3, 0, N, lo, hi)
ts = np.linspace(0, 0.63, 64)
out = torch.empty((64,) + tuple(y0.shape), dtype=torch.float64, device=y0.device)
def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: s.batch_evolve(y0, ts, integ=L.RKF45_GSL, out=out), 5)
y1 = s.batch_init_random(4, 0, 1 << 20, lo, hi); o1 = torch.empty_like(y1)
ms2 = timed(lambda: s.batch_step(y1, 0.01, 1, integ=L.RKF45_GSL, out=o1), 10)
print("defines=%r: evolveHam 64 pts %.3f ms (%.4g intervals/s); stepHam %.4f ms (%.4g /s)" % (os.environ.get("HB_JIT_DEFINES", ""), ms, N * 63 / ms * 1e3, ms2, (1 << 20) / ms2 * 1e3))
