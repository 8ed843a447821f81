"""A/B of an engine compile-time switch on the reference-semantics kernel (stepHam = fresh GSL RKF45 solve per step).
usage: HB_JIT_DEFINES=... python profiles/exp/exp_jit_rkf45.py <system name> [log2N]"""
import sys, os
sys.path.insert(0, ".")
import torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES
name = sys.argv[1]
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
sid, lo, hi = BOXES[name]
s = hb.systems.from_def(hb.systems.DEFS[sid]())
ring = 5
ins = [s.batch_init_random(7 + r, 0, N, lo, hi) for r in range(ring)]
outs = [torch.empty_like(b) for b in ins]
for _ in range(2): s.batch_step(ins[0], 0.01, 1, integ=L.RKF45_GSL, out=outs[0])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for i in range(reps): s.batch_step(ins[i % ring], 0.01, 1, integ=L.RKF45_GSL, out=outs[i % ring])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("%s defines=%r stepHam(RKF45_GSL): %.4f ms/launch, %.4g stepHam/s" % (name, os.environ.get("HB_JIT_DEFINES", ""), ms, N / ms * 1e3))
