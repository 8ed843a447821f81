"""A/B of an engine compile-time switch on a JIT-compiled built-in definition.
usage: HB_JIT_DEFINES=... python profiles/exp/exp_jit_ab.py <system name> [log2N]"""
import sys, os
sys.path.insert(0, ".")
import torch
import hamilton_b200 as hb
from tests.common import BOXES
name = sys.argv[1]
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
sid, lo, hi = BOXES[name]
s = hb.systems.from_def(hb.systems.DEFS[sid]())
ring = 5
ins = [s.batch_init_random(7 + r, 0, N, lo, hi) for r in range(ring)]
outs = [torch.empty_like(b) for b in ins]
for nsteps in (1, 16):
    for _ in range(3): s.batch_step(ins[0], 0.01, nsteps, out=outs[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40 if nsteps == 1 else 5
    e0.record()
    for i in range(reps): s.batch_step(ins[i % ring], 0.01, nsteps, out=outs[i % ring])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s defines=%r nsteps=%d: %.4f ms/launch, %.4g steps/s" % (name, os.environ.get("HB_JIT_DEFINES", ""), nsteps, ms, N * nsteps / ms * 1e3))
