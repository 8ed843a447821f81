import torch, time, numpy as np, sys
sys.path.insert(0, '.')
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
def bench(name, sid, N, lo, hi, layout, nsteps, reps=20, jit=False):
    s = hb.systems.from_def(hb.systems.DEFS[sid]()) if jit else hb.systems.builtin(sid)
    bufs = [s.batch_init_random(1, 0, N, lo, hi, layout=layout) for _ in range(6)]
    outs = [torch.empty_like(b) for b in bufs]
    for i in range(3): s.batch_step(bufs[i%6], 0.01, nsteps, out=outs[i%6], layout=layout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): s.batch_step(bufs[i%6], 0.01, nsteps, out=outs[i%6], layout=layout)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    sps = N*nsteps/(ms*1e-3)
    print(f"{name:18s} jit={jit} layout={layout} N={N} nsteps={nsteps}: {ms:.4f} ms/launch  {sps:.3e} steps/s  hbm={sps*16*s.n*2/1e9:.1f} GB/s", flush=True)
PI=np.pi
for jit in (False, True):
  for layout in (0,1):
    for nsteps in (1, 16):
        bench("double_pendulum", 1, 1<<20, [-PI,-PI,-1,-1],[PI,PI,1,1], layout, nsteps, jit=jit)
bench("pendulum", 0, 1<<21, [-PI,-1],[PI,1], 0, 1)
bench("two_body", 3, 1<<21, [1,-PI,-1,1],[3,PI,1,5], 0, 1)
bench("triple", 6, 1<<20, [-PI]*3+[-1]*3,[PI]*3+[1]*3, 0, 1)
bench("chain12", 7, 1<<18, [-PI]*12+[-1]*12,[PI]*12+[1]*12, 0, 1, reps=5)
bench("chain12", 7, 1<<18, [-PI]*12+[-1]*12,[PI]*12+[1]*12, 1, 1, reps=5)
