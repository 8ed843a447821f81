import torch, time, numpy as np, sys, os
sys.path.insert(0, '.')
import hamilton_b200 as hb
PI=np.pi
def bench(tag, s, N, lo, hi, layout, nsteps, reps=30):
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
    print(f"{tag:40s} N={N} nsteps={nsteps}: {ms:.4f} ms/launch  {sps:.3e} steps/s  hbm={sps*16*s.n*2/1e9:.1f} GB/s", flush=True)
box=([-PI,-PI,-1,-1],[PI,PI,1,1])
aot = hb.systems.builtin(1)
for ns in (1,16): bench("dp aot", aot, 1<<20, *box, 0, ns)
for defs in ("HB_MINB_RK4=1","HB_MINB_RK4=8","HB_MINB_RK4=10","HB_MINB_RK4=12"):
    os.environ["HB_JIT_DEFINES"]=defs
    s = hb.systems.from_def(hb.systems.double_pendulum_def())
    for ns in (1,16): bench("dp jit "+defs, s, 1<<20, *box, 0, ns)
