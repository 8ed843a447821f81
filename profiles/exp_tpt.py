import torch, numpy as np, sys, os
sys.path.insert(0, '.')
import hamilton_b200 as hb
PI=np.pi
def bench(tag, s, N, lo, hi, nsteps, reps=200, graph=True):
    R=9
    bufs = [s.batch_init_random(1+i, 0, N, lo, hi) for i in range(R)]
    outs = [torch.empty_like(b) for b in bufs]
    for i in range(3): s.batch_step(bufs[i%R], 0.01, nsteps, out=outs[i%R])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(reps): s.batch_step(bufs[i%R], 0.01, nsteps, out=outs[i%R])
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print(f"{tag:36s} N={N} nsteps={nsteps}: {ms:.4f} ms/launch  {N*nsteps/(ms*1e-3):.3e} steps/s", flush=True)
tag = "tpt=%s" % os.environ.get("HB_TRAJ_PER_THREAD","1")
box=([-PI,-PI,-1,-1],[PI,PI,1,1])
aot = hb.systems.builtin(1)
bench("dp aot graph "+tag, aot, 1<<20, *box, 1)
bench("dp aot graph "+tag, aot, 1<<20, *box, 16, reps=20)
bench("pendulum graph "+tag, hb.systems.builtin(0), 1<<21, [-PI,-1],[PI,1], 1)
bench("triple graph "+tag, hb.systems.builtin(6), 1<<20, [-PI]*3+[-1]*3,[PI]*3+[1]*3, 1, reps=50)
