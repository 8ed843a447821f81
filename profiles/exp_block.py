import torch, numpy as np, sys, os
sys.path.insert(0, '.')
import hamilton_b200 as hb
PI=np.pi
def bench(tag, s, N, lo, hi, layout, nsteps, reps=40, graph=False):
    R=9
    bufs = [s.batch_init_random(1+i, 0, N, lo, hi, layout=layout) for i in range(R)]
    outs = [torch.empty_like(b) for b in bufs]
    for i in range(3): s.batch_step(bufs[i%R], 0.01, nsteps, out=outs[i%R], layout=layout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                for i in range(reps): s.batch_step(bufs[i%R], 0.01, nsteps, out=outs[i%R], layout=layout)
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    else:
        e0.record()
        for i in range(reps): s.batch_step(bufs[i%R], 0.01, nsteps, out=outs[i%R], layout=layout)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    sps = N*nsteps/(ms*1e-3)
    print(f"{tag:44s} N={N} nsteps={nsteps}: {ms:.4f} ms/launch  {sps:.3e} steps/s  hbm={sps*16*s.n*2/1e9:.1f} GB/s", flush=True)
box=([-PI,-PI,-1,-1],[PI,PI,1,1])
if __name__ != "__main__": sys.argv = [sys.argv[0]]
tag = "blk=%s" % os.environ.get("HB_LAUNCH_BLOCK","128")
aot = hb.systems.builtin(1)
bench("dp aot "+tag, aot, 1<<20, *box, 0, 1)
bench("dp aot "+tag, aot, 1<<20, *box, 0, 16)
if len(sys.argv) > 1:
    try:
        bench("dp aot graph "+tag, aot, 1<<20, *box, 0, 1, graph=True)
    except Exception as e:
        print("graph failed:", repr(e)[:300])
    jit = hb.systems.from_def(hb.systems.double_pendulum_def())
    bench("dp jit "+tag, jit, 1<<20, *box, 0, 1)
    bench("dp jit "+tag, jit, 1<<20, *box, 1, 1)
    bench("dp jit "+tag, jit, 1<<22, *box, 0, 1)
    for name, sid, lo, hi, lg in (("pendulum",0,[-PI,-1],[PI,1],21),("two_body",3,[1,-PI,-1,1],[3,PI,1,5],21),("triple",6,[-PI]*3+[-1]*3,[PI]*3+[1]*3,20),("chain12",7,[-PI]*12+[-1]*12,[PI]*12+[1]*12,18)):
        s = hb.systems.builtin(sid)
        bench(name+" aot", s, 1<<lg, lo, hi, 0, 1, reps=10)
        bench(name+" aot rkf45", s, 1<<lg, lo, hi, 0, 1, reps=3) if False else None
