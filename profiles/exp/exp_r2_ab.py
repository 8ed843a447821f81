"""Round-2 A/B of engine compile-time switches on the NVRTC build of a built-in definition (same engine source as the
ahead-of-time kernels).  One variant per subprocess: HB_JIT_DEFINES / HB_BLOCK / HB_GRID_WAVES are read at library load.
  python profiles/exp/exp_r2_ab.py run <system> [log2N]       (worker)
  python profiles/exp/exp_r2_ab.py sweep <system> [log2N]     (driver: prints one line per variant)
Timing: a CUDA graph of K one-step launches over a ring of 9 (in, out) buffer pairs (inputs larger than L2, as bench.py),
a graph of K launches ping-ponging between two buffers (a real stepping loop: L2-resident), and 16 steps fused per launch."""
import os, subprocess, sys
VARIANTS = [
    ("default", {}),
    ("round-1 final build", {"HB_LIB_PATH": "profiles/ab_libs/lib_r1_final.so"}),
    ("round-2 build 8af0be2 (unrolled stages)", {"HB_LIB_PATH": "profiles/ab_libs/lib_r2_8af0be2.so"}),
]
for kv in filter(None, os.environ.get("HB_AB_LIBS", "").split(",")):   # extra saved builds: HB_AB_LIBS=label=path,label=path
    VARIANTS.append((kv.split("=")[0], {"HB_LIB_PATH": kv.split("=", 1)[1]}))
def worker(name, log2n):
    sys.path.insert(0, ".")
    import torch
    import hamilton_b200 as hb
    from tests.common import BOXES
    N = 1 << log2n
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid) if os.environ.get("HB_AB_BUILTIN") else hb.systems.from_def(hb.systems.DEFS[sid]())
    fl = torch.zeros(N, dtype=torch.int32, device="cuda") if os.environ.get("HB_AB_FLAGS") else None
    ring = 9
    ins = [s.batch_init_random(7 + r, 0, N, lo, hi) for r in range(ring)]
    outs = [torch.empty_like(b) for b in ins]
    K = int(os.environ.get('HB_AB_K', '200'))
    def timed(fn, reps=3):
        if os.environ.get("HB_AB_MEAN"):          # mean over back-to-back repetitions (bench.py's way) instead of the best of 3
            reps = int(os.environ["HB_AB_MEAN"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    def graph_of(body):
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                body()
        torch.cuda.current_stream().wait_stream(st)
        g.replay(); torch.cuda.synchronize()
        return g
    for i in range(3): s.batch_step(ins[i], 0.01, 1, out=outs[i])
    torch.cuda.synchronize()
    g_ring = graph_of(lambda: [s.batch_step(ins[i % ring], 0.01, 1, out=outs[i % ring], flags=fl) for i in range(K)])
    a, b = ins[0].clone(), outs[0]
    g_chain = graph_of(lambda: [s.batch_step(a if i % 2 == 0 else b, 0.01, 1, out=b if i % 2 == 0 else a) for i in range(K)])
    t_ring = timed(g_ring.replay) / K
    t_chain = timed(g_chain.replay) / K
    t_16 = timed(lambda: [s.batch_step(ins[i % ring], 0.01, 16, out=outs[i % ring]) for i in range(8)]) / 8
    from hamilton_b200 import _lib as LL
    t_rkf = timed(lambda: [s.batch_step(ins[i % ring], 0.01, 1, integ=LL.RKF45_GSL, out=outs[i % ring]) for i in range(4)]) / 4
    # e2e: the same call on pinned HOST arrays (blocking), as bench.py's e2e leg
    import time
    src = ins[0].cpu()
    h_in = [src.clone().pin_memory() for _ in range(2)]
    h_out = [torch.empty_like(src).pin_memory() for _ in range(2)]
    for i in range(3): s.batch_step(h_in[i % 2], 0.01, 1, out=h_out[i % 2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20): s.batch_step(h_in[i % 2], 0.01, 1, out=h_out[i % 2])
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / 20
    print("RESULT ring %.3f us/launch (%.3e steps/s)  chain %.3f us  fused16 %.3e  stepHam(RKF45) %.3e /s  e2e %.3f ms/call (%.3e steps/s)" %
          (t_ring * 1e3, N / t_ring * 1e3, t_chain * 1e3, N * 16 / t_16 * 1e3, N / t_rkf * 1e3, t_e2e * 1e3, N / t_e2e))
if sys.argv[1] == "run":
    worker(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20)
else:
    name, log2n = sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "20")
    only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    for label, env in VARIANTS:
        if only and not any(o in label for o in only): continue
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, __file__, "run", name, log2n], env=e, capture_output=True, text=True, timeout=600)
        res = [l for l in p.stdout.split("\n") if l.startswith("RESULT")]
        print("%-42s %s" % (label, res[0][7:] if res else "FAILED: " + (p.stderr.strip().split("\n") or ["?"])[-1][:200]), flush=True)
