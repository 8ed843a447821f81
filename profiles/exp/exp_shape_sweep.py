"""Launch-shape sweep of the HBM-bound launches (one-evaluation kernels, light stepping kernels) through the library's
experiment knobs HB_BLOCK / HB_GRID_WAVES / HB_CONTIGUOUS (read at library load: one variant per subprocess).
  python profiles/exp/exp_shape_sweep.py run <system> <log2N>      (worker)
  python profiles/exp/exp_shape_sweep.py sweep <system> <log2N>    (driver)
Timing: one CUDA graph of 200 launches over a ring of buffers larger than L2, best of 3 replays."""
import os, subprocess, sys
VARIANTS = [("default", {})]
FULL = os.environ.get("HB_SWEEP_FULL")
for b in (128, 256, 512):
    for w in (1, 2, 4):
        for c in (0, 1):
            if not FULL and not (c == 1 and (w, b) in ((1, 128), (4, 128), (1, 256), (2, 256), (1, 512), (2, 512), (4, 512))) and not (c == 0 and (w, b) == (1, 512)): continue
            VARIANTS.append(("block %d, %d wave(s), %s" % (b, w, "contiguous" if c else "spread"), {"HB_BLOCK": str(b), "HB_GRID_WAVES": str(w), "HB_CONTIGUOUS": str(c)}))
if os.environ.get("HB_SWEEP_WAVES"):   # large systems: only the number of waves is a knob (grid = min(tiles, 2 x waves x resident CTAs))
    VARIANTS = [("default", {})] + [("HB_GRID_WAVES=%s" % w, {"HB_GRID_WAVES": w}) for w in os.environ["HB_SWEEP_WAVES"].split(",")]
def worker(name, log2n):
    sys.path.insert(0, ".")
    import torch
    import hamilton_b200 as hb
    from tests.common import BOXES
    N = 1 << log2n
    sid, lo, hi = BOXES[name]
    s = hb.systems.builtin(sid)
    ring = max(2, int(600e6 // (2 * N * 2 * s.n * 8)) + 1)
    ins = [s.batch_init_random(7 + r, 0, N, lo, hi) for r in range(ring)]
    outs = [torch.empty_like(b) for b in ins]
    K = int(os.environ.get('HB_SWEEP_K', '200'))
    def graph_time(body):
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
        return best / K * 1e3
    for r in range(2): s.batch_ham_eqs(ins[r], out=outs[r]); s.batch_step(ins[r], 0.01, 1, out=outs[r]); s.batch_from_phase(ins[r], out=outs[r])
    torch.cuda.synchronize()
    t_h = graph_time(lambda i: s.batch_ham_eqs(ins[i % ring], out=outs[i % ring]))
    t_f = graph_time(lambda i: s.batch_from_phase(ins[i % ring], out=outs[i % ring]))
    t_s = graph_time(lambda i: s.batch_step(ins[i % ring], 0.01, 1, out=outs[i % ring]))
    gb = N * 32 * s.n / 1e3
    print("RESULT hamEqs %.2f us (%.0f GB/s)  fromPhase %.2f us (%.0f GB/s)  step_rk4 %.2f us (%.0f GB/s)" % (t_h, gb / t_h, t_f, gb / t_f, t_s, gb / t_s))
if sys.argv[1] == "run":
    worker(sys.argv[2], int(sys.argv[3]))
else:
    name, log2n = sys.argv[2], sys.argv[3]
    for label, env in VARIANTS:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, __file__, "run", name, log2n], env=e, capture_output=True, text=True, timeout=300)
        res = [l for l in p.stdout.split("\n") if l.startswith("RESULT")]
        print("%-36s %s" % (label, res[0][7:] if res else "FAILED: " + (p.stderr.strip().split("\n") or ["?"])[-1][:200]), flush=True)
