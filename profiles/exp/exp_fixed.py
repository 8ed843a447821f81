"""Per-launch fixed cost of the RK4 step kernel: time per launch (CUDA graph of 200 dependent launches, ping-pong buffers
so every launch consumes the previous launch's output) against batch size and steps per launch.
T(N, s) ~ a + N * (b * s + c):  a = launch/ramp/drain, c = per-trajectory load/store/bookkeeping, b = one RK4 step."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES
name = sys.argv[1] if len(sys.argv) > 1 else "double_pendulum"
sid, lo, hi = BOXES[name]
s = hb.systems.builtin(sid)
res = []
for N in (1 << 12, 1 << 14, 1 << 16, 113664, 227328, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    a = s.batch_init_random(1, 0, N, lo, hi); b = torch.empty_like(a)
    row = []
    for ns in (1, 2, 4):
        for _ in range(3): s.batch_step(a, 1e-3, ns, out=b)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                for k in range(100):
                    s.batch_step(a, 1e-3, ns, out=b); s.batch_step(b, 1e-3, ns, out=a)
        torch.cuda.current_stream().wait_stream(st); torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 200 * 1e3)
    res.append((N, row))
    print("N=%8d  us/launch: 1 step %.2f   2 steps %.2f   4 steps %.2f   -> per step %.2f, per launch beyond the steps %.2f" % (
        N, row[0], row[1], row[2], (row[2] - row[0]) / 3, row[0] - (row[2] - row[0]) / 3))
