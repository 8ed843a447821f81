"""e2e leg of bench.py alone: hb_batch_step(HB_MEM_HOST) on pinned host arrays, config 2 (1,048,576 double-pendulum Phases).
usage: [HB_HOST_CHUNKS=c] [HB_HOST_GRAPH=0|1] python profiles/exp/exp_e2e.py [log2N] [steps]"""
import os, sys, time
sys.path.insert(0, ".")
import torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sid, lo, hi = BOXES["double_pendulum"]
s = hb.systems.builtin(sid)
src = s.batch_init_random(7, 0, N, lo, hi).cpu()
h_in = [src.clone().pin_memory() for _ in range(2)]
h_out = [torch.empty_like(src).pin_memory() for _ in range(2)]
for i in range(4):
    s.batch_step(h_in[i % 2], 0.01, 1, integ=L.RK4, out=h_out[i % 2])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(steps):
        s.batch_step(h_in[i % 2], 0.01, 1, integ=L.RK4, out=h_out[i % 2])
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / steps)
print("N=%d direct=%s chunks=%s graph=%s: %.3f ms/call, %.4g steps/s, %.1f GB/s each way" % (
    N, os.environ.get("HB_HOST_DIRECT", "1"), os.environ.get("HB_HOST_CHUNKS", "auto"), os.environ.get("HB_HOST_GRAPH", "1"), best * 1e3, N / best, N * 32 / best / 1e9))
