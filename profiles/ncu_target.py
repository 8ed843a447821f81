"""Profiling target: double pendulum (System 4 2), batch 1,048,576, RK4 — the BASELINE config-2 kernel.
usage: python profiles/ncu_target.py [nsteps_per_launch] [launches] [system_id] [log2N]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import hamilton_b200 as hb
from tests.common import BOXES
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sid = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 20)
name = [k for k, v in BOXES.items() if v[0] == sid][0]
s = hb.systems.builtin(sid)
y = s.batch_init_random(0x48414D49, 0, N, BOXES[name][1], BOXES[name][2])
out = torch.empty_like(y)
for _ in range(launches):
    s.batch_step(y, 0.01, nsteps, out=out)
torch.cuda.synchronize()
print("done", name, N, nsteps)
