"""Profiling target: reference-semantics stepHam (adaptive GSL RKF45) of the double pendulum, batch 1,048,576.
usage: python profiles/ncu_target_rkf45.py [launches]"""
import sys
sys.path.insert(0, ".")
import torch
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES
s = hb.systems.builtin(1)
y = s.batch_init_random(0x48414D49, 0, 1 << 20, BOXES["double_pendulum"][1], BOXES["double_pendulum"][2])
out = torch.empty_like(y)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    s.batch_step(y, 0.01, 1, integ=L.RKF45_GSL, out=out)
torch.cuda.synchronize()
print("done")
