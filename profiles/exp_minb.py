import torch, numpy as np, sys, os
sys.path.insert(0, '.')
import hamilton_b200 as hb
from profiles.exp_block import bench, box
for defs in ("", "HB_MINB_RK4=6", "HB_MINB_RK4=7", "HB_MINB_RK4=8", "HB_MINB_RK4=10"):
    if defs: os.environ["HB_JIT_DEFINES"] = defs
    s = hb.systems.from_def(hb.systems.double_pendulum_def())
    for ns in (1, 16): bench("dp jit [" + defs + "]", s, 1 << 20, *box, 0, ns)
