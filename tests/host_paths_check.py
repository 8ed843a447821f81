"""Run by tests/test_gpu_parity.py::test_host_paths_match_device_path in a fresh interpreter (the library reads
HB_HOST_DIRECT / HB_HOST_GRAPH once per process).  HB_MEM_HOST results must be bit-identical to the device-pointer path:
first call (graph capture where that applies), repeat call (replay), changed dt, flags round trip, pageable numpy
arrays, in-place stepping, ragged batch sizes, SOA, evolve output, and singular-matrix flags."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from tests.common import BOXES, SEED


def pin(t):
    return t.cpu().contiguous().pin_memory()


s = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)
lo, hi = BOXES["double_pendulum"][1:]
for N in (200_003, 1000, 31):                     # not multiples of the chunk count, the CTA size or the warp size
    y0 = s.batch_init_random(SEED + 5, 0, N, lo, hi)
    want = {dt: s.batch_step(y0, dt, 2, integ=L.RK4).cpu() for dt in (0.01, 0.02)}
    h_in, h_out, h_fl = pin(y0), pin(torch.empty_like(y0)), pin(torch.zeros(N, dtype=torch.int32))
    for dt in (0.01, 0.01, 0.02, 0.01):
        h_out.zero_()
        s.batch_step(h_in, dt, 2, integ=L.RK4, out=h_out, flags=h_fl)
        assert torch.equal(h_out, want[dt]), (N, dt)
        assert int(h_fl.sum()) == 0
    got = s.batch_step(y0.cpu().numpy(), 0.01, 2, integ=L.RK4)            # pageable memory
    assert np.array_equal(got, want[0.01].numpy())
    buf = pin(y0)                                                          # in place, twice
    s.batch_step(buf, 0.01, 2, integ=L.RK4, out=buf)
    assert torch.equal(buf, want[0.01])
    s.batch_step(buf, 0.01, 2, integ=L.RK4, out=buf)
    assert torch.equal(buf, s.batch_step(want[0.01].cuda(), 0.01, 2, integ=L.RK4).cpu())
    soa = pin(y0.t())
    out = s.batch_step(soa, 0.01, 2, integ=L.RK4, layout=L.SOA)
    assert torch.equal(out.t().contiguous(), want[0.01])
    rk = s.batch_step(h_in, 0.01, 1, integ=L.RKF45_GSL)
    assert torch.equal(rk, s.batch_step(y0, 0.01, 1, integ=L.RKF45_GSL).cpu())
    e_h, e_d = s.batch_energies(h_in), s.batch_energies(y0).cpu()
    assert torch.equal(torch.as_tensor(e_h), e_d)
    x_h, x_d = s.batch_underlying_pos(pin(y0[:, :2])), s.batch_underlying_pos(y0[:, :2].contiguous()).cpu()
    assert torch.equal(torch.as_tensor(x_h), x_d)

ts = np.linspace(0.0, 0.5, 6)
y0 = s.batch_init_random(SEED, 0, 5000, lo, hi)
ev_h = s.batch_evolve(pin(y0), ts, integ=L.RKF45_GSL)
ev_d = s.batch_evolve(y0, ts, integ=L.RKF45_GSL)
assert torch.equal(torch.as_tensor(ev_h), ev_d.cpu())

# singular mass matrices are flagged through every host path (flags are OR-ed into the caller's array)
tb = hb.systems.builtin(hb.systems.TWO_BODY)
lo2, hi2 = BOXES["two_body"][1:]
n2 = 1 << 17
t_in = pin(tb.batch_init_random(SEED, 0, n2, lo2, hi2))
t_in[5, 0] = 0.0
t_in[-1, 0] = 0.0                                  # r = 0: J^T W J singular
t_fl = pin(torch.zeros(n2, dtype=torch.int32))
t_fl[7] = 64                                       # a bit the caller set earlier must survive
tb.batch_step(t_in, 0.01, 1, integ=L.RK4, out=pin(torch.empty_like(t_in)), flags=t_fl)
assert sorted(torch.nonzero(t_fl).flatten().tolist()) == [5, 7, n2 - 1] and int(t_fl[7]) == 64

# a wide system (chain12: 24 doubles per Phase) takes the staged path even for page-locked buffers
ch = hb.systems.builtin(hb.systems.CHAIN12)
loc, hic = BOXES["chain12"][1:]
c0 = ch.batch_init_random(SEED, 0, 300, loc, hic)
assert torch.equal(torch.as_tensor(ch.batch_step(pin(c0), 0.01, 1, integ=L.RK4)), ch.batch_step(c0, 0.01, 1, integ=L.RK4).cpu())
print("host paths ok", {k: v for k, v in os.environ.items() if k.startswith("HB_HOST")})
