"""Multi-GPU ensembles: trajectories are independent (a pure function of each Phase,
src/Numeric/Hamilton.hs:390-399), so an ensemble of N initial conditions is block-split across GPUs with no exchange
during integration and a single all-gather to collect the final Phases.

Two front ends over the same block split:
  * `Ensemble` — binding of the C ABI's hb_ensemble_* (include/hamilton_b200.h): ONE process drives all GPUs, one host
    thread per device inside the library, ncclCommInitAll + one ncclAllGather on registered buffers.  This is what a
    Haskell / C++ host links against; no torch involved.
  * `shard` / `gather_final` / `run_ensemble` — one process per GPU under torchrun (torch.distributed, NCCL), the launch
    model bench.py's contract prescribes."""
import ctypes as C

import numpy as np

from . import _lib as L


class Ensemble:
    """hb_ensemble: N trajectories of `system` over `ndev` GPUs of this process (devices: ordinals, default 0..ndev-1)."""

    def __init__(self, system, n_total, ndev, devices=None):
        self.system, self.N, self.ndev, self.d = system, int(n_total), int(ndev), 2 * system.n
        dv = (C.c_int32 * ndev)(*devices) if devices is not None else None
        self._h = C.c_void_p()
        L.check(L.lib().hb_ensemble_create(system._h, ndev, dv, self.N, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().hb_ensemble_free(self._h)
            self._h = None

    __del__ = close

    def shards(self):
        first = (C.c_int64 * (self.ndev + 1))()
        L.check(L.lib().hb_ensemble_dims(self._h, None, None, first))
        return list(first)

    def init_random(self, seed, lo, hi):
        lo, hi = (np.ascontiguousarray(x, dtype=np.float64) for x in (lo, hi))
        assert lo.size == self.d and hi.size == self.d
        dp = C.POINTER(C.c_double)
        L.check(L.lib().hb_ensemble_init_random(self._h, seed, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp)))

    def upload(self, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.shape != (self.N, self.d):
            raise ValueError("expected an array of %d Phases of %d doubles" % (self.N, self.d))
        L.check(L.lib().hb_ensemble_upload(self._h, y.ctypes.data))

    def step(self, dt, nsteps=1, launches=1, integ=L.RK4):
        """`launches` x hb_batch_step(integ, dt, nsteps) on every shard; returns the device time in ms (max over devices)."""
        ms = C.c_double()
        L.check(L.lib().hb_ensemble_step(self._h, integ, dt, nsteps, launches, C.byref(ms)))
        return ms.value

    def gather(self, to_host=True):
        """One all-gather of the current Phases; returns (array of all N Phases on the host or None, device ms)."""
        ms = C.c_double()
        y = np.empty((self.N, self.d)) if to_host else None
        L.check(L.lib().hb_ensemble_gather(self._h, y.ctypes.data if to_host else None, C.byref(ms)))
        return y, ms.value

    def flags(self):
        f = np.zeros(self.N, dtype=np.int32)
        L.check(L.lib().hb_ensemble_flags(self._h, f.ctypes.data))
        return f


def shard(n_total, rank, world):
    """Contiguous block split: rank g owns [g*N/G, (g+1)*N/G).  Returns (first, count)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    first = n_total * rank // world
    return first, n_total * (rank + 1) // world - first


def gather_final(local, n_total=None, group=None, out=None):
    """All-gather of per-rank AOS batches [count_r, d] into the full [N, d] array on every rank
    (NCCL over NVLink on GPU tensors, gloo on CPU tensors).  Handles uneven shards by padding.
    `out` (optional, even shards only): preallocated [world * count, d] destination, so that a timed gather holds no allocation."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n_total is not None:      # block split known to every rank: no count exchange, no host synchronisation
        all_counts = [shard(n_total, r, world)[1] for r in range(world)]
        if all_counts[rank] != local.shape[0]:
            raise ValueError("local batch does not match shard(n_total, rank, world)")
    else:
        counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        all_counts = [int(c.item()) for c in all_counts]
    cmax = max(all_counts)
    if all(c == cmax for c in all_counts):
        shape = (world * cmax,) + tuple(local.shape[1:])
        if out is None or tuple(out.shape) != shape or out.dtype != local.dtype or out.device != local.device:
            out = torch.empty(shape, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, all_counts)], dim=0)


def run_ensemble(system, seed, n_total, lo, hi, dt, nsteps, integ=0, gather=True):
    """Generates this rank's shard on its GPU (counter-based RNG: no scatter), steps it, gathers the result."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    first, count = shard(n_total, rank, world)
    y = system.batch_init_random(seed, first, count, lo, hi)
    out = system.batch_step(y, dt, nsteps, integ=integ)
    return gather_final(out, n_total=n_total) if gather else out
