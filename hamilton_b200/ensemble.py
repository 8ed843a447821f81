"""Multi-GPU ensembles: trajectories are independent (a pure function of each Phase,
src/Numeric/Hamilton.hs:390-399), so an ensemble of N initial conditions is block-split across ranks
(one process per GPU, torch.distributed for the plumbing) with no exchange during integration and a
single all-gather to collect the final Phases."""


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
