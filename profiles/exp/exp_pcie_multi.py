"""Is the host the ceiling of the multi-GPU e2e leg?  Under torchrun, every rank moves 32 MiB host->device and 32 MiB
device->host at once with the COPY ENGINES only (no kernel of ours involved), all ranks together, and alone.
  python -m torch.distributed.run --nproc-per-node N profiles/exp/exp_pcie_multi.py
Prints the aggregate GB/s each way: if it saturates at the same level as the e2e leg of bench.py, the host memory / PCIe
root complex is the limit, not the zero-copy kernels."""
import os, time
import torch
import torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import pynvml
    pynvml.nvmlInit()
    uuid = str(torch.cuda.get_device_properties(local).uuid)
    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
    pynvml.nvmlDeviceSetCpuAffinity(h)
except Exception:
    pass
n = 32 << 20
hin = torch.empty(n, dtype=torch.uint8).pin_memory(); hout = torch.empty(n, dtype=torch.uint8).pin_memory()
din = torch.empty(n, dtype=torch.uint8, device="cuda"); dout = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(reps):
    for _ in range(reps):
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
def timed(active):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active: both(40)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    return time.perf_counter() - t0
both(5)
t_all = timed(True)
t_solo = timed(rank == 0)
if rank == 0:
    print("%d ranks at once : %.1f GB/s each way per rank, %.1f GB/s each way aggregate" % (world, 40 * n / t_all / 1e9, world * 40 * n / t_all / 1e9))
    print("rank 0 alone     : %.1f GB/s each way" % (40 * n / t_solo / 1e9))
    print("cpus of rank 0   : %s" % sorted(os.sched_getaffinity(0))[:4], "... (%d)" % len(os.sched_getaffinity(0)))
if world > 1:
    dist.destroy_process_group()
