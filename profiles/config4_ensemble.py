"""BASELINE configs[3]: triple pendulum (System 6 3), 8,388,608 initial conditions block-sharded across the ranks of one
node, K RK4 steps with no data-path collective, ONE NCCL all-gather of the final Phases at the end.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P profiles/config4_ensemble.py [steps]
Prints one JSON line from rank 0 (device-timed, max over ranks) and verifies the gathered array against a
recomputation of another rank's first trajectories."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import hamilton_b200 as hb
from hamilton_b200 import _lib as L
from hamilton_b200 import ensemble
from tests.common import BOXES, SEED

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_TOTAL = 8_388_608
sid, lo, hi = BOXES["triple_pendulum"]
s = hb.systems.builtin(sid)
first, count = ensemble.shard(N_TOTAL, rank, world)
y = s.batch_init_random(SEED, first, count, lo, hi)
a, b = y, torch.empty_like(y)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    s.batch_step(a, 0.01, 1, integ=L.RK4, out=b)
full = ensemble.gather_final(b, n_total=N_TOTAL)
for _ in range(3):                       # warm the communicator at this message size
    full = ensemble.gather_final(b, n_total=N_TOTAL, out=full)
sync()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
src, dst = y, b
scratch = torch.empty_like(y)
for k in range(steps):                   # step k+1 consumes step k's output (ping-pong), 1 step per launch
    s.batch_step(src, 0.01, 1, integ=L.RK4, out=dst)
    src, dst = dst, (scratch if dst is b else b)
e1.record()
full = ensemble.gather_final(src, n_total=N_TOTAL, out=full)
e2.record()
sync()
t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
step_ms, gather_ms = t.tolist()
# verification: recompute the first 4096 trajectories of the NEXT rank's shard here and compare with the gathered rows
other = (rank + 1) % world
ofirst, _ = ensemble.shard(N_TOTAL, other, world)
chk = s.batch_init_random(SEED, ofirst, 4096, lo, hi)
chk = s.batch_step(chk, 0.01, steps, integ=L.RK4)
ok = torch.equal(chk, full[ofirst:ofirst + 4096])
okt = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"config": "triple pendulum (System 6 3), %d trajectories over %d GPU(s), %d RK4 steps dt=0.01, 1 step/launch, all-gather at end" % (N_TOTAL, world, steps),
                      "n_gpus": world, "steps_per_s": N_TOTAL * steps / (step_ms * 1e-3), "step_ms_total": step_ms, "gather_ms": gather_ms,
                      "gather_GBps_bus": (world - 1) / max(world, 1) * N_TOTAL * 48 / (gather_ms * 1e-3) / 1e9 if world > 1 else None,
                      "steps_per_s_with_gather": N_TOTAL * steps / ((step_ms + gather_ms) * 1e-3),
                      "gathered_rows_match_recomputation": bool(int(okt.item()))}))
if world > 1:
    dist.destroy_process_group()
