#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its configuration: phase-space RK4 steps/sec of the double pendulum
(System 4 2), batch 1,048,576 random initial Phases per GPU, fp64 (configs[1]).

A bench "step" is ONE pass of the hot path over one batch: one `hb_batch_step(RK4, dt=0.01, nsteps=1)` call that reads
every Phase of the batch from HBM, advances it by one classical RK4 step (4 hamEqs evaluations) and writes it back —
the I/O-honest mode SURVEY.md §8(d) defines (32·n = 64 algorithmic bytes per trajectory-step), so `value` and
`roofline` describe the same launches.  Batches rotate through a ring of buffers larger than L2 (see config.l2).

  value     steps/s with the batch resident in HBM (device pointers through the C ABI), CUDA-event timed.
  e2e       the same call through the C ABI with HOST buffers (pinned): H2D + kernel + D2H inside the timed region.
  roofline  HBM roofline of the dominant kernel (hbk_double_pendulum_dflt_step_rk4) + the FP64-pipe view that actually binds.
  cpu_baseline  the CPU oracle (restatement of the reference algorithm) timed on this box's host cores, bounded sample.

`--impl reference` times the reference's CPU implementation of the path: the Haskell+GSL binary cannot be built in this
image (no GHC/GSL), so it is the oracle port (oracle/hamilton_oracle.c) on all host threads.

N > 1 (torchrun, one rank per GPU): trajectories are independent, so each rank owns its own 1,048,576-trajectory shard
(weak scaling, no data-path collective); the single NCCL all-gather that collects the final Phases is executed after the
timed steps and reported separately (gather_ms, value_with_gather).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("HB_JIT_CACHE_DIR", os.path.join(ROOT, ".jit_cache", "gpu" if os.path.exists("/dev/nvidiactl") else "cpu"))   # NVRTC output cache stays inside the repository
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PER_GPU = 1 << 20
DT = 0.01
SEED = 0x48414D49
LO = [-np.pi, -np.pi, -1.0, -1.0]
HI = [np.pi, np.pi, 1.0, 1.0]
ALGO_BYTES_PER_STEP = 64          # 32·n, n = 2
RING = 9                          # (in,out) buffer pairs: 9 × 64 MiB = 576 MiB touched per cycle >> 126 MB L2
METRIC = "phase-space RK4 steps/sec (batched trajectories)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region is tens of milliseconds long, so NVML is polled
    from a thread every ~1 ms (nvidia-smi -lms cannot sample that fast); falls back to nvidia-smi when NVML is missing."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.h, self.stop_flag, self.mx = index, [], None, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), clk, mask))
            except Exception:
                pass
            time.sleep(0.0005)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                mask = sum(bit for (name, bit), v in zip((("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                                                          ("sw_power_cap", 0x4)), f[3:7]) if v.lower().startswith("active"))
                self.mx = float(f[1])
                self.rows.append((time.time(), float(f[0]), mask))
            except (ValueError, IndexError):
                pass

    def stop(self, t0, t1):
        if self.h is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        if self.h is None:
            time.sleep(0.05)
            self.proc.terminate()
        self.stop_flag = True
        inside = [(c, m) for ts, c, m in self.rows if t0 <= ts <= t1]
        where = "timed region"
        if not inside:   # region shorter than the sampling period: nearest samples around it
            inside = [(c, m) for ts, c, m in self.rows if t0 - 0.1 <= ts <= t1 + 0.1]
            where = "timed region +- 100 ms"
        reasons = set()
        for _c, m in inside:
            for name, bit in self.REASONS:
                if m & bit:
                    reasons.add(name)
        sm = [c for c, _m in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons), "samples": len(sm),
                "sampled": ("NVML every ~1 ms, " if self.h is not None else "nvidia-smi -lms 20, ") + where}


def bind_near_gpu(index):
    """Pins this process to the CPUs NVML reports as local to GPU `index` (one rank per GPU: without it the ranks' host
    buffers pile up on one socket and the e2e leg measures the inter-socket link).  Returns a short description."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return "%d cpus near gpu %d (%d..%d)" % (len(cpus), index, cpus[0], cpus[-1])
    except Exception as ex:
        return "unbound (%s)" % type(ex).__name__


def cpu_oracle_steps_per_sec(target_seconds, threads):
    """Times the oracle's RK4 on a bounded sample of the same workload (same RNG stream, first trajectories)."""
    from oracle import oracle as O
    S = O.OracleSystem.builtin(O.DOUBLE_PENDULUM)
    n_s = 16384 * max(1, threads)
    y = S.init_random(SEED, 0, n_s, LO, HI)
    t = time.perf_counter(); S.batch_step(y, 0, DT, 1, threads=threads); cal = time.perf_counter() - t
    reps = max(1, min(64, int(target_seconds / max(cal, 1e-3))))
    t = time.perf_counter(); _, bad = S.batch_step(y, 0, DT, reps, threads=threads); el = time.perf_counter() - t
    return n_s * reps / el, {"trajectories": n_s, "rk4_steps_each": reps, "seconds": round(el, 3), "failed": int(bad)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; Haskell+GSL unbuildable here), rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.max_threads()
    S = O.OracleSystem.builtin(O.DOUBLE_PENDULUM)
    n_s = 16384 * threads                       # bounded sample of the 1,048,576-trajectory batch
    y = S.init_random(SEED, 0, n_s, LO, HI)
    for _ in range(args.warmup):
        S.batch_step(y, 0, DT, 1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y, _bad = S.batch_step(y, 0, DT, 1, threads=threads)
    el = time.perf_counter() - t0
    v = n_s * args.steps / el
    sample = "%d of the %d trajectories per step (same splitmix64 stream), %d RK4 steps, %d host threads" % (n_s, N_PER_GPU, args.steps, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "double pendulum (System 4 2), batch 1,048,576 random Phases, RK4 dt=0.01 — CPU sample", "integrator": "rk4"},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference algorithm (hamEqs via dense forward-mode jets + explicit inverse, classical RK4); "
                "the ad+hmatrix+GSL Haskell binary cannot be built in this image (no GHC, no libgsl)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph of K launches")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import hamilton_b200 as hb
    from hamilton_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(local_rank)       # page-locked host buffers of the e2e leg land on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sysm = hb.systems.builtin(hb.systems.DOUBLE_PENDULUM)          # ahead-of-time sm_100a kernels, m1 = m2 = 1
    N = N_PER_GPU
    first = rank * N                                                # contiguous block split of the global ensemble
    ring_in = [sysm.batch_init_random(SEED + r, first, N, LO, HI) for r in range(RING)]
    ring_out = [torch.empty_like(b) for b in ring_in]
    flags = torch.zeros(N, dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        sysm.batch_step(ring_in[i % RING], DT, 1, integ=L.RK4, out=ring_out[i % RING], flags=flags)

    # ---------------- device-resident throughput (value, roofline) ----------------
    for i in range(args.warmup):
        step(i)
    barrier()
    # The K timed launches are captured once into a CUDA graph (the C ABI is stream-ordered and capture-safe), so the
    # timed region holds exactly K kernel launches and no per-call host overhead; falls back to eager launches.
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                with torch.cuda.graph(graph, stream=cap):
                    for i in range(args.steps):
                        step(i)
            torch.cuda.current_stream().wait_stream(cap)
        except Exception as ex:   # pragma: no cover
            sys.stderr.write("bench: CUDA-graph capture failed (%r); timing eager launches\n" % (ex,))
            graph = None
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(args.steps):
            step(i)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    assert int(flags.sum().item()) == 0, "numerical failure flags raised during the bench"

    # explanation only: the same kernel with 16 RK4 steps fused per launch (state stays in registers between steps)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sysm.batch_step(ring_in[0], DT, 16, integ=L.RK4, out=ring_out[0])
    f0.record()
    for i in range(8):
        sysm.batch_step(ring_in[i % RING], DT, 16, integ=L.RK4, out=ring_out[i % RING])
    f1.record()
    torch.cuda.synchronize()
    fused_value = world * N * 16 * 8 / (f0.elapsed_time(f1) * 1e-3)

    # ---------------- final collection: one NCCL all-gather of the final Phases ----------------
    gather_ms = 0.0
    if world > 1:
        final = ring_out[(args.steps - 1) % RING]
        allp = torch.empty((world * N, final.shape[1]), dtype=final.dtype, device=dev)
        dist.all_gather_into_tensor(allp, final)                  # warm the communicator
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); dist.all_gather_into_tensor(allp, final); g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    # ---------------- end to end through the C ABI with host buffers ----------------
    h_in = [torch.empty((N, 4), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_out = [torch.empty((N, 4), dtype=torch.float64).pin_memory() for _ in range(2)]
    for b, src in zip(h_in, ring_in):
        b.copy_(src.cpu())
    e2e_steps = max(3, min(args.steps, 50))
    for i in range(3):
        sysm.batch_step(h_in[i % 2], DT, 1, integ=L.RK4, out=h_out[i % 2])
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        sysm.batch_step(h_in[i % 2], DT, 1, integ=L.RK4, out=h_out[i % 2])   # blocking: returns when the result is on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    chk = float(h_out[0][0, 0])                                    # result is read on the host
    assert np.isfinite(chk)

    # ---------------- reduce over ranks (max time) ----------------
    times = torch.tensor([ms, gather_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, gather_ms, e2e_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        total_steps = world * N * args.steps
        value = total_steps / (ms * 1e-3)
        launch_ms = ms / args.steps
        peak, peak_src = peaks()
        achieved = N * ALGO_BYTES_PER_STEP / (launch_ms * 1e-3) / 1e9
        traffic = None
        fp64 = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_r1.json")) as f:
                prof = json.load(f)
            traffic = prof.get("dram_bytes_per_launch")
            fp64 = prof.get("fp64")
            # live FP64 view: issue cycles the step needs on the FP64 pipe (static SASS count, 3-operand DFMAs at 3 clocks)
            # against the cycles the schedulers had — the resource that actually binds this kernel
            cyc = fp64.get("fp64_issue_cycles_per_rk4_step") if fp64 else None
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
            if cyc and mhz:
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                bound = sms * 4 * 32 * mhz * 1e6 / cyc                # steps/s per GPU at 100 % FP64 issue
                fp64 = dict(fp64, fp64_bound_steps_per_s_at_measured_clock=bound, achieved_frac_of_fp64_bound=value / world / bound)
        except Exception:
            pass
        cpu_v = cpu_sample = cpu_threads = None
        if world == 1:                                            # rank 0 at N=1 only
            from oracle import oracle as O
            os.sched_setaffinity(0, all_cpus)                     # the CPU baseline uses every host core again
            cpu_threads = O.max_threads()
            cpu_v, cpu_sample = cpu_oracle_steps_per_sec(args.cpu_seconds, cpu_threads)
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "double pendulum (System 4 2), batch 1,048,576 random initial Phases per GPU, RK4 dt=0.01, fp64 (BASELINE configs[1])",
                       "batch_per_gpu": N, "global_batch": world * N, "integrator": "rk4", "steps_per_launch": 1, "layout": "AOS (array of Phases)",
                       "launch": "cuda_graph of K kernel launches" if graph is not None else "eager",
                       "parallelism": "%d independent shards" % world,
                       "l2": "inputs larger than L2: ring of %d (in,out) batch pairs = %d MiB touched per cycle" % (RING, RING * 64)},
            "clocks": clocks,
            "gpu_launches": args.steps,
            "e2e": {"value": world * N * e2e_steps / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": N * 32, "d2h_bytes_per_step": N * 32,
                    "steps": e2e_steps, "call": "hb_batch_step(HB_INTEG_RK4, nsteps=1, HB_MEM_HOST) on pinned host arrays, blocking: one kernel reads the Phases from and writes the results to host memory over PCIe (DESIGN.md section 2)",
                    "host_binding": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "hbk_double_pendulum_dflt_step_rk4",
                         "algorithmic_bytes_per_launch": N * ALGO_BYTES_PER_STEP,
                         "note": "the binding resource is the FP64 pipe, not HBM (SURVEY.md §8(d)); see fp64", "fp64": fp64},
        }
        out["fused16"] = {"value": fused_value, "unit": "steps/s", "note": "16 RK4 steps per launch, no per-step HBM traffic (FP64-pipe view)"}
        if world > 1:
            out["gather_ms"] = gather_ms
            out["value_with_gather"] = total_steps / ((ms + gather_ms) * 1e-3)
        if cpu_v is not None:
            out["cpu_baseline"] = {"value": cpu_v, "unit": "steps/s", "cores": cpu_threads, "kind": "port",
                                   "sample": "%(trajectories)d trajectories x %(rk4_steps_each)d RK4 steps in %(seconds)s s (oracle/hamilton_oracle.c, pthreads)" % cpu_sample}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
