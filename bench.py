#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: phase-space RK4 steps/sec of batched trajectories, fp64.

A bench "step" is ONE pass of the hot path over one batch: one `hb_batch_step(RK4, dt=0.01, nsteps=1)` call that reads
every Phase of the batch from HBM, advances it by one classical RK4 step (4 hamEqs evaluations) and writes it back —
the I/O-honest mode SURVEY.md §8(d) defines (32·n algorithmic bytes per trajectory-step), so `value` and `roofline`
describe the same launches.  Batches rotate through a ring of buffers larger than L2 (see config.l2).

Headline workload (N = 1 and the per-GPU shard at N > 1): BASELINE configs[1] — double pendulum (System 4 2), batch
1,048,576 random initial Phases per GPU.  K (--steps) launches estimate the rate; the timed region is ONE replay of a CUDA graph of
L = K * ceil(100 ms / (K launches)) launches (ms_per_step = region / L) — long enough for the clock sampler, and free of the
gaps between consecutive replays of one executable graph, which are launch plumbing.

  value      steps/s with the batch resident in HBM (device pointers through the C ABI), CUDA-event timed, max over ranks.
  e2e        the same call through the C ABI with HOST buffers (pinned): H2D + kernel + D2H inside the timed region;
             e2e.by_nsteps repeats it with 16 and 100 RK4 steps per call (what an evolveHam-style caller does).
  roofline   HBM roofline of the dominant kernel (hbk_double_pendulum_dflt_step_rk4) + the instruction-issue view that binds it.
  ham_eqs    explanation only: the path's central function on its own — ONE hamEqs evaluation per trajectory per launch
             (hb_batch_ham_eqs) over the same ring of batches: the HBM-bound kernel of the path against the same roofline.
  fused16 / chain / burst   explanation only: 16 steps per launch; a real stepping loop (L2-resident state); the K launches alone.
  cpu_baseline   the CPU oracle (restatement of the reference algorithm) on this box's host cores, the SAME 1,048,576 batch.
  configs    BASELINE configs[2], [3], [4] measured the same way (`--config 3|4|5` makes one of them the whole run):
             "3" 2,097,152 pendulums (System 2 1) + 2,097,152 two-body orbits (System 4 2) (two streams or one, whichever is faster),
             "4" triple pendulum (System 6 3), 8,388,608 initial conditions split over the N GPUs (STRONG scaling), 1000 steps,
                 one NCCL all-gather of the final Phases INSIDE the timed region,
             "5" 12-link chain (System 24 12), batch 262,144.
`--impl reference` times the reference's CPU implementation of the path: the Haskell+GSL binary cannot be built in this
image (no GHC/GSL), so it is the oracle port (oracle/hamilton_oracle.c) on all host threads, on the full batch.

N > 1 (torchrun, one rank per GPU): trajectories are independent, so for `value` each rank owns its own 1,048,576-trajectory
shard (weak scaling, no data-path collective; the all-gather of the final Phases is timed separately: gather_ms,
value_with_gather); configs["4"] is the strong-scaling ensemble with the gather inside.
"""
import argparse
import json
import math
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

PI = float(np.pi)
N_PER_GPU = 1 << 20
DT = 0.01
SEED = 0x48414D49
METRIC = "phase-space RK4 steps/sec (batched trajectories)"
MIN_REGION_MS = 100.0
NVML_POLL_S = float(os.environ.get("HB_BENCH_NVML_MS", "10")) * 1e-3   # ~10 clock samples per 100 ms timed region
# name -> (builtin id, n, lo, hi): sampling boxes of SURVEY.md §8(d)
SYS = {
    "pendulum": (0, 1, [-PI, -1.0], [PI, 1.0]),
    "double_pendulum": (1, 2, [-PI, -PI, -1.0, -1.0], [PI, PI, 1.0, 1.0]),
    "two_body": (3, 2, [1.0, -PI, -1.0, 1.0], [3.0, PI, 1.0, 5.0]),
    "triple_pendulum": (6, 3, [-PI] * 3 + [-1.0] * 3, [PI] * 3 + [1.0] * 3),
    "chain12": (7, 12, [-PI] * 12 + [-1.0] * 12, [PI] * 12 + [1.0] * 12),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every ~10 ms, started well before the
    region (falls back to `nvidia-smi -lms 20`)."""

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
            time.sleep(NVML_POLL_S)

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
                "sampled": ("NVML every ~10 ms, " if self.h is not None else "nvidia-smi -lms 20, ") + where}


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
    """Times the oracle's RK4 on the SAME 1,048,576-trajectory batch (same RNG stream); as many whole-batch steps as fit
    the time target."""
    from oracle import oracle as O
    S = O.OracleSystem.builtin(O.DOUBLE_PENDULUM)
    _id, _n, lo, hi = SYS["double_pendulum"]
    y = S.init_random(SEED, 0, N_PER_GPU, lo, hi)
    t = time.perf_counter(); y, _ = S.batch_step(y, 0, DT, 1, threads=threads); cal = time.perf_counter() - t
    reps = max(1, min(200, int(target_seconds / max(cal, 1e-3))))
    t = time.perf_counter(); _, bad = S.batch_step(y, 0, DT, reps, threads=threads); el = time.perf_counter() - t
    return N_PER_GPU * reps / el, {"trajectories": N_PER_GPU, "rk4_steps_each": reps, "seconds": round(el, 3), "failed": int(bad)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; Haskell+GSL unbuildable here), rank 0 only, on the FULL
    1,048,576-trajectory batch of the b200 arm's config."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.max_threads()
    S = O.OracleSystem.builtin(O.DOUBLE_PENDULUM)
    _id, _n, lo, hi = SYS["double_pendulum"]
    y = S.init_random(SEED, 0, N_PER_GPU, lo, hi)
    for _ in range(args.warmup):
        y, _bad = S.batch_step(y, 0, DT, 1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y, _bad = S.batch_step(y, 0, DT, 1, threads=threads)
    el = time.perf_counter() - t0
    v = N_PER_GPU * args.steps / el
    sample = "all %d trajectories of the batch per step (same splitmix64 stream), %d RK4 steps, %d host threads" % (N_PER_GPU, args.steps, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": main_config(1, None),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference algorithm (hamEqs via dense forward-mode jets + explicit inverse, classical RK4); "
                "the ad+hmatrix+GSL Haskell binary cannot be built in this image (no GHC, no libgsl)"}))


def main_config(world, extra):
    c = {"workload": "double pendulum (System 4 2), batch 1,048,576 random initial Phases per GPU, RK4 dt=0.01, fp64 (BASELINE configs[1])",
         "batch_per_gpu": N_PER_GPU, "global_batch": world * N_PER_GPU, "integrator": "rk4", "steps_per_launch": 1,
         "layout": "AOS (array of Phases)", "parallelism": "%d independent shards" % world}
    if extra:
        c.update(extra)
    return c


class Harness:
    """Device-resident timing of hb_batch_step(RK4, nsteps = 1) launches: K launches captured in one CUDA graph over a ring
    of (in, out) buffer pairs larger than L2, replayed until the region is >= MIN_REGION_MS."""

    def __init__(self, torch, hb, L, dev, rank, world, dist):
        self.torch, self.hb, self.L, self.dev, self.rank, self.world, self.dist = torch, hb, L, dev, rank, world, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def ring_for(self, name, N, first=0, min_bytes=576 << 20):
        sid, n, lo, hi = SYS[name]
        s = self.hb.systems.builtin(sid)
        pair = 2 * N * 2 * n * 8
        ring = max(3, min(12, int(math.ceil(min_bytes / pair))))
        ins = [s.batch_init_random(SEED + r, first, N, lo, hi) for r in range(ring)]
        outs = [self.torch.empty_like(b) for b in ins]
        return s, ins, outs

    def capture(self, launches, stream=None):
        """launches: list of callables, each enqueues work on the current stream; returns a replay() callable."""
        torch = self.torch
        try:
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                with torch.cuda.graph(g, stream=cap):
                    for f in launches:
                        f()
            torch.cuda.current_stream().wait_stream(cap)
            return g.replay, "cuda_graph of K kernel launches"
        except Exception as ex:   # pragma: no cover
            sys.stderr.write("bench: CUDA-graph capture failed (%r); timing eager launches\n" % (ex,))
            return (lambda: [f() for f in launches]), "eager"

    def time_launches(self, launch, K, sampler=None):
        """launch(i) enqueues the i-th launch on the current stream.  Times L = K * ceil(100 ms / (K launches)) launches
        captured in ONE CUDA graph and replayed ONCE: consecutive replays of the same executable graph do not overlap (the
        second waits for the first and pays the graph's submission again — measured 1.3 us per kernel node,
        profiles/r2i/ab_graph_size.txt), which is launch plumbing, not the kernel.  Returns (ms per launch, L, region ms,
        clocks, how)."""
        torch = self.torch
        small, how = self.capture([(lambda i=i: launch(i)) for i in range(K)])
        small(); self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); small(); e1.record()
        self.barrier()
        est = max(e0.elapsed_time(e1), 1e-3)
        self.last_burst_ms = est / K                             # the K-launch region on its own (a burst: no power capping yet)
        R = max(1, int(math.ceil(MIN_REGION_MS / est)))
        if self.world > 1:   # every rank times the same number of launches
            t = torch.tensor([R], dtype=torch.int64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            R = int(t.item())
        L = K * R
        if R > 1:
            del small
            big, how = self.capture([(lambda i=i: launch(i)) for i in range(L)])
        else:
            big = small
        big()                                                    # warm: uploads the graph
        self.barrier()
        t0 = time.time()
        e0.record(); big(); e1.record()
        self.barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t0, t1) if sampler is not None else None
        return ms / L, L, ms, clocks, how

    def reduce_max(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def roofline_obj(n, N, launch_ms, kernel):
    peak, peak_src = peaks()
    algo = N * 32 * n
    achieved = algo / (launch_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_launch": algo}


def bench_simple(H, name, N, K, kernel):
    """One system, one batch, one step per launch."""
    sid, n, lo, hi = SYS[name]
    s, ins, outs = H.ring_for(name, N)
    ring = len(ins)
    L = H.L
    for i in range(3):
        s.batch_step(ins[i % ring], DT, 1, integ=L.RK4, out=outs[i % ring])
    per, nl, _ms, _, how = H.time_launches(lambda i: s.batch_step(ins[i % ring], DT, 1, integ=L.RK4, out=outs[i % ring]), K)
    per = H.reduce_max(per)
    return {"workload": "%s (System %d %d), batch %d, RK4 dt=0.01, one step per launch" % (name, 2 * n, n, N), "value": N / (per * 1e-3),
            "unit": "steps/s", "ms_per_step": per, "launches_timed": nl, "launch": how,
            "l2": "ring of %d (in,out) pairs = %d MiB" % (ring, ring * 2 * N * 2 * n * 8 >> 20), "roofline": roofline_obj(n, N, per, kernel)}


def bench_config3(H, K):
    """2,097,152 pendulums (System 2 1) + 2,097,152 two-body orbits (System 4 2): two kernels per step (independent chains on two streams, or alternating on one)."""
    torch, L = H.torch, H.L
    Np = 1 << 21
    sp, pin, pout = H.ring_for("pendulum", Np, min_bytes=288 << 20)
    st, tin, tout = H.ring_for("two_body", Np, min_bytes=288 << 20)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(k, two_streams):
        """k steps of both batches.  two_streams: the two systems are independent, so each runs its own chain of k launches on
        its own stream (forked once, joined once) and the kernels of the two chains overlap freely; otherwise the 2k launches
        alternate on ONE stream (each kernel has the whole GPU, the next one's prologue overlaps its tail through PDL)."""
        cur = torch.cuda.current_stream()
        if not two_streams:
            for i in range(k):
                sp.batch_step(pin[i % len(pin)], DT, 1, integ=L.RK4, out=pout[i % len(pin)])
                st.batch_step(tin[i % len(tin)], DT, 1, integ=L.RK4, out=tout[i % len(tin)])
            return
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            for i in range(k):
                sp.batch_step(pin[i % len(pin)], DT, 1, integ=L.RK4, out=pout[i % len(pin)])
        with torch.cuda.stream(s2):
            for i in range(k):
                st.batch_step(tin[i % len(tin)], DT, 1, integ=L.RK4, out=tout[i % len(tin)])
        cur.wait_stream(s1); cur.wait_stream(s2)

    def timed(two_streams):
        run(3, two_streams)
        H.barrier()
        small, how = H.capture([lambda: run(K, two_streams)])
        small(); H.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); small(); e1.record()
        H.barrier()
        R = max(1, int(math.ceil(MIN_REGION_MS / max(e0.elapsed_time(e1), 1e-3))))
        nl = K * R
        big, how = H.capture([lambda: run(nl, two_streams)]) if R > 1 else (small, how)
        big(); H.barrier()
        e0.record(); big(); e1.record()
        H.barrier()
        return e0.elapsed_time(e1) / nl, nl, how
    per2, nl2, how2 = timed(True)
    per1, nl1, how1 = timed(False)
    per, nl, how, mode = (per2, nl2, how2, "two streams") if per2 <= per1 else (per1, nl1, how1, "one stream, alternating")
    algo = Np * 32 + Np * 64
    peak, peak_src = peaks()
    ach = algo / (per * 1e-3) / 1e9
    return {"workload": "2,097,152 pendulums (System 2 1) + 2,097,152 two-body orbits (System 4 2), mixed batch 4,194,304 (BASELINE configs[2])",
            "value": 2 * Np / (per * 1e-3), "unit": "steps/s", "ms_per_step": per, "launches_timed": 2 * nl, "launch": how, "mode": mode,
            "ms_per_step_two_streams": per2, "ms_per_step_one_stream": per1,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "hbk_pendulum_step_rk4 + hbk_two_body_dflt_step_rk4 (" + mode + ")", "algorithmic_bytes_per_launch": algo}}


def bench_config4(H, steps=1000):
    """Triple pendulum, 8,388,608 initial conditions split over the ranks (strong scaling); `steps` one-step launches, then ONE
    all-gather of the final Phases (NCCL over NVLink), all inside the timed region."""
    torch, L, dist, world, rank = H.torch, H.L, H.dist, H.world, H.rank
    total = 8 << 20
    N = total // world
    sid, n, lo, hi = SYS["triple_pendulum"]
    s = H.hb.systems.builtin(sid)
    a = s.batch_init_random(SEED, rank * N, N, lo, hi)
    b = torch.empty_like(a)
    allp = torch.empty((total, 2 * n), dtype=a.dtype, device=H.dev) if world > 1 else None
    bufs = [a, b]
    for i in range(4):
        s.batch_step(bufs[i % 2], DT, 1, integ=L.RK4, out=bufs[(i + 1) % 2])
    K = steps                                                # ONE graph of all the launches (no gaps between replays)
    replay, how = H.capture([(lambda i=i: s.batch_step(bufs[i % 2], DT, 1, integ=L.RK4, out=bufs[(i + 1) % 2])) for i in range(K)])
    if world > 1:
        for _ in range(2):                                   # warm the communicator and its buffer registration
            dist.all_gather_into_tensor(allp, bufs[0])
    replay()
    H.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps // K):
        replay()
    e1.record()
    if world > 1:
        dist.all_gather_into_tensor(allp, bufs[0])
    e2.record()
    H.barrier()
    ms_steps, ms_all = H.reduce_max(e0.elapsed_time(e1)), H.reduce_max(e0.elapsed_time(e2))
    done = (steps // K) * K
    per = ms_steps / done
    out = {"workload": "triple pendulum (System 6 3), 8,388,608 initial conditions over %d GPU(s), %d RK4 steps dt=0.01 (one per launch, state in HBM every step), all-gather of the final Phases inside the timed region (BASELINE configs[3])" % (world, done),
           "scaling": "strong", "batch_per_gpu": N, "value": total * done / (ms_all * 1e-3), "unit": "steps/s", "ms_total": ms_all,
           "ms_steps": ms_steps, "gather_ms": ms_all - ms_steps, "gather_bytes": total * 2 * n * 8 if world > 1 else 0,
           "value_without_gather": total * done / (ms_steps * 1e-3), "ms_per_step": per, "launch": how,
           "l2": "two buffers of %d MiB ping-pong (a real stepping loop); per-GPU state %s L2" % (N * 2 * n * 8 >> 20, "larger than" if 2 * N * 2 * n * 8 > (126 << 20) else "within"),
           "roofline": roofline_obj(n, N, per, "hbk_triple_pendulum_dflt_step_rk4")}
    if world > 1:
        out["gather_bus_GBps"] = total * 2 * n * 8 * (world - 1) / world / ((ms_all - ms_steps) * 1e-3) / 1e9
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5], help="0/2: the headline line (+ the other configs as sub-objects); 3|4|5: only that BASELINE config")
    ap.add_argument("--no-extras", action="store_true", help="skip configs 3-5 and the e2e sweep (quick runs)")
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
    H = Harness(torch, hb, L, dev, rank, world, dist)
    K = args.steps

    if args.config in (3, 4, 5):          # one secondary config as the whole run
        if args.config == 3:
            r = bench_config3(H, K)
        elif args.config == 4:
            r = bench_config4(H)
        else:
            r = bench_simple(H, "chain12", 1 << 18, K, "hbk_chain12_step_rk4")
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": r["value"], "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
                              "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r.get("scaling", "weak"), "vs_baseline": None,
                              "dtype": "f64", "data": "synthetic", "config": {"workload": r["workload"]}, "roofline": r["roofline"], "detail": r}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- headline: device-resident throughput (value, roofline) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()                                                 # polling starts seconds before the timed region
    sysm, ring_in, ring_out = H.ring_for("double_pendulum", N_PER_GPU, first=rank * N_PER_GPU)
    RING = len(ring_in)
    N = N_PER_GPU
    flags = torch.zeros(N, dtype=torch.int32, device=dev)

    def step(i):
        sysm.batch_step(ring_in[i % RING], DT, 1, integ=L.RK4, out=ring_out[i % RING], flags=flags)

    for i in range(args.warmup):
        step(i)
    H.barrier()
    launch_ms, n_launches, region_ms, clocks, how = H.time_launches(step, K, sampler=sampler)
    burst_ms = H.reduce_max(H.last_burst_ms)
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
    # explanation only: a real stepping loop (each launch reads what the previous one wrote: L2-resident 32 MiB state)
    ca, cb = ring_in[0].clone(), ring_out[0]
    chain_ms, _, _, _, _ = H.time_launches(lambda i: sysm.batch_step(ca if i % 2 == 0 else cb, DT, 1, integ=L.RK4, out=cb if i % 2 == 0 else ca), K)

    # the path's central function on its own: ONE hamEqs evaluation per trajectory per launch (hb_batch_ham_eqs, src/Numeric/Hamilton.hs:370-387)
    # over the same ring of batches — the HBM-bound kernel of the path, against the same roofline
    for i in range(3):
        sysm.batch_ham_eqs(ring_in[i % RING], out=ring_out[i % RING])
    hameqs_ms, _, _, _, _ = H.time_launches(lambda i: sysm.batch_ham_eqs(ring_in[i % RING], out=ring_out[i % RING]), K)
    hameqs_ms = H.reduce_max(hameqs_ms)

    # ---------------- final collection: one NCCL all-gather of the final Phases ----------------
    gather_ms = 0.0
    if world > 1:
        final = ring_out[(K - 1) % RING]
        allp = torch.empty((world * N, final.shape[1]), dtype=final.dtype, device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(allp, final)                  # warm the communicator
        H.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); dist.all_gather_into_tensor(allp, final); g1.record()
        H.barrier()
        gather_ms = g0.elapsed_time(g1)
        del allp

    # ---------------- end to end through the C ABI with host buffers ----------------
    h_in = [torch.empty((N, 4), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_out = [torch.empty((N, 4), dtype=torch.float64).pin_memory() for _ in range(2)]
    for b, src in zip(h_in, ring_in):
        b.copy_(src.cpu())

    def e2e(nsteps, calls):
        for i in range(3):
            sysm.batch_step(h_in[i % 2], DT, nsteps, integ=L.RK4, out=h_out[i % 2])
        H.barrier()
        w0 = time.perf_counter()
        for i in range(calls):
            sysm.batch_step(h_in[i % 2], DT, nsteps, integ=L.RK4, out=h_out[i % 2])   # blocking: returns when the result is on the host
        torch.cuda.synchronize()
        el = time.perf_counter() - w0
        assert np.isfinite(float(h_out[0][0, 0]))                 # result is read on the host
        return H.reduce_max(el * 1e3), calls
    e2e_calls = max(3, min(K, 50))
    e2e_ms, _ = e2e(1, e2e_calls)
    by_nsteps = {}
    if not args.no_extras:
        for ns in (16, 100):
            ms_ns, calls = e2e(ns, 20)
            by_nsteps[str(ns)] = {"value": world * N * ns * calls / (ms_ns * 1e-3), "unit": "steps/s", "rk4_steps_per_call": ns, "calls": calls,
                                  "h2d_bytes_per_call": N * 32, "d2h_bytes_per_call": N * 32}
    del h_in, h_out

    # ---------------- the other BASELINE configs ----------------
    configs = {}
    if not args.no_extras:
        del ring_in, ring_out
        torch.cuda.empty_cache()
        try:
            configs["4"] = bench_config4(H)
            if world == 1:
                configs["3"] = bench_config3(H, min(K, 200))
                configs["5"] = bench_simple(H, "chain12", 1 << 18, min(K, 100), "hbk_chain12_step_rk4")
        except Exception as ex:   # pragma: no cover
            configs["error"] = repr(ex)

    # ---------------- reduce over ranks (max time) ----------------
    times = torch.tensor([launch_ms, gather_ms, chain_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    launch_ms, gather_ms, chain_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        total_steps = world * N * n_launches
        value = world * N / (launch_ms * 1e-3)
        roof = roofline_obj(2, N, launch_ms, "hbk_double_pendulum_dflt_step_rk4")
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_r2.json")) as f:
                prof = json.load(f)
            roof["traffic"] = prof.get("dram_bytes_per_launch")
            issue = prof.get("issue")
            # live issue view: issue clocks one trajectory-step needs (static SASS count under the measured issue model: FP64
            # instruction 2 clocks, 3 with three distinct register operands, any other instruction 1) against the clocks the
            # schedulers had — the resource that binds this kernel
            cyc = issue.get("issue_clocks_per_trajectory_step") if issue else None
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
            if cyc and mhz:
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                bound = sms * 4 * 32 * mhz * 1e6 / cyc
                issue = dict(issue, issue_bound_steps_per_s_at_measured_clock=bound, achieved_frac_of_issue_bound=value / world / bound)
                f64 = issue.get("fp64_issue_clocks_per_trajectory_step")
                if f64:
                    b64 = sms * 4 * 32 * mhz * 1e6 / f64
                    issue = dict(issue, fp64_bound_steps_per_s_at_measured_clock=b64, achieved_frac_of_fp64_bound=value / world / b64)
            roof["issue"] = issue
            roof["note"] = "the binding resource is instruction issue on the FP64-heavy stream, not HBM (SURVEY.md §8(d), DESIGN.md §4); see issue"
        except Exception:
            pass
        cpu_v = cpu_sample = cpu_threads = None
        if world == 1:                                            # rank 0 at N=1 only
            from oracle import oracle as O
            os.sched_setaffinity(0, all_cpus)                     # the CPU baseline uses every host core again
            cpu_threads = O.max_threads()
            cpu_v, cpu_sample = cpu_oracle_steps_per_sec(args.cpu_seconds, cpu_threads)
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": main_config(world, {"launch": how, "timed_region_ms": region_ms, "launches_timed": n_launches,
                                          "timing": "K = --steps launches estimate the rate; launches_timed = K * ceil(100 ms / that) launches captured in ONE graph, replayed once warm and once timed",
                                          "l2": "inputs larger than L2: ring of %d (in,out) batch pairs = %d MiB touched per cycle" % (RING, RING * 64)}),
            "clocks": clocks,
            "gpu_launches": n_launches,
            "e2e": {"value": world * N * e2e_calls / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": N * 32, "d2h_bytes_per_step": N * 32,
                    "steps": e2e_calls, "call": "hb_batch_step(HB_INTEG_RK4, nsteps=1, HB_MEM_HOST) on pinned host arrays, blocking: one kernel reads the Phases from and writes the results to host memory over PCIe (DESIGN.md section 2)",
                    "host_binding": numa, "by_nsteps": by_nsteps},
            "roofline": roof,
            "fused16": {"value": fused_value, "unit": "steps/s", "note": "16 RK4 steps per launch, no per-step HBM traffic (issue-bound view)"},
            "chain": {"value": world * N / (chain_ms * 1e-3), "unit": "steps/s", "ms_per_step": chain_ms,
                      "note": "explanation only: K dependent one-step launches ping-ponging between two buffers (32 MiB state stays in L2)"},
            "burst": {"value": world * N / (burst_ms * 1e-3), "unit": "steps/s", "ms_per_step": burst_ms, "launches": K,
                      "note": "explanation only: the K = --steps launches alone (a sub-10-ms region, as round 1 timed it); `value` is the >= 100 ms region, which on this part runs into the power cap (clocks.reasons) when every launch streams 64 MiB through HBM"},
            "ham_eqs": {"value": world * N / (hameqs_ms * 1e-3), "unit": "hamEqs evaluations/s", "ms_per_launch": hameqs_ms,
                        "roofline": roofline_obj(2, N, hameqs_ms, "hbk_double_pendulum_dflt_ham_eqs"),
                        "note": "explanation only: one hamEqs evaluation per trajectory per launch (hb_batch_ham_eqs), the HBM-bound kernel of the path; same batches, same ring"},
            "configs": configs,
        }
        if world > 1:
            out["gather_ms"] = gather_ms
            out["value_with_gather"] = total_steps / ((launch_ms * n_launches + gather_ms) * 1e-3)
        if cpu_v is not None:
            out["cpu_baseline"] = {"value": cpu_v, "unit": "steps/s", "cores": cpu_threads, "kind": "port",
                                   "sample": "all %(trajectories)d trajectories x %(rk4_steps_each)d RK4 steps in %(seconds)s s (oracle/hamilton_oracle.c, pthreads)" % cpu_sample}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
