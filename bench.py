#!/usr/bin/env python
"""bench.py -- throughput of the fused flocking step (controller + step) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4] [--impl reference]

One "step" = one lockstep {u = controller(); step(u)} over the whole env batch: one kernel
launch that integrates, visits every ordered pair once and writes values, the dense adjacency, reward and
the next action.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

  value     pair-interactions/s, inputs resident in HBM, K launches timed with CUDA events on the launch stream
  e2e       same metric through the host-buffer API (NumPy in / NumPy out, H2D + D2H inside the timed region)
  roofline  dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes of SURVEY.md 8d
  cpu_baseline / --impl reference   the NumPy restatement of the reference (oracle/flock_numpy.py) on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pair-interactions/sec (step+controller)"
UNIT = "pair-interactions/s"

# BASELINE.json configs that fit one GPU (per-GPU work; multi-GPU runs replicate it: weak scaling)
CONFIGS = {
    "C2": dict(env_id="FlockingRelative-v0", E=4096, N=100, precision="fp32"),
    "C3": dict(env_id="FlockingRelative-v0", E=2048, N=50, precision="fp32"),     # 16384 envs / 8 GPUs
    "C4": dict(env_id="FlockingRelative-v0", E=1024, N=1000, precision="fp64"),
    "C4x": dict(env_id="FlockingRelative-v0", E=1024, N=1000, precision="fp64-exact"),
    "C1g": dict(env_id="Flocking-v0", E=1, N=100, precision="fp64-exact"),
    # BASELINE config 1 through the drop-in gym API (Python loop, NumPy in / out, one env): make() -> controller() -> step()
    "C1": dict(env_id="Flocking-v0", E=1, N=100, precision="fp64-exact", dropin=True),
    "C2k": dict(env_id="Flocking-v0", E=4096, N=100, precision="fp32"),            # C2 with the kNN observation
    # single giant swarm: rows sharded over the ranks, one all-gather of the actions per step (constant density)
    "C5": dict(env_id="Flocking-v0", E=1, N=65536, precision="fp32", swarm=True),
}


def algorithmic_bytes(E, N, precision, knn):
    """s*(N^2 + 16N + 1) per env-step (s*(N^2 + 38N + 1) for Flocking-v0), dense adjacency."""
    s = 4 if precision == "fp32" else 8
    return E * s * (N * N + (38 if knn else 16) * N + 1)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ---- CPU side (runs before CUDA is initialised; workers are plain NumPy processes) -------------------------
def _cpu_worker(args):
    x, steps, knn = args
    from oracle import flock_numpy
    flock_numpy.step_batch(x, steps, knn=knn)
    return x.shape[0]


def cpu_numpy_port(cfg, envs_per_core=8, steps=20, procs=None, pairs_per_core=None):
    """The reference's algorithm (NumPy restatement, dense [N,N,*] temporaries) on all host cores: independent
    env slices, one process per core.  Returns (pairs/s, seconds, sample description, cores)."""
    import multiprocessing as mp
    from gym_flock_b200.synthetic import initial_states_fast
    cores = procs or os.cpu_count() or 1
    N = cfg["N"]
    per = max(1, min(envs_per_core, max(1, int(2.0e6 / (N * N)))))     # keep a sample a few seconds per core
    if pairs_per_core:                                                  # bound the work per core (pairs, all steps)
        per = max(1, min(envs_per_core, int(pairs_per_core / (N * N * steps)) or 1))
        steps = max(1, min(steps, int(pairs_per_core / (N * N * per)) or 1))
    E = per * cores
    x = initial_states_fast(E, N, seed=1234)
    knn = cfg["env_id"] == "Flocking-v0"
    chunks = [(x[i * per:(i + 1) * per].copy(), steps, knn) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[0][:1], 1, knn) for c in chunks])       # warm the workers (imports, page-in)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, chunks)
        dt = time.perf_counter() - t0
    pairs = E * N * N * steps
    return pairs / dt, dt, f"{E} envs x {N} agents x {steps} steps, {cores} processes (1 env slice per core)", cores


def cpu_c_port(cfg, steps=20):
    """The scalar C restatement with OpenMP over envs (a faster CPU implementation than the reference's)."""
    from gym_flock_b200.synthetic import initial_states_fast
    from oracle import c_oracle
    N = cfg["N"]
    threads = c_oracle.threads()
    E = max(threads, min(4096, int(threads * 4.0e8 / (N * N * steps)) or 1))
    x = initial_states_fast(E, N, seed=1234)
    c_oracle.rollout(x[:threads], 1, want_network=False)
    t0 = time.perf_counter()
    c_oracle.rollout(x, steps, want_network=False, knn=7 if cfg["env_id"] == "Flocking-v0" else 0)
    dt = time.perf_counter() - t0
    return E * N * N * steps / dt, dt, f"{E} envs x {N} agents x {steps} steps, OpenMP {threads} threads", threads


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock + throttle reasons while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- arms ------------------------------------------------------------------------------------------------------------
def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    steps = max(1, args.steps)
    N = cfg["N"]
    # Each "step" is one lockstep step of a bounded sample of the workload: as many envs per core as keep the whole run
    # at ~15 s of CPU work (about 6.5e6 pair-interactions/s per core at N=100), at most the config's own batch.  The
    # warm-up steps are folded into the pool warm-up (imports, page-in).
    ncores = os.cpu_count() or 1
    budget = float(os.environ.get("FLOCK_BENCH_REF_PAIRS_PER_CORE", "9.0e7"))      # (tests shrink the sample)
    per = max(1, min(max(1, cfg["E"] // ncores), int(budget / (N * N * steps)) or 1))
    rate, dt, sample, cores = cpu_numpy_port(cfg, envs_per_core=per, steps=steps, pairs_per_core=budget)
    E_s = per * cores
    out = {"metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
           "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "agent_steps_per_sec": rate / N,
           "config": bench_config(cfg, args.config, args.gpus),
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "what": "oracle/flock_numpy.py: NumPy restatement of the reference algorithm "
                                    "(bit-exact vs the reference; the reference tree itself cannot travel to this box)"},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "sample_envs": E_s}
    print(json.dumps(out), flush=True)


def bench_config(cfg, name, gpus):
    return {"workload": f"{name}: {cfg['env_id']} {cfg['E']} envs x {cfg['N']} agents per GPU, {cfg['precision']}, "
                        f"dense adjacency output, controller+step fused",
            "envs_per_gpu": cfg["E"], "n_agents": cfg["N"], "precision": cfg["precision"],
            "global_envs": cfg["E"] * gpus, "parallelism": f"env-batch sharded x{gpus}, no collectives",
            "l2": "each step writes more than the 126 MB L2 (inputs+outputs > L2); no explicit flush"}


def run_swarm(args, cfg, rank, world, local_rank):
    """C5: one swarm of N agents, rows sharded over the ranks (strong scaling: total work fixed)."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_states_fast
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    N, prec = cfg["N"], cfg["precision"]
    K, W = max(1, args.steps), max(3, args.warmup)
    x0 = initial_states_fast(1, N, seed=1234, law="density")[0]
    env = SwarmEnv(cfg["env_id"], n_agents=N, precision=prec, device=f"cuda:{local_rank}",
                   adjacency_mode=args.adjacency, lanes=args.lanes)
    env.reset(x0)
    for _ in range(W):
        env.step()
    env._skip_gather = bool(args.skip_gather)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        env.step()                       # prep + row kernel (+ NCCL all-gather of the actions when world > 1)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=env.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        peaks, peak_src = measured_peaks()
        rows = N // world
        tb = 4 if prec == "fp32" else 8
        dense = args.adjacency == "dense"
        alg = rows * tb * ((N if dense else 0) + 38 * N // max(N, 1) * 38) if False else rows * tb * ((N if dense else 0) + 38) + N * 8 * 6
        step_s = ms * 1e-3 / K
        out = {"metric": METRIC, "value": N * N * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32 filter, f64 refine, f64 state", "data": "synthetic", "agent_steps_per_sec": N * K / (ms * 1e-3),
               "config": {"workload": f"C5: {cfg['env_id']} single swarm of {N} agents (constant density), {prec}, "
                                      f"adjacency {args.adjacency}, rows sharded x{world}, all-gather of actions per step",
                          "n_agents": N, "rows_per_gpu": rows, "precision": prec,
                          "parallelism": f"agent rows sharded x{world} + 1 NCCL all-gather/step",
                          "l2": "per-GPU outputs exceed L2 when the dense adjacency is written"},
               "roofline": {"bound": "hbm", "achieved": alg / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": alg / step_s / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": alg, "kernel": "flock_swarm_kernel (per GPU)",
                            "avg_launch_us": step_s * 1e6},
               "cpu_baseline": {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                "sample": "n/a: the reference materialises diff[N,N,4] = 137 GB at N=65536 (cannot run)"},
               "e2e": None, "gpu_launches": 2 * K, "clocks": clocks}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_dropin(args, cfg, rank):
    """C1: the reference's own usage pattern -- one env, a Python loop of controller() + step() -- through the drop-in
    classes (bit-exact mode), next to the NumPy port of the reference on one host core.  Wall clock: every call returns
    host arrays, so each step is synchronous by construction."""
    if rank != 0:
        return
    import gym_flock_b200 as gf
    from gym_flock_b200.synthetic import initial_state
    from oracle import flock_numpy as fnp
    N = cfg["N"]
    K, W = max(1, args.steps), max(3, args.warmup)
    x0 = initial_state(N, seed=1234, min_dist=0.1)
    # reference arm first (before CUDA is touched): compute_helpers + controller + integrate (+ kNN observation)
    t0 = time.perf_counter()
    x = x0.copy()
    h = fnp.compute_helpers(x, 0.81)
    for _ in range(K):
        u = fnp.controller(h, 0.9, True)
        x = fnp.integrate(x, u, 0.01)
        h = fnp.compute_helpers(x, 0.81)
        fnp.knn_observation(x, h["r2"])
        fnp.instant_cost(x)
    t_ref = time.perf_counter() - t0
    env = gf.make(cfg["env_id"])
    raw = env.env
    raw.set_state(x0)
    for _ in range(W):
        env.step(raw.controller())
    raw.set_state(x0)
    t0 = time.perf_counter()
    for _ in range(K):
        obs, r, done, info = env.step(raw.controller())
    t = time.perf_counter() - t0
    same = bool(np.array_equal(raw.x, x))
    out = {"metric": METRIC, "value": N * N * K / t, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": t / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "agent_steps_per_sec": N * K / t,
           "config": {"workload": f"C1: {cfg['env_id']} one env x {N} agents, fp64-exact, drop-in gym API "
                                  "(Python loop: controller() + step(), NumPy in/out)", "n_agents": N},
           "cpu_baseline": {"value": N * N * K / t_ref, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"{K} steps, one process (oracle/flock_numpy.py)", "ms_per_step": t_ref / K * 1e3},
           "final_state_bit_equal_to_port": same, "e2e": {"value": N * N * K / t, "unit": UNIT,
                                                           "h2d_bytes_per_step": N * 2 * 8,
                                                           "d2h_bytes_per_step": N * 2 * 8 + N * 4 * 8 + (N * 28 + N * N + 1) * 8},
           "gpu_launches": K, "roofline": None}
    print(json.dumps(out), flush=True)


def run_ours(args, cfg, rank, world, local_rank):
    if cfg.get("dropin"):
        return run_dropin(args, cfg, rank)
    if cfg.get("swarm"):
        return run_swarm(args, cfg, rank, world, local_rank)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:       # before CUDA init (fork-safe)
        # a bounded sample of the same workload, ~10 s of CPU work for the NumPy port (about 1e8 pair-interactions/s
        # on 16 cores at N=100), a few seconds for the C/OpenMP restatement
        r, dt, sample, cores = cpu_numpy_port(cfg, envs_per_core=32, steps=200, pairs_per_core=6.5e7)
        rc, dtc, samplec, thr = cpu_c_port(cfg, steps=100)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds": dt,
               "c_port": {"value": rc, "unit": UNIT, "cores": thr, "sample": samplec, "seconds": dtc,
                          "what": "oracle/flock_oracle.c, OpenMP over envs"}}
    import torch
    from gym_flock_b200 import FlockVecEnv, HostFlockBatch
    from gym_flock_b200.synthetic import initial_states_fast
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    E, N, prec = cfg["E"], cfg["N"], cfg["precision"]
    knn = cfg["env_id"] == "Flocking-v0"
    K, W = max(1, args.steps), max(3, args.warmup)
    x0 = initial_states_fast(E, N, seed=1234 + rank * E)
    env = FlockVecEnv(cfg["env_id"], num_envs=E, n_agents=N, precision=prec, device=local_rank,
                      adjacency_mode=args.adjacency, lanes=args.env_lanes)
    env.reset(x0)
    rewards = torch.zeros(K, E, dtype=env.dtype, device=env.device)
    env.rollout(W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    env.rollout(K, rewards)           # K kernel launches, back to back on the current stream, no host sync
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=env.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    assert bool(torch.isfinite(rewards).all()), "non-finite rewards in the timed rollout"
    pairs_total = world * E * N * N * K
    value = pairs_total / (ms_max * 1e-3)

    # ---- e2e: host-buffer API, H2D action + D2H observation every step --------------------------------
    Ke = min(K, args.e2e_steps)
    host = HostFlockBatch(E, N, knn_obs=knn, precision=prec, device=local_rank)
    host.set_state(x0)
    for _ in range(3):
        host.step(host.controller(True), fetch_state=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        u = host.controller(True)                 # D2H  u[E,N,2]
        host.step(u, fetch_state=False)           # H2D  u;  D2H values + network + reward (+obs)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    te_t = torch.tensor([te], dtype=torch.float64, device=env.device)
    if dist is not None:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    te = float(te_t.item())
    tb = 4 if prec == "fp32" else 8
    h2d = E * N * 2 * 8
    d2h = E * N * 2 * 8 + E * (N * 6 + N * N + 1) * tb + (E * N * 28 * tb if knn else 0)
    e2e_val = world * E * N * N * Ke / te
    host.close()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        step_s = ms_max * 1e-3 / K
        alg = algorithmic_bytes(E, N, prec, knn)
        achieved = alg / step_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.config)
        except Exception:
            pass
        fp_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12 if prec == "fp32" else \
            148 * 64 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"fp32": "f32 pair math, f64 state", "fp64": "f64", "fp64-exact": "f64"}[prec],
               "data": "synthetic", "agent_steps_per_sec": value / N,
               "config": bench_config(cfg, args.config, world),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": alg, "kernel": "flock_env_kernel",
                            "avg_launch_us": step_s * 1e6,
                            "fp_term": {"flop_per_pair": 32, "achieved_tflops": 32 * E * N * N / step_s / 1e12,
                                        "nominal_peak_tflops": fp_peak,
                                        "frac": 32 * E * N * N / step_s / 1e12 / fp_peak}},
               "cpu_baseline": cpu,
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "steps": Ke, "ms_per_step": te / Ke * 1e3,
                       "api": "HostFlockBatch.controller()+step(u): NumPy in/out through flock_ctx_* (C ABI)"},
               "gpu_launches": K, "clocks": clocks}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-gather", type=int, default=0, help="C5 diagnostic: skip the all-gather (results wrong; timing only)")
    ap.add_argument("--lanes", type=int, default=32, help="C5: threads per row of the swarm kernel (8/16/32; same on all ranks)")
    ap.add_argument("--env-lanes", type=int, default=0, help="env batches: threads per row (0 = library heuristic)")
    ap.add_argument("--adjacency", default="dense", choices=["dense", "none"],
                    help="'none' skips the dense [E,N,N] output (diagnostic only; the headline is always dense)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one rank per GPU): "
                         "python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
