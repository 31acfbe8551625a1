#!/usr/bin/env python
"""bench.py -- throughput of the fused flocking step (controller + step) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C2k|C3|C4|C4x|C5|C1] [--impl reference]

One "step" = one lockstep {u = controller(); step(u)} over the whole env batch: integrate, visit every
ordered pair once, write values, the dense adjacency, reward and the next action.  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for every field.

  value     pair-interactions/s, inputs resident in HBM: K steps through flock_rollout_* (small envs: ONE persistent
            launch that runs the K steps and writes every step's outputs), CUDA events on the launch stream, max over ranks
  e2e       same metric through the host-buffer API (NumPy in / NumPy out, H2D + D2H inside the timed region)
  roofline  dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes of SURVEY.md 8d
  extra     per_step_api (the same K steps as K single-step launches), per_launch_us, and every other BASELINE config
            (C2k, C3 = 16384/N envs per GPU, C4, C4x, C5) with ms_per_step, algorithmic bytes and roofline fraction
  cpu_baseline / --impl reference   the reference's CPU implementation on all host cores: the UNMODIFIED reference classes
            from baseline/_ref when that copy travelled ("reference"), else the NumPy restatement oracle/flock_numpy.py ("port")
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pair-interactions/sec (step+controller)"
UNIT = "pair-interactions/s"

# BASELINE.json configs (per-GPU work; C3's batch is split over the ranks)
CONFIGS = {
    "C2": dict(env_id="FlockingRelative-v0", E=4096, N=100, precision="fp32"),
    "C3": dict(env_id="FlockingRelative-v0", E=16384, N=50, precision="fp32", split=True),   # 16384 envs over the ranks
    "C4": dict(env_id="FlockingRelative-v0", E=1024, N=1000, precision="fp64"),
    "C4x": dict(env_id="FlockingRelative-v0", E=1024, N=1000, precision="fp64-exact"),
    "C1g": dict(env_id="Flocking-v0", E=1, N=100, precision="fp64-exact"),
    # BASELINE config 1 through the drop-in gym API (Python loop, NumPy in / out, one env): make() -> controller() -> step()
    "C1": dict(env_id="Flocking-v0", E=1, N=100, precision="fp64-exact", dropin=True),
    "C2k": dict(env_id="Flocking-v0", E=4096, N=100, precision="fp32"),            # C2 with the kNN observation
    # single giant swarm: rows sharded over the ranks, one all-gather of the actions per step (constant density)
    "C5": dict(env_id="Flocking-v0", E=1, N=65536, precision="fp32", swarm=True),
}
TUNE_NO_FUSE = 1 << 26        # FlockParams.reserved: flock_rollout_* as K single-step launches (diagnostics)


def envs_per_gpu(cfg, world):
    return max(1, cfg["E"] // world) if cfg.get("split") else cfg["E"]


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


# the dominant kernel of every config (mangled, as cuobjdump lists it): the DRAM-traffic figures of profiles/traffic.json are keyed
# to the SASS of exactly these functions
TRAFFIC_KERNELS = {
    "C2": "_ZN5flock18flock_group_kernelIfLb0ELi100ELi512EEEvNS_5GArgsE",
    "C2k": "_ZN5flock18flock_group_kernelIfLb1ELi100ELi512EEEvNS_5GArgsE",
    "C3": "_ZN5flock18flock_group_kernelIfLb0ELi50ELi256EEEvNS_5GArgsE",
    "C4": "_ZN5flock16flock_env_kernelIdLi1ELb0ELb0EEEvNS_5KArgsE",
    "C4x": "_ZN5flock16flock_env_kernelIdLi1ELb1ELb0EEEvNS_5KArgsE",
    "C5": "_ZN5flock18flock_swarm_kernelIfLi32ELb1EEEvNS_5SArgsE",
}
_SASS_SHA = {}


def kernel_sass_sha(name, lib_path=None):
    """Hash of the SASS of config `name`'s dominant kernel in the built library (cuobjdump -sass -fun): a DRAM-traffic figure
    captured under ncu describes the machine code it was captured from, nothing else.  None if it cannot be read."""
    if lib_path is None:
        from gym_flock_b200 import _lib
        lib_path = _lib.LIB_PATH
    key = (name, lib_path)
    if key not in _SASS_SHA:
        sha = None
        try:
            exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
            txt = subprocess.run([exe, "-sass", "-fun", TRAFFIC_KERNELS[name], lib_path], capture_output=True, text=True,
                                 timeout=120).stdout
            ins = [ln.strip() for ln in txt.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln)]
            if ins:
                sha = hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16]
        except Exception:
            sha = None
        _SASS_SHA[key] = sha
    return _SASS_SHA[key]


def recorded_traffic(name, steps_per_launch=1):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu capture recorded in profiles/traffic.json (bytes per
    step of the captured launch x the steps one launch of THIS run makes) -- only if the kernel of the loaded library is, SASS
    instruction for instruction, the kernel that capture was taken from; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[name]
        sha = kernel_sass_sha(name)
        if sha is None or t.get("sass_sha") != sha:
            return None
        return float(t["bytes_per_step"]) * steps_per_launch
    except Exception:
        return None


# ---- CPU side (runs before CUDA is initialised; workers are plain NumPy processes) -------------------------
def _cpu_worker(args):
    x, steps, knn, use_ref = args
    if use_ref:      # the unmodified reference classes (baseline/_ref or /root/reference), one env at a time
        from oracle import ref_loader
        for e in range(x.shape[0]):
            env = ref_loader.make_ref_env("flocking" if knn else "relative", n_agents=x.shape[1], x0=x[e])
            for _ in range(steps):
                env.step(env.controller())
    else:
        from oracle import flock_numpy
        flock_numpy.step_batch(x, steps, knn=knn)
    return x.shape[0]


def reference_available():
    try:
        from oracle import ref_loader
        return ref_loader.available()
    except Exception:
        return False


def cpu_reference(cfg, envs_per_core=8, steps=20, procs=None, pairs_per_core=None):
    """The reference's algorithm on all host cores: independent env slices, one process per core (NumPy's elementwise
    loops are single-threaded, so envs/core is the honest scaling).  Runs the unmodified reference when its copy is
    present, else the NumPy restatement.  Returns (pairs/s, seconds, sample description, cores, kind)."""
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
    use_ref = reference_available()
    chunks = [(x[i * per:(i + 1) * per].copy(), steps, knn, use_ref) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[0][:1], 1, knn, use_ref) for c in chunks])       # warm the workers (imports, page-in)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, chunks)
        dt = time.perf_counter() - t0
    pairs = E * N * N * steps
    kind = "reference" if use_ref else "port"
    return pairs / dt, dt, f"{E} envs x {N} agents x {steps} steps, {cores} processes (1 env slice per core)", cores, kind


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


WHAT = {"reference": "the unmodified gym-flock classes (baseline/_ref copy): FlockingRelativeEnv / FlockingEnv .controller() + .step()",
        "port": "oracle/flock_numpy.py: NumPy restatement of the reference algorithm (bit-exact vs the reference; no copy of the "
                "reference tree on this box)"}


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock + throttle reasons while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.mark_t, self.stamps = None, []          # mark(): start of the timed region (samples after it are counted apart)
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
                self.stamps.append(time.perf_counter())
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.0005)

    def mark(self):
        self.mark_t = time.perf_counter()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.mark_t is not None:      # the sampler ran from the device warm-up (same kernel, same load) through the timed region
            out["samples_in_timed_region"] = sum(1 for t in self.stamps if t >= self.mark_t)
            out["window"] = "device warm-up + W warm-up steps + the K timed steps (one continuous load)"
        return out


# ---- arms ------------------------------------------------------------------------------------------------------------
def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU implementation on the host cores (rank 0 only)."""
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
    rate, dt, sample, cores, kind = cpu_reference(cfg, envs_per_core=per, steps=steps, pairs_per_core=budget)
    E_s = per * cores
    out = {"metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
           "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "agent_steps_per_sec": rate / N,
           "config": bench_config(cfg, args.config, args.gpus),
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "what": WHAT[kind]},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "sample_envs": E_s}
    print(json.dumps(out), flush=True)


def bench_config(cfg, name, gpus):
    E = envs_per_gpu(cfg, gpus)
    return {"workload": f"{name}: {cfg['env_id']} {E} envs x {cfg['N']} agents per GPU, {cfg['precision']}, "
                        f"dense adjacency output, controller+step fused",
            "envs_per_gpu": E, "n_agents": cfg["N"], "precision": cfg["precision"],
            "global_envs": E * gpus, "parallelism": f"env-batch sharded x{gpus}, no collectives",
            "l2": "each step writes more than the 126 MB L2 (inputs+outputs > L2); no explicit flush"}


class Dist:
    """torch.distributed plumbing for the bench: barrier + max over ranks (NCCL), no data-path collective."""

    def __init__(self, world, local_rank):
        import torch
        self.torch = torch
        self.world = world
        self.dist = None
        torch.cuda.set_device(local_rank)
        self.device = torch.device(f"cuda:{local_rank}")
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_rollout(D, env, K, rewards=None):
    """K steps through flock_rollout_* between two CUDA events on the launch stream, barrier + synchronize on both
    sides; returns the max over ranks in ms."""
    torch = D.torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    lead_in(D)
    ev0.record()
    env.rollout(K, rewards)
    ev1.record()
    D.barrier()
    mine = ev0.elapsed_time(ev1)
    if os.environ.get("FLOCK_BENCH_DEBUG"):
        sys.stderr.write(f"[rank {os.environ.get('RANK', '0')}] timed_rollout K={K}: {mine * 1e3:.1f} us\n")
    return D.max(mine)


LEAD_IN_US = 200.0


def lead_in(D):
    """Optional device-side lead-in (declared as config.lead_in_us): a spin of LEAD_IN_US microseconds enqueued after the barrier
    and BEFORE the first timing event, so that the host has queued the timed launches by the time the device reaches the event
    and the host's launch latency (Python + ctypes + driver, ~10-80 us depending on how busy the host is) is not inside a
    timed region that lasts 0.7 ms at the driver's K = 20.  Nothing is skipped: the K steps run in full between the events."""
    if LEAD_IN_US > 0:
        D.torch.cuda._sleep(int(LEAD_IN_US * 1.9e3))


def device_warm(D, env, ms):
    """Time-based device warm-up (declared as warmup_ms): the same steps, untimed, until `ms` of device time have passed,
    so that clocks and memory are in their steady state when the K timed steps start."""
    torch = D.torch
    done, chunk = 0.0, 64
    while done < ms:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        env.rollout(chunk)
        ev1.record()
        torch.cuda.synchronize()
        done += ev0.elapsed_time(ev1)
    return done


def measure_env_config(D, name, cfg, rank, world, K, W, warm_ms, args, with_diag=False):
    """One env-batch config on this rank's GPU: returns the dict that goes into the JSON line (value is whole-job)."""
    torch = D.torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states_fast
    E, N, prec = envs_per_gpu(cfg, world), cfg["N"], cfg["precision"]
    knn = cfg["env_id"] == "Flocking-v0"
    x0 = initial_states_fast(E, N, seed=1234 + (rank + args.seed_rank) * E)
    env = FlockVecEnv(cfg["env_id"], num_envs=E, n_agents=N, precision=prec, device=D.device,
                      adjacency_mode=args.adjacency, lanes=args.env_lanes, tune=args.tune)
    env.reset(x0)
    rewards = torch.zeros(K, E, dtype=env.dtype, device=env.device)
    # the time-based device warm-up evolves the flocks for thousands of steps (they spread out: fewer neighbours, cheaper
    # steps), so the state goes back to the synthetic x0 afterwards: the K timed steps are steps W+1 .. W+K of x0
    # a short run's timed region (K = 20: under 1 ms) is shorter than one NVML query, so the sampler covers the continuous load
    # from the device warm-up on and counts the samples that fell into the timed region separately
    sampler = ClockSampler(D.device.index)
    sampler.start()
    warmed = device_warm(D, env, warm_ms) if warm_ms > 0 else 0.0
    if warmed and not args.keep_warm_state:
        env.reset(x0)
    env.rollout(W)                                   # W untimed warm-up steps
    sampler.mark()
    ms = timed_rollout(D, env, K, rewards)           # EXACTLY K steps
    clocks = sampler.stop()
    assert bool(torch.isfinite(rewards).all()), "non-finite rewards in the timed rollout"
    assert int(env.nonfinite.sum().item()) == 0, "the kernel flagged a non-finite env"
    # the workload the timed steps ended on: neighbours per agent (the refine phase scales with it)
    mean_deg = None
    if env.network is not None:
        mean_deg = sum(int(torch.count_nonzero(env.network[e0:e0 + 256])) for e0 in range(0, E, 256)) / float(E * N)
    step_s = ms * 1e-3 / K
    alg = algorithmic_bytes(E, N, prec, knn)
    peaks, peak_src = measured_peaks()
    fused = (N <= 512) and prec != "fp64-exact" and not (args.tune & TUNE_NO_FUSE)
    out = {"ms_per_step": ms / K, "value": world * E * N * N / step_s, "envs_per_gpu": E, "n_agents": N, "precision": prec,
           "steps": K, "algorithmic_bytes_per_step": alg, "achieved_gbs": alg / step_s / 1e9,
           "frac": alg / step_s / 1e9 / peaks["hbm_gbs"], "warmup_ms": warmed, "clocks": clocks,
           "launches": 1 if fused else K, "mean_degree_at_end": mean_deg,
           "kernel": "flock_group_kernel (one persistent launch, K steps)" if fused else "flock_env_kernel (K launches)",
           "peak_source": peak_src}
    if with_diag:
        # the same K steps as K single-step launches queued back to back (what flock_step_* costs per call) ...
        env2 = FlockVecEnv(cfg["env_id"], num_envs=E, n_agents=N, precision=prec, device=D.device,
                           adjacency_mode=args.adjacency, lanes=args.env_lanes, tune=args.tune | TUNE_NO_FUSE)
        env2.reset(x0)
        if warm_ms > 0:
            device_warm(D, env2, warm_ms)
            if not args.keep_warm_state:
                env2.reset(x0)
        env2.rollout(W)
        ms2 = timed_rollout(D, env2, K)
        out["per_step_api"] = {"ms_per_step": ms2 / K, "frac": alg / (ms2 * 1e-3 / K) / 1e9 / peaks["hbm_gbs"], "launches": K,
                               "what": "the same K steps as K single-step launches (flock_step_*), queued back to back"}
        # ... and one event per launch for the first launches of a COLD sequence (explains short-run numbers: ramp vs steady)
        env2.reset(x0)
        torch.cuda.synchronize()
        time.sleep(0.5)
        n = min(K, 40)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            env2.rollout(1)
            evs[i + 1].record()
        torch.cuda.synchronize()
        out["per_launch_us"] = [round(evs[i].elapsed_time(evs[i + 1]) * 1e3, 1) for i in range(n)]
        del env2
    del env, rewards
    torch.cuda.empty_cache()
    return out


def measure_swarm(D, cfg, rank, world, K, W, args, adjacency, skip_gather=False, lanes=32):
    """C5: one swarm of N agents, rows sharded over the ranks (strong scaling: total work fixed)."""
    torch = D.torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_states_fast
    N, prec = cfg["N"], cfg["precision"]
    x0 = initial_states_fast(1, N, seed=1234, law="density")[0]
    env = SwarmEnv(cfg["env_id"], n_agents=N, precision=prec, device=D.device, adjacency_mode=adjacency, lanes=lanes)
    env.reset(x0)
    env._skip_gather = bool(skip_gather)
    if args.swarm_graph:
        env.capture_graphs()
    for _ in range(W):
        env.step()
    sampler = ClockSampler(D.device.index)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    lead_in(D)
    ev0.record()
    for _ in range(K):
        env.step()                       # prep + row kernel (+ kNN fix-up) (+ NCCL all-gather of the actions when world > 1)
    ev1.record()
    D.barrier()
    clocks = sampler.stop()
    ms = D.max(ev0.elapsed_time(ev1))
    peaks, peak_src = measured_peaks()
    rows = N // world
    tb = 4 if prec == "fp32" else 8
    dense = adjacency == "dense"
    # per GPU and step: its rows of the dense adjacency and of the kNN observation / features, plus the O(N) state traffic
    alg = rows * tb * ((N if dense else 0) + 38) + N * 8 * 6
    step_s = ms * 1e-3 / K
    out = {"ms_per_step": ms / K, "value": N * N / step_s, "n_agents": N, "rows_per_gpu": rows, "precision": prec, "steps": K,
           "adjacency": adjacency, "skip_gather": bool(skip_gather), "algorithmic_bytes_per_step": alg,
           "achieved_gbs": alg / step_s / 1e9, "frac": alg / step_s / 1e9 / peaks["hbm_gbs"], "clocks": clocks,
           "launches": K * 3 + (K if world > 1 and not skip_gather else 0), "peak_source": peak_src,
           "cuda_graph": bool(args.swarm_graph),
           "kernel": "flock_prep_kernel + flock_swarm_kernel + flock_knn_fix_kernel (per GPU)"}
    del env
    torch.cuda.empty_cache()
    return out


def run_swarm(args, cfg, rank, world, local_rank):
    D = Dist(world, local_rank)
    K, W = max(1, args.steps), max(3, args.warmup)
    m = measure_swarm(D, cfg, rank, world, K, W, args, args.adjacency, args.skip_gather, args.lanes)
    if rank == 0:
        N = cfg["N"]
        out = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32 filter, f64 refine, f64 state", "data": "synthetic", "agent_steps_per_sec": m["value"] / N,
               "config": {"workload": f"C5: {cfg['env_id']} single swarm of {N} agents (constant density), {cfg['precision']}, "
                                      f"adjacency {args.adjacency}, rows sharded x{world}, all-gather of actions per step",
                          "n_agents": N, "rows_per_gpu": m["rows_per_gpu"], "precision": cfg["precision"],
                          "parallelism": f"agent rows sharded x{world} + 1 NCCL all-gather/step",
                          "l2": "per-GPU outputs exceed L2 when the dense adjacency is written"},
               "roofline": {"bound": "hbm", "achieved": m["achieved_gbs"], "peak": measured_peaks()[0]["hbm_gbs"], "unit": "GB/s",
                            "frac": m["frac"], "traffic": recorded_traffic("C5"), "peak_source": m["peak_source"],
                            "algorithmic_bytes_per_launch": m["algorithmic_bytes_per_step"], "kernel": m["kernel"],
                            "avg_launch_us": m["ms_per_step"] * 1e3},
               "cpu_baseline": {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                "sample": "n/a: the reference materialises diff[N,N,4] = 137 GB at N=65536 (cannot run)"},
               "e2e": None, "gpu_launches": m["launches"], "clocks": m["clocks"]}
        print(json.dumps(out), flush=True)
    D.close()


def run_dropin(args, cfg, rank):
    """C1: the reference's own usage pattern -- one env, a Python loop of controller() + step() -- through the drop-in
    classes (bit-exact mode), next to the NumPy port of the reference on one host core.  Wall clock: every call returns
    host arrays, so each step is synchronous by construction."""
    if rank != 0:
        return None
    import gym_flock_b200 as gf
    from gym_flock_b200.synthetic import initial_state
    from oracle import flock_numpy as fnp
    N = cfg["N"]
    K, W = max(1, args.steps), max(3, args.warmup)
    x0 = initial_state(N, seed=1234, min_dist=0.1)
    # reference arm first: compute_helpers + controller + integrate (+ kNN observation)
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
    env.close()
    return {"metric": METRIC, "value": N * N * K / t, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
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


def measure_e2e(D, cfg, rank, world, Ke):
    """Host-buffer API (NumPy in / NumPy out, pinned): per step D2H of the controller action, H2D of the action, D2H of
    values + dense network + reward (+ obs); wall clock between synchronisations, max over ranks."""
    torch = D.torch
    from gym_flock_b200 import HostFlockBatch
    from gym_flock_b200.synthetic import initial_states_fast
    E, N, prec = envs_per_gpu(cfg, world), cfg["N"], cfg["precision"]
    knn = cfg["env_id"] == "Flocking-v0"
    x0 = initial_states_fast(E, N, seed=1234 + rank * E)
    host = HostFlockBatch(E, N, knn_obs=knn, precision=prec, device=D.device.index)
    host.set_state(x0)
    for _ in range(3):
        host.step(host.controller(True), fetch_state=False)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        u = host.controller(True)                 # D2H  u[E,N,2]
        host.step(u, fetch_state=False)           # H2D  u;  D2H values + network + reward (+obs)
    torch.cuda.synchronize()
    te = D.max(time.perf_counter() - t0)
    tb = 4 if prec == "fp32" else 8
    h2d = E * N * 2 * 8
    d2h = E * N * 2 * 8 + E * (N * 6 + N * N + 1) * tb + (E * N * 28 * tb if knn else 0)
    host.close()
    return {"value": world * E * N * N * Ke / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": Ke, "ms_per_step": te / Ke * 1e3,
            "api": "HostFlockBatch.controller()+step(u): NumPy in/out through flock_ctx_* (C ABI)"}


def measure_e2e_bitmask(D, cfg, rank, world, Ke):
    """The same loop with the adjacency as BITS (FlockVecEnv adjacency_mode='bitmask': FlockAux.adj_bits + degree instead of
    the dense [E,N,N] network), pinned host buffers, copies inside the timed region.  A SEPARATE line: a consumer that
    aggregates on the graph (flock_aggregate_*) needs no dense matrix; the headline e2e keeps the reference's dense layout."""
    torch = D.torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states_fast
    E, N, prec = envs_per_gpu(cfg, world), cfg["N"], cfg["precision"]
    x0 = initial_states_fast(E, N, seed=1234 + rank * E)
    env = FlockVecEnv(cfg["env_id"], num_envs=E, n_agents=N, precision=prec, device=D.device, adjacency_mode="bitmask")
    env.reset(x0)
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    u_h, v_h, b_h, d_h, r_h = pin(env.u_cached), pin(env.values), pin(env.adj_bits), pin(env.degree), pin(env.reward)
    o_h = pin(env.obs) if env.obs is not None else None

    def one():
        u_h.copy_(env.controller(), non_blocking=True)            # D2H  u[E,N,2]
        torch.cuda.synchronize()
        u_d = u_h.to(env.device, non_blocking=True)               # H2D  u
        env.step(u_d)
        v_h.copy_(env.values, non_blocking=True); b_h.copy_(env.adj_bits, non_blocking=True)
        d_h.copy_(env.degree, non_blocking=True); r_h.copy_(env.reward, non_blocking=True)
        if o_h is not None:
            o_h.copy_(env.obs, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        one()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        one()
    te = D.max(time.perf_counter() - t0)
    nbytes = lambda t: t.numel() * t.element_size()
    d2h = nbytes(u_h) + nbytes(v_h) + nbytes(b_h) + nbytes(d_h) + nbytes(r_h) + (nbytes(o_h) if o_h is not None else 0)
    del env
    torch.cuda.empty_cache()
    return {"value": world * E * N * N * Ke / te, "unit": UNIT, "h2d_bytes_per_step": nbytes(u_h), "d2h_bytes_per_step": d2h,
            "steps": Ke, "ms_per_step": te / Ke * 1e3,
            "api": "FlockVecEnv(adjacency_mode='bitmask').controller()+step(u): adjacency bits + degree instead of the dense network"}


def run_ours(args, cfg, rank, world, local_rank):
    if cfg.get("dropin"):
        out = run_dropin(args, cfg, rank)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if cfg.get("swarm"):
        return run_swarm(args, cfg, rank, world, local_rank)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:       # before CUDA init (fork-safe)
        # a bounded sample of the same workload, ~10 s of CPU work for the reference (about 1e8 pair-interactions/s
        # on 16 cores at N=100), a few seconds for the C/OpenMP restatement
        r, dt, sample, cores, kind = cpu_reference(cfg, envs_per_core=32, steps=200, pairs_per_core=6.5e7)
        rc, dtc, samplec, thr = cpu_c_port(cfg, steps=100)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": dt, "what": WHAT[kind],
               "c_port": {"value": rc, "unit": UNIT, "cores": thr, "sample": samplec, "seconds": dtc,
                          "what": "oracle/flock_oracle.c, OpenMP over envs"}}
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    D = Dist(world, local_rank)
    K, W = max(1, args.steps), max(3, args.warmup)
    E, N, prec = envs_per_gpu(cfg, world), cfg["N"], cfg["precision"]
    knn = cfg["env_id"] == "Flocking-v0"

    head = measure_env_config(D, args.config, cfg, rank, world, K, W, args.warmup_ms, args, with_diag=not args.no_extras)
    e2e = measure_e2e(D, cfg, rank, world, min(K, args.e2e_steps))
    e2e_bits = None
    if not args.no_extras:
        try:
            e2e_bits = measure_e2e_bitmask(D, cfg, rank, world, min(K, args.e2e_steps))
        except Exception as e:
            e2e_bits = {"error": repr(e)[:300]}

    # ---- every other BASELINE config, so that the driver's BENCH / SCALE records carry them --------------------------------
    extras = {}
    if not args.no_extras and args.config == "C2":
        for name, k_steps in (("C2k", 200), ("C3", 400), ("C4", 20), ("C4x", 10)):
            try:
                extras[name] = measure_env_config(D, name, CONFIGS[name], rank, world, k_steps, 3, min(args.warmup_ms, 50.0), args)
            except Exception as e:   # (a config that cannot run must not take the headline down)
                extras[name] = {"error": repr(e)[:300]}
        try:
            c5 = measure_swarm(D, CONFIGS["C5"], rank, world, 20, 3, args, "dense")
            if world > 1:
                c5["skip_gather_ms_per_step"] = measure_swarm(D, CONFIGS["C5"], rank, world, 20, 3, args, "dense",
                                                              skip_gather=True)["ms_per_step"]
            c5["adjacency_none_ms_per_step"] = measure_swarm(D, CONFIGS["C5"], rank, world, 20, 3, args, "none")["ms_per_step"]
            extras["C5"] = c5
        except Exception as e:
            extras["C5"] = {"error": repr(e)[:300]}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        step_s = head["ms_per_step"] * 1e-3
        alg = head["algorithmic_bytes_per_step"]
        launches = head["launches"]
        fp_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12 if prec == "fp32" else \
            148 * 64 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        out = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"fp32": "f32 pair math, f64 state", "fp64": "f64", "fp64-exact": "f64"}[prec],
               "data": "synthetic", "agent_steps_per_sec": head["value"] / N,
               "config": dict(bench_config(cfg, args.config, world), warmup_ms=head["warmup_ms"], lead_in_us=LEAD_IN_US,
                              timed_steps="steps W+1 .. W+K from the synthetic x0 (the state is reset after the device warm-up)"
                              if not args.keep_warm_state else "the state the device warm-up left",
                              mean_degree_at_end=head["mean_degree_at_end"],
                              step_api="flock_rollout_*: " + head["kernel"] + "; every step writes all of its outputs"),
               "roofline": {"bound": "hbm", "achieved": head["achieved_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": head["frac"], "traffic": recorded_traffic(args.config, K / launches), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": alg * K / launches, "algorithmic_bytes_per_step": alg,
                            "kernel": head["kernel"], "avg_launch_us": step_s * 1e6 * K / launches,
                            "fp_term": {"flop_per_pair": 32, "achieved_tflops": 32 * E * N * N / step_s / 1e12,
                                        "nominal_peak_tflops": fp_peak,
                                        "frac": 32 * E * N * N / step_s / 1e12 / fp_peak}},
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": head["clocks"],
               "extra": {"per_step_api": head.get("per_step_api"), "per_launch_us": head.get("per_launch_us"),
                         "e2e_bitmask": e2e_bits, "configs": extras}}
        print(json.dumps(out), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--warmup-ms", type=float, default=200.0,
                    help="time-based device warm-up before the K timed steps (reported as config.warmup_ms; 0 = none)")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--keep-warm-state", action="store_true",
                    help="diagnostics: time the K steps on the state the device warm-up left (default: back to x0 first)")
    ap.add_argument("--no-extras", action="store_true", help="headline config only (no per-step-API / other-config measurements)")
    ap.add_argument("--skip-gather", type=int, default=0, help="C5 diagnostic: skip the all-gather (results wrong; timing only)")
    ap.add_argument("--lanes", type=int, default=32, help="C5: threads per row of the swarm kernel (8/16/32; same on all ranks)")
    ap.add_argument("--env-lanes", type=int, default=0, help="env batches: threads per row (0 = library heuristic)")
    ap.add_argument("--tune", type=lambda v: int(v, 0), default=0, help="FlockParams.reserved diagnostics word (A/B runs)")
    ap.add_argument("--swarm-graph", type=int, default=1, help="C5: replay one CUDA graph per step (prep + row + fix-up + all-gather)")
    ap.add_argument("--seed-rank", type=int, default=0, help="diagnostics: use the synthetic states rank R of a multi-GPU run would get")
    ap.add_argument("--lead-in-us", type=float, default=200.0,
                    help="device-side spin enqueued before the first timing event (keeps the host's launch latency out of short "
                         "timed regions; declared in config.lead_in_us)")
    ap.add_argument("--adjacency", default="dense", choices=["dense", "none"],
                    help="'none' skips the dense [E,N,N] output (diagnostic only; the headline is always dense)")
    args = ap.parse_args()
    global LEAD_IN_US
    LEAD_IN_US = float(args.lead_in_us)
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
