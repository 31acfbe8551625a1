"""One shard of the 65 536-agent swarm (rank r of w, emulated on ONE GPU, no collective): ms per step of the streaming
kernels with and without the dense adjacency rows.  python tools/shard_time.py [r,w] [lanes]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_flock_b200 import SwarmEnv
from gym_flock_b200.synthetic import initial_states_fast
shard = sys.argv[1] if len(sys.argv) > 1 else "3,8"
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 32
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tune = int(sys.argv[4], 0) if len(sys.argv) > 4 else 0
sync_each = int(sys.argv[5]) if len(sys.argv) > 5 else 0
n = 65536
x0 = initial_states_fast(1, n, seed=1234, law="density")[0]
for adjacency in ("dense", "none"):
    os.environ["FLOCK_EMULATE_SHARD"] = shard
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision="fp32", adjacency_mode=adjacency, distributed=False, lanes=lanes, tune=tune)
    del os.environ["FLOCK_EMULATE_SHARD"]
    sw.reset(x0)
    if graph:
        sw.capture_graphs()
    for _ in range(5):
        sw.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if sync_each:          # isolated steps: one event pair per step, a device synchronisation in between
        tot = 0.0
        for _ in range(100):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sw.step(); b.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms = tot / 100
    else:
        for _ in range(100):
            sw.step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
    rows = sw.hi - sw.lo
    gb = rows * 4 * ((n if adjacency == "dense" else 0) + 38) / 1e9
    print(f"shard {shard} lanes {lanes} graph {graph} tune {tune:#x} sync_each {sync_each} adjacency {adjacency}: {ms*1e3:.1f} us/step  ({gb/ms*1e3:.0f} GB/s of row output)", flush=True)
    del sw
    torch.cuda.empty_cache()
