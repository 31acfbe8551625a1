"""Fused-rollout time vs K from the same state: separates the per-launch cost from the per-step cost.
usage: python tools/launch_overhead.py [config]"""
import sys
import torch
sys.path.insert(0, ".")
from bench import CONFIGS
from gym_flock_b200 import FlockVecEnv
from gym_flock_b200.synthetic import initial_states_fast

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
E, N = cfg["E"], cfg["N"]
x0 = initial_states_fast(E, N, seed=1234)
env = FlockVecEnv(cfg["env_id"], num_envs=E, n_agents=N, precision=cfg["precision"])
env.reset(x0)
env.rollout(2000)          # device warm-up
torch.cuda.synchronize()
for K in (1, 2, 5, 10, 20, 40, 100):
    ts = []
    for rep in range(5):
        env.reset(x0)
        env.rollout(5)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); env.rollout(K); ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1) * 1e3)
    ts.sort()
    print(f"{name} K={K:4d}: launch {ts[2]:9.1f} us  = {ts[2] / K:7.2f} us/step (min {ts[0] / K:.2f})")
