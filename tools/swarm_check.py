"""torchrun --nproc-per-node R tools/swarm_check.py [N]: R-rank SwarmEnv (NCCL all-gather) vs the single-GPU run of the
same swarm, bit for bit, plus a timing of the sharded step (CUDA events, max over ranks)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_flock_b200 import FlockVecEnv, SwarmEnv
from gym_flock_b200.synthetic import initial_states_fast

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lrank}"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = 20
x0 = initial_states_fast(1, n, seed=7, law="density")[0]
for precision in ("fp64-exact", "fp32"):
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision)
    sw.reset(x0)
    for _ in range(steps):
        (obs, net), rew, _, _ = sw.step()
    ref = FlockVecEnv("Flocking-v0", num_envs=1, n_agents=n, precision=precision, device=lrank)
    ref.reset(x0[None])
    for _ in range(steps):
        (obs_r, net_r), rew_r, _, _ = ref.step(ref.controller())
    torch.cuda.synchronize()
    ok = bool(torch.equal(sw.x, ref.x[0]) and torch.equal(obs, obs_r[0, sw.lo:sw.hi]) and
              torch.equal(net, net_r[0, sw.lo:sw.hi]) and torch.equal(rew, rew_r) and torch.equal(sw.u, ref.controller()[0]))
    t = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{lrank}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        sw.step()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 50], device=f"cuda:{lrank}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"swarm N={n} world={world} {precision}: sharded == single-GPU bit for bit: {bool(t.item())};  "
              f"{ms.item()*1e3:.1f} us/step  ({n*n/ms.item()*1e3:.3e} pairs/s)", flush=True)
dist.barrier(); dist.destroy_process_group()
