"""torchrun --nproc-per-node R tools/swarm_check.py N [N ...]: the R-rank SwarmEnv (rows sharded, NCCL all-gather of the
actions) against a single-rank run of the SAME swarm on every rank's own GPU, bit for bit over `steps` steps, plus a timing of
the sharded step (CUDA events, max over ranks).  N <= 2048 takes the env-resident kernel, larger N the streaming kernels
(prep + TMA-tiled row kernel + kNN fix-up) -- the path BASELINE config 5 (N = 65 536) runs."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_flock_b200 import SwarmEnv
from gym_flock_b200.synthetic import initial_states_fast

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lrank}"))
sizes = [int(a) for a in sys.argv[1:]] or [4096]
steps = 10
for n in sizes:
    x0 = initial_states_fast(1, n, seed=7, law="density")[0]
    for precision in (("fp64-exact", "fp32") if n <= 2048 else ("fp64", "fp32")):
        sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision)
        ref = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, distributed=False)     # all rows on this GPU
        sw.reset(x0); ref.reset(x0)
        ok = True
        for _ in range(steps):
            (obs, net), rew, _, _ = sw.step()
            (obs_r, net_r), rew_r, _, _ = ref.step()
            ok = ok and bool(torch.equal(sw.x, ref.x) and torch.equal(sw.u, ref.u) and torch.equal(rew, rew_r) and
                             torch.equal(obs, obs_r[sw.lo:sw.hi]) and torch.equal(net, net_r[sw.lo:sw.hi]) and
                             torch.equal(sw.values, ref.values[sw.lo:sw.hi]))
        torch.cuda.synchronize()
        streaming = bool(getattr(sw, "streaming", False))
        del ref, net_r, obs_r
        torch.cuda.empty_cache()
        t = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{lrank}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            sw.step()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 30], device=f"cuda:{lrank}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"swarm N={n} world={world} {precision} ({'streaming' if streaming else 'env-resident'} kernel): sharded == "
                  f"single-rank bit for bit over {steps} steps: {bool(t.item())};  {ms.item()*1e3:.1f} us/step  "
                  f"({n*n/ms.item()*1e3:.3e} pairs/s)", flush=True)
        del sw, net, obs
        torch.cuda.empty_cache()
dist.barrier(); dist.destroy_process_group()
