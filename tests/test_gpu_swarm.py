"""GPU: row-sharded calls (the swarm path) agree bit for bit with the unsharded call, for every mode; the ranks are
emulated as sequential launches over the row ranges on one GPU (no kernel waits on another)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp64-exact", "fp64", "fp32"])
@pytest.mark.parametrize("n,world,knn", [(256, 2, True), (1000, 8, True), (1024, 4, False), (96, 2, False)])
def test_row_shards_equal_full(precision, n, world, knn):
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.swarm import shard_rows
    from gym_flock_b200.synthetic import initial_state
    x0 = initial_state(n, seed=n, law="density", min_dist=0.1)
    env_id = "Flocking-v0" if knn else "FlockingRelative-v0"
    full = FlockVecEnv(env_id, num_envs=1, n_agents=n, precision=precision)
    full.reset(x0[None])
    (obs_f, net_f), rew_f, _, _ = full.step(full.controller())
    u_f = full.controller().clone()
    torch.cuda.synchronize()
    for rank in range(world):
        lo, hi = shard_rows(n, world, rank)
        part = FlockVecEnv(env_id, num_envs=1, n_agents=n, precision=precision, rows=(lo, hi))
        part.reset(x0[None])
        u0 = full_u0 = None
        # every rank needs the FULL action: take it from the unsharded env's first controller() call
        ref0 = FlockVecEnv(env_id, num_envs=1, n_agents=n, precision=precision)
        ref0.reset(x0[None])
        (obs_p, net_p), rew_p, _, _ = part.step(ref0.controller())
        u_p = part.u_cached
        torch.cuda.synchronize()
        assert torch.equal(part.x, full.x)                              # every rank integrates ALL rows
        assert torch.equal(obs_p[0, lo:hi], obs_f[0, lo:hi])
        assert torch.equal(net_p[0, lo:hi], net_f[0, lo:hi])
        assert torch.equal(u_p[0, lo:hi], u_f[0, lo:hi])
        assert torch.equal(rew_p, rew_f)


def test_swarm_env_single_rank_matches_vec_env():
    import torch
    from gym_flock_b200 import FlockVecEnv, SwarmEnv
    from gym_flock_b200.synthetic import initial_state
    n = 512
    x0 = initial_state(n, seed=3, law="density", min_dist=0.1)
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision="fp32")
    ve = FlockVecEnv("Flocking-v0", num_envs=1, n_agents=n, precision="fp32")
    sw.reset(x0); ve.reset(x0[None])
    for _ in range(4):
        (o1, n1), r1, _, _ = sw.step()
        (o2, n2), r2, _, _ = ve.step(ve.controller())
    torch.cuda.synchronize()
    assert torch.equal(sw.x, ve.x[0]) and torch.equal(o1, o2[0]) and torch.equal(n1, n2[0]) and torch.equal(r1, r2)
