"""CPU, world_size 2 over gloo: the row-sharded swarm protocol (shard rows -> step own rows -> all-gather actions)
reproduces the unsharded rollout exactly.  The CUDA launch is replaced by an oracle-based stand-in (test
infrastructure) so the distributed host logic is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _oracle_backend(x, u, lo, hi, kw):
    from oracle import c_oracle
    xn = c_oracle.integrate(x, u, kw["dt"]) if u is not None else np.array(x, copy=True)
    h = c_oracle.helpers(xn, kw["comm_radius"])
    uu = c_oracle.controller(xn, kw["comm_radius"], kw["centralized"])
    out = dict(x=xn, u_rows=uu[lo:hi], values_rows=h["values"][lo:hi], network_rows=h["adj_mean"][lo:hi],
               reward=c_oracle.reward(xn))
    if kw["knn"]:
        out["obs_rows"] = c_oracle.knn_obs(xn)[lo:hi]
    return out


def _worker(rank, world, port, n, steps, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_flock_b200.swarm import SwarmEnv, shard_rows
    from gym_flock_b200.synthetic import initial_state
    x0 = initial_state(n, seed=5, law="density", min_dist=0.1)
    env = SwarmEnv("Flocking-v0", n_agents=n, precision="fp64-exact", row_backend=_oracle_backend)
    assert (env.lo, env.hi) == shard_rows(n, world, rank)
    env.reset(x0)
    rewards = []
    for _ in range(steps):
        (obs, net), r, done, _ = env.step()
        rewards.append(float(r))
    np.savez(f"{path}.{rank}.npz", x=env.x.numpy(), u=env.u.numpy(), obs=obs.numpy(), net=net.numpy(),
             values=env.values.numpy(), rewards=np.array(rewards), lo=env.lo, hi=env.hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_swarm_equals_unsharded(tmp_path, world):
    from oracle import c_oracle
    from gym_flock_b200.synthetic import initial_state
    n, steps = 64, 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "swarm")
    mp.spawn(_worker, args=(world, port, n, steps, path), nprocs=world, join=True)
    x0 = initial_state(n, seed=5, law="density", min_dist=0.1)
    ref = c_oracle.rollout(x0, steps, knn=7)
    for rank in range(world):
        g = np.load(f"{path}.{rank}.npz")
        lo, hi = int(g["lo"]), int(g["hi"])
        assert np.array_equal(g["x"], ref["x"])                       # full state replicated on every rank
        assert np.array_equal(g["u"], c_oracle.controller(ref["x"]))  # gathered next action, all rows
        assert np.array_equal(g["obs"], ref["obs"][lo:hi])
        assert np.array_equal(g["net"], ref["adj_mean"][lo:hi])
        assert np.array_equal(g["values"], ref["values"][lo:hi])
        assert np.array_equal(g["rewards"], ref["rewards"])


def test_shard_rows_partition():
    from gym_flock_b200.swarm import shard_rows
    assert [shard_rows(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 65536)]
    with pytest.raises(ValueError):
        shard_rows(100, 8, 0)
