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


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n,knn", [(512, True), (5000, True), (4097, False)])
def test_streaming_swarm_kernel_vs_oracle_and_shards(precision, n, knn):
    """The streaming swarm path (prep kernel + TMA-tiled row kernel): parity with the oracle (teacher-forced), and row
    shards (emulated ranks, sequential launches) bit-identical to the unsharded call."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_state
    from parity_util import check_pass
    from oracle import c_oracle
    x0 = initial_state(n, seed=n, law="density", min_dist=0.1)
    env_id = "Flocking-v0" if knn else "FlockingRelative-v0"
    sw = SwarmEnv(env_id, n_agents=n, precision=precision, streaming=True)
    sw.reset(x0)
    u0 = sw.u.clone()
    (obs, net), rew, _, _ = sw.step()
    torch.cuda.synchronize()
    x1 = c_oracle.integrate(x0, u0.cpu().numpy())
    assert np.array_equal(sw.x.cpu().numpy(), x1)                     # the integrator is the reference's, exactly
    check_pass(x1, sw.values.cpu().numpy(), net.cpu().numpy(), sw.u.cpu().numpy(), float(rew), precision)
    if knn:
        from parity_util import check_knn_obs
        check_knn_obs(obs.cpu().numpy(), x1, precision, select_kind="fp64")
    if n % 4 == 0:      # equal shards
        lo, hi = n // 4, n // 2
        from gym_flock_b200 import _lib
        part = SwarmEnv.__new__(SwarmEnv)
        part.__dict__.update(sw.__dict__)
        part.lo, part.hi = lo, hi
        part._params = _lib.make_params(1, n, knn_obs=knn, rows=(lo, hi), lanes=32)
        part.values = torch.zeros(hi - lo, 6, dtype=sw.dtype, device=sw.device)
        part.network = torch.zeros(hi - lo, n, dtype=sw.dtype, device=sw.device)
        part.obs = torch.zeros(hi - lo, 28, dtype=sw.dtype, device=sw.device) if knn else None
        part._x = [torch.as_tensor(x0, device=sw.device).clone(), torch.zeros_like(sw.x)]
        part._u = [u0.clone(), torch.zeros_like(u0)]
        part._cur = 0
        part.reward = torch.zeros_like(sw.reward)
        part._launch(part._x[0], part._u[0], part._x[1], part._u[1])
        torch.cuda.synchronize()
        assert torch.equal(part._x[1], sw.x)
        assert torch.equal(part.values, sw.values[lo:hi]) and torch.equal(part.network, net[lo:hi])
        assert torch.equal(part._u[1][lo:hi], sw.u[lo:hi]) and torch.equal(part.reward, rew)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_streaming_swarm_sparse_rows_use_knn_fixup(precision):
    """A sparse swarm: most rows have fewer than 7 agents inside the inflated candidate radius, so their kNN
    observation comes from the fix-up kernel (full rescan).  Must equal the oracle's argsort-based observation."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from oracle import c_oracle
    rs = np.random.RandomState(11)
    n = 3000
    x0 = np.zeros((n, 4))
    x0[:, :2] = rs.uniform(-60, 60, size=(n, 2))          # ~0.2 agents per unit area
    x0[:, 2:] = rs.uniform(-1, 1, size=(n, 2))
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, streaming=True)
    obs, net = sw.reset(x0)
    torch.cuda.synchronize()
    ref = c_oracle.knn_obs(x0)
    tol = 1e-4 if precision == "fp32" else 1e-10
    err = np.abs(obs.cpu().numpy() - ref).max(axis=1) / np.abs(ref).max()
    assert (err <= tol).all(), (int((err > tol).sum()), float(err.max()))
    h = c_oracle.helpers(x0)
    assert np.array_equal(net.cpu().numpy() > 0, h["adj_mean"] > 0)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n,density,min_dist,floor", [(800, 177.0, 0.05, 16 * 32), (4300, 1690.0, 0.008, 128 * 32)])
def test_streaming_swarm_dense_rows_overflow_the_candidate_lists(precision, n, density, min_dist, floor):
    """Swarms far denser than the reference's.  (800 agents in a disk of radius 1.2) every row has more candidates inside the
    inflated radius than the shared-memory lists hold: they spill into the second tier in global scratch.  (4300 agents in a
    disk of radius 0.9) every lane's list overflows both tiers: the row kernel's full-scan fallback.  Same parity bars."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_state
    from parity_util import check_knn_obs, check_pass
    x0 = initial_state(n, seed=n, law="density", density=density, min_dist=min_dist)
    d = x0[:, None, :2] - x0[None, :, :2]
    r2 = (d * d).sum(-1)
    assert ((r2 <= 3.6).sum(1) - 1).min() > floor
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, streaming=True)
    obs, net = sw.reset(x0)
    torch.cuda.synchronize()
    check_pass(x0, sw.values.cpu().numpy(), net.cpu().numpy(), sw.u.cpu().numpy(), float(sw.reward), precision)
    check_knn_obs(obs.cpu().numpy(), x0, precision, select_kind="fp64")


def test_swarm_graph_replay_equals_eager_steps():
    """SwarmEnv.capture_graphs(): one CUDA graph per state parity replaces the launches of a step; same bits as eager steps."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_state
    n = 3000
    x0 = initial_state(n, seed=12, law="density", min_dist=0.1)
    a = SwarmEnv("Flocking-v0", n_agents=n, precision="fp32", streaming=True, distributed=False)
    b = SwarmEnv("Flocking-v0", n_agents=n, precision="fp32", streaming=True, distributed=False)
    a.reset(x0); b.reset(x0)
    b.capture_graphs(warmup=2)              # (warm-up steps advance b: bring a to the same state)
    for _ in range(2):
        a.step()
    for _ in range(5):
        (oa, na), ra, _, _ = a.step()
        (ob, nb), rb, _, _ = b.step()
    torch.cuda.synchronize()
    assert torch.equal(a.x, b.x) and torch.equal(a.u, b.u) and torch.equal(oa, ob) and torch.equal(na, nb) and torch.equal(ra, rb)
    assert torch.equal(a.values, b.values)
