"""GPU: the whole-envs-per-CTA kernel (flock_group_kernel: fused K-step rollouts), the optional outputs of a step
(FlockAux: non-finite guard, adjacency bits, degree, get_stats), and the two adjacent components built on the adjacency
bits / the all-pairs shape (graph-signal aggregation, shepherding repulsion).  All through the C ABI."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR
from parity_util import TOL, normwise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("env_id,e,n", [("FlockingRelative-v0", 23, 100), ("Flocking-v0", 12, 50), ("Flocking-v0", 5, 37),
                                        ("FlockingRelative-v0", 3, 300)])
def test_fused_rollout_is_bitwise_the_per_step_api(precision, env_id, e, n):
    """flock_rollout_* as ONE persistent launch (state in shared memory between steps) == K flock_step_* launches:
    state, action, every observation output and all K rewards, bit for bit; odd and even K (the final state lands in
    either ping-pong buffer)."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(e, n, seed=31 + n, min_dist=0.1)
    for K in (1, 2, 7):
        a = FlockVecEnv(env_id, num_envs=e, n_agents=n, precision=precision, lanes=1)
        b = FlockVecEnv(env_id, num_envs=e, n_agents=n, precision=precision, lanes=1)
        a.reset(x0); b.reset(x0)
        ra = torch.zeros(K, e, dtype=a.dtype, device=a.device)
        rb = torch.zeros(K, e, dtype=a.dtype, device=a.device)
        (oa, na), rew_a = a.rollout(K, ra)
        for t in range(K):
            (ob, nb), rew_b, _, _ = b.step(b.controller())
            rb[t] = rew_b
        torch.cuda.synchronize()
        assert torch.equal(a.x, b.x), K
        assert torch.equal(a.u_cached, b.u_cached), K
        assert torch.equal(oa, ob) and torch.equal(na, nb) and torch.equal(a.values, b.values), K
        assert torch.equal(ra, rb) and torch.equal(rew_a, rb[-1]), K
        # ... and a second rollout continues from where the first one ended
        a.rollout(3); [b.step(b.controller()) for _ in range(3)]
        torch.cuda.synchronize()
        assert torch.equal(a.x, b.x) and torch.equal(a.network, b.network)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_fused_rollout_tracks_the_oracle_200_steps(precision):
    """north_star's bar on the fused path itself: 200 steps in one launch, state and rewards within the mode's tolerance."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    from oracle import c_oracle
    e, n = 11, 100
    x0 = initial_states(e, n, seed=77, min_dist=0.1)
    ref = c_oracle.rollout(x0, 200)
    env = FlockVecEnv("FlockingRelative-v0", num_envs=e, n_agents=n, precision=precision, lanes=1)
    env.reset(x0)
    rewards = torch.zeros(200, e, dtype=env.dtype, device=env.device)
    env.rollout(200, rewards)
    torch.cuda.synchronize()
    tol = TOL[precision]
    x = env.x.cpu().numpy()
    for k in range(e):
        assert normwise(x[k], ref["x"][k]) <= tol
    assert normwise(rewards.cpu().numpy().T, ref["rewards"]) <= tol
    assert int(env.nonfinite.sum()) == 0


def test_variants_through_the_group_kernel_match_the_env_kernel():
    """Leader / Stochastic step variants: the group kernel (lanes=1) and the env kernel (lanes=2) integrate identically."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(9, 40, seed=5, min_dist=0.1)
    for kw in (dict(step_scale=1.0, n_leaders=2), dict(step_scale=6.0, state_scale=6.0, u_clip=0.5), dict(action_scalar=7.5)):
        a = FlockVecEnv("FlockingRelative-v0", num_envs=9, n_agents=40, precision="fp64", lanes=1, **kw)
        b = FlockVecEnv("FlockingRelative-v0", num_envs=9, n_agents=40, precision="fp64", lanes=2, **kw)
        a.reset(x0); b.reset(x0)
        a.rollout(5)
        for _ in range(5):
            b.step(b.controller())
        torch.cuda.synchronize()
        assert normwise(a.x.cpu().numpy(), b.x.cpu().numpy()) <= 1e-12, kw
        assert torch.equal(a.network, b.network), kw


def test_action_scalar_reaches_the_kernel():
    """env.action_scalar is a user-settable attribute of the reference (flocking_relative.py:64, used at :96 and :211)."""
    import gym_flock_b200 as gf
    from oracle import flock_numpy as fnp, ref_loader
    g = np.load(os.path.join(GOLDEN_DIR, "rel_n50_cen.npz"))
    env = gf.FlockingRelativeEnv()
    env.n_agents = 50
    env.set_state(g["x0"])
    u10 = env.controller()
    env.action_scalar = 4.0                        # changed AFTER the engine exists: must still reach the kernel
    u = env.controller()
    assert not np.array_equal(u, u10) and np.max(np.abs(u)) <= 2.5 and np.max(np.abs(u)) > 1.0      # clip(+-10) / 4
    h = fnp.compute_helpers(g["x0"], 0.81)
    assert np.array_equal(u, fnp.controller(h, 0.9, True, action_scalar=4.0))
    (v, net), r, _, _ = env.step(u)
    assert np.array_equal(env.u, u * 4.0)
    assert np.array_equal(env.x, fnp.integrate(g["x0"], u, 0.01, action_scalar=4.0))
    if ref_loader.available():                     # and the unmodified reference agrees
        renv = ref_loader.make_ref_env("relative", n_agents=50, x0=g["x0"])
        renv.action_scalar = 4.0
        ur = renv.controller()
        assert np.array_equal(u, ur)
        renv.step(ur)
        assert np.array_equal(env.x, renv.x) and np.array_equal(v, renv.state_values)
    env.close()


@pytest.mark.parametrize("precision", ["fp64-exact", "fp64", "fp32"])
@pytest.mark.parametrize("n,lanes", [(100, 1), (50, 1), (33, 1), (257, 1), (100, 4)])
def test_aux_outputs_bits_degree_stats_and_guard(precision, n, lanes):
    """FlockAux: adjacency bits == (dense network > 0), degree == row sums, get_stats() == the reference's, non-finite flag."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    from oracle import c_oracle, flock_numpy as fnp
    if precision == "fp64-exact" and lanes > 1:
        pytest.skip("exact mode is one thread per row")
    e = 6
    x0 = initial_states(e, n, seed=900 + n, min_dist=0.1)
    dense = FlockVecEnv("FlockingRelative-v0", num_envs=e, n_agents=n, precision=precision, lanes=lanes)
    dense.reset(x0)
    st = dense.get_stats()
    torch.cuda.synchronize()
    for k in range(e):
        h = fnp.compute_helpers(x0[k], 0.81)
        ref = fnp.get_stats(x0[k], h["r2"])
        if precision == "fp64-exact":
            assert np.array_equal(st["vel_diffs"][k].cpu().numpy(), ref["vel_diffs"])
            assert np.array_equal(st["min_dists"][k].cpu().numpy(), ref["min_dists"])
        else:
            assert normwise(st["vel_diffs"][k].cpu().numpy(), ref["vel_diffs"]) <= TOL[precision]
            assert normwise(st["min_dists"][k].cpu().numpy(), ref["min_dists"]) <= TOL[precision]
    if lanes == 1:
        bits = FlockVecEnv("FlockingRelative-v0", num_envs=e, n_agents=n, precision=precision, adjacency_mode="bitmask")
        (vals, words), = (bits.reset(x0),)
        torch.cuda.synchronize()
        w = words.cpu().numpy().view(np.uint32)
        adj = ((w[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(e, n, -1)[:, :, :n]
        assert np.array_equal(adj.astype(bool), dense.network.cpu().numpy() > 0)
        assert np.array_equal(bits.degree.cpu().numpy(), adj.sum(-1))
        assert torch.equal(bits.values, dense.values)
        (vals, words), rew, _, _ = bits.step(bits.controller())
        dense.step(dense.controller())
        torch.cuda.synchronize()
        w = words.cpu().numpy().view(np.uint32)
        adj = ((w[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(e, n, -1)[:, :, :n]
        assert np.array_equal(adj.astype(bool), dense.network.cpu().numpy() > 0) and torch.equal(bits.x, dense.x)
    # coincident agents: 0/0 in the reference (flocking_relative.py:124) -> the env is flagged, the others are not
    xb = x0.copy()
    xb[2, 1, :2] = xb[2, 0, :2]
    dense.reset(xb)
    torch.cuda.synchronize()
    flags = dense.nonfinite.cpu().numpy()
    assert flags[2] == 1 and flags.sum() == 1


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n,filter_len", [(100, 3), (37, 4)])
def test_graph_signal_aggregation_from_the_bits(precision, n, filter_len):
    """flock_aggregate_*: delayed aggregation and K-hop on the adjacency BITS vs the restatement of the reference's
    aggregate() (old/flocking_multi.py:180-196); f64 bit for bit."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    from oracle import flock_numpy as fnp
    e = 5
    x0 = initial_states(e, n, seed=40 + n, min_dist=0.1)
    env = FlockVecEnv("FlockingRelative-v0", num_envs=e, n_agents=n, precision=precision, adjacency_mode="bitmask")
    env.reset(x0)
    rs = np.random.RandomState(1)
    prev = rs.uniform(-1, 1, size=(e, n, 6 * filter_len))
    sig = env.values.cpu().numpy().astype(np.float64)
    out_d = env.aggregate(filter_len, prev=torch.as_tensor(prev)).cpu().numpy()
    out_k = env.aggregate(filter_len).cpu().numpy()
    out_s = env.aggregate(2, pool="sum").cpu().numpy()
    torch.cuda.synchronize()
    for k in range(e):
        a_net = fnp.compute_helpers(x0[k], 0.81)["adj"]
        pk = prev[k] if precision == "fp64" else prev[k].astype(np.float32).astype(np.float64)
        ref_d = fnp.aggregate_delayed(sig[k], pk, a_net)
        ref_k = fnp.aggregate_khop(sig[k], a_net, filter_len)
        ref_s = fnp.aggregate_khop(sig[k], a_net, 2, mean=False)
        if precision == "fp64":
            assert np.array_equal(out_d[k], ref_d) and np.array_equal(out_k[k], ref_k) and np.array_equal(out_s[k], ref_s)
        else:
            assert normwise(out_d[k], ref_d) <= 1e-5 and normwise(out_k[k], ref_k) <= 1e-5 and normwise(out_s[k], ref_s) <= 1e-5


def test_sheep_repulsion_vs_reference_restatement():
    """flock_repulsion_*: shepherding.py:164-178 for a batch of flocks (f64 bit-exact, f32 within 1e-5)."""
    import ctypes
    import torch
    from gym_flock_b200 import _lib
    from oracle import flock_numpy as fnp
    L = _lib.lib()
    e, n_sh, n_sheep = 7, 10, 20
    n = n_sh + n_sheep
    rs = np.random.RandomState(9)
    x = np.zeros((e, n, 3))
    x[:, :, :2] = rs.uniform(-2.0, 2.0, size=(e, n, 2))
    x[:, :, 2] = rs.uniform(-3, 3, size=(e, n))
    w = 0.15 * np.hstack((3.0 * np.ones((1, n_sh, 1)), 0.5 * np.ones((1, n_sheep, 1))))       # shepherding.py:50
    p = _lib.make_params(e, n)
    xd = torch.as_tensor(x, device="cuda")
    for dt, fn, tol in ((torch.float64, L.flock_repulsion_f64, 0.0), (torch.float32, L.flock_repulsion_f32, 1e-5)):
        wd = torch.as_tensor(w.reshape(-1), device="cuda", dtype=dt)
        out = torch.zeros(e, n, 2, device="cuda", dtype=dt)
        _lib.check(fn(ctypes.byref(p), ctypes.c_void_p(xd.data_ptr()), 3, ctypes.c_void_p(wd.data_ptr()), 2.0,
                      ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        for k in range(e):
            ref = fnp.sheep_repulsion(x[k], w)
            if tol == 0.0:
                assert np.array_equal(out[k].cpu().numpy(), ref)
            else:
                assert normwise(out[k].cpu().numpy(), ref) <= tol


def test_leader_reset_keeps_the_reference_stale_helpers():
    """FlockingLeaderEnv.reset() overwrites the leaders' velocities AFTER compute_helpers() (flocking_leader.py:36-40): the
    first controller() still sees the old velocities, step() integrates the new ones.  Golden from the reference class."""
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, "reset_leader_n20.npz"))
    env = gf.FlockingLeaderEnv()
    env.n_agents = 20
    env.mask = np.ones((20,)); env.mask[0:2] = 0
    vals, net = env.set_state(g["x_pre"])
    assert np.array_equal(vals, g["values_reset"])
    env._set_leader_velocity(g["leader_vel"])
    assert np.array_equal(env.x, g["x_post"])
    u = env.controller()
    assert np.array_equal(u, g["u_first"])                     # computed from the helpers of x_pre
    for t in range(3):
        (v, nw), r, _, _ = env.step(u)
        assert np.array_equal(env.x, g[f"x_{t + 1}"]) and np.array_equal(v, g[f"values_{t + 1}"])
        u = env.controller()
        assert np.array_equal(u, g[f"u_{t + 1}"])
    env.close()


def test_swarm_env_odd_row_offset_fp32():
    """Shard-sized output buffers shifted back by an ODD number of rows (fp32 values: 24 B per row -> 8-byte aligned only)."""
    import torch
    from gym_flock_b200 import FlockVecEnv, SwarmEnv
    from gym_flock_b200.synthetic import initial_state
    n = 1000
    x0 = initial_state(n, seed=21, law="density", min_dist=0.1)
    full = FlockVecEnv("Flocking-v0", num_envs=1, n_agents=n, precision="fp32")
    full.reset(x0[None])
    os.environ["FLOCK_EMULATE_SHARD"] = "1,8"                  # rows [125, 250)
    try:
        part = SwarmEnv("Flocking-v0", n_agents=n, precision="fp32", distributed=False)
    finally:
        del os.environ["FLOCK_EMULATE_SHARD"]
    assert (part.lo, part.hi) == (125, 250)
    obs, net = part.reset(x0)
    torch.cuda.synchronize()
    assert torch.equal(part.values, full.values[0, 125:250]) and torch.equal(net, full.network[0, 125:250])
    assert torch.equal(obs, full.obs[0, 125:250])


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_c5_giant_swarm_row_sample_vs_oracle(precision):
    """BASELINE config 5 at its own size: 65 536 agents through the STREAMING kernels (prep + TMA-tiled row kernel + kNN
    fix-up), checked against the oracle on a sample of 512 rows x all 65 536 columns (3e7 pairs): integrator, features,
    adjacency bits, 1/deg, action, kNN observation; plus a row shard (rank 3 of 8) bit-identical to the full run."""
    import torch
    from gym_flock_b200 import SwarmEnv
    from gym_flock_b200.synthetic import initial_states_fast
    from oracle import c_oracle
    from parity_util import check_knn_obs, check_rows
    n = 65536
    x0 = initial_states_fast(1, n, seed=1234, law="density")[0]
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, distributed=False)
    assert sw.streaming
    rs = np.random.RandomState(0)
    idx = np.unique(np.concatenate(([0, 1, n - 1, 8191, 8192, 24576, 32767], rs.choice(n, 505, replace=False)))).astype(np.int32)
    sw.reset(x0)
    ti = torch.as_tensor(idx.astype(np.int64), device=sw.device)
    u0 = sw.u.clone()
    torch.cuda.synchronize()
    # the pass reset() ran (no integration) on the sample ...
    check_rows(x0, idx, sw.values[ti].cpu().numpy(), sw.network[ti].cpu().numpy(), u0[ti].cpu().numpy(), precision)
    # ... and one full step
    (obs, net), rew, _, _ = sw.step()
    torch.cuda.synchronize()
    x1 = c_oracle.integrate(x0, u0.cpu().numpy())
    assert np.array_equal(sw.x.cpu().numpy(), x1)                       # the integrator is the reference's, exactly
    errs = check_rows(x1, idx, sw.values[ti].cpu().numpy(), net[ti].cpu().numpy(), sw.u[ti].cpu().numpy(), precision)
    print("C5 row-sample errors:", errs)
    check_knn_obs(obs[ti].cpu().numpy(), x1, precision, rows=idx, select_kind="fp64")
    rw = c_oracle.reward(x1)
    assert abs(float(rew) - rw) <= TOL[precision] * abs(rw)
    # one shard of eight, emulated on this GPU: bit-identical rows
    os.environ["FLOCK_EMULATE_SHARD"] = "3,8"
    try:
        part = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, distributed=False)
    finally:
        del os.environ["FLOCK_EMULATE_SHARD"]
    lo, hi = part.lo, part.hi
    assert (lo, hi) == (3 * 8192, 4 * 8192)
    part.reset(x0)
    (pobs, pnet), prew, _, _ = part.step(u0)
    torch.cuda.synchronize()
    assert torch.equal(part.x, sw.x) and torch.equal(prew, rew)
    assert torch.equal(part.values, sw.values[lo:hi]) and torch.equal(pobs, obs[lo:hi]) and torch.equal(pnet, net[lo:hi])
    assert torch.equal(part.u[lo:hi], sw.u[lo:hi])


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n", [100, 37, 9])
def test_knn_observation_does_not_depend_on_the_hint(precision, n):
    """Flocking-v0 in the group kernel: last step's neighbours (FlockAux.knn_hint) only choose the search path -- the fast
    path (row + 7 hints + <= 2 newcomers in the radius mask), the mask search, or the scan of all N.  Garbage hints
    (out of range, duplicates, the row itself), hints of an unrelated state (every neighbour a newcomer) and no hints give
    the observation of the tracked run, bit for bit, and that observation is the oracle's."""
    import torch
    from gym_flock_b200 import FlockVecEnv
    from gym_flock_b200.synthetic import initial_states
    from parity_util import check_knn_obs
    e = 6
    x0 = initial_states(e, n, seed=5 + n, min_dist=0.1)
    envs = {k: FlockVecEnv("Flocking-v0", num_envs=e, n_agents=n, precision=precision, lanes=1)
            for k in ("tracked", "garbage", "foreign", "none")}
    other = FlockVecEnv("Flocking-v0", num_envs=e, n_agents=n, precision=precision, lanes=1)
    other.reset(initial_states(e, n, seed=900 + n, min_dist=0.1))
    for v in envs.values():
        v.reset(x0)
    g = torch.Generator(device="cpu").manual_seed(3)
    for t in range(6):
        other.step(other.controller())
        h = envs["garbage"]._knn_hint
        junk = torch.randint(-32768, 32767, h.shape, generator=g, dtype=torch.int16)
        junk[:, ::3, :] = junk[:, ::3, :1]                               # duplicates
        junk[:, 1::5, :] = torch.arange(n, dtype=torch.int16)[1::5, None]   # the row itself
        h.copy_(junk)
        envs["foreign"]._knn_hint.copy_(other._knn_hint)
        envs["none"]._knn_hint.zero_()
        obs = {}
        for k, v in envs.items():
            (o, _), _, _, _ = v.step(v.controller())
            obs[k] = o
        torch.cuda.synchronize()
        for k in ("garbage", "foreign", "none"):
            assert torch.equal(envs[k].x, envs["tracked"].x), (k, t)
            assert torch.equal(obs[k], obs["tracked"]), (k, t)
    x = envs["tracked"].x.cpu().numpy()
    for env in range(e):
        check_knn_obs(obs["tracked"][env].cpu().numpy(), x[env], precision)
