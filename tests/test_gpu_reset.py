"""Batched device reset() (flock_reset_sample): every accepted configuration satisfies the reference's acceptance test
exactly, the sampler is deterministic and batch-size independent, and its law matches the reference's rejection sampler
(flocking_relative.py:156-192) statistically."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair_stats(x, r2_comm=0.81):
    p = x[..., 0:2]
    d = p[:, :, None, :] - p[:, None, :, :]
    d2 = np.sum(d * d, axis=3)
    n = x.shape[1]
    d2[:, np.arange(n), np.arange(n)] = np.inf
    return np.sqrt(d2.min(axis=(1, 2))), (d2 < r2_comm).sum(axis=2)


def _reference_reset(n, r_max, rs, v_max=5.0, v_bias=5.0, r2_comm=0.81):
    """flocking_relative.py:156-186 verbatim in structure, on a private RandomState."""
    tries = 0
    x = np.zeros((n, 4))
    degree, min_dist = 0, 0
    while degree < 2 or min_dist < 0.1:
        tries += 1
        length = np.sqrt(rs.uniform(0, r_max, size=(n,)))
        angle = np.pi * rs.uniform(0, 2, size=(n,))
        x[:, 0] = length * np.cos(angle)
        x[:, 1] = length * np.sin(angle)
        bias = rs.uniform(low=-v_bias, high=v_bias, size=(2,))
        x[:, 2] = rs.uniform(low=-v_max, high=v_max, size=(n,)) + bias[0]
        x[:, 3] = rs.uniform(low=-v_max, high=v_max, size=(n,)) + bias[1]
        x_loc = np.reshape(x[:, 0:2], (n, 2, 1))
        a_net = np.sum(np.square(np.transpose(x_loc, (0, 2, 1)) - np.transpose(x_loc, (2, 0, 1))), axis=2)
        np.fill_diagonal(a_net, np.inf)
        min_dist = np.sqrt(np.min(np.min(a_net)))
        degree = np.min(np.sum((a_net < r2_comm).astype(int), axis=1))
    return x.copy(), tries


def test_accepted_configurations_pass_the_reference_test_and_are_deterministic():
    from gym_flock_b200 import FlockVecEnv
    E, N = 96, 50
    env = FlockVecEnv("FlockingRelative-v0", num_envs=E, n_agents=N, precision="fp64")
    (values, net), tries = env.reset_random(seed=11)
    x = env.x.cpu().numpy()
    t = tries.cpu().numpy()
    assert (t > 0).all()
    md, deg = _pair_stats(x)
    assert (md >= 0.1).all() and (deg.min(axis=1) >= 2).all()
    assert (np.hypot(x[..., 0], x[..., 1]) <= np.sqrt(np.sqrt(N)) * (1 + 1e-12)).all()     # disk of radius sqrt(r_max)
    assert ((x[..., 2:4].max(axis=1) - x[..., 2:4].min(axis=1)) <= 10.0).all()            # U(-5, 5) + one bias per env
    assert np.abs(x[..., 2:4]).max() <= 10.0
    # the observation belongs to the sampled state: the network's support is the sampled adjacency
    assert np.array_equal(net.cpu().numpy() > 0, _adj(x) > 0)
    # deterministic, and env e does not depend on the batch it is drawn in
    (_, _), tries2 = env.reset_random(seed=11)
    assert np.array_equal(env.x.cpu().numpy(), x) and np.array_equal(tries2.cpu().numpy(), t)
    small = FlockVecEnv("FlockingRelative-v0", num_envs=8, n_agents=N, precision="fp64")
    small.reset_random(seed=11)
    assert np.array_equal(small.x.cpu().numpy(), x[:8])
    (_, _), _ = env.reset_random(seed=12)
    assert not np.array_equal(env.x.cpu().numpy(), x)


def _adj(x, r2_comm=0.81):
    p = x[..., 0:2]
    d = p[:, :, None, :] - p[:, None, :, :]
    d2 = np.sum(d * d, axis=3)
    n = x.shape[1]
    d2[:, np.arange(n), np.arange(n)] = np.inf
    return (d2 < r2_comm).astype(np.float64)


def test_law_matches_the_reference_sampler_statistically():
    from gym_flock_b200 import FlockVecEnv
    N = 50
    rs = np.random.RandomState(2024)
    ref = [_reference_reset(N, np.sqrt(N), rs) for _ in range(300)]
    xr = np.stack([r[0] for r in ref]); tr = np.array([r[1] for r in ref], dtype=np.float64)
    env = FlockVecEnv("FlockingRelative-v0", num_envs=2048, n_agents=N, precision="fp32")
    (_, _), tries = env.reset_random(seed=5)
    xg = env.x.cpu().numpy(); tg = tries.cpu().numpy().astype(np.float64)
    assert (tg > 0).all()
    # acceptance probability (mean of a geometric law: relative s.e. ~ 1/sqrt(300) on the reference side)
    assert abs(tg.mean() - tr.mean()) <= 0.25 * tr.mean(), (tg.mean(), tr.mean())
    rad_r, rad_g = np.hypot(xr[..., 0], xr[..., 1]).ravel(), np.hypot(xg[..., 0], xg[..., 1]).ravel()
    assert abs(rad_g.mean() - rad_r.mean()) <= 0.01 * rad_r.mean()
    # two-sample Kolmogorov-Smirnov distance of the radial law (agents of one env are not independent: loose bound)
    grid = np.linspace(0, np.sqrt(np.sqrt(N)), 400)
    ks = np.max(np.abs(np.searchsorted(np.sort(rad_r), grid) / rad_r.size - np.searchsorted(np.sort(rad_g), grid) / rad_g.size))
    assert ks <= 0.02, ks
    md_r, deg_r = _pair_stats(xr); md_g, deg_g = _pair_stats(xg)
    assert abs(deg_g.mean() - deg_r.mean()) <= 0.03 * deg_r.mean()
    assert abs(md_g.mean() - md_r.mean()) <= 0.05 * md_r.mean()
    # velocities: U(-5,5) + bias, bias ~ U(-5,5)^2: mean 0, variance 25/3 + 25/3
    v = xg[..., 2:4]
    assert abs(v.mean()) <= 0.25 and abs(v.var() - 50.0 / 3.0) <= 1.0


def test_inadmissible_configuration_reports_failure():
    from gym_flock_b200 import FlockVecEnv
    env = FlockVecEnv("FlockingRelative-v0", num_envs=4, n_agents=100, precision="fp32")
    # constructor defaults of the reference (r_max = 1.0): 100 agents in a unit disk never pass min_dist
    (_, _), tries = env.reset_random(seed=1, r_max=1.0, max_tries=50)
    assert (tries.cpu().numpy() == -50).all()
