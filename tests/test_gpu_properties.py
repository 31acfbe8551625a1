"""GPU: size-independent properties at BASELINE.json's full sizes: symmetric adjacency with a
zero diagonal, rows of adj_mean summing to 1 (or 0), translation / velocity-shift invariance, permutation equivariance,
batch independence.  These hold for any N, so they are checked on the bench workload itself (4096 x 100)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(x0, precision, env_id="FlockingRelative-v0", steps=0):
    import torch
    from gym_flock_b200 import FlockVecEnv
    e, n = x0.shape[0], x0.shape[1]
    env = FlockVecEnv(env_id, num_envs=e, n_agents=n, precision=precision)
    obs, net = env.reset(x0)
    for _ in range(steps):
        (obs, net), rew, _, _ = env.step(env.controller())
    torch.cuda.synchronize()
    return dict(values=env.values.cpu().numpy(), net=net.cpu().numpy(), u=env.controller().cpu().numpy(),
                reward=env.reward.cpu().numpy(), x=env.x.cpu().numpy(), obs=None if env.obs is None else env.obs.cpu().numpy())


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_full_size_adjacency_structure(precision):
    from gym_flock_b200.synthetic import initial_states_fast
    x0 = initial_states_fast(4096, 100, seed=1234)                 # the bench workload (C2)
    out = _run(x0, precision, steps=3)
    net = out["net"]
    adj = net > 0
    assert np.array_equal(adj, adj.transpose(0, 2, 1))                       # r2 is symmetric
    assert not adj[:, np.arange(100), np.arange(100)].any()                  # diagonal = +inf in the reference
    deg = adj.sum(2)
    rows = net.sum(2, dtype=np.float64)
    assert np.allclose(rows[deg > 0], 1.0, rtol=0, atol=2e-6 if precision == "fp32" else 1e-14)
    assert np.all(rows[deg == 0] == 0.0)
    inv = np.where(deg > 0, 1.0 / np.maximum(deg, 1), 0.0)
    assert np.allclose(net, adj * inv[:, :, None], rtol=2e-7 if precision == "fp32" else 1e-15, atol=0)
    assert np.all(np.abs(out["u"]) <= 1.0) and np.all(np.isfinite(out["values"])) and np.all(out["reward"] <= 0)


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-9), ("fp32", 1e-4)])
def test_translation_and_velocity_shift_invariance(precision, tol):
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(64, 100, seed=77, min_dist=0.1)
    base = _run(x0, precision)
    moved = x0.copy()
    moved[:, :, 0] += 37.25; moved[:, :, 1] -= 11.5                          # exactly representable offsets
    moved[:, :, 2] += 3.0; moved[:, :, 3] -= 2.0
    out = _run(moved, precision)
    assert np.array_equal(out["net"] > 0, base["net"] > 0)
    scale = np.abs(base["values"]).max()
    assert np.abs(out["values"] - base["values"]).max() <= tol * scale
    assert np.abs(out["u"] - base["u"]).max() <= tol * max(np.abs(base["u"]).max(), 1e-2)
    assert np.abs(out["reward"] - base["reward"]).max() <= tol * np.abs(base["reward"]).max()


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-11), ("fp32", 1e-4)])
def test_permutation_equivariance(precision, tol):
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(16, 100, seed=5, min_dist=0.1)
    perm = np.random.RandomState(0).permutation(100)
    base = _run(x0, precision, env_id="Flocking-v0")
    out = _run(x0[:, perm], precision, env_id="Flocking-v0")
    assert np.array_equal(out["net"] > 0, (base["net"] > 0)[:, perm][:, :, perm])
    assert np.abs(out["values"] - base["values"][:, perm]).max() <= tol * np.abs(base["values"]).max()
    assert np.abs(out["u"] - base["u"][:, perm]).max() <= tol
    assert np.abs(out["obs"] - base["obs"][:, perm]).max() <= tol * np.abs(base["obs"]).max()


@pytest.mark.parametrize("precision", ["fp64-exact", "fp64", "fp32"])
def test_batch_independence_is_bitwise(precision):
    """Env e does not care who else is in the batch (windows pack envs back to back)."""
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(37, 50, seed=9, min_dist=0.1)
    full = _run(x0, precision, steps=5)
    for sl in (slice(0, 1), slice(5, 12), slice(36, 37)):
        part = _run(x0[sl], precision, steps=5)
        for k in ("values", "net", "u", "reward", "x"):
            assert np.array_equal(part[k], full[k][sl]), (k, sl)


def test_c4_sized_fp64_modes_agree():
    """1000-agent envs (row tiles + bulk zero-fill/scatter output path): fast fp64 vs exact fp64 within 1e-10."""
    from gym_flock_b200.synthetic import initial_states_fast
    x0 = initial_states_fast(8, 1000, seed=3, law="density")
    a = _run(x0, "fp64-exact")
    b = _run(x0, "fp64")
    assert np.array_equal(a["net"] > 0, b["net"] > 0)
    assert np.abs(a["values"] - b["values"]).max() <= 1e-10 * np.abs(a["values"]).max()
    assert np.abs(a["u"] - b["u"]).max() <= 1e-10 and np.abs(a["reward"] - b["reward"]).max() <= 1e-10 * np.abs(a["reward"]).max()
