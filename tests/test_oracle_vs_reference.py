"""CPU, only where the reference tree exists (the build container): the restatements are bit-exact
against the UNMODIFIED reference on fresh seeded inputs (this is what pins the oracle)."""
import numpy as np
import pytest

from gym_flock_b200.synthetic import initial_state
from oracle import c_oracle, flock_numpy as fnp, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("n", [2, 3, 7, 8, 50, 100, 257])
@pytest.mark.parametrize("centralized", [True, False])
def test_restatements_bit_exact_vs_reference(n, centralized):
    x0 = initial_state(n, seed=4000 + n, min_dist=0.1 if n <= 100 else 0.0)
    kind = "flocking" if n >= 7 else "relative"
    env = ref_loader.make_ref_env(kind, n_agents=n, x0=x0)
    rewards = []
    for _ in range(200):
        u = env.controller(centralized=centralized)
        (obs, net), r, done, info = env.step(u)
        rewards.append(r)
    ref = dict(x=env.x, values=env.state_values, adj_mean=env.adj_mat_mean, u=u, rewards=np.array(rewards))
    if kind == "flocking":
        ref["obs"] = obs
    a = fnp.rollout(x0, 200, centralized=centralized, knn=(kind == "flocking"))
    b = c_oracle.rollout(x0, 200, centralized=centralized, knn=7 if kind == "flocking" else 0)
    for k, v in ref.items():
        assert np.array_equal(a[k], v), ("numpy port", k)
        assert np.array_equal(b[k], v), ("C oracle", k)


def test_reference_reset_cfg_path_matches_golden_x0():
    """The committed smoke x0 is what the reference's own reset() produces here (libm-dependent)."""
    import os
    from golden_util import GOLDEN_DIR
    Rel, _ = ref_loader.load()
    env = Rel()
    env.params_from_cfg(ref_loader.CfgSection(comm_radius=0.9, n_agents=100, v_max=5.0, dt=0.01))
    np.random.seed(0)
    env.reset()
    g = np.load(os.path.join(GOLDEN_DIR, "smoke_seed0.npz"))
    assert np.allclose(env.x, g["x0"], rtol=0, atol=1e-12)


def test_two_flocks_reset_law_matches_reference():
    """FlockingTwoFlocks-v0 differs from FlockingRelative-v0 only in reset(): same lattice, same RNG draw."""
    import importlib.util, os, sys
    root = ref_loader.reference_root()
    ref_loader.load()
    for name in ("utils", "flocking_twoflocks"):
        spec = importlib.util.spec_from_file_location(f"gym_flock.envs.flocking.{name}",
                                                      os.path.join(root, "gym_flock/envs/flocking", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"gym_flock.envs.flocking.{name}"] = mod
        spec.loader.exec_module(mod)
    ref = sys.modules["gym_flock.envs.flocking.flocking_twoflocks"].FlockingTwoFlocksEnv()
    np.random.seed(4)
    ref.reset()
    from gym_flock_b200.envs import _grid
    np.random.seed(4)
    bias = np.random.uniform(low=-ref.v_bias / 2.0, high=ref.v_bias / 2.0, size=(2,))
    g = _grid(100, side=10)
    x = np.zeros((100, 4)); x[:, :2] = g; x[:, 2:] = -g; x[:, 2] += bias[0]; x[:, 3] += bias[1]
    assert np.array_equal(ref.x, x)
