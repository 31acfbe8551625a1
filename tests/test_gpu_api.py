"""GPU: the drop-in gym API (make/reset/step/controller/...) behaves like the reference classes."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR

pytestmark = pytest.mark.gpu


class Cfg(dict):
    def getfloat(self, k): return float(self[k])
    def getint(self, k): return int(self[k])


def test_make_reset_rollout_matches_smoke_golden():
    """SURVEY.md section 4 smoke: params_from_cfg + np.random.seed(0) + reset() + 200 x {controller, step}."""
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, "smoke_seed0.npz"))
    env = gf.make('FlockingRelative-v0')
    env.env.params_from_cfg(Cfg(comm_radius=0.9, n_agents=100, v_max=5.0, dt=0.01))
    assert env.env.r_max == g["r_max"]
    np.random.seed(0)
    values0, net0 = env.reset()
    assert np.allclose(env.env.x, g["x0"], rtol=0, atol=1e-12)      # libm cos/sin may differ in the last ulp
    # step() integrates env.x in place (flocking_relative.py:97-105), so the view reset() leaves in init_vel (:188) stays live
    x_obj, iv = env.env.x, env.env.init_vel
    env.step(env.env.controller())
    assert env.env.x is x_obj and np.shares_memory(iv, env.env.x) and np.array_equal(iv, env.env.x[:, 2:4])
    env.env.set_state(g["x0"])                                       # continue from the committed BYTES
    assert np.array_equal(env.env.state_values, g["values_0"]) and np.array_equal(env.env.state_network, g["adj_mean_0"])
    rewards = []
    for t in range(200):
        u = env.env.controller()
        if t == 0:
            assert np.array_equal(u, g["u_first"])
        (values, net), r, done, info = env.step(u)
        assert done is False and info == {} and isinstance(r, float)
        assert values.shape == (100, 6) and net.shape == (100, 100) and values.dtype == np.float64
        rewards.append(r)
    assert np.array_equal(np.array(rewards), g["rewards"])
    assert np.array_equal(env.env.x, g["x_200"])
    st = env.env.get_stats()
    assert np.array_equal(st["vel_diffs"], g["stats_vel_diffs"]) and np.array_equal(st["min_dists"], g["stats_min_dists"])
    env.close()


def test_flocking_v0_observation_and_errors():
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, "flk_n8_cen.npz"))
    env = gf.FlockingEnv()
    assert env.n_neighbors == 7 and env.n_f == 28 and env.nx_system == 4 and env.n_features == 6 and env.nu == 2
    assert env.action_space.shape == (200,) and env.observation_space.shape == (100, 6)
    env.n_agents = 8
    obs, net = env.set_state(g["x0"])
    assert obs.shape == (8, 28) and net.shape == (8, 8)
    with pytest.raises(AssertionError):
        env.step(np.zeros(16))                     # flocking_relative.py:94: only (N, 2) is accepted
    with pytest.raises(AssertionError):
        env.step(np.zeros((1, 8, 2)))
    for t in range(1, 3):
        (obs, net), r, done, info = env.step(env.controller())
        assert np.array_equal(obs, g[f"obs_{t}"])
    # state injection by attribute, the reference idiom
    env.x = g["x_50"]
    env.compute_helpers()
    assert np.array_equal(env.state_values, g["values_50"])
    # in-place edits of env.x are picked up as well
    env.x[:] = g["x_100"]
    assert np.array_equal(env.controller(), np.array(env.controller(centralized=True)))
    (obs, net), r, _, _ = env.step(env.controller())
    env.n_agents = 5
    env.x = g["x0"][:5]
    env.compute_helpers()
    with pytest.raises(IndexError):
        env.get_observation()                      # flocking.py:23-24 raises for N < 7
    env.close()


def test_decentralized_controller_and_float32_actions():
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, "rel_n100_dec.npz"))
    env = gf.FlockingRelativeEnv()
    env.set_state(g["x0"])
    assert np.array_equal(env.controller(centralized=False), g["u_0"])
    u = env.controller(centralized=False)
    (v, n), r, _, _ = env.step(u.astype(np.float32).astype(np.float64))   # promoted actions keep state float64
    assert env.x.dtype == np.float64
    env.close()


def test_time_limit_and_reset_escape():
    import gym_flock_b200 as gf
    env = gf.make('Flocking-v0')
    assert env._max_episode_steps == 1000
    env.env.reset_max_tries = 50
    with pytest.raises(RuntimeError):
        env.reset()                                 # constructor defaults never admit a configuration
    env.env.params_from_cfg(Cfg(comm_radius=0.9, n_agents=20, v_max=5.0, dt=0.01))
    env.env.reset_max_tries = 20000
    np.random.seed(3)
    obs, net = env.reset()
    assert obs.shape == (20, 28)
    env._max_episode_steps = 3
    dones = [env.step(env.controller())[2] for _ in range(3)]
    assert dones == [False, False, True]
    env.close()


def test_vec_env_matches_host_engine_and_rollout():
    import torch
    from gym_flock_b200 import FlockVecEnv, HostFlockBatch
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(6, 50, seed=5, min_dist=0.1)
    for precision in ("fp64-exact", "fp64", "fp32"):
        for env_id in ("FlockingRelative-v0", "Flocking-v0"):
            vec = FlockVecEnv(env_id, num_envs=6, n_agents=50, precision=precision)
            host = HostFlockBatch(6, 50, knn_obs=(env_id == "Flocking-v0"), precision=precision)
            obs, net = vec.reset(x0)
            host.set_state(x0)
            for _ in range(5):
                (obs, net), rew, done, _ = vec.step(vec.controller())
                host.step(None)
            torch.cuda.synchronize()
            assert np.array_equal(vec.x.cpu().numpy(), host.x)
            assert np.array_equal(net.cpu().numpy(), host.network)
            assert np.array_equal(rew.cpu().numpy(), host.reward)
            ref = host.obs if env_id == "Flocking-v0" else host.values
            assert np.array_equal(obs.cpu().numpy(), ref)
            # C-side rollout == per-step API
            vec2 = FlockVecEnv(env_id, num_envs=6, n_agents=50, precision=precision)
            vec2.reset(x0)
            rewards = torch.zeros(5, 6, dtype=vec2.dtype, device=vec2.device)
            vec2.rollout(5, rewards)
            torch.cuda.synchronize()
            assert torch.equal(vec2.x, vec.x) and torch.equal(vec2.values, vec.values)
            assert torch.equal(rewards[-1], vec.reward)
            host.close()


def test_two_flocks_variant_runs_and_matches_oracle():
    import gym_flock_b200 as gf
    from oracle import c_oracle
    env = gf.make('FlockingTwoFlocks-v0')
    assert env._max_episode_steps == 500
    np.random.seed(4)
    values, net = env.reset()
    x0 = env.env.x.copy()
    assert np.array_equal(values, c_oracle.helpers(x0)["values"])
    for _ in range(5):
        (values, net), r, done, _ = env.step(env.env.controller())
    ref = c_oracle.rollout(x0, 5)
    assert np.array_equal(env.env.x, ref["x"]) and np.array_equal(values, ref["values"]) and np.array_equal(net, ref["adj_mean"])
    env.close()


@pytest.mark.parametrize("name", ["variant_leader_n100", "variant_leader_n20", "variant_stochastic_n100",
                                  "variant_stochastic_n33"])
def test_leader_and_stochastic_variants_match_reference_golden(name, monkeypatch):
    """FlockingLeader-v0 / FlockingStochastic-v0 through the same kernel (FlockParams.n_leaders / step_scale /
    state_scale / u_clip), bit-exact against the reference's own rollouts (committed goldens)."""
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = int(g["meta"][0])
    leader = "leader" in name
    env = (gf.FlockingLeaderEnv if leader else gf.FlockingStochasticEnv)()
    env.n_agents = n
    if leader:
        env.mask = np.ones((n,)); env.mask[0:env.n_leaders] = 0
    else:
        dts = iter(g["dts"].tolist())
        monkeypatch.setattr(np.random, "normal", lambda *a, **k: next(dts))   # the reference's draws, replayed
    env.set_state(g["x0"])
    assert np.array_equal(env.state_values, g["values_0"])
    assert np.array_equal(env.controller(), g["u_0"])
    rewards = np.zeros(200)
    for t in range(1, 201):
        u = env.controller()
        (values, net), r, done, info = env.step(u)
        rewards[t - 1] = r
        if t in (1, 2, 50, 100, 200):
            assert np.array_equal(env.x, g[f"x_{t}"]), t
            assert np.array_equal(u, g[f"u_{t}"]), t
            assert np.array_equal(values, g[f"values_{t}"]), t
            assert np.array_equal((net > 0).astype(np.uint8), np.unpackbits(g[f"adj_{t}"], axis=1)[:, :n]), t
    assert np.array_equal(rewards, g["rewards"])
    assert np.array_equal(net, g["adj_mean_200"])
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["variant_obstacle_n100", "variant_obstacle_n40"])
def test_obstacle_variant_matches_reference_golden(name):
    """FlockingObstacle-v0 through the same kernel (FlockParams.n_obstacles, step_scale = 1), bit-exact against the
    reference's own rollouts: n100 starts from the reference's deterministic reset(), n40 from an injected state."""
    import gym_flock_b200 as gf
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = int(g["meta"][0])
    env = gf.FlockingObstacleEnv()
    if n == 100:
        values0, net0 = env.reset()                                    # deterministic lattice (flocking_obstacle.py:57-72)
        assert np.array_equal(env.x, g["x0"])
    else:
        env.n_agents = n
        env.mask = np.ones((n,)); env.mask[0:env.n_obstacles] = 0
        env.set_state(g["x0"])
    assert np.array_equal(env.state_values, g["values_0"])
    assert np.array_equal(env.controller(), g["u_0"])
    rewards = np.zeros(200)
    for t in range(1, 201):
        u = env.controller()
        (values, net), r, done, info = env.step(u)
        rewards[t - 1] = r
        if t in (1, 2, 50, 100, 200):
            assert np.array_equal(env.x, g[f"x_{t}"]), t
            assert np.array_equal(u, g[f"u_{t}"]), t
            assert np.array_equal(values, g[f"values_{t}"]), t
            assert np.array_equal((net > 0).astype(np.uint8), np.unpackbits(g[f"adj_{t}"], axis=1)[:, :n]), t
    assert np.array_equal(rewards, g["rewards"])
    assert np.array_equal(net, g["adj_mean_200"])
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp64", 1e-10), ("fp32", 1e-4)])
@pytest.mark.parametrize("centralized", [True, False])
def test_obstacle_variant_fast_modes(precision, tol, centralized):
    """The fast modes take the obstacle terms out in the row epilogue: teacher-forced against the reference golden states
    (features, adjacency, controller action) for both controller kinds."""
    from gym_flock_b200 import HostFlockBatch
    from oracle import flock_numpy as fnp
    g = np.load(os.path.join(GOLDEN_DIR, "variant_obstacle_n100.npz"))
    eng = HostFlockBatch(1, 100, centralized=centralized, precision=precision, step_scale=1.0, n_obstacles=4)
    for key in ("x0", "x_1", "x_50", "x_100", "x_200"):
        x = g[key]
        eng.set_state(x[None])
        h = fnp.compute_helpers_obstacle(x, 0.81, 4)
        u_ref = fnp.controller(h, 0.9, centralized)
        v = eng.values[0].astype(np.float64)
        assert np.max(np.abs(v - h["values"])) <= tol * np.max(np.abs(h["values"])), key
        assert np.array_equal(eng.network[0] > 0, h["adj"] > 0), key
        u = eng.controller(centralized)[0]
        assert np.max(np.abs(u - u_ref)) <= tol * max(np.max(np.abs(u_ref)), 1e-2), key
    eng.close()
