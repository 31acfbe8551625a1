"""Generates tests/golden/*.npz by running the UNMODIFIED gym-flock reference (needs /root/reference
or $GYM_FLOCK_REF; run in the build container:  python tests/golden/make_golden.py).

Each file holds the initial state as raw float64 (x0 BYTES, not a seed -- reset() uses libm cos/sin)
and what the reference returned along a 200-step {controller(); step()} rollout:
  x_t, u_t (action applied in step t), values_t, adj_t (packed bits), rewards[200],
  adj_mean_200, obs_t (Flocking-v0 only) for t in RECORD.
`smoke_seed0.npz` is the SURVEY.md section 4 smoke golden: params_from_cfg + np.random.seed(0) + reset().
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from gym_flock_b200.synthetic import initial_state  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RECORD = (1, 2, 50, 100, 200)
STEPS = 200

# (name, kind, N, centralized, comm_radius, dt, seed, min_dist)
CASES = [
    ("rel_n100_cen", "relative", 100, True, 0.9, 0.01, 11, 0.1),
    ("rel_n100_dec", "relative", 100, False, 0.9, 0.01, 12, 0.1),
    ("flk_n100_cen", "flocking", 100, True, 0.9, 0.01, 13, 0.1),
    ("rel_n50_cen", "relative", 50, True, 0.9, 0.01, 14, 0.1),
    ("flk_n50_dec", "flocking", 50, False, 0.9, 0.01, 15, 0.1),
    ("flk_n8_cen", "flocking", 8, True, 0.9, 0.01, 16, 0.1),
    ("flk_n7_cen", "flocking", 7, True, 0.9, 0.01, 17, 0.1),
    ("rel_n3_cen", "relative", 3, True, 0.9, 0.01, 18, 0.1),
    ("rel_n2_dec", "relative", 2, False, 0.9, 0.01, 19, 0.1),
    ("rel_n33_r15_dt02", "relative", 33, True, 1.5, 0.02, 20, 0.1),
    ("flk_n257_cen", "flocking", 257, True, 0.9, 0.01, 21, 0.0),
]


def run_case(kind, n, centralized, comm_radius, dt, x0):
    env = ref_loader.make_ref_env(kind, n_agents=n, comm_radius=comm_radius, dt=dt, x0=x0)
    out = {"x0": x0.copy(), "values_0": env.state_values.copy(),
           "adj_0": np.packbits(env.adj_mat.astype(np.uint8), axis=1),
           "u_0": env.controller(centralized=centralized)}
    rewards = np.zeros(STEPS)
    for t in range(1, STEPS + 1):
        u = env.controller(centralized=centralized)
        (obs, net), r, done, info = env.step(u)
        assert done is False and info == {}
        rewards[t - 1] = r
        if t in RECORD:
            out[f"x_{t}"] = env.x.copy()
            out[f"u_{t}"] = u.copy()
            out[f"values_{t}"] = env.state_values.copy()
            out[f"adj_{t}"] = np.packbits(env.adj_mat.astype(np.uint8), axis=1)
            out[f"r2gap_{t}"] = np.min(np.abs(env.r2[np.isfinite(env.r2)] - env.comm_radius2)) if n > 1 else np.inf
            if kind == "flocking":
                out[f"obs_{t}"] = obs.copy()
    out["rewards"] = rewards
    out["adj_mean_200"] = env.adj_mat_mean.copy()
    stats = env.get_stats()
    out["stats_vel_diffs"] = stats["vel_diffs"]
    out["stats_min_dists"] = stats["min_dists"]
    out["meta"] = np.array([n, int(centralized), comm_radius, dt, 1.0 if kind == "flocking" else 0.0])
    return out


def smoke_seed0():
    Rel, _ = ref_loader.load()
    env = Rel()
    env.params_from_cfg(ref_loader.CfgSection(comm_radius=0.9, n_agents=100, v_max=5.0, dt=0.01))
    np.random.seed(0)
    values0, net0 = env.reset()
    x0 = env.x.copy()
    out = {"x0": x0, "values_0": values0.copy(), "adj_mean_0": net0.copy(), "r_max": np.array(env.r_max)}
    rewards = np.zeros(STEPS)
    for t in range(STEPS):
        u = env.controller()
        if t == 0:
            out["u_first"] = u.copy()
        (_, _), r, _, _ = env.step(u)
        rewards[t] = r
    out["rewards"] = rewards
    out["x_200"] = env.x.copy()
    st = env.get_stats()
    out["stats_vel_diffs"], out["stats_min_dists"] = st["vel_diffs"], st["min_dists"]
    return out


def run_variant(which, n, seed):
    """FlockingLeader-v0 / FlockingStochastic-v0 of the reference, state injected, 200-step expert rollout."""
    if which == "leader":
        env = ref_loader.load_variant("flocking_leader", "FlockingLeaderEnv")()
    elif which == "obstacle":
        env = ref_loader.load_variant("flocking_obstacle", "FlockingObstacleEnv")()
    else:
        env = ref_loader.load_variant("flocking_stoch", "FlockingStochasticEnv")()
    x0 = initial_state(n, seed=seed, min_dist=0.1)
    env.n_agents = n
    if which == "leader":
        env.mask = np.ones((n,)); env.mask[0:env.n_leaders] = 0
        x0[0:env.n_leaders, 2:4] = 1.75
    if which == "obstacle":
        env.mask = np.ones((n,)); env.mask[0:env.n_obstacles] = 0
        if seed == 0:                      # the reference's own deterministic reset() (flocking_obstacle.py:57-72)
            env.reset()
            x0 = env.x.copy()
        else:                              # injected state: obstacles at rest
            x0[0:env.n_obstacles, 2:4] = 0
    env.x = x0.copy()
    env.compute_helpers()
    out = {"x0": x0.copy(), "values_0": env.state_values.copy(), "u_0": env.controller()}
    np.random.seed(seed)
    rewards, dts = np.zeros(STEPS), np.zeros(STEPS)
    for t in range(1, STEPS + 1):
        u = env.controller()
        (values, net), r, done, info = env.step(u)
        rewards[t - 1], dts[t - 1] = r, env.dt
        if t in RECORD:
            out[f"x_{t}"] = env.x.copy(); out[f"u_{t}"] = u.copy(); out[f"values_{t}"] = values.copy()
            out[f"adj_{t}"] = np.packbits(env.adj_mat.astype(np.uint8), axis=1)
    out["rewards"], out["dts"] = rewards, dts
    out["adj_mean_200"] = env.adj_mat_mean.copy()
    out["meta"] = np.array([n, seed])
    return out


def leader_reset(n=20, seed=41):
    """FlockingLeaderEnv.reset() of the reference (flocking_leader.py:36-40) followed by controller() / step(): the
    leaders' velocities are overwritten AFTER compute_helpers(), so the returned features and the first controller()
    describe the state before the overwrite.  x_pre is what the base-class reset() leaves with the same seed (the
    leader draw comes after the rejection loop)."""
    Rel, _ = ref_loader.load()
    env = ref_loader.load_variant("flocking_leader", "FlockingLeaderEnv")()
    env.params_from_cfg(ref_loader.CfgSection(comm_radius=0.9, n_agents=n, v_max=5.0, dt=0.01))
    np.random.seed(seed)
    Rel.reset(env)
    x_pre = env.x.copy()
    np.random.seed(seed)
    values, net = env.reset()
    x_post = env.x.copy()
    assert np.array_equal(x_pre[:, :2], x_post[:, :2]) and np.array_equal(x_pre[env.n_leaders:], x_post[env.n_leaders:])
    out = {"x_pre": x_pre, "x_post": x_post, "leader_vel": np.array(x_post[0, 2]), "values_reset": values.copy()}
    u = env.controller()
    out["u_first"] = u.copy()
    for t in range(1, 4):
        (values, net), r, _, _ = env.step(u)
        out[f"x_{t}"] = env.x.copy(); out[f"values_{t}"] = values.copy()
        u = env.controller()
        out[f"u_{t}"] = u.copy()
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "reset_leader_n20.npz"), **leader_reset())
    print("wrote reset_leader_n20")
    if "--only-new" in sys.argv:
        return
    for which, n, seed in (("leader", 100, 31), ("stochastic", 100, 32), ("leader", 20, 33), ("stochastic", 33, 34),
                           ("obstacle", 100, 0), ("obstacle", 40, 35)):
        np.savez_compressed(os.path.join(HERE, f"variant_{which}_n{n}.npz"), **run_variant(which, n, seed))
        print("wrote variant", which, n)
    for name, kind, n, cen, cr, dt, seed, md in CASES:
        x0 = initial_state(n, seed=seed, min_dist=md)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(kind, n, cen, cr, dt, x0))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "smoke_seed0.npz"), **smoke_seed0())
    print("wrote smoke_seed0")


if __name__ == "__main__":
    main()
