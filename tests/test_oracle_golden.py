"""CPU: both oracle restatements reproduce the committed outputs of the reference BIT-exactly."""
import numpy as np
import pytest

from golden_util import RECORD, golden_cases, load, unpack_adj
from oracle import c_oracle, flock_numpy as fnp


@pytest.mark.parametrize("name", golden_cases())
def test_numpy_port_matches_reference_golden(name):
    g, m = load(name)
    out = fnp.rollout(g["x0"], 200, comm_radius=m["comm_radius"], dt=m["dt"], centralized=m["centralized"],
                      knn=m["knn"], record_every=1)
    tr = out["traj"]
    for t in RECORD:
        assert np.array_equal(tr["x"][t - 1], g[f"x_{t}"])
        assert np.array_equal(tr["u"][t - 1], g[f"u_{t}"])
        assert np.array_equal(tr["values"][t - 1], g[f"values_{t}"])
        assert np.array_equal(tr["adj"][t - 1], unpack_adj(g[f"adj_{t}"], m["n"]))
    assert np.array_equal(out["rewards"], g["rewards"])
    assert np.array_equal(out["adj_mean"], g["adj_mean_200"])
    if m["knn"]:
        assert np.array_equal(out["obs"], g["obs_200"])
    st = fnp.get_stats(out["x"], out["r2"])
    assert np.array_equal(st["vel_diffs"], g["stats_vel_diffs"])
    assert np.array_equal(st["min_dists"], g["stats_min_dists"])


@pytest.mark.parametrize("name", golden_cases())
def test_c_oracle_matches_reference_golden(name):
    g, m = load(name)
    kw = dict(comm_radius=m["comm_radius"], dt=m["dt"], centralized=m["centralized"])
    prev = 0
    x = g["x0"]
    assert np.array_equal(c_oracle.helpers(x, m["comm_radius"])["values"], g["values_0"])
    assert np.array_equal(c_oracle.controller(x, m["comm_radius"], m["centralized"]), g["u_0"])
    rewards = []
    for t in RECORD:
        out = c_oracle.rollout(x, t - prev, knn=7 if m["knn"] else 0, **kw)
        x, prev = out["x"], t
        rewards.append(out["rewards"])
        assert np.array_equal(x, g[f"x_{t}"])
        assert np.array_equal(out["u"], g[f"u_{t}"])
        assert np.array_equal(out["values"], g[f"values_{t}"])
        assert np.array_equal((out["adj_mean"] > 0).astype(float), unpack_adj(g[f"adj_{t}"], m["n"]))
        if m["knn"]:
            assert np.array_equal(out["obs"], g[f"obs_{t}"])
    assert np.array_equal(np.concatenate(rewards), g["rewards"])
    assert np.array_equal(out["adj_mean"], g["adj_mean_200"])


def test_smoke_seed0_golden():
    """SURVEY.md section 4 smoke golden (reference reset() with np.random.seed(0) via the cfg path)."""
    import os
    from golden_util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "smoke_seed0.npz"))
    assert g["x0"][0].tolist() == [0.1778536750997756, -1.0798168495725435, 8.174058699089354, 0.37089636162829187]
    out = c_oracle.rollout(g["x0"], 200)
    assert np.array_equal(out["x"], g["x_200"])
    assert np.array_equal(out["rewards"], g["rewards"])
    assert out["rewards"][0] == -15.534897470315922 and out["rewards"][-1] == -0.003264501952654671
    assert np.array_equal(c_oracle.controller(g["x0"]), g["u_first"])
    h = c_oracle.helpers(g["x0"])
    assert np.array_equal(h["values"], g["values_0"]) and np.array_equal(h["adj_mean"], g["adj_mean_0"])
    deg = (h["adj"].sum(1))
    assert (deg.min(), deg.max()) == (2.0, 11.0) and abs(deg.mean() - 6.84) < 1e-12


import glob, os
from golden_util import GOLDEN_DIR


def _obstacle_variant(g):
    """FlockingObstacle-v0: unscaled action, obstacles not actuated, velocity differences with obstacles zeroed."""
    x = g["x0"].copy()
    h = fnp.compute_helpers_obstacle(x, 0.9 * 0.9, 4)
    assert np.array_equal(h["values"], g["values_0"])
    assert np.array_equal(fnp.controller(h, 0.9, True), g["u_0"])
    rewards = np.zeros(200)
    for t in range(1, 201):
        u = fnp.controller(h, 0.9, True)
        x = fnp.integrate_leader(x, u, 0.01, 4)                         # flocking_obstacle.py:40-46 == the leader step
        h = fnp.compute_helpers_obstacle(x, 0.9 * 0.9, 4)
        rewards[t - 1] = fnp.instant_cost(x)
        if t in RECORD:
            assert np.array_equal(x, g[f"x_{t}"]), t
            assert np.array_equal(u, g[f"u_{t}"]), t
            assert np.array_equal(h["values"], g[f"values_{t}"]), t
            assert np.array_equal(np.packbits(h["adj"].astype(np.uint8), axis=1), g[f"adj_{t}"]), t
    assert np.array_equal(rewards, g["rewards"])
    assert np.array_equal(h["adj_mean"], g["adj_mean_200"])


def _variant_files():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "variant_*.npz")))


@pytest.mark.parametrize("name", _variant_files())
def test_numpy_port_variants_match_reference_golden(name):
    """FlockingLeader-v0 / FlockingStochastic-v0: the restated step variants reproduce the reference bit for bit."""
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    leader = "leader" in name
    if "obstacle" in name:
        return _obstacle_variant(g)
    x = g["x0"].copy()
    h = fnp.compute_helpers(x, 0.81)
    assert np.array_equal(h["values"], g["values_0"])
    rewards = np.zeros(200)
    for t in range(1, 201):
        u = fnp.controller(h, 0.9, True)
        if leader:
            x = fnp.integrate_leader(x, u, 0.01, 2)
        else:
            u = np.clip(u, -0.5, 0.5)                                   # flocking_stoch.py:43
            x = fnp.integrate_stochastic(x, u, float(g["dts"][t - 1]))
        h = fnp.compute_helpers(x, 0.9 * 0.9)
        rewards[t - 1] = fnp.instant_cost(x)
        if t in RECORD:
            assert np.array_equal(x, g[f"x_{t}"]), t
            assert np.array_equal(u, g[f"u_{t}"]), t
            assert np.array_equal(h["values"], g[f"values_{t}"]), t
    assert np.array_equal(rewards, g["rewards"])


def test_c_oracle_row_subset_equals_full_matrix():
    """flk_rows_f64 (what the giant-swarm GPU tests compare against) is the same arithmetic as the pinned full-matrix
    entry points: bit-equal on a row sample."""
    from gym_flock_b200.synthetic import initial_state
    from oracle import c_oracle
    n = 300
    x = initial_state(n, seed=5, law="density", min_dist=0.1)
    idx = np.array([0, 1, 17, 150, 298, 299], dtype=np.int32)
    for cen in (True, False):
        part = c_oracle.rows(x, idx, centralized=cen, knn=7)
        h = c_oracle.helpers(x)
        assert np.array_equal(part["values"], h["values"][idx])
        assert np.array_equal(part["adj_mean"], h["adj_mean"][idx])
        assert np.array_equal(part["deg"], h["adj"][idx].sum(1))
        assert np.array_equal(part["u"], c_oracle.controller(x, centralized=cen)[idx])
        assert np.array_equal(part["obs"], c_oracle.knn_obs(x)[idx])
        assert np.all(np.diff(part["r2_sorted"], axis=1) >= 0)
