"""GPU: fp64 exact-order mode reproduces the reference BIT for bit (golden vectors + fresh batches vs the oracle)."""
import numpy as np
import pytest

from golden_util import RECORD, golden_cases, load, unpack_adj

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_cases())
def test_exact_mode_matches_reference_golden(name):
    from gym_flock_b200 import HostFlockBatch
    g, m = load(name)
    n = m["n"]
    knn = m["knn"]
    eng = HostFlockBatch(1, n, comm_radius=m["comm_radius"], dt=m["dt"], centralized=m["centralized"],
                         knn_obs=knn, precision="fp64-exact")
    eng.set_state(g["x0"][None])
    assert np.array_equal(eng.values[0], g["values_0"])
    assert np.array_equal((eng.network[0] > 0).astype(np.uint8), unpack_adj(g["adj_0"], n).astype(np.uint8))
    assert np.array_equal(eng.controller(m["centralized"])[0], g["u_0"])
    rewards = np.zeros(200)
    for t in range(1, 201):
        u = eng.controller(m["centralized"]).copy()
        eng.step(u)
        rewards[t - 1] = eng.reward[0]
        if t in RECORD:
            assert np.array_equal(eng.x[0], g[f"x_{t}"]), t
            assert np.array_equal(u[0], g[f"u_{t}"]), t
            assert np.array_equal(eng.values[0], g[f"values_{t}"]), t
            assert np.array_equal((eng.network[0] > 0).astype(np.float64), unpack_adj(g[f"adj_{t}"], n)), t
            if knn:
                assert np.array_equal(eng.obs[0], g[f"obs_{t}"]), t
    assert np.array_equal(rewards, g["rewards"])
    assert np.array_equal(eng.network[0], g["adj_mean_200"])
    eng.close()


@pytest.mark.parametrize("e,n,cen,knn", [(37, 100, True, False), (64, 50, False, True), (5, 257, True, True),
                                          (2, 1000, True, False), (3, 301, False, False), (2, 300, False, True), (3, 1, True, False), (9, 7, False, True)])
def test_exact_mode_batch_vs_oracle(e, n, cen, knn):
    """Cached-action rollouts (u=None) over a batch: every env bit-equal to the scalar C oracle."""
    from gym_flock_b200 import HostFlockBatch
    from gym_flock_b200.synthetic import initial_states
    from oracle import c_oracle
    x0 = initial_states(e, n, seed=777, min_dist=0.1 if n <= 100 else 0.0)
    steps = 200                    # north_star: "over a 200-step rollout", also at N = 1000 (BASELINE config 4)
    ref = c_oracle.rollout(x0, steps, centralized=cen, knn=7 if knn else 0)
    eng = HostFlockBatch(e, n, centralized=cen, knn_obs=knn, precision="fp64-exact")
    eng.set_state(x0)
    rewards = np.zeros((e, steps))
    for t in range(steps):
        eng.step(None, fetch_state=(t == steps - 1))
        rewards[:, t] = eng.reward
    assert np.array_equal(eng.x, ref["x"])
    assert np.array_equal(eng.values, ref["values"])
    assert np.array_equal(eng.network, ref["adj_mean"])
    assert np.array_equal(rewards, ref["rewards"])
    if knn:
        assert np.array_equal(eng.obs, ref["obs"])
    # the action applied in the last step is what the oracle reports
    eng2 = HostFlockBatch(e, n, centralized=cen, precision="fp64-exact")
    ref_prev = c_oracle.rollout(x0, steps - 1, centralized=cen)
    eng2.set_state(ref_prev["x"])
    assert np.array_equal(eng2.controller(cen), ref["u"])
    eng.close(); eng2.close()
