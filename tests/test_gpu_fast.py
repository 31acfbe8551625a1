"""GPU: fast fp64 (<= 1e-10) and fp32 (<= 1e-4) modes against the oracle, SURVEY.md 8a metric:
teacher-forced at points of the reference trajectory + a free-running 200-step rollout."""
import numpy as np
import pytest

from parity_util import TOL, check_knn_obs, check_pass, normwise

pytestmark = pytest.mark.gpu

CASES = [  # (E, N, centralized, knn, lanes)
    (8, 100, True, False, 0), (8, 100, False, True, 1), (6, 50, True, True, 4), (3, 257, True, True, 32),
    (2, 1000, True, False, 0), (4, 33, False, False, 2), (5, 7, True, True, 8), (3, 2, True, False, 0),
    # row-expansion shapes (one thread per row): long rows with 16-byte groups (nv = 50, 33), 8-byte groups (66 -> nv 33,
    # 130 -> nv 65), short rows handled several per iteration (N = 12, 20), partial last warp
    (3, 200, True, False, 1), (3, 132, False, False, 1), (5, 66, True, False, 1), (2, 130, True, False, 1),
    (40, 12, True, False, 1), (27, 20, False, False, 1),
    # kNN with one thread per row (fp32 mode: packed-key min/max network), incl. fewer than 7 neighbours and odd N
    (9, 100, True, True, 1), (4, 7, True, True, 1), (3, 9, False, True, 1), (2, 257, True, True, 1),
    # whole envs per CTA (flock_group_kernel): compile-time sizes N = 100 / 50, a last CTA with fewer envs, odd N, N = 512
    (7, 100, True, False, 1), (11, 50, False, False, 1), (13, 50, True, True, 1), (3, 101, True, True, 1), (2, 512, True, False, 1),
    (1, 1, True, False, 1), (600, 10, True, False, 1),
]


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("e,n,cen,knn,lanes", CASES)
def test_teacher_forced_and_free_running(precision, e, n, cen, knn, lanes):
    from gym_flock_b200 import HostFlockBatch
    from gym_flock_b200.synthetic import initial_states
    from oracle import c_oracle
    kind = precision
    # states obey the reference's own reset() invariant (min pairwise distance >= 0.1, flocking_relative.py:160).
    x0 = initial_states(e, n, seed=4321 + n, min_dist=0.1)
    steps = 200 if n <= 257 else 40
    # The centralized alignment term is -N*(v_i - mean v); explicit Euler multiplies a velocity deviation by
    # (1 - N*dt) per unclipped step, so for N*dt > 2 the REFERENCE dynamics amplify any 1-ulp difference
    # exponentially (1.57^200 at N=257): only the bit-exact mode can track a free-running rollout there.
    # Non-exact modes are compared free-running in the stable regime N*dt <= 1 (dt = 0.01 up to N = 100).
    dt = min(0.01, 1.0 / n)
    eng = HostFlockBatch(e, n, centralized=cen, knn_obs=knn, precision=precision, lanes=lanes, dt=dt)

    # (i) teacher-forced along the reference trajectory
    x_ref, t_prev = x0, 0
    worst = {}
    for t in (0, 1, 2, steps // 4, steps // 2, steps):
        x_ref = c_oracle.rollout(x_ref, t - t_prev, centralized=cen, dt=dt)["x"] if t > t_prev else x_ref
        t_prev = t
        eng.set_state(x_ref)
        u = eng.controller(cen)
        for k in range(e):
            errs = check_pass(x_ref[k], eng.values[k], eng.network[k], u[k], eng.reward[k], kind, centralized=cen,
                              got_obs=None)
            for kk, vv in errs.items():
                worst[kk] = max(worst.get(kk, 0), vv)
            if knn and n >= 7:
                # every row, exactly the reference's neighbours in the reference's order (the fast modes re-rank their
                # survivors with the reference's own keys); only a provable membership near-tie is exempt
                worst["knn_near_tie_rows"] = max(worst.get("knn_near_tie_rows", 0), check_knn_obs(eng.obs[k], x_ref[k], kind))

    # (ii) free-running with the device-side controller
    print("teacher-forced worst errors:", worst)
    ref = c_oracle.rollout(x0, steps, centralized=cen, dt=dt)
    eng.set_state(x0)
    rewards = np.zeros((e, steps))
    for t in range(steps):
        eng.step(None, fetch_state=(t == steps - 1), want_network=(t == steps - 1))
        rewards[:, t] = eng.reward
    tol = TOL[kind]
    for k in range(e):
        assert normwise(eng.x[k], ref["x"][k]) <= tol, (k, normwise(eng.x[k], ref["x"][k]))
    assert normwise(rewards, ref["rewards"]) <= tol
    eng.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_lanes_agree(precision):
    """Row arithmetic split over 1..32 lanes: same adjacency, values within the mode's tolerance."""
    from gym_flock_b200 import HostFlockBatch
    from gym_flock_b200.synthetic import initial_states
    x0 = initial_states(3, 100, seed=99, min_dist=0.1)
    outs = []
    for lanes in (1, 2, 4, 8, 16, 32):
        eng = HostFlockBatch(3, 100, knn_obs=True, precision=precision, lanes=lanes)
        eng.set_state(x0)
        outs.append((eng.values.copy(), eng.network.copy(), eng.controller(True).copy(), eng.obs.copy()))
        eng.close()
    v0, n0, u0, o0 = outs[0]
    for v, nw, u, o in outs[1:]:
        assert np.array_equal(nw, n0)
        assert normwise(v, v0) <= TOL[precision] and normwise(u, u0) <= TOL[precision]
        assert normwise(o, o0) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp64-exact", "fp64", "fp32"])
@pytest.mark.parametrize("n,lanes", [(100, 0), (100, 1), (257, 4), (64, 2)])
def test_dense_flock_overflows_candidate_lists(precision, n, lanes):
    """A tight cluster (everybody within the radius of everybody): more candidates per row than the kernel's
    candidate lists hold, so the all-pairs fallback path runs.  Same parity bars as everywhere else."""
    from gym_flock_b200 import HostFlockBatch
    from oracle import c_oracle
    rs = np.random.RandomState(n)
    e = 3
    x0 = np.zeros((e, n, 4))
    side = int(np.ceil(np.sqrt(n)))
    grid = np.stack(np.meshgrid(np.arange(side), np.arange(side)), -1).reshape(-1, 2)[:n] * (0.8 / side)
    x0[:, :, :2] = grid[None] + rs.uniform(-0.01, 0.01, size=(e, n, 2))
    x0[:, :, 2:] = rs.uniform(-1, 1, size=(e, n, 2))
    if precision == "fp64-exact" and lanes > 1:
        pytest.skip("exact mode is one thread per row")
    eng = HostFlockBatch(e, n, knn_obs=True, precision=precision, lanes=lanes)
    eng.set_state(x0)
    u = eng.controller(True)
    for k in range(e):
        h = c_oracle.helpers(x0[k])
        assert h["adj"].sum(1).min() > 48
        if precision == "fp64-exact":
            assert np.array_equal(eng.values[k], h["values"]) and np.array_equal(eng.network[k], h["adj_mean"])
            assert np.array_equal(u[k], c_oracle.controller(x0[k]))
            assert np.array_equal(eng.obs[k], c_oracle.knn_obs(x0[k]))
        else:
            check_pass(x0[k], eng.values[k], eng.network[k], u[k], eng.reward[k], precision)
    eng.close()
