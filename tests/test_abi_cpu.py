"""CPU: the C-ABI library loads, exports every symbol include/flock_b200.h declares, and rejects bad
arguments without touching a GPU.  (No compute calls here.)"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from gym_flock_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "flock_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(flock_[a-z0-9_]+)\s*\(", hdr))
    assert {"flock_step_f64", "flock_step_f32", "flock_rollout_f64", "flock_rollout_f32", "flock_ctx_create",
            "flock_ctx_step", "flock_ctx_controller", "flock_last_error", "flock_step_aux_f32", "flock_aggregate_f64",
            "flock_repulsion_f64", "flock_ctx_get_stats", "flock_ctx_upload_state", "flock_adj_words"} <= names
    for n in sorted(names):
        assert hasattr(L, n), n
    from gym_flock_b200 import _lib
    assert L.flock_abi_version() == _lib.ABI_VERSION == 4


def test_params_struct_layout_matches_header():
    from gym_flock_b200._lib import FlockParams
    assert ctypes.sizeof(FlockParams) == 96
    assert FlockParams.comm_radius.offset == 8 and FlockParams.flags.offset == 40 and FlockParams.lanes.offset == 56
    assert FlockParams.step_scale.offset == 64 and FlockParams.n_leaders.offset == 88


def test_aux_struct_layout_and_validation(L):
    from gym_flock_b200 import _lib
    assert ctypes.sizeof(_lib.FlockAux) == 48 and _lib.FlockAux.vel_diff.offset == 32 and _lib.FlockAux.knn_hint.offset == 40
    assert L.flock_adj_words(100) == 4 and L.flock_adj_words(32) == 1 and L.flock_adj_words(33) == 2
    buf = np.zeros(64, dtype=np.float64)
    ptr = ctypes.c_void_p((buf.ctypes.data + 15) & ~15)
    aux = _lib.FlockAux(ptr.value + 2, None, None, None, None, None)           # misaligned int32 flag buffer
    p = _lib.make_params(1, 4)
    assert L.flock_step_aux_f64(ctypes.byref(p), ptr, None, None, None, None, None, None, None, ctypes.byref(aux), None) == -2
    assert p.reserved == 0                                                      # bindings never set the diagnostics word
    assert L.flock_aggregate_f64(ctypes.byref(p), ptr, 0, None, ptr, 2, 0, ptr, None) == -1       # F < 1
    assert L.flock_aggregate_f64(ctypes.byref(p), ptr, 6, None, ptr, 2, 7, ctypes.c_void_p(ptr.value + 64), None) == -1   # pool
    assert L.flock_repulsion_f64(ctypes.byref(p), ptr, 1, ptr, 2.0, ptr, None) == -1               # stride < 2
    os.environ["FLOCK_TUNE"] = "4096"                                          # a stray environment variable changes nothing
    try:
        assert _lib.make_params(1, 4).reserved == 0
    finally:
        del os.environ["FLOCK_TUNE"]


def test_argument_validation_needs_no_gpu(L):
    from gym_flock_b200 import _lib
    buf = np.zeros(64, dtype=np.float64)
    ptr = ctypes.c_void_p((buf.ctypes.data + 15) & ~15)

    def call(p, x=ptr, xo=None):
        return L.flock_step_f64(ctypes.byref(p), x, None, xo, None, None, None, None, None, None)

    assert call(_lib.make_params(0, 4)) == -1                      # FLOCK_EINVAL: E < 1
    assert b"n_envs" in L.flock_last_error()
    assert call(_lib.make_params(1, 4, lanes=3)) == -1             # lanes not a power of two
    assert call(_lib.make_params(1, 4, exact=True, lanes=4)) == -1  # exact needs sequential j
    assert call(_lib.make_params(1, 4, rows=(3, 2))) == -1
    assert call(_lib.make_params(1, 100000)) == -3                 # FLOCK_EUNSUPPORTED: beyond resident smem
    assert call(_lib.make_params(1, 4), x=None) == -1
    assert call(_lib.make_params(1, 4), xo=ptr) == -1              # x_out aliases x_in
    assert call(_lib.make_params(1, 4), x=ctypes.c_void_p(ptr.value + 8)) == -2   # FLOCK_EALIGN
    p32 = _lib.make_params(1, 4, exact=True)
    assert L.flock_step_f32(ctypes.byref(p32), ptr, None, None, None, None, None, None, None, None) == -1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import gym_flock_b200 as gf
    with pytest.raises(RuntimeError):
        gf.HostFlockBatch(1, 10)
    with pytest.raises(RuntimeError):
        gf.FlockVecEnv(num_envs=1, n_agents=10)
    env = gf.FlockingRelativeEnv()
    env.n_agents = 4
    with pytest.raises(RuntimeError):
        env.set_state(np.zeros((4, 4)))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gym_flock_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gym_flock_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("__oracle_stub__", ""), os.path.join(dirpath, f)


def test_drop_in_surface_and_synthetic_states():
    import gym_flock_b200 as gf
    from gym_flock_b200.synthetic import initial_state, initial_states
    env = gf.FlockingRelativeEnv()
    for attr, val in dict(n_agents=100, comm_radius=0.9, dt=0.01, v_max=5.0, r_max=1.0, action_scalar=10.0,
                          nx_system=4, n_features=6, nu=2, mean_pooling=True, centralized=True).items():
        assert getattr(env, attr) == val
    assert abs(env.comm_radius2 - 0.81) < 1e-15 and isinstance(env.seed(5), list)
    assert set(gf.envs.REGISTRY) == {"Flocking-v0", "FlockingRelative-v0", "FlockingTwoFlocks-v0", "FlockingLeader-v0",
                                        "FlockingStochastic-v0", "FlockingObstacle-v0"}
    with pytest.raises(KeyError):
        gf.make("Shepherding-v0")
    a, b = initial_state(50, seed=7, min_dist=0.1), initial_states(2, 50, seed=7, min_dist=0.1)
    assert np.array_equal(a, b[0]) and not np.array_equal(b[0], b[1])
    d = a[:, None, :2] - a[None, :, :2]
    r2 = (d ** 2).sum(-1) + np.eye(50) * 9
    assert r2.min() >= 0.01


def test_render_overlays_of_the_leader_and_obstacle_envs(monkeypatch):
    """render() is a matplotlib pass-through (flocking_relative.py:234-257); the leader env adds an arrow per leader
    (flocking_leader.py:42-57), the obstacle env draws its obstacles over the scatter (flocking_obstacle.py:106-118).
    Host-only: a recording stand-in for matplotlib, no device."""
    import sys
    import types
    import numpy as np
    calls = []

    class Rec:
        def __init__(self, name): self.name = name
        def __getattr__(self, attr):
            def f(*a, **k):
                calls.append((self.name, attr, a))
                if attr in ("plot",):
                    return (Rec("line"),)
                return Rec(attr)
            return f

    plt = types.ModuleType("matplotlib.pyplot")
    ax = Rec("ax")
    fig = Rec("fig")
    fig.canvas = Rec("canvas")
    fig.add_subplot = lambda *a: ax
    plt.ion = lambda: None
    plt.figure = lambda: fig
    plt.ylim = plt.xlim = plt.title = lambda *a, **k: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    from gym_flock_b200.envs import FlockingLeaderEnv, FlockingObstacleEnv
    rng = np.random.default_rng(0)
    lead = FlockingLeaderEnv()
    lead._x = rng.standard_normal((lead.n_agents, 4))
    lead.render()
    q = [c for c in calls if c[1] == "quiver"]
    assert len(q) == 1 and np.array_equal(q[0][2][0], lead._x[:2, 0]) and np.array_equal(q[0][2][3], lead._x[:2, 3])
    lead.render()                                            # second frame: the arrows are moved, not re-created
    assert len([c for c in calls if c[1] == "quiver"]) == 1
    assert any(c[1] == "set_UVC" for c in calls) and any(c[1] == "set_offsets" for c in calls)
    calls.clear()
    obst = FlockingObstacleEnv()
    obst._x = rng.standard_normal((obst.n_agents, 4))
    obst.render()
    red = [c for c in calls if c[1] == "plot" and len(c[2]) == 3 and c[2][2] == 'ro']
    assert len(red) == 1 and np.array_equal(red[0][2][0], obst._x[:4, 0])
