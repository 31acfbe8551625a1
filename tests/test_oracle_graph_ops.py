"""CPU: the restatements of the two adjacent components (SURVEY.md 8f #3 graph-signal aggregation, #4 shepherding
repulsion) are bit-exact against the UNMODIFIED reference code, where the reference tree exists."""
import importlib.util
import os
import types

import numpy as np
import pytest

from oracle import flock_numpy as fnp, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _load(relpath, name):
    ref_loader.load()                      # installs the gym / matplotlib / np.Inf shims
    import sys
    for stub in ("matplotlib.patches",):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except Exception:
                sys.modules[stub] = types.ModuleType(stub)
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_loader.reference_root(), relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n,filter_len", [(30, 3), (100, 2), (57, 4)])
def test_delayed_aggregation_matches_reference_aggregate(n, filter_len):
    """FlockingMultiEnv.aggregate() (old/flocking_multi.py:180-196) called on a duck-typed self: its constructor reads a
    config file that is not in the tree, but the method only needs a handful of attributes."""
    mod = _load("gym_flock/envs/old/flocking_multi.py", "ref_flocking_multi")
    cls = mod.FlockingMultiEnv
    rs = np.random.RandomState(n)
    x = np.zeros((n, 4))
    x[:, :2] = rs.uniform(-2.5, 2.5, size=(n, 2))
    x[:, 2:] = rs.uniform(-1, 1, size=(n, 2))
    nx = 6
    me = types.SimpleNamespace(n_nodes=n, nx=nx, filter_len=filter_len, n_pools=1, pooling=[np.nanmean],
                               init_vel=rs.uniform(-1, 1, size=(n, 2)), comm_radius2=0.81, n_features=nx * filter_len)
    for name in ("get_x_features", "get_connectivity", "get_features", "get_comms", "get_pool"):
        setattr(me, name, types.MethodType(getattr(cls, name), me))
    x_agg = rs.uniform(-1, 1, size=(n, nx * filter_len, 1))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = cls.aggregate(me, x, x_agg.copy())[:, :, 0]
    a_net = me.get_connectivity(x)
    got = fnp.aggregate_delayed(me.get_x_features(x), x_agg[:, :, 0], a_net)
    assert np.array_equal(got, ref)
    # a frozen graph and signal: L-1 delayed steps from zero reach the K-hop aggregate
    xf = me.get_x_features(x)
    agg = np.zeros((n, nx * filter_len))
    for _ in range(filter_len):
        agg = fnp.aggregate_delayed(xf, agg, a_net)
    assert np.array_equal(agg, fnp.aggregate_khop(xf, a_net, filter_len))


def test_sheep_repulsion_matches_reference():
    mod = _load("gym_flock/envs/shepherding/shepherding.py", "ref_shepherding")
    env = mod.ShepherdingEnv()
    rs = np.random.RandomState(3)
    for _ in range(5):
        env.x[:, 0:2] = rs.uniform(-2.0, 2.0, size=(env.n_agents, 2))
        env.x[:, 2] = rs.uniform(-3, 3, size=env.n_agents)
        ref = env._compute_sheep_controller()
        got = fnp.sheep_repulsion(env.x, env.force_weights)[env.n_shepherds:, 0:2]
        assert np.array_equal(got, ref)
