"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/flock_oracle.c (see its header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libflock_oracle.so")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flock_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.flk_reward_f64.restype = ctypes.c_double
        _lib.flk_oracle_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_D)


def helpers(x, comm_radius=0.9, want_adj=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    values = np.empty((n, 6)); adj = np.empty((n, n)) if want_adj else None
    adj_mean = np.empty((n, n)) if want_adj else None
    lib().flk_helpers_f64(_p(x), ctypes.c_int(n), ctypes.c_double(comm_radius * comm_radius),
                          _p(values), _p(adj), _p(adj_mean))
    return {"values": values, "adj": adj, "adj_mean": adj_mean}


def controller(x, comm_radius=0.9, centralized=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    u = np.empty((n, 2))
    lib().flk_controller_f64(_p(x), ctypes.c_int(n), ctypes.c_double(comm_radius),
                             ctypes.c_double(comm_radius * comm_radius), ctypes.c_int(int(centralized)), _p(u))
    return u


def integrate(x, u, dt=0.01):
    y = np.array(x, dtype=np.float64, order="C").copy()
    u = np.ascontiguousarray(u, dtype=np.float64)
    lib().flk_integrate_f64(_p(y), _p(u), ctypes.c_int(y.shape[0]), ctypes.c_double(dt), ctypes.c_double(10.0))
    return y


def reward(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(lib().flk_reward_f64(_p(x), ctypes.c_int(x.shape[0])))


def knn_obs(x, k=7):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    obs = np.empty((n, 4 * k))
    lib().flk_knn_obs_f64(_p(x), ctypes.c_int(n), ctypes.c_int(k), _p(obs))
    return obs


def rows(x, row_idx, comm_radius=0.9, centralized=True, knn=0, want_network=True):
    """A subset of rows of ONE env x[N,4] (giant-swarm tests): dict of values[R,6], deg[R], adj_mean[R,N], u[R,2],
    obs[R,4*knn], r2_sorted[R,knn+2] -- the same per-row C functions as the full-matrix entry points."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    idx = np.ascontiguousarray(row_idx, dtype=np.int32)
    n, r = x.shape[0], idx.shape[0]
    values = np.empty((r, 6)); deg = np.empty(r); u = np.empty((r, 2))
    net = np.empty((r, n)) if want_network else None
    obs = np.empty((r, 4 * knn)) if knn else None
    r2s = np.empty((r, knn + 2)) if knn else None
    lib().flk_rows_f64(_p(x), ctypes.c_int(n), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_int(r),
                       ctypes.c_double(comm_radius), ctypes.c_double(comm_radius * comm_radius),
                       ctypes.c_int(int(centralized)), ctypes.c_int(knn), _p(values), _p(deg), _p(net), _p(u), _p(obs),
                       _p(r2s))
    return {"values": values, "deg": deg, "adj_mean": net, "u": u, "obs": obs, "r2_sorted": r2s}


def rollout(x0, steps, comm_radius=0.9, dt=0.01, centralized=True, knn=0, want_network=True, nthreads=0):
    """x0[E,N,4] (or [N,4]) -> dict of final x, values, adj_mean, u (last action), rewards[E,steps], obs."""
    x = np.array(x0, dtype=np.float64, order="C").copy()
    single = x.ndim == 2
    if single:
        x = x[None]
    e, n = x.shape[0], x.shape[1]
    values = np.empty((e, n, 6)); u = np.zeros((e, n, 2)); rewards = np.empty((e, steps))
    net = np.empty((e, n, n)) if want_network else None
    obs = np.empty((e, n, 4 * knn)) if knn else None
    lib().flk_rollout_f64(_p(x), ctypes.c_int(e), ctypes.c_int(n), ctypes.c_int(steps),
                          ctypes.c_double(comm_radius), ctypes.c_double(dt), ctypes.c_int(int(centralized)),
                          ctypes.c_int(knn), _p(values), _p(net), _p(u), _p(rewards), _p(obs),
                          ctypes.c_int(nthreads))
    out = {"x": x, "values": values, "adj_mean": net, "u": u, "rewards": rewards, "obs": obs}
    if single:
        out = {k: (v[0] if v is not None else None) for k, v in out.items()}
    return out


def threads() -> int:
    return int(lib().flk_oracle_threads())
