"""Synthetic initial flock states (host side, NumPy, deterministic per env seed).

The reference's own ``reset()`` (flocking_relative.py:156-192) is a whole-configuration
rejection sampler that never terminates with constructor defaults and is hopeless for
N >= 1000, so benchmarks and parity tests inject states instead.
Laws:

* ``cfg``      -- the law of the reference's cfg path without rejection:
                  radius = sqrt(U(0, sqrt(N))), angle = pi*U(0,2), v = U(-v_max, v_max) + bias,
                  bias ~ U(-v_max, v_max)^2  (flocking_relative.py:167-174 with r_max = sqrt(N), :75).
* ``density``  -- uniform in a disk of constant density rho (default 3.18 agents / unit area, the
                  N=100 cfg density) -- used for the giant swarm so the degree stays ~8.
* ``min_dist`` -- optional dart-throwing filter (agent by agent) so that no pair starts closer than
                  ``min_dist``; parity runs use 0.1 like the reference's acceptance test (:160).

Env e of a batch uses ``np.random.RandomState(seed + e)``; the bytes are platform independent up to
libm's cos/sin, so golden tests commit x0 bytes rather than seeds.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 1234
CFG_DENSITY = 100.0 / (np.pi * np.sqrt(100.0))  # N=100 agents in a disk of radius 100**0.25


def _positions(rs: np.random.RandomState, n: int, r_disk: float) -> np.ndarray:
    # radius = sqrt(U(0, r_disk^2)) is uniform in the disk; the reference's law is this with r_disk^2 = r_max
    length = np.sqrt(rs.uniform(0.0, r_disk * r_disk, size=n))
    angle = np.pi * rs.uniform(0.0, 2.0, size=n)
    return np.stack((length * np.cos(angle), length * np.sin(angle)), axis=1)


def initial_state(n_agents: int, seed: int = DEFAULT_SEED, law: str = "cfg", v_max: float = 5.0,
                  min_dist: float = 0.0, density: float = CFG_DENSITY) -> np.ndarray:
    """One env: float64 x0[N,4] = (px, py, vx, vy)."""
    rs = np.random.RandomState(seed)
    n = int(n_agents)
    if law == "cfg":
        r_disk = float(n) ** 0.25          # r_max = sqrt(N); radius = sqrt(U(0, r_max)) <= N**0.25
    elif law == "density":
        r_disk = float(np.sqrt(n / (density * np.pi)))
    else:
        raise ValueError(f"unknown law {law!r}")
    p = _positions(rs, n, r_disk)
    if min_dist > 0.0:
        # dart throwing: re-draw agents that land within min_dist of an already accepted one
        md2 = min_dist * min_dist
        for i in range(1, n):
            for _ in range(10000):
                d = p[:i] - p[i]
                if np.min(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) >= md2:
                    break
                p[i] = _positions(rs, 1, r_disk)[0]
            else:
                raise RuntimeError("min_dist filter failed: disk too dense")
    bias = rs.uniform(-v_max, v_max, size=2)
    v = rs.uniform(-v_max, v_max, size=(n, 2)) + bias
    return np.ascontiguousarray(np.concatenate((p, v), axis=1), dtype=np.float64)


def initial_states(n_envs: int, n_agents: int, seed: int = DEFAULT_SEED, **kw) -> np.ndarray:
    """Batch: float64 x0[E,N,4]; env e is initial_state(n_agents, seed + e)."""
    return np.stack([initial_state(n_agents, seed + e, **kw) for e in range(int(n_envs))], axis=0)


def initial_states_fast(n_envs: int, n_agents: int, seed: int = DEFAULT_SEED, law: str = "cfg",
                        v_max: float = 5.0, density: float = CFG_DENSITY) -> np.ndarray:
    """Same laws, one vectorised draw for the whole batch (benchmarks with 10^4+ envs or a 65k swarm).
    Not byte-identical to initial_states(); deterministic in (seed, E, N)."""
    rs = np.random.RandomState(seed)
    e, n = int(n_envs), int(n_agents)
    r_disk = float(n) ** 0.25 if law == "cfg" else float(np.sqrt(n / (density * np.pi)))
    length = np.sqrt(rs.uniform(0.0, r_disk * r_disk, size=(e, n)))
    angle = np.pi * rs.uniform(0.0, 2.0, size=(e, n))
    bias = rs.uniform(-v_max, v_max, size=(e, 1, 2))
    v = rs.uniform(-v_max, v_max, size=(e, n, 2)) + bias
    x = np.empty((e, n, 4), dtype=np.float64)
    x[:, :, 0] = length * np.cos(angle)
    x[:, :, 1] = length * np.sin(angle)
    x[:, :, 2:4] = v
    return x
