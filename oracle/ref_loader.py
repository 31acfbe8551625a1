"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* gym-flock reference.

Only ``tests/``, ``tests/golden/make_golden.py``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module.
The product package (``gym_flock_b200``) never does.

The reference (``gym_flock/envs/flocking/flocking_relative.py`` and
``flocking.py``) is plain Python/NumPy but imports ``gym`` and ``matplotlib``
(absent in this image) and uses ``np.Inf`` (removed in NumPy 2).  This module
installs three shims -- stub ``gym``/``gym.spaces``/``gym.utils.seeding``
modules, a stub ``matplotlib.pyplot`` and ``np.Inf = np.inf`` -- and then
executes the two reference files *by path*, without touching them.

Search order for the reference tree: ``$GYM_FLOCK_REF``, ``/root/reference``,
``<repo>/baseline/_ref`` (the git-ignored ``pip install --target`` copy of the
unmodified reference that travels to the GPU box).  Where none
exists ``available()`` is False and callers fall back to the committed golden
vectors + the restatements in ``oracle/flock_numpy.py`` / ``oracle/flock_oracle.c``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

_HOT_FILES = ("gym_flock/envs/flocking/flocking_relative.py", "gym_flock/envs/flocking/flocking.py")
_cache: dict = {}


def reference_root() -> str | None:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("GYM_FLOCK_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and all(os.path.isfile(os.path.join(cand, f)) for f in _HOT_FILES):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


class _Box:
    """Stand-in for gym.spaces.Box: records its arguments, nothing else."""

    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape) if shape is not None else None, dtype


class _Env:
    """Stand-in for gym.Env (the reference only subclasses it)."""

    metadata: dict = {}


def _np_random(seed=None):
    # gym 0.11-era contract: returns (RandomState-like, seed). Flocking never draws from it.
    if seed is None:
        seed = 0
    return np.random.RandomState(int(seed) % (2 ** 32)), seed


def _install_shims() -> None:
    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # removed in NumPy 2.x; reference uses it at flocking_relative.py:59,80,115,179
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = _Env
        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Box
        error = types.ModuleType("gym.error")
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")
        seeding.np_random = _np_random
        utils.seeding = seeding
        gym.spaces, gym.error, gym.utils = spaces, error, utils
        gym.__oracle_stub__ = True
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.error": error,
                            "gym.utils": utils, "gym.utils.seeding": seeding})
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            plt.gca = lambda *a, **k: None
            mpl.pyplot = plt
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


def _exec(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns (FlockingRelativeEnv, FlockingEnv) classes of the unmodified reference."""
    if "classes" in _cache:
        return _cache["classes"]
    root = reference_root()
    if root is None:
        raise RuntimeError("gym-flock reference tree not found (set $GYM_FLOCK_REF); "
                           "use the committed golden vectors instead")
    _install_shims()
    # flocking.py does `from gym_flock.envs.flocking.flocking_relative import ...`; provide that
    # dotted path without executing gym_flock/__init__.py (which registers 15 unrelated envs).
    for pkg in ("gym_flock", "gym_flock.envs", "gym_flock.envs.flocking"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []  # mark as package
            sys.modules[pkg] = m
    rel = _exec(os.path.join(root, _HOT_FILES[0]), "gym_flock.envs.flocking.flocking_relative")
    flk = _exec(os.path.join(root, _HOT_FILES[1]), "gym_flock.envs.flocking.flocking")
    _cache["classes"] = (rel.FlockingRelativeEnv, flk.FlockingEnv)
    return _cache["classes"]


def load_variant(module: str, cls: str):
    """Load another class of the reference's flocking package (e.g. 'flocking_leader', 'FlockingLeaderEnv')."""
    load()
    root = reference_root()
    for dep in ("utils", module):
        name = f"gym_flock.envs.flocking.{dep}"
        if name not in sys.modules:
            _exec(os.path.join(root, "gym_flock/envs/flocking", dep + ".py"), name)
    return getattr(sys.modules[f"gym_flock.envs.flocking.{module}"], cls)


class CfgSection(dict):
    """Minimal configparser-section look-alike for params_from_cfg (flocking_relative.py:68-85)."""

    def getfloat(self, k):
        return float(self[k])

    def getint(self, k):
        return int(self[k])


def make_ref_env(kind: str = "relative", n_agents: int = 100, comm_radius: float = 0.9,
                 dt: float = 0.01, v_max: float = 5.0, x0: np.ndarray | None = None):
    """Instantiate a reference env with the given parameters and (optionally) inject state x0.

    Parameters are set by attribute (what params_from_cfg does minus the r_max side effect);
    state injection is `env.x = x0.copy(); env.compute_helpers()`.
    """
    Rel, Flk = load()
    env = Rel() if kind == "relative" else Flk()
    env.n_agents = int(n_agents)
    env.comm_radius = float(comm_radius)
    env.comm_radius2 = env.comm_radius * env.comm_radius
    env.dt = float(dt)
    env.v_max = float(v_max)
    env.v_bias = env.v_max
    if x0 is not None:
        env.x = np.array(x0, dtype=np.float64).copy()
        env.compute_helpers()
    return env
