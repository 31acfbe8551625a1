"""Host-buffer batch engine: E lockstep flocking envs behind NumPy arrays, on one GPU.

Thin Python over the C ABI's context API (include/flock_b200.h: flock_ctx_*): every call copies its
host inputs to the device, launches the fused step kernel and copies the requested outputs back.
This is the layer the gym drop-in classes (envs.py) sit on and what bench.py's `e2e` number times.
torch is NOT needed here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

_PRECISIONS = {"fp64-exact": (8, True), "fp64": (8, False), "fp32": (4, False)}


class PinnedArray:
    """NumPy view over page-locked host memory from flock_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        nbytes = max(int(np.prod(self.shape)) * self.dtype.itemsize, 16)
        self._ptr = _lib.lib().flock_host_alloc(nbytes)
        if not self._ptr:
            raise MemoryError(f"flock_host_alloc({nbytes}) failed")
        buf = (ctypes.c_char * nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr:
            try:
                _lib.lib().flock_host_free(ptr)
            except Exception:
                pass


class HostFlockBatch:
    """E independent envs x N agents; state and outputs live in (pinned) NumPy arrays.

    precision: 'fp64-exact' (bit-exact vs the reference NumPy code), 'fp64' (fast, <=1e-10), 'fp32'.
    """

    def __init__(self, n_envs: int, n_agents: int, comm_radius: float = 0.9, dt: float = 0.01,
                 centralized: bool = True, knn_obs: bool = False, precision: str = "fp64-exact",
                 device: int = 0, lanes: int = 0, comm_radius2: float | None = None, pinned: bool = True,
                 step_scale: float = 0.0, state_scale: float = 0.0, u_clip: float = 0.0, n_leaders: int = 0,
                 n_obstacles: int = 0, action_scalar: float = 10.0, tune: int = 0):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        _lib.require_device()
        self.E, self.N = int(n_envs), int(n_agents)
        self.tbytes, self.exact = _PRECISIONS[precision]
        self.precision = precision
        self.dtype = np.float64 if self.tbytes == 8 else np.float32
        self.knn_obs = bool(knn_obs)
        self._kw = dict(comm_radius=comm_radius, dt=dt, centralized=centralized, exact=self.exact,
                        knn_obs=knn_obs, lanes=lanes, comm_radius2=comm_radius2, step_scale=step_scale,
                        state_scale=state_scale, u_clip=u_clip, n_leaders=n_leaders, n_obstacles=n_obstacles,
                        action_scalar=action_scalar, tune=tune)
        self._params = _lib.make_params(self.E, self.N, **self._kw)
        ctx = ctypes.c_void_p()
        _lib.check(_lib.lib().flock_ctx_create(ctypes.byref(ctx), ctypes.byref(self._params),
                                               self.tbytes, int(device)))
        self._ctx = ctx
        E, N = self.E, self.N
        mk = (lambda s, d: PinnedArray(s, d)) if pinned else None
        self._bufs = {}

        def alloc(name, shape, dtype):
            if mk:
                pa = mk(shape, dtype)
                self._bufs[name] = pa
                return pa.array
            return np.empty(shape, dtype)

        self.x = alloc("x", (E, N, 4), np.float64)
        self.u_in = alloc("u_in", (E, N, 2), np.float64)
        self.u_ctrl = alloc("u_ctrl", (E, N, 2), np.float64)
        self.values = alloc("values", (E, N, 6), self.dtype)
        self.network = alloc("network", (E, N, N), self.dtype)
        self.reward = alloc("reward", (E,), self.dtype)
        self.obs = alloc("obs", (E, N, 28), self.dtype) if knn_obs else None

    # -- plumbing -------------------------------------------------------------------------------------
    @staticmethod
    def _p(a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)

    def close(self):
        ctx, self._ctx = getattr(self, "_ctx", None), None
        if ctx:
            _lib.lib().flock_ctx_destroy(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        self._kw.update(kw)
        self._params = _lib.make_params(self.E, self.N, **self._kw)
        _lib.check(_lib.lib().flock_ctx_set_params(self._ctx, ctypes.byref(self._params)))

    # -- the three operations of the path ----------------------------------------------------------------
    def set_state(self, x0=None, want_network: bool = True):
        """x <- x0 (host, [E,N,4] float64) and run the pair pass on it: reference `env.x = x0;
        env.compute_helpers()` (flocking_relative.py:111-134).  Fills values/network/reward/obs."""
        if x0 is not None:
            self.x[...] = np.asarray(x0, dtype=np.float64).reshape(self.E, self.N, 4)
        _lib.check(_lib.lib().flock_ctx_set_state(self._ctx, self._p(self.x), self._p(self.values),
                                                  self._p(self.network) if want_network else None,
                                                  self._p(self.reward), self._p(self.obs)))

    def step(self, u=None, want_network: bool = True, fetch_state: bool = True):
        """One env step (flocking_relative.py:91-109) for all E envs.  u: host [E,N,2] in [-1,1], or None
        to apply the controller action cached on the device."""
        up = None
        if u is not None:
            self.u_in[...] = np.asarray(u, dtype=np.float64).reshape(self.E, self.N, 2)
            up = self._p(self.u_in)
        _lib.check(_lib.lib().flock_ctx_step(self._ctx, up, self._p(self.values),
                                             self._p(self.network) if want_network else None,
                                             self._p(self.reward), self._p(self.obs)))
        if fetch_state:
            _lib.check(_lib.lib().flock_ctx_get_state(self._ctx, self._p(self.x)))

    def controller(self, centralized: bool = True) -> np.ndarray:
        """controller() for the current state (flocking_relative.py:194-212): host [E,N,2]."""
        _lib.check(_lib.lib().flock_ctx_controller(self._ctx, int(bool(centralized)), self._p(self.u_ctrl)))
        return self.u_ctrl

    def get_state(self) -> np.ndarray:
        _lib.check(_lib.lib().flock_ctx_get_state(self._ctx, self._p(self.x)))
        return self.x

    def upload_state(self, x) -> None:
        """x <- host state WITHOUT the pair pass: the reference's in-place edit of env.x (flocking_leader.py:38-39).
        The cached controller action and the observation outputs keep describing the previous state until the next step."""
        self.x[...] = np.asarray(x, dtype=np.float64).reshape(self.E, self.N, 4)
        _lib.check(_lib.lib().flock_ctx_upload_state(self._ctx, self._p(self.x)))

    def get_stats(self) -> dict:
        """get_stats() of every env for the current state (flocking_relative.py:136-143): vel_diffs[E,N], min_dists[E,N]."""
        vd = np.empty((self.E, self.N), self.dtype)
        md = np.empty((self.E, self.N), self.dtype)
        _lib.check(_lib.lib().flock_ctx_get_stats(self._ctx, self._p(vd), self._p(md)))
        return {"vel_diffs": vd, "min_dists": md}
