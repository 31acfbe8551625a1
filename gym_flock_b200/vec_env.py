"""FlockVecEnv: E lockstep flocking envs on one GPU, PyTorch CUDA tensors in and out.

torch is used for device memory and streams only; every numeric of the step is the hand-written
kernel behind the C ABI (include/flock_b200.h: flock_step_*, flock_rollout_*), called through ctypes
with raw device pointers on torch's current stream.  No CPU fallback: constructing one without a CUDA
device, or without the built library, raises.

Layout (all contiguous):
  x[E,N,4] float64 (px,py,vx,vy) -- canonical state, double in every precision mode
  u[E,N,2] float64 in [-1,1]
  values[E,N,6], network[E,N,N], reward[E], obs[E,N,28]: float64 ('fp64', 'fp64-exact') or float32 ('fp32')
  adjacency_mode='bitmask': instead of the dense network, adj_bits[E,N,ceil(N/32)] int32 (bit j&31 of word j>>5 of row i:
  agents i and j are neighbours) and degree[E,N] int32 -- 32x fewer bytes; `aggregate()` consumes them directly.
  nonfinite[E] int32: set when an env produced NaN/Inf (coincident agents give 0/0 in the reference too).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

_PREC = {"fp64-exact": (torch.float64, True), "fp64": (torch.float64, False), "fp32": (torch.float32, False)}
_IDS = {"FlockingRelative-v0": False, "Flocking-v0": True}


class FlockVecEnv:
    def __init__(self, env_id: str = "FlockingRelative-v0", num_envs: int = 1, n_agents: int = 100,
                 precision: str = "fp32", device: str | int | torch.device = "cuda:0",
                 comm_radius: float = 0.9, dt: float = 0.01, centralized: bool = True,
                 adjacency_mode: str = "dense", max_episode_steps: int = 1000, lanes: int = 0,
                 rows: tuple | None = None, action_scalar: float = 10.0, step_scale: float = 0.0, state_scale: float = 0.0,
                 u_clip: float = 0.0, n_leaders: int = 0, n_obstacles: int = 0, tune: int = 0):
        if env_id not in _IDS:
            raise KeyError(f"unknown env id {env_id!r}; provided: {sorted(_IDS)}")
        if precision not in _PREC:
            raise ValueError(f"precision must be one of {sorted(_PREC)}")
        if adjacency_mode not in ("dense", "bitmask", "none"):
            raise ValueError("adjacency_mode must be 'dense' (reference layout), 'bitmask' (exact adjacency bits + degree) "
                             "or 'none' (skip the adjacency output)")
        if adjacency_mode == "bitmask" and rows is not None:
            raise ValueError("adjacency_mode='bitmask' is offered for full rows only")
        if not torch.cuda.is_available():
            raise RuntimeError("FlockVecEnv needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.E, self.N = int(num_envs), int(n_agents)
        self.knn = _IDS[env_id]
        self.env_id = env_id
        self.precision = precision
        self.dtype, self.exact = _PREC[precision]
        self.centralized = bool(centralized)
        self.comm_radius, self.dt = float(comm_radius), float(dt)
        self.max_episode_steps = int(max_episode_steps)
        self.adjacency_mode = adjacency_mode
        self.rows = rows
        self.lanes = int(lanes)
        # flocking variants through the same kernel (FlockParams.step_scale / state_scale / u_clip / n_leaders / n_obstacles)
        self._variant = dict(action_scalar=float(action_scalar), step_scale=float(step_scale), state_scale=float(state_scale),
                             u_clip=float(u_clip), n_leaders=int(n_leaders), n_obstacles=int(n_obstacles), tune=int(tune))
        E, N, dev = self.E, self.N, self.device
        f64 = torch.float64
        self._x = [torch.zeros(E, N, 4, dtype=f64, device=dev) for _ in range(2)]
        self._u = [torch.zeros(E, N, 2, dtype=f64, device=dev) for _ in range(2)]
        self._cur = 0
        self.values = torch.zeros(E, N, 6, dtype=self.dtype, device=dev)
        self.network = torch.zeros(E, N, N, dtype=self.dtype, device=dev) if adjacency_mode == "dense" else None
        self.reward = torch.zeros(E, dtype=self.dtype, device=dev)
        self.obs = torch.zeros(E, N, 28, dtype=self.dtype, device=dev) if self.knn else None
        self.nonfinite = torch.zeros(E, dtype=torch.int32, device=dev)
        self.words = int(self.lib.flock_adj_words(N))
        self.adj_bits = torch.zeros(E, N, self.words, dtype=torch.int32, device=dev) if adjacency_mode == "bitmask" else None
        self.degree = torch.zeros(E, N, dtype=torch.int32, device=dev) if adjacency_mode == "bitmask" else None
        # Flocking-v0: last step's 7 nearest per row -- only a search-radius hint for the next step's kNN (never affects results)
        self._knn_hint = torch.zeros(E, N, 8, dtype=torch.int16, device=dev) if self.knn else None
        self._aux = _lib.FlockAux(self.nonfinite.data_ptr(),
                                  self.adj_bits.data_ptr() if self.adj_bits is not None else None,
                                  self.degree.data_ptr() if self.degree is not None else None, None, None,
                                  self._knn_hint.data_ptr() if self._knn_hint is not None else None)
        self.elapsed = 0
        self._fn_step = self.lib.flock_step_aux_f64 if self.dtype == torch.float64 else self.lib.flock_step_aux_f32
        self._fn_roll = self.lib.flock_rollout_f64 if self.dtype == torch.float64 else self.lib.flock_rollout_f32
        self._u_valid = False
        self._refresh_params()

    # ---- plumbing ----------------------------------------------------------------------------------------
    def _refresh_params(self):
        self._params = _lib.make_params(self.E, self.N, comm_radius=self.comm_radius, dt=self.dt,
                                        centralized=self.centralized, exact=self.exact, knn_obs=self.knn,
                                        rows=self.rows, lanes=self.lanes, **self._variant)
        self._pp = ctypes.byref(self._params)

    @staticmethod
    def _ptr(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _device_ctx(self):
        """torch.cuda.device(...) only when this thread's current device is not the env's (the context manager costs more
        than a launch)."""
        import contextlib
        return contextlib.nullcontext() if torch.cuda.current_device() == self.device.index else torch.cuda.device(self.device)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def x(self) -> torch.Tensor:
        return self._x[self._cur]

    @property
    def u_cached(self) -> torch.Tensor:
        """controller() action for the current state (valid after reset/step)."""
        return self._u[self._cur]

    def _observation(self):
        """(values | obs, network): the reference's observation tuple; in 'bitmask' mode the second item is adj_bits."""
        return (self.obs if self.knn else self.values), (self.adj_bits if self.adjacency_mode == "bitmask" else self.network)

    def _launch(self, x_in, u_in, x_out, u_next, want_obs=True, aux=None):
        if aux is None:
            aux = self._aux if want_obs else _lib.FlockAux(self.nonfinite.data_ptr(), None, None, None, None, None)
        with self._device_ctx():
            _lib.check(self._fn_step(self._pp, self._ptr(x_in), self._ptr(u_in), self._ptr(x_out),
                                     self._ptr(self.values) if want_obs else None,
                                     self._ptr(self.network) if want_obs else None,
                                     self._ptr(self.reward) if want_obs else None,
                                     self._ptr(u_next), self._ptr(self.obs) if want_obs else None,
                                     ctypes.byref(aux), self._stream()))

    # ---- API -----------------------------------------------------------------------------------------------
    def reset(self, x0=None, seed: int | None = None):
        """Inject initial states (x0[E,N,4], torch or NumPy; default: the synthetic cfg law, seed per env) and
        run the pair pass on them.  Returns (values|obs, network)."""
        if x0 is None:
            from .synthetic import DEFAULT_SEED, initial_states_fast
            x0 = initial_states_fast(self.E, self.N, DEFAULT_SEED if seed is None else seed)
        x0 = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0, dtype=torch.float64)
        self._x[self._cur].copy_(x0.reshape(self.E, self.N, 4), non_blocking=True)
        self._launch(self.x, None, None, self._u[self._cur])
        self._u_valid = True
        self.elapsed = 0
        return self._observation()

    def reset_random(self, seed: int = 0, r_max: float | None = None, v_max: float = 5.0, v_bias: float | None = None,
                     min_dist: float = 0.1, max_tries: int = 200000):
        """The reference's reset() (flocking_relative.py:156-192) for every env at once, on the device: each env draws
        whole configurations until no pair is closer than min_dist and every agent has two neighbours
        (flock_reset_sample: one CTA per env, Philox keyed by (seed, env)).  Defaults follow the reference's cfg path:
        r_max = sqrt(n_agents) (:75), v_bias = v_max (:41,49).  Returns ((values|obs, network), tries[E] int32); an env
        that found no admissible configuration has tries < 0 (the reference would loop forever there)."""
        r_max = float(np.sqrt(self.N)) if r_max is None else float(r_max)
        v_bias = float(v_max) if v_bias is None else float(v_bias)
        tries = torch.zeros(self.E, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.flock_reset_sample(self._pp, r_max, float(v_max), v_bias, float(min_dist),
                                                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(max_tries),
                                                   self._ptr(self._x[self._cur]), self._ptr(tries), self._stream()))
        self._launch(self.x, None, None, self._u[self._cur])
        self._u_valid = True
        self.elapsed = 0
        return self._observation(), tries

    def set_state(self, x0):
        return self.reset(x0)

    def get_state(self) -> torch.Tensor:
        return self.x.clone()

    def controller(self, centralized: bool | None = None) -> torch.Tensor:
        """flocking_relative.py:194-212 for every env: u[E,N,2] float64 (a view of the cached action)."""
        cen = self.centralized if centralized is None else bool(centralized)
        if cen == self.centralized and self._u_valid:
            return self.u_cached
        saved = self.centralized
        self.centralized = cen
        self._refresh_params()
        out = torch.empty_like(self.u_cached) if cen != saved else self._u[self._cur]
        self._launch(self.x, None, None, out, want_obs=False)
        self.centralized = saved
        self._refresh_params()
        if cen == saved:
            self._u_valid = True
        return out

    def step(self, u: torch.Tensor | None = None):
        """One lockstep env step.  u[E,N,2] (any float dtype, on the device) or None = apply controller().
        Returns ((values|obs, network), reward[E], done[E] bool, {})."""
        if u is None:
            u_in = self.controller()
        else:
            u_in = u.to(device=self.device, dtype=torch.float64).reshape(self.E, self.N, 2).contiguous()
        nxt = self._cur ^ 1
        u_next = self._u[nxt]
        if u_in.data_ptr() == u_next.data_ptr():     # caller handed us our own buffer back
            u_in = u_in.clone()
        self._launch(self.x, u_in, self._x[nxt], u_next)
        self._cur = nxt
        self._u_valid = True
        self.elapsed += 1
        done = torch.full((self.E,), self.elapsed >= self.max_episode_steps, dtype=torch.bool, device=self.device)
        return self._observation(), self.reward, done, {}

    def rollout(self, steps: int, rewards_out: torch.Tensor | None = None):
        """`steps` x {u = controller(); step(u)} launched back to back from C (no per-step Python).
        Observation tensors hold the last step; rewards_out[steps,E] (optional) gets every step's reward."""
        if not self._u_valid:
            self.controller()
        if self.adjacency_mode == "bitmask":         # the fused rollout has no bit output: step by step
            for t in range(int(steps)):
                self.step(None)
                if rewards_out is not None:
                    rewards_out[t].copy_(self.reward)
            return self._observation(), self.reward
        a, b = self._cur, self._cur ^ 1
        fin = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self._fn_roll(self._pp, int(steps), self._ptr(self._x[a]), self._ptr(self._x[b]),
                                     self._ptr(self._u[a]), self._ptr(self._u[b]), self._ptr(self.values),
                                     self._ptr(self.network), self._ptr(self.reward), self._ptr(rewards_out),
                                     self._ptr(self.obs), ctypes.byref(fin), self._stream()))
        self._cur = b if fin.value else a
        self.elapsed += int(steps)
        return self._observation(), self.reward

    def get_stats(self) -> dict:
        """get_stats() of every env for the current state (flocking_relative.py:136-143), computed by the kernel:
        {'vel_diffs': [E,N], 'min_dists': [E,N]} in the observation dtype."""
        vd = torch.empty(self.E, self.N, dtype=self.dtype, device=self.device)
        md = torch.empty(self.E, self.N, dtype=self.dtype, device=self.device)
        aux = _lib.FlockAux(None, None, None, md.data_ptr(), vd.data_ptr(), None)
        with torch.cuda.device(self.device):
            _lib.check(self._fn_step(self._pp, self._ptr(self.x), None, None, None, None, None, None, None,
                                     ctypes.byref(aux), self._stream()))
        return {"vel_diffs": vd, "min_dists": md}

    def aggregate(self, filter_len: int, prev: torch.Tensor | None = None, pool: str = "mean",
                  signal: torch.Tensor | None = None) -> torch.Tensor:
        """Graph-signal aggregation on the current adjacency, from its bits (adjacency_mode='bitmask'), never the dense
        [N,N] (gym_flock/envs/old/flocking_multi.py:180-196).  signal[E,N,F] defaults to `values`.
          prev given   : delayed aggregation, out[i] = [signal[i] | pool_{j in N(i)} prev[j][:F*(L-1)]]
          prev is None : K-hop, out[i] = [signal | S signal | ... | S^(L-1) signal], S = D^-1 A ('mean') or A ('sum')"""
        if self.adj_bits is None:
            raise RuntimeError("aggregate() needs adjacency_mode='bitmask'")
        x = self.values if signal is None else signal.to(device=self.device, dtype=self.dtype).contiguous()
        F, L = int(x.shape[-1]), int(filter_len)
        out = torch.empty(self.E, self.N, F * L, dtype=self.dtype, device=self.device)
        if prev is not None:
            prev = prev.to(device=self.device, dtype=self.dtype).contiguous()
            assert prev.shape == out.shape, (prev.shape, out.shape)
        fn = self.lib.flock_aggregate_f64 if self.dtype == torch.float64 else self.lib.flock_aggregate_f32
        with torch.cuda.device(self.device):
            _lib.check(fn(self._pp, self._ptr(x), F, self._ptr(prev), self._ptr(self.adj_bits), L,
                          {"mean": 0, "sum": 1}[pool], self._ptr(out), self._stream()))
        return out

    def close(self):
        pass
