"""SwarmEnv: ONE giant flock whose rows (agents i) are sharded over the ranks of a torch.distributed group.

Protocol per step: every rank holds the full state x[N,4] and the full
action u[N,2]; rank r runs the fused step kernel for rows [r*N/R, (r+1)*N/R) only (the kernel integrates all N agents
into its staging copy, but computes features / adjacency rows / kNN obs / the next action for its rows only) and the
ranks all-gather the new actions (N*2 float64: half the bytes of gathering x) so that every rank can integrate all rows
at the next step.  No other data-path collective.  Row arithmetic never depends on the shard: world sizes 1 and 8
produce identical bits.

Outputs stay row-sharded: values[rows,6] / obs[rows,28], network[rows,N], and the (replicated) scalar reward.

`row_backend` lets tests replace the CUDA launch by a CPU stand-in to exercise the sharding / all-gather logic under
gloo without a GPU; the product path (row_backend=None) is the CUDA kernel and raises without it.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

_IDS = {"FlockingRelative-v0": False, "Flocking-v0": True}
_PREC = {"fp64-exact": (torch.float64, True), "fp64": (torch.float64, False), "fp32": (torch.float32, False)}


def shard_rows(n_agents: int, world: int, rank: int) -> tuple:
    """Equal contiguous row blocks (the in-place all-gather needs equal shards)."""
    if n_agents % world:
        raise ValueError(f"n_agents={n_agents} must be divisible by the world size {world}")
    per = n_agents // world
    return rank * per, (rank + 1) * per


class SwarmEnv:
    def __init__(self, env_id: str = "Flocking-v0", n_agents: int = 65536, precision: str = "fp32",
                 device=None, group=None, comm_radius: float = 0.9, dt: float = 0.01, centralized: bool = True,
                 adjacency_mode: str = "dense", row_backend=None, lanes: int = 32, streaming: bool | None = None,
                 distributed: bool | None = None, tune: int = 0):
        if env_id not in _IDS:
            raise KeyError(env_id)
        self.knn = _IDS[env_id]
        self.N = int(n_agents)
        self.group = group
        # distributed=False: a single-rank swarm (all rows on this GPU) even inside an initialised process group --
        # the reference run the multi-rank checks compare against
        self.distributed = (dist.is_available() and dist.is_initialized()) if distributed is None else bool(distributed)
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.lo, self.hi = shard_rows(self.N, self.world, self.rank)
        emu = os.environ.get("FLOCK_EMULATE_SHARD")          # diagnostics: "rank,world" -> time one shard on one GPU
        if emu and not self.distributed:
            r, w = (int(v) for v in emu.split(","))
            self.lo, self.hi = shard_rows(self.N, w, r)
        self.dtype, self.exact = _PREC[precision]
        self.precision = precision
        self.centralized, self.comm_radius, self.dt = bool(centralized), float(comm_radius), float(dt)
        self.row_backend = row_backend
        self._skip_gather = False        # diagnostics only
        self.adjacency_mode = adjacency_mode
        if row_backend is None:
            if not torch.cuda.is_available():
                raise RuntimeError("SwarmEnv needs a CUDA device (no CPU fallback); tests may inject row_backend")
            from . import _lib
            self._lib = _lib
            self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            self._params = _lib.make_params(1, self.N, comm_radius=comm_radius, dt=dt, centralized=centralized,
                                            exact=self.exact, knn_obs=self.knn, rows=(self.lo, self.hi),
                                            lanes=1 if self.exact else lanes, tune=tune)
            # Swarms that one CTA can stage use the env-resident kernel; larger ones (and streaming=True) the
            # streaming swarm kernel.  `lanes` is fixed (not derived from the shard size) so that every world size
            # performs bit-identical row arithmetic.
            self.streaming = (self.N > 2048) if streaming is None else bool(streaming)
            if self.streaming and self.exact:
                raise ValueError("fp64-exact is offered by the env-resident path only (N <= ~3000)")
            L = _lib.lib()
            if self.streaming:
                self._fn = L.flock_swarm_step_f64 if self.dtype == torch.float64 else L.flock_swarm_step_f32
                nbytes = int(L.flock_swarm_scratch_bytes(ctypes.byref(self._params)))
                self._scratch = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
                self._scratch_ptr = (self._scratch.data_ptr() + 255) & ~255
                self._scratch_bytes = nbytes
            else:
                self._fn = L.flock_step_f64 if self.dtype == torch.float64 else L.flock_step_f32
        else:
            self.device = torch.device("cpu")
        n, rows, dev, f64 = self.N, self.hi - self.lo, self.device, torch.float64
        self._x = [torch.zeros(n, 4, dtype=f64, device=dev) for _ in range(2)]
        self._u = [torch.zeros(n, 2, dtype=f64, device=dev) for _ in range(2)]
        self._cur = 0
        self.values = torch.zeros(rows, 6, dtype=self.dtype, device=dev)
        self.network = torch.zeros(rows, n, dtype=self.dtype, device=dev) if adjacency_mode == "dense" else None
        self.obs = torch.zeros(rows, 28, dtype=self.dtype, device=dev) if self.knn else None
        self.reward = torch.zeros(1, dtype=self.dtype, device=dev)

    # ---- plumbing ----------------------------------------------------------------------------------------------
    @property
    def x(self) -> torch.Tensor:
        return self._x[self._cur]

    @property
    def u(self) -> torch.Tensor:
        """controller() action for the current state, all N rows (gathered)."""
        return self._u[self._cur]

    def _shifted(self, t, row_elems):
        """Device pointer of a shard-sized buffer shifted back by row_lo rows (the kernel indexes by global row)."""
        if t is None:
            return None
        return ctypes.c_void_p(t.data_ptr() - self.lo * row_elems * t.element_size())

    def _launch(self, x_in, u_in, x_out, u_next):
        if self.row_backend is not None:
            out = self.row_backend(x_in.numpy(), None if u_in is None else u_in.numpy(), self.lo, self.hi,
                                   dict(comm_radius=self.comm_radius, dt=self.dt, centralized=self.centralized, knn=self.knn))
            if x_out is not None:
                x_out.copy_(torch.from_numpy(out["x"]))
            u_next[self.lo:self.hi].copy_(torch.from_numpy(out["u_rows"]))
            self.values.copy_(torch.from_numpy(out["values_rows"]).to(self.dtype))
            if self.network is not None:
                self.network.copy_(torch.from_numpy(out["network_rows"]).to(self.dtype))
            if self.obs is not None:
                self.obs.copy_(torch.from_numpy(out["obs_rows"]).to(self.dtype))
            self.reward.fill_(float(out["reward"]))
            return
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        # (the per-step host path matters at 8 ranks x ~0.4 ms steps: skip the device context manager when this thread's
        #  current device already is the env's)
        import contextlib
        ctx = contextlib.nullcontext() if torch.cuda.current_device() == self.device.index else torch.cuda.device(self.device)
        with ctx:
            if self.streaming:
                if x_out is None:                      # reset(): the streaming path always publishes x'
                    x_out = self._x[self._cur ^ 1]
                self._lib.check(self._fn(ctypes.byref(self._params), p(x_in), p(u_in), p(x_out),
                                         self._shifted(self.values, 6), self._shifted(self.network, self.N),
                                         p(self.reward), p(u_next), self._shifted(self.obs, 28),
                                         ctypes.c_void_p(self._scratch_ptr), ctypes.c_size_t(self._scratch_bytes), stream))
            else:
                self._lib.check(self._fn(ctypes.byref(self._params), p(x_in), p(u_in), p(x_out),
                                         self._shifted(self.values, 6), self._shifted(self.network, self.N),
                                         p(self.reward), p(u_next), self._shifted(self.obs, 28), stream))

    def _gather_actions(self, u_full):
        """In-place all-gather of the rows each rank just computed (equal shards, rank order = row order)."""
        if self.world > 1 and not self._skip_gather:
            mine = u_full[self.lo:self.hi]
            if u_full.is_cuda:      # NCCL gathers in place (send buffer = this rank's slice of the receive buffer)
                dist.all_gather_into_tensor(u_full, mine, group=self.group)
            else:                   # gloo (CPU tests) wants a separate send buffer
                dist.all_gather_into_tensor(u_full, mine.clone(), group=self.group)

    def _observation(self):
        return (self.obs if self.knn else self.values), self.network

    # ---- API -----------------------------------------------------------------------------------------------------
    def reset(self, x0):
        """x0[N,4]: the same full state on every rank."""
        x0 = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0, dtype=torch.float64)
        self.x.copy_(x0.reshape(self.N, 4))
        self._launch(self.x, None, None, self._u[self._cur])
        self._gather_actions(self._u[self._cur])
        return self._observation()

    def capture_graphs(self, warmup: int = 3):
        """Capture one CUDA graph per state parity for the internal-controller step (prep + row kernel + kNN fix-up + the
        NCCL all-gather): `step()` then replays a graph instead of issuing four launches -- the per-launch gaps between the
        small kernels of a step are a tenth of a sharded 65 536-agent step.  Call after reset(); every rank must call it."""
        if self.row_backend is not None:
            raise RuntimeError("graphs need the CUDA path")
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                  # (kernels configured, NCCL communicator up, before any capture)
                self._step_eager(None)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        start, graphs = self._cur, {}
        for parity in (0, 1):
            self._cur = parity
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager(None)
            graphs[parity] = g
        self._cur = start
        self._graphs = graphs

    def step(self, u=None):
        """u[N,2] full action (same on every rank) or None = the gathered controller action."""
        graphs = getattr(self, "_graphs", None)
        if u is None and graphs is not None:
            graphs[self._cur].replay()
            self._cur ^= 1
            return self._observation(), self.reward, False, {}
        return self._step_eager(u)

    def _step_eager(self, u=None):
        u_in = self.u if u is None else torch.as_tensor(u, dtype=torch.float64, device=self.device).reshape(self.N, 2)
        nxt = self._cur ^ 1
        if u_in.data_ptr() == self._u[nxt].data_ptr():
            u_in = u_in.clone()
        self._launch(self.x, u_in, self._x[nxt], self._u[nxt])
        self._cur = nxt
        self._gather_actions(self._u[nxt])
        return self._observation(), self.reward, False, {}
