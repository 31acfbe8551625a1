"""Drop-in gym API for gym-flock's 'FlockingRelative-v0' and 'Flocking-v0', computed on a B200.

Same names, argument meaning, return layout and error behaviour as the reference classes
(gym_flock/envs/flocking/flocking_relative.py:18-304, flocking.py:5-24, registry gym_flock/__init__.py:53-63);
the numerics of step() / compute_helpers() / controller() / instant_cost() / get_observation() run in the
fused CUDA kernel behind the C ABI (csrc/, include/flock_b200.h).  Default precision 'fp64-exact' returns
the reference's float64 values bit for bit.  There is no CPU fallback for those methods.

Differences, all opt-in or forced by the reference's own behaviour:
  * `reset()` with constructor defaults never terminates in the reference; here the same
    rejection loop runs for at most `reset_max_tries` draws (default 20000) and then raises RuntimeError.
  * `set_state(x0)` is sugar for the reference idiom `env.x = x0; env.compute_helpers()` (also supported).
  * `diff`, `r2`, `x_features` are never materialised by the kernel; they are derived lazily from `x` on
    attribute access (debug/compat only, O(N^2) NumPy temporaries).
"""
from __future__ import annotations

import numpy as np

from .host_env import HostFlockBatch

font = {'family': 'sans-serif', 'weight': 'bold', 'size': 14}


class Box:
    """Minimal stand-in for gym.spaces.Box when gym is not installed (low/high/shape/dtype + sample/contains)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e6)
        hi = np.where(np.isfinite(self.high), self.high, 1e6)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box{self.shape}"


def _make_box(low, high, shape, dtype):
    try:  # use the real thing when the user's stack has it
        from gym import spaces  # type: ignore
        if not getattr(__import__("gym"), "__oracle_stub__", False):
            return spaces.Box(low=low, high=high, shape=shape, dtype=dtype)
    except Exception:
        pass
    return Box(low, high, shape, dtype)


class FlockingRelativeEnv:
    """gym-flock's FlockingRelativeEnv (flocking_relative.py:18) on the GPU."""

    metadata: dict = {}

    def __init__(self, precision: str = "fp64-exact", device: int = 0):
        self.mean_pooling = True      # flocking_relative.py:27
        self.centralized = True       # :28
        self.nx_system = 4            # :31
        self.n_features = 6           # :33
        self.nu = 2                   # :35
        self.n_agents = 100           # :38
        self.comm_radius = 0.9        # :39
        self.dt = 0.01                # :40
        self.v_max = 5.0              # :41
        self.r_max = 1.0              # :42
        self.comm_radius2 = self.comm_radius * self.comm_radius     # :44
        self.vr = 1 / self.comm_radius2 + np.log(self.comm_radius2)  # :45
        self.v_bias = self.v_max      # :46
        self.u = None
        self.mean_vel = None
        self.init_vel = None
        self.max_accel = 1
        self._make_spaces()
        self.fig = None
        self.line1 = None
        self.action_scalar = 10.0     # :64
        self.reset_max_tries = 20000
        self._precision = precision
        self._device = int(device)
        self._eng = None
        self._eng_key = None
        self._x = None
        self._shadow = None           # what the device holds; detects in-place edits of env.x
        self.state_values = None
        self.state_network = None
        self._reward = None
        self.seed()

    # ---- parameters -------------------------------------------------------------------------------------
    def _make_spaces(self):
        self.action_space = _make_box(-self.max_accel, self.max_accel, (2 * self.n_agents,), np.float32)   # :56
        self.observation_space = _make_box(-np.inf, np.inf, (self.n_agents, self.n_features), np.float32)  # :59

    def params_from_cfg(self, args):
        """flocking_relative.py:68-85 (including the compounding r_max side effect, :75)."""
        self.comm_radius = args.getfloat('comm_radius')
        self.comm_radius2 = self.comm_radius * self.comm_radius
        self.vr = 1 / self.comm_radius2 + np.log(self.comm_radius2)
        self.n_agents = args.getint('n_agents')
        self.r_max = self.r_max * np.sqrt(self.n_agents)
        self._make_spaces()
        self.v_max = args.getfloat('v_max')
        self.v_bias = self.v_max
        self.dt = args.getfloat('dt')

    def seed(self, seed=None):
        """flocking_relative.py:87-89: seeds an RNG the flocking envs never draw from (reset uses np.random)."""
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 31))
        self.np_random = np.random.RandomState(int(seed) % (2 ** 32))
        return [seed]

    # ---- engine -----------------------------------------------------------------------------------------
    def _knn(self) -> bool:
        return False

    def _variant(self) -> dict:
        """Step-variant parameters of the C ABI (FlockParams.step_scale / state_scale / u_clip / n_leaders / n_obstacles)."""
        return {}

    def _engine(self) -> HostFlockBatch:
        """The device engine for the env's CURRENT attributes: every parameter a user may change between calls, as the
        reference allows (n_agents, comm_radius(2), dt, action_scalar, centralized and the variant's own knobs), reaches
        the kernel -- a changed n_agents rebuilds the engine, anything else is a parameter update."""
        var = self._variant()
        key = (self.n_agents, self.comm_radius, self.comm_radius2, self.dt, self.action_scalar, self.centralized,
               tuple(sorted(var.items())))
        if self._eng is None or self._eng_key[0] != self.n_agents:
            if self._eng is not None:
                self._eng.close()
            self._eng = HostFlockBatch(1, self.n_agents, comm_radius=self.comm_radius, dt=self.dt,
                                       centralized=bool(self.centralized), knn_obs=self._knn(),
                                       precision=self._precision, device=self._device,
                                       comm_radius2=self.comm_radius2, pinned=False,
                                       action_scalar=float(self.action_scalar), **var)
            self._shadow = None
        elif key != self._eng_key:
            self._eng.set_params(comm_radius=self.comm_radius, comm_radius2=self.comm_radius2, dt=self.dt,
                                 centralized=bool(self.centralized), action_scalar=float(self.action_scalar), **var)
        self._eng_key = key
        return self._eng

    @property
    def x(self):
        return self._x

    @x.setter
    def x(self, value):
        self._x = None if value is None else np.array(value, dtype=np.float64).reshape(self.n_agents, self.nx_system)
        self._shadow = None

    def _sync_device(self, for_step: bool = False):
        """Upload env.x if the user replaced or edited it since the device last saw it.  step() only needs the state (it
        recomputes everything after integrating); everything else gets the helpers recomputed -- except after an in-place
        edit the reference makes itself without recomputing them (`_helpers_stale`, FlockingLeaderEnv.reset)."""
        if self._x is None:
            raise RuntimeError("no state: call reset() or set_state(x0) first")
        if self._shadow is None or not np.array_equal(self._x, self._shadow):
            if for_step and self._shadow is not None:
                self._engine().upload_state(self._x[None])
                self._shadow = self._x.copy()
                self._helpers_stale = False
            elif not getattr(self, "_helpers_stale", False):
                self.compute_helpers()

    def _pull(self, eng):
        self.state_values = eng.values[0].copy()
        net = eng.network[0].copy()
        # mean_pooling=False returns the 0/1 adjacency (:131-134)
        self.state_network = net if self.mean_pooling else (net > 0).astype(net.dtype)
        self._adj_mean = net
        self._reward = eng.reward[0].item()

    # ---- the hot path -------------------------------------------------------------------------------------
    def compute_helpers(self):
        """flocking_relative.py:111-134 on the current env.x (state injection entry point)."""
        eng = self._engine()
        eng.set_state(self._x[None])
        self._shadow = self._x.copy()
        self._helpers_stale = False
        self._pull(eng)

    def set_state(self, x0):
        self.x = x0
        self.compute_helpers()
        return (self.state_values, self.state_network)

    def step(self, u):
        """flocking_relative.py:91-109."""
        assert u.shape == (self.n_agents, self.nu)
        eng = self._engine()
        self._sync_device(for_step=True)
        self.u = u * self.action_scalar                      # :96 (attribute only; the kernel redoes it)
        eng.step(np.asarray(u, dtype=np.float64)[None])
        self._x[...] = eng.x[0]                              # in place, as the reference (:97-105): views such as init_vel stay live
        self._shadow = self._x.copy()
        self._pull(eng)
        return (self.state_values, self.state_network), self.instant_cost(), False, {}

    def controller(self, centralized=None):
        """flocking_relative.py:194-212: Tanner-style controller, actions in [-1, 1]."""
        if centralized is None:
            centralized = self.centralized
        self._sync_device()
        return self._engine().controller(bool(centralized))[0].copy()

    def instant_cost(self):
        """flocking_relative.py:145-147 (computed by the kernel with the state it belongs to)."""
        self._sync_device()
        return self._reward

    def reset(self):
        """flocking_relative.py:156-192: the reference's rejection sampler on the global np.random stream
        (same draws in the same order), bounded by reset_max_tries."""
        n = self.n_agents
        x = np.zeros((n, self.nx_system))
        degree, min_dist, thresh, tries = 0, 0, 0.1, 0
        while degree < 2 or min_dist < thresh:
            if tries >= self.reset_max_tries:
                raise RuntimeError(
                    f"reset(): no admissible configuration in {tries} draws (n_agents={n}, r_max={self.r_max}); "
                    "the reference loops forever here -- use params_from_cfg() or set_state(x0)")
            tries += 1
            length = np.sqrt(np.random.uniform(0, self.r_max, size=(n,)))
            angle = np.pi * np.random.uniform(0, 2, size=(n,))
            x[:, 0] = length * np.cos(angle)
            x[:, 1] = length * np.sin(angle)
            bias = np.random.uniform(low=-self.v_bias, high=self.v_bias, size=(2,))
            x[:, 2] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[0]
            x[:, 3] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[1]
            p = x[:, 0:2]
            d = p[:, None, :] - p[None, :, :]
            d2 = np.sum(np.square(d), axis=2)
            np.fill_diagonal(d2, np.inf)
            min_dist = np.sqrt(np.min(d2))
            degree = np.min(np.sum((d2 < self.comm_radius2).astype(int), axis=1))
        self.mean_vel = np.mean(x[:, 2:4], axis=0)
        self._x = x
        self._shadow = None
        self.init_vel = self._x[:, 2:4]
        self.compute_helpers()
        return (self.state_values, self.state_network)

    # ---- compat attributes the kernel never materialises (derived from x on access) -------------------------------
    @property
    def diff(self):
        x = self._x
        return x.reshape((self.n_agents, 1, self.nx_system)) - x.reshape((1, self.n_agents, self.nx_system))

    @property
    def r2(self):
        d = self.diff
        r2 = d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]
        np.fill_diagonal(r2, np.inf)
        return r2

    @property
    def adj_mat(self):
        self._sync_device()
        return (self._adj_mean > 0).astype(float)

    @property
    def adj_mat_mean(self):
        self._sync_device()
        return self._adj_mean

    @property
    def x_features(self):
        d, r2 = self.diff, self.r2
        return np.dstack((d[:, :, 2], d[:, :, 0] / (r2 * r2), d[:, :, 0] / r2,
                          d[:, :, 3], d[:, :, 1] / (r2 * r2), d[:, :, 1] / r2))

    def get_stats(self):
        """flocking_relative.py:136-143, computed by the kernel from the adjacency bits of the current state
        (FlockAux.vel_diff / min_dist; bit-exact in the default 'fp64-exact' precision)."""
        self._sync_device()
        st = self._engine().get_stats()
        return {'vel_diffs': st['vel_diffs'][0].astype(np.float64), 'min_dists': st['min_dists'][0].astype(np.float64)}

    def potential(self, r2):
        """flocking_relative.py:228-232 (unused by the env itself)."""
        p = np.reciprocal(r2) + np.log(r2)
        p[r2 > self.comm_radius2] = self.vr
        np.fill_diagonal(p, 0)
        return np.sum(np.sum(p))

    def render(self, mode='human'):
        """flocking_relative.py:234-257; needs matplotlib (not a dependency of the compute path)."""
        import matplotlib.pyplot as plt
        if self.fig is None:
            plt.ion()
            fig = plt.figure()
            self.ax = fig.add_subplot(111)
            line1, = self.ax.plot(self._x[:, 0], self._x[:, 1], 'bo')
            self.ax.plot([0], [0], 'kx')
            plt.ylim(-1.0 * self.r_max, 1.0 * self.r_max)
            plt.xlim(-1.0 * self.r_max, 1.0 * self.r_max)
            plt.title('GNN Controller')
            self.fig, self.line1 = fig, line1
        self.line1.set_xdata(self._x[:, 0])
        self.line1.set_ydata(self._x[:, 1])
        self.fig.canvas.draw()
        self.fig.canvas.flush_events()

    def close(self):
        if self._eng is not None:
            self._eng.close()
            self._eng = None


class FlockingEnv(FlockingRelativeEnv):
    """gym-flock's FlockingEnv / 'Flocking-v0' (flocking.py:5-24): observation = state differences to the
    7 nearest neighbours, computed in the same kernel pass (top-7 kept in registers)."""

    def __init__(self, precision: str = "fp64-exact", device: int = 0):
        super().__init__(precision=precision, device=device)
        self.n_neighbors = 7                                   # flocking.py:9
        self.n_f = self.nx_system * self.n_neighbors           # flocking.py:10
        self._obs = None

    def _knn(self) -> bool:
        return True

    def _pull(self, eng):
        super()._pull(eng)
        self._obs = eng.obs[0].copy()

    def get_observation(self):
        """flocking.py:20-24.  The reference indexes argsort(...)[:, 6] and raises IndexError when N < 7."""
        if self.n_agents < self.n_neighbors:
            raise IndexError(f"index {self.n_neighbors - 1} is out of bounds for axis 1 with size {self.n_agents}")
        self._sync_device()
        return self._obs

    def step(self, u):
        super().step(u)
        return (self.get_observation(), self.state_network), self.instant_cost(), False, {}

    def reset(self):
        super().reset()
        return self.get_observation(), self.state_network

    def set_state(self, x0):
        super().set_state(x0)
        return self.get_observation(), self.state_network


def _grid(n, side=5):
    """Initial lattice of the two-flocks variant (gym_flock/envs/flocking/utils.py:23-31): 0.8 x (col - side/2, row - rows/2)."""
    rows = int(n / side)
    xs, ys = np.meshgrid(np.arange(0, side) - side / 2.0, np.arange(0, rows) - rows / 2.0)
    return 0.8 * np.hstack((xs.reshape((n, 1)), ys.reshape((n, 1))))


class FlockingTwoFlocksEnv(FlockingRelativeEnv):
    """gym-flock's 'FlockingTwoFlocks-v0' (flocking_twoflocks.py:8-26): same dynamics and kernel, lattice reset with
    velocities pointing at the origin plus a common random bias (one draw of size 2 from the global np.random)."""

    def reset(self):
        n = self.n_agents
        x = np.zeros((n, self.nx_system))
        bias = np.random.uniform(low=-self.v_bias / 2.0, high=self.v_bias / 2.0, size=(2,))
        grids = _grid(n, side=int(n / 10))
        x[:, 0:2] = grids
        x[:, 2:4] = -grids
        x[:, 2] = x[:, 2] + bias[0]
        x[:, 3] = x[:, 3] + bias[1]
        self.mean_vel = np.mean(x[:, 2:4], axis=0)
        self._x = x
        self._shadow = None
        self.init_vel = self._x[:, 2:4]
        self.compute_helpers()
        return (self.state_values, self.state_network)


class FlockingLeaderEnv(FlockingRelativeEnv):
    """gym-flock's 'FlockingLeader-v0' (flocking_leader.py:5-57): the first n_leaders agents are not actuated and the
    action is applied unscaled (self.u = u, :24); same kernel with FlockParams.n_leaders / step_scale = 1."""

    def __init__(self, precision: str = "fp64-exact", device: int = 0):
        super().__init__(precision=precision, device=device)
        self.n_leaders = 2                                           # :11
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_leaders] = 0
        self.quiver = None
        self.half_leaders = int(self.n_leaders / 2.0)

    def params_from_cfg(self, args):
        super().params_from_cfg(args)
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_leaders] = 0

    def _variant(self) -> dict:
        return dict(step_scale=1.0, n_leaders=int(self.n_leaders))

    def step(self, u):
        """flocking_leader.py:21-34."""
        assert u.shape == (self.n_agents, self.nu)
        eng = self._engine()
        self._sync_device(for_step=True)
        self.u = u
        eng.step(np.asarray(u, dtype=np.float64)[None])
        self._x[...] = eng.x[0]                              # in place, as the reference (:97-105): views such as init_vel stay live
        self._shadow = self._x.copy()
        self._pull(eng)
        return (self.state_values, self.state_network), self.instant_cost(), False, {}

    def render(self, mode='human'):
        """flocking_leader.py:42-57: the base scatter plus one red arrow per leader (position, velocity)."""
        super().render(mode)
        lead = self._x[:self.n_leaders]
        if self.quiver is None:
            self.quiver = self.ax.quiver(lead[:, 0], lead[:, 1], lead[:, 2], lead[:, 3], color='r')
        else:
            self.quiver.set_offsets(lead[:, 0:2])
            self.quiver.set_UVC(lead[:, 2], lead[:, 3])
        self.fig.canvas.draw()
        self.fig.canvas.flush_events()

    def _set_leader_velocity(self, vel):
        """The in-place edit of flocking_leader.py:38-39.  The reference does NOT recompute its helpers afterwards, so the
        features it returns and the first controller() call still describe the state before the edit; step() then
        integrates the edited state.  Reproduced: the device keeps the old helpers until the next step()."""
        self._x[0:self.n_leaders, 2:4] = np.ones((self.n_leaders, 2)) * vel
        self._helpers_stale = True

    def reset(self):
        """flocking_leader.py:36-40: the leaders get one common random velocity AFTER the helpers were computed."""
        out = super().reset()
        self._set_leader_velocity(np.random.uniform(low=-self.v_max, high=self.v_max, size=(1, 1)))
        return out


class FlockingObstacleEnv(FlockingRelativeEnv):
    """gym-flock's 'FlockingObstacle-v0' (flocking_obstacle.py:14-104): the first n_obstacles agents are static
    obstacles -- not actuated (:40-46), and every velocity difference with an obstacle row or column is zero (:81-82);
    the action is applied unscaled (:37).  Same kernel with FlockParams.n_obstacles / step_scale = 1."""

    def __init__(self, precision: str = "fp64-exact", device: int = 0):
        super().__init__(precision=precision, device=device)
        self.n_obstacles = 4                                         # :19
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_obstacles] = 0
        self.r_max = 3.0                                             # :23
        self.line1 = None
        self.line2 = None

    def params_from_cfg(self, args):
        super().params_from_cfg(args)
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_obstacles] = 0

    def _variant(self) -> dict:
        return dict(step_scale=1.0, n_obstacles=int(self.n_obstacles))

    def render(self, mode='human'):
        """flocking_obstacle.py:106-118: the base scatter with the obstacles drawn over it in red."""
        super().render(mode)
        obst = self._x[:self.n_obstacles]
        if self.line2 is None:
            self.line2, = self.ax.plot(obst[:, 0], obst[:, 1], 'ro')
        self.line2.set_xdata(obst[:, 0])
        self.line2.set_ydata(obst[:, 1])
        self.fig.canvas.draw()
        self.fig.canvas.flush_events()

    def _engine(self):
        # the engine is rebuilt when the number of obstacles changes (it is a launch parameter)
        if self._eng is not None and getattr(self, "_eng_nobs", None) != int(self.n_obstacles):
            self._eng.close()
            self._eng = None
        self._eng_nobs = int(self.n_obstacles)
        return super()._engine()

    def step(self, u):
        """flocking_obstacle.py:33-50."""
        assert u.shape == (self.n_agents, self.nu)
        self._sync_device()
        eng = self._engine()
        self.u = u
        eng.step(np.asarray(u, dtype=np.float64)[None])
        self._x[...] = eng.x[0]                              # in place, as the reference (:97-105): views such as init_vel stay live
        self._shadow = self._x.copy()
        self._pull(eng)
        return (self.state_values, self.state_network), self.instant_cost(), False, {}

    def reset(self):
        """flocking_obstacle.py:57-72: deterministic lattice moving down at 7 m/s towards a 2x2 block of obstacles."""
        n, no = self.n_agents, self.n_obstacles
        x = np.zeros((n, self.nx_system))
        x[:, 0:2] = _grid(n)
        x[:, 2:4] = [0, -7.0]
        x[0:no, 0:2] = _grid(no, side=2) * 0.5
        x[0:no, 1] -= 10.0
        x[0:no, 2:4] = 0
        self.mean_vel = np.mean(x[no:, 2:4], axis=0)
        self._x = x
        self._shadow = None
        self.init_vel = self._x[no:, 2:4]
        self.compute_helpers()
        return (self.state_values, self.state_network)


class FlockingStochasticEnv(FlockingRelativeEnv):
    """gym-flock's 'FlockingStochastic-v0' (flocking_stoch.py:5-45): actions clipped to +-0.5 and scaled by 6, the state is
    scaled by 6 around the Euler step, dt ~ N(0.12, 0.018) drawn per step from the global np.random stream."""

    def __init__(self, precision: str = "fp64-exact", device: int = 0):
        super().__init__(precision=precision, device=device)
        self.dt_mean = 0.12
        self.dt_sigma = 0.018
        self.max_accel = 0.5
        self.scale = 6.0

    def _variant(self) -> dict:
        return dict(step_scale=float(self.scale), state_scale=float(self.scale), u_clip=float(self.max_accel))

    def step(self, u):
        """flocking_stoch.py:14-35."""
        assert u.shape == (self.n_agents, self.nu)
        self._sync_device()
        uc = np.clip(u, a_min=-self.max_accel, a_max=self.max_accel)
        self.u = uc * self.scale
        self.dt = np.random.normal(self.dt_mean, self.dt_sigma)      # one draw per step, as in the reference (:20)
        eng = self._engine()                                          # picks up the new dt
        eng.step(np.asarray(u, dtype=np.float64)[None])               # the kernel clips and scales (bit-identical)
        self._x = eng.x[0].copy()
        self._shadow = self._x.copy()
        self._pull(eng)
        return (self.state_values, self.state_network), self.instant_cost(), False, {}

    # controller(): the kernel clips its output to +-u_clip (flocking_stoch.py:37-45)


class TimeLimit:
    """What gym.make() wraps the env in (gym 0.11-era API): counts steps, forces done=True once
    max_episode_steps is reached, forwards everything else; the raw env is `.env`."""

    def __init__(self, env, max_episode_steps: int):
        self.env = env
        self._max_episode_steps = int(max_episode_steps)
        self._elapsed_steps = 0

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info = dict(info)
            info['TimeLimit.truncated'] = not done
            done = True
        return obs, reward, done, info

    @property
    def unwrapped(self):
        return self.env

    def __getattr__(self, name):
        return getattr(self.env, name)


REGISTRY = {   # gym_flock/__init__.py:53-63, 90-94
    'Flocking-v0': (FlockingEnv, 1000),
    'FlockingRelative-v0': (FlockingRelativeEnv, 1000),
    'FlockingTwoFlocks-v0': (FlockingTwoFlocksEnv, 500),   # SURVEY.md 8f "next #1": same kernel, different reset
    'FlockingLeader-v0': (FlockingLeaderEnv, 200),         # gym_flock/__init__.py:65-69
    'FlockingStochastic-v0': (FlockingStochasticEnv, 500), # gym_flock/__init__.py:84-88
    'FlockingObstacle-v0': (FlockingObstacleEnv, 200),     # gym_flock/__init__.py:72-76
}


def make(env_id: str, **kwargs):
    """gym.make() for the two hot-path ids (works without gym installed)."""
    if env_id not in REGISTRY:
        raise KeyError(f"{env_id!r} is not provided by gym_flock_b200 (ids: {sorted(REGISTRY)})")
    cls, steps = REGISTRY[env_id]
    return TimeLimit(cls(**kwargs), steps)


def register_with_gym() -> list:
    """Register the ids with gym / gymnasium when the user's stack has one of them; returns what was done."""
    done = []
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname)
            if getattr(mod, "__oracle_stub__", False):
                continue
            from importlib import import_module
            reg = import_module(modname + ".envs.registration")
            for env_id, (cls, steps) in REGISTRY.items():
                try:
                    reg.register(id=env_id, entry_point=f"gym_flock_b200.envs:{cls.__name__}", max_episode_steps=steps)
                    done.append((modname, env_id))
                except Exception:
                    pass
        except Exception:
            continue
    return done
