"""gym_flock_b200 -- gym-flock's flocking step (FlockingRelative-v0 / Flocking-v0) on B200 GPUs.

    from gym_flock_b200 import make, FlockVecEnv
    env = make('FlockingRelative-v0')          # drop-in gym API, NumPy in/out, bit-exact fp64 by default
    vec = FlockVecEnv('FlockingRelative-v0', num_envs=4096, n_agents=100, precision='fp32')   # torch CUDA tensors

The numerics live in libflock_b200.so (hand-written sm_100a CUDA behind the C ABI of include/flock_b200.h).
"""
from .envs import (FlockingEnv, FlockingLeaderEnv, FlockingObstacleEnv, FlockingRelativeEnv, FlockingStochasticEnv,  # noqa: F401
                   FlockingTwoFlocksEnv, TimeLimit, make, register_with_gym)
from .host_env import HostFlockBatch  # noqa: F401

__all__ = ["FlockingEnv", "FlockingRelativeEnv", "FlockingTwoFlocksEnv", "FlockingLeaderEnv", "FlockingStochasticEnv", "FlockingObstacleEnv", "TimeLimit", "make", "register_with_gym", "HostFlockBatch",
           "FlockVecEnv", "SwarmEnv"]


def __getattr__(name):   # torch-dependent classes are imported on first use
    if name == "FlockVecEnv":
        from .vec_env import FlockVecEnv
        return FlockVecEnv
    if name == "SwarmEnv":
        from .swarm import SwarmEnv
        return SwarmEnv
    raise AttributeError(name)


register_with_gym()
