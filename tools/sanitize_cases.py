"""The smallest case of every kernel family, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_cases.py`
(one tool per gpurun call; logs go to profiles/).  Results are also checked against the oracle so a silent fault shows."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_flock_b200 import FlockVecEnv, SwarmEnv, _lib
from gym_flock_b200.synthetic import initial_state, initial_states
from oracle import c_oracle

x0 = initial_states(6, 50, seed=3, min_dist=0.1)
ref = c_oracle.rollout(x0, 2, knn=7)
for precision, lanes, tol in (("fp64-exact", 0, 0.0), ("fp64", 4, 1e-10), ("fp32", 2, 1e-4), ("fp64", 1, 1e-10), ("fp32", 1, 1e-4)):
    env = FlockVecEnv("Flocking-v0", num_envs=6, n_agents=50, precision=precision, lanes=lanes)   # env kernel / group kernel
    env.reset(x0)
    if lanes == 1:
        env.rollout(2)                                # fused persistent launch
    else:
        for _ in range(2):
            env.step(env.controller())
    torch.cuda.synchronize()
    err = np.max(np.abs(env.x.cpu().numpy() - ref["x"])) / np.max(np.abs(ref["x"]))
    assert err <= tol, (precision, lanes, err)
    st = env.get_stats()
    print("env family", precision, lanes, "state err %.2e" % err, flush=True)
bits = FlockVecEnv("FlockingRelative-v0", num_envs=6, n_agents=50, precision="fp32", adjacency_mode="bitmask")
bits.reset(x0)
agg = bits.aggregate(3)                                # aggregation kernel
torch.cuda.synchronize()
print("aggregate", tuple(agg.shape), flush=True)
n = 2600
xs = initial_state(n, seed=5, law="density")
for precision in ("fp64", "fp32"):
    sw = SwarmEnv("Flocking-v0", n_agents=n, precision=precision, streaming=True, distributed=False)   # prep + row kernel + kNN fix-up
    sw.reset(xs)
    sw.step()
    torch.cuda.synchronize()
    x1 = c_oracle.integrate(xs, c_oracle.controller(xs))
    assert np.max(np.abs(sw.x.cpu().numpy() - x1)) <= 1e-9
    print("swarm family", precision, flush=True)
rv = FlockVecEnv("FlockingRelative-v0", num_envs=4, n_agents=20, precision="fp64")
obs, tries = rv.reset_random(seed=1, max_tries=20000)   # reset kernel
torch.cuda.synchronize()
print("reset kernel tries", tries.cpu().numpy(), flush=True)
L = _lib.lib()
p = _lib.make_params(3, 30)
pos = torch.rand(3, 30, 3, dtype=torch.float64, device="cuda") * 4 - 2
w = torch.ones(30, dtype=torch.float64, device="cuda")
out = torch.zeros(3, 30, 2, dtype=torch.float64, device="cuda")
_lib.check(L.flock_repulsion_f64(ctypes.byref(p), ctypes.c_void_p(pos.data_ptr()), 3, ctypes.c_void_p(w.data_ptr()), 2.0,
                                 ctypes.c_void_p(out.data_ptr()), None))
torch.cuda.synchronize()
print("sanitize cases OK", flush=True)
