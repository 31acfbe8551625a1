"""ctypes binding of libflock_b200.so (the C ABI declared in include/flock_b200.h).

The product path has NO CPU fallback: if the library is missing or no CUDA device is present,
loading / creating a context raises loudly.  ``build()`` compiles the library in-tree with nvcc for
sm_100a (cross-compiles on a box without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("FLOCK_LIB") or os.path.join(_PKG, "libflock_b200.so")
# translation units (compiled in parallel, then linked) and the headers they depend on
_UNITS = ("flock_abi.cu", "launch_env_f32.cu", "launch_env_f64.cu", "launch_group.cu", "launch_swarm.cu", "launch_graph.cu",
          "nvtx_ranges.cu")
_HEADERS = ("flock_host.hpp", "flock_common.cuh", "flock_env_kernel.cuh", "flock_group_kernel.cuh", "flock_swarm_kernel.cuh",
            "flock_reset_kernel.cuh", "flock_graph_kernels.cuh")
_SOURCES = [os.path.join(_PKG, "csrc", f) for f in _UNITS + _HEADERS] + [os.path.join(_ROOT, "include", "flock_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]

# flags / error codes (include/flock_b200.h)
FLOCK_CENTRALIZED = 0x1
FLOCK_EXACT_ORDER = 0x2
FLOCK_KNN_OBS = 0x4
ABI_VERSION = 4


class FlockParams(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_agents", ctypes.c_int32),
                ("comm_radius", ctypes.c_double), ("comm_radius2", ctypes.c_double),
                ("dt", ctypes.c_double), ("action_scalar", ctypes.c_double),
                ("flags", ctypes.c_int32), ("knn", ctypes.c_int32),
                ("row_lo", ctypes.c_int32), ("row_hi", ctypes.c_int32),
                ("lanes", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("step_scale", ctypes.c_double), ("state_scale", ctypes.c_double), ("u_clip", ctypes.c_double),
                ("n_leaders", ctypes.c_int32), ("n_obstacles", ctypes.c_int32)]


class FlockAux(ctypes.Structure):
    """Optional extra outputs of a step (include/flock_b200.h); device pointers, all nullable."""
    _fields_ = [("nonfinite", ctypes.c_void_p), ("adj_bits", ctypes.c_void_p), ("degree", ctypes.c_void_p),
                ("min_dist", ctypes.c_void_p), ("vel_diff", ctypes.c_void_p), ("knn_hint", ctypes.c_void_p)]


class FlockError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libflock_b200 error {code}: {msg}")
        self.code = code


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libflock_b200.so")


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _SOURCES if os.path.isfile(s))


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> gym_flock_b200/libflock_b200.so (in-tree).
    One object per translation unit, compiled concurrently, then one link."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(_PKG, "csrc", "_obj" + ("" if out == LIB_PATH else "_" + os.path.basename(out)))
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def compile_unit(unit):
        obj = os.path.join(objdir, unit[:-3] + ".o")
        subprocess.run([nvcc] + flags + ["-c", "-o", obj, os.path.join(_PKG, "csrc", unit)], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(_UNITS)) as pool:
        objs = list(pool.map(compile_unit, _UNITS))
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-ldl"], check=True)   # (-ldl: NVTX v3)
    return out


_lib = None
_VP = ctypes.c_void_p
_PP = ctypes.POINTER(FlockParams)


def lib() -> ctypes.CDLL:
    """Load the library (never builds implicitly on a GPU box: a missing .so is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the flocking step)")
    L = ctypes.CDLL(LIB_PATH)
    L.flock_abi_version.restype = ctypes.c_int
    L.flock_last_error.restype = ctypes.c_char_p
    L.flock_device_count.restype = ctypes.c_int
    L.flock_host_alloc.restype = _VP
    L.flock_host_alloc.argtypes = [ctypes.c_size_t]
    L.flock_host_free.argtypes = [_VP]
    step_args = [_PP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]
    for name in ("flock_step_f64", "flock_step_f32"):
        getattr(L, name).argtypes = step_args
        getattr(L, name).restype = ctypes.c_int
    for name in ("flock_step_aux_f64", "flock_step_aux_f32"):
        getattr(L, name).argtypes = step_args[:-1] + [ctypes.POINTER(FlockAux), _VP]
        getattr(L, name).restype = ctypes.c_int
    L.flock_adj_words.argtypes = [ctypes.c_int32]
    L.flock_adj_words.restype = ctypes.c_int
    roll_args = [_PP, ctypes.c_int32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                 ctypes.POINTER(ctypes.c_int32), _VP]
    for name in ("flock_rollout_f64", "flock_rollout_f32"):
        getattr(L, name).argtypes = roll_args
        getattr(L, name).restype = ctypes.c_int
    L.flock_swarm_scratch_bytes.argtypes = [_PP]
    L.flock_swarm_scratch_bytes.restype = ctypes.c_size_t
    for name in ("flock_swarm_step_f64", "flock_swarm_step_f32"):
        getattr(L, name).argtypes = step_args[:-1] + [_VP, ctypes.c_size_t, _VP]
        getattr(L, name).restype = ctypes.c_int
    for name in ("flock_aggregate_f64", "flock_aggregate_f32"):
        getattr(L, name).argtypes = [_PP, _VP, ctypes.c_int32, _VP, _VP, ctypes.c_int32, ctypes.c_int32, _VP, _VP]
        getattr(L, name).restype = ctypes.c_int
    for name in ("flock_repulsion_f64", "flock_repulsion_f32"):
        getattr(L, name).argtypes = [_PP, _VP, ctypes.c_int32, _VP, ctypes.c_double, _VP, _VP]
        getattr(L, name).restype = ctypes.c_int
    L.flock_reset_sample.argtypes = [_PP, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                     ctypes.c_int32, _VP, _VP, _VP]
    L.flock_reset_sample.restype = ctypes.c_int
    L.flock_ctx_create.argtypes = [ctypes.POINTER(_VP), _PP, ctypes.c_int32, ctypes.c_int32]
    L.flock_ctx_create.restype = ctypes.c_int
    L.flock_ctx_destroy.argtypes = [_VP]
    L.flock_ctx_destroy.restype = None
    L.flock_ctx_set_params.argtypes = [_VP, _PP]
    L.flock_ctx_set_state.argtypes = [_VP, _VP, _VP, _VP, _VP, _VP]
    L.flock_ctx_step.argtypes = [_VP, _VP, _VP, _VP, _VP, _VP]
    L.flock_ctx_controller.argtypes = [_VP, ctypes.c_int32, _VP]
    L.flock_ctx_get_state.argtypes = [_VP, _VP]
    L.flock_ctx_upload_state.argtypes = [_VP, _VP]
    L.flock_ctx_get_stats.argtypes = [_VP, _VP, _VP]
    for name in ("flock_ctx_set_params", "flock_ctx_set_state", "flock_ctx_step", "flock_ctx_controller",
                 "flock_ctx_get_state", "flock_ctx_upload_state", "flock_ctx_get_stats"):
        getattr(L, name).restype = ctypes.c_int
    if L.flock_abi_version() != ABI_VERSION:
        raise ImportError(f"libflock_b200.so ABI {L.flock_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise FlockError(rc, (lib().flock_last_error() or b"").decode("utf-8", "replace"))


def make_params(n_envs: int, n_agents: int, comm_radius: float = 0.9, dt: float = 0.01,
                action_scalar: float = 10.0, centralized: bool = True, exact: bool = False,
                knn_obs: bool = False, knn: int = 7, rows: tuple | None = None, lanes: int = 0,
                comm_radius2: float | None = None, step_scale: float = 0.0, state_scale: float = 0.0,
                u_clip: float = 0.0, n_leaders: int = 0, n_obstacles: int = 0, tune: int = 0) -> FlockParams:
    """`tune` is the diagnostics word of tools/ and bench.py A/B runs (FlockParams.reserved); bindings leave it 0."""
    flags = (FLOCK_CENTRALIZED if centralized else 0) | (FLOCK_EXACT_ORDER if exact else 0) | \
            (FLOCK_KNN_OBS if knn_obs else 0)
    r2 = comm_radius * comm_radius if comm_radius2 is None else comm_radius2   # flocking_relative.py:44
    lo, hi = rows if rows is not None else (0, 0)
    return FlockParams(int(n_envs), int(n_agents), float(comm_radius), float(r2), float(dt),
                       float(action_scalar), flags, int(knn), int(lo), int(hi), int(lanes),
                       int(tune), float(step_scale), float(state_scale), float(u_clip),
                       int(n_leaders), int(n_obstacles))


def require_device() -> int:
    n = lib().flock_device_count()
    if n < 1:
        raise RuntimeError("gym_flock_b200 needs a CUDA device (B200, sm_100a); none is visible and there "
                           "is no CPU fallback")
    return n
