// flock_common.cuh -- shared device helpers for the flocking step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace flock {

constexpr int NXS = 4;   // state per agent  (flocking_relative.py:31)
constexpr int NFE = 6;   // features / agent (flocking_relative.py:33)
constexpr int KNN_K = 7; // flocking.py:9

// ---- per-type traits ---------------------------------------------------------------------------
template <typename T> struct Tr;
template <> struct Tr<float> {
    using V4 = float4;
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ float tiny() { return 1e-12f; }   // 1/tiny^2 stays finite
    static __device__ __forceinline__ float rcp(float x) {
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));           // one MUFU.RCP, <= 1 ulp
        return y;
    }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
};
template <> struct Tr<double> {
    using V4 = double4;
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    static __device__ __forceinline__ double tiny() { return 1e-150; }
    static __device__ __forceinline__ double rcp(double x) {
        // MUFU.RCP64H seed (~20 bits) + two Newton steps: full double accuracy for normal x, 4 DFMA.
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
        e = fma(-x, y, 1.0);
        y = fma(y, e, y);
        return y;
    }
    static __device__ __forceinline__ double rcp1(double x) {   // seed (2^-23) + ONE Newton step: ~2^-46
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double e = fma(-x, y, 1.0);
        return fma(y, e, y);
    }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
};

template <typename T> __device__ __forceinline__ T shfl_xor(T v, int m);
template <> __device__ __forceinline__ float shfl_xor<float>(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ double shfl_xor<double>(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ int shfl_xor<int>(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// programmatic dependent launch: every kernel waits here before its first global access, so a launch may be scheduled
// while its predecessor in the stream drains (no-ops when the launch did not ask for it)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- exact-order Euler step, flocking_relative.py:96-105 ----------------------------------------
// ((p + (v*dt)) + (((a*dt)*dt)*0.5)),  v + (a*dt),  a = u*action_scalar; every op rounded separately.
__device__ __forceinline__ double4 integrate_exact(double4 x, double2 u, double dt, double ascale) {
    const double ax = __dmul_rn(u.x, ascale), ay = __dmul_rn(u.y, ascale);
    const double axdt = __dmul_rn(ax, dt), aydt = __dmul_rn(ay, dt);
    double4 y;
    y.x = __dadd_rn(__dadd_rn(x.x, __dmul_rn(x.z, dt)), __dmul_rn(__dmul_rn(axdt, dt), 0.5));
    y.y = __dadd_rn(__dadd_rn(x.y, __dmul_rn(x.w, dt)), __dmul_rn(__dmul_rn(aydt, dt), 0.5));
    y.z = __dadd_rn(x.z, axdt);
    y.w = __dadd_rn(x.w, aydt);
    return y;
}

// The step of every flocking variant, in the reference's operation order:
//   FlockingRelative/Flocking : a = u*10                                    flocking_relative.py:96-105
//   FlockingLeader            : a = u, leaders (j < n_leaders) get a*0       flocking_leader.py:21-31
//   FlockingStochastic        : u = clip(u, +-c); a = u*6; x = x*6; step; x = x/6      flocking_stoch.py:14-33
//   FlockingObstacle          : a = u, obstacles (j < n_obstacles) are not actuated; velocity differences with an
//                               obstacle row or column are zero                flocking_obstacle.py:33-50,78-82
struct StepVariant { double step_scale, state_scale, u_clip; int n_leaders, n_obstacles; };
__device__ __forceinline__ double4 integrate_variant(double4 x, double2 u, const int j, const double dt, const StepVariant v) {
    if (v.u_clip > 0.0) { u.x = fmin(fmax(u.x, -v.u_clip), v.u_clip); u.y = fmin(fmax(u.y, -v.u_clip), v.u_clip); }
    if (j < v.n_leaders || j < v.n_obstacles) { u.x = 0.0; u.y = 0.0; }          // (u*dt*dt*0.5)*0 = 0: the position/velocity keep their drift only
    const bool scaled = (v.state_scale != 1.0);
    if (scaled) {
        x.x = __dmul_rn(x.x, v.state_scale); x.y = __dmul_rn(x.y, v.state_scale);
        x.z = __dmul_rn(x.z, v.state_scale); x.w = __dmul_rn(x.w, v.state_scale);
    }
    double4 y = integrate_exact(x, u, dt, v.step_scale);
    if (scaled) {
        y.x = __ddiv_rn(y.x, v.state_scale); y.y = __ddiv_rn(y.y, v.state_scale);
        y.z = __ddiv_rn(y.z, v.state_scale); y.w = __ddiv_rn(y.w, v.state_scale);
    }
    return y;
}

__device__ __forceinline__ double4 ld_state(const double *p) {   // p 32-byte aligned (row of x)
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_state(double *p, double4 v) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v.z, v.w);
}

// streaming (evict-first) stores for the write-once dense adjacency
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float *p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void st_stream2(double *p, double a, double b) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
}

}  // namespace flock
