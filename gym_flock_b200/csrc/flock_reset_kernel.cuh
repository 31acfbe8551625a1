// flock_reset_kernel.cuh -- batched reset(): the reference's whole-configuration rejection sampler, one CTA per env.
//
// flocking_relative.py:156-192 draws a complete configuration (positions uniform in a disk of radius sqrt(r_max),
// velocities U(-v_max, v_max) plus a common bias U(-v_bias, v_bias)^2) and rejects it until the closest pair is at
// least 0.1 apart AND every agent has at least 2 neighbours inside the communication radius.  On the host that is
// ~750 draws x an [N,N] distance matrix per env at N = 100; here every env retries on its own CTA.
//
// The draws come from a counter-based generator (Philox4x32-10) keyed by (seed; agent, try, env, stream), so an env's
// result depends only on (seed, env index) -- not on the batch size, the grid or the order CTAs run in -- and the
// velocities of rejected configurations are never generated.  Parity with the reference is STATISTICAL (same law, same
// acceptance test), not bitwise: NumPy's MT19937 stream cannot be reproduced in parallel.
#pragma once
#include "flock_common.cuh"

namespace flock {

struct RArgs {
    double *x_out;            // [E][N][4]
    int *tries_out;           // [E] draws used (> 0) or -max_tries when no admissible configuration was found; nullable
    int E, N;
    double r_max, v_max, v_bias, R2, min_dist;
    unsigned long long seed;
    int max_tries;
};


}  // namespace flock
#ifndef FLOCK_ARGS_ONLY
namespace flock {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniform doubles in [0, 1) with 53 random bits each (the construction of NumPy's random_double)
__device__ __forceinline__ double2 uniform2(const unsigned agent, const unsigned tr, const unsigned env, const unsigned stream,
                                            const unsigned long long seed) {
    unsigned o[4];
    philox4x32_10(agent, tr, env, stream, (unsigned)seed, (unsigned)(seed >> 32), o);
    const double a = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
    const double b = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
    return make_double2(a, b);
}

constexpr int RESET_THREADS = 256;

__global__ void __launch_bounds__(RESET_THREADS) flock_reset_kernel(const RArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *pos = reinterpret_cast<double2 *>(smem);                  // [N] candidate positions
    __shared__ double s_min[RESET_THREADS / 32];
    __shared__ int s_deg[RESET_THREADS / 32];
    __shared__ int s_accept;
    const int env = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N;
    int tr = 0;
    bool ok = false;
    for (; tr < a.max_tries; ++tr) {
        // positions: length = sqrt(U(0, r_max)), angle = pi * U(0, 2)                         (:167-170)
        for (int i = tid; i < N; i += RESET_THREADS) {
            const double2 u = uniform2((unsigned)i, (unsigned)tr, (unsigned)env, 0u, a.seed);
            const double len = sqrt(u.x * a.r_max);
            double sn, cs;
            sincospi(2.0 * u.y, &sn, &cs);
            pos[i] = make_double2(len * cs, len * sn);
        }
        __syncthreads();
        // closest pair and smallest degree                                                    (:177-184)
        double mn = CUDART_INF;
        int dg = 0x7fffffff;
        for (int i = tid; i < N; i += RESET_THREADS) {
            const double2 pi_ = pos[i];
            int d = 0;
            for (int j = 0; j < N; ++j) {
                const double2 pj = pos[j];
                const double dx = pi_.x - pj.x, dy = pi_.y - pj.y;
                const double r2 = (j == i) ? CUDART_INF : dx * dx + dy * dy;
                mn = fmin(mn, r2);
                d += (r2 < a.R2) ? 1 : 0;
            }
            dg = min(dg, d);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, m));
            dg = min(dg, __shfl_xor_sync(0xffffffffu, dg, m));
        }
        if (lane == 0) { s_min[warp] = mn; s_deg[warp] = dg; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < RESET_THREADS / 32; ++w) { mn = fmin(mn, s_min[w]); dg = min(dg, s_deg[w]); }
            s_accept = !(dg < 2 || sqrt(mn) < a.min_dist);                                   // the reference's loop test (:164)
        }
        __syncthreads();
        ok = s_accept != 0;
        if (ok) break;
    }
    // velocities of the configuration that is kept (the last one when none was admissible)   (:172-174)
    const unsigned trk = (unsigned)(ok ? tr : a.max_tries - 1);
    const double2 ub = uniform2(0xffffffffu, trk, (unsigned)env, 1u, a.seed);
    const double bx = (2.0 * ub.x - 1.0) * a.v_bias, by = (2.0 * ub.y - 1.0) * a.v_bias;
    for (int i = tid; i < N; i += RESET_THREADS) {
        const double2 uv = uniform2((unsigned)i, trk, (unsigned)env, 2u, a.seed);
        const double2 p = pos[i];
        st_state(a.x_out + ((size_t)env * N + i) * 4,
                 make_double4(p.x, p.y, (2.0 * uv.x - 1.0) * a.v_max + bx, (2.0 * uv.y - 1.0) * a.v_max + by));
    }
    if (tid == 0 && a.tries_out) a.tries_out[env] = ok ? tr + 1 : -a.max_tries;
}

}  // namespace flock
#endif  // FLOCK_ARGS_ONLY
