// flock_env_kernel.cuh -- the env-resident fused flocking step (sm_100a).
//
// One CTA owns either G whole environments (small N) or one row tile of one environment (large N);
// either way the whole state x[N,4] of its environment(s) is staged in shared memory, so each ordered
// pair (i,j) is touched exactly once per step, in registers.  One launch does, per env:
//
//   prologue : x' = integrate(x_in, u_in)                         flocking_relative.py:96-105
//              reward = -(var vx' + var vy')                      flocking_relative.py:145-147
//              (fp32 mode) centre x' on the env centroid / mean velocity in fp64, convert to fp32
//   pair loop: r2, adjacency bit, 6 masked feature sums           flocking_relative.py:113-129
//              potential-gradient + alignment sums (controller)   flocking_relative.py:194-226
//              running 7 nearest neighbours (Flocking-v0)         flocking.py:20-24
//   epilogue : values, u_next, obs; then the dense adj_mean rows are expanded from the bit masks
//              held in shared memory and written with coalesced, vectorised streaming stores.
//
// LANES threads cooperate on one row (j interleaved across them, warp-shuffle reduction) so small
// batches still fill 148 SMs.  EXACT (fp64, LANES=1) reproduces NumPy's arithmetic bit for bit:
// sequential j, separately rounded mul/add (__d*_rn intrinsics are never contracted), IEEE divide.
#pragma once
#include <type_traits>
#include "flock_common.cuh"

namespace flock {

struct KArgs {
    const double *x_in;
    const double *u_in;
    double *x_out;
    double *u_next;
    void *values;
    void *network;
    void *reward;
    void *obs;
    int E, N;
    int G;        // envs per CTA
    int RT;       // rows per CTA
    int tiles;    // CTAs per env group
    int row_lo, row_hi;
    int centralized;
    int W;        // 32-bit mask words per (row, lane)
    double R, R2, dt, ascale;
    double Rm;    // gradient mask threshold for the fast paths: pair contributes iff r2 <= Rm, where
    float Rm_f;   //   centralized: Rm = R; decentralized: r2 <= R && r2 < R2 folded into one compare
    unsigned off_xs, off_stage, off_masks, off_invdeg, off_stats;
};

template <typename T> struct TopK {
    T d[KNN_K];
    int id[KNN_K];
    __device__ __forceinline__ void init(int self) {
#pragma unroll
        for (int m = 0; m < KNN_K; ++m) { d[m] = Tr<T>::inf(); id[m] = self; }
    }
    // strict < everywhere: among equal r2 the lower index (seen first) stays in front
    __device__ __forceinline__ void insert(T r2, int j) {
        if (r2 < d[KNN_K - 1]) {
            d[KNN_K - 1] = r2; id[KNN_K - 1] = j;
#pragma unroll
            for (int p = KNN_K - 1; p > 0; --p) {
                if (d[p] < d[p - 1]) {
                    const T td = d[p]; d[p] = d[p - 1]; d[p - 1] = td;
                    const int ti = id[p]; id[p] = id[p - 1]; id[p - 1] = ti;
                }
            }
        }
    }
    __device__ __forceinline__ void pop_front(int self) {
#pragma unroll
        for (int p = 0; p < KNN_K - 1; ++p) { d[p] = d[p + 1]; id[p] = id[p + 1]; }
        d[KNN_K - 1] = Tr<T>::inf(); id[KNN_K - 1] = self;
    }
};

// ---- accumulators ------------------------------------------------------------------------------
template <typename T> struct Acc {
    T f0, f1, f2, f3, f4, f5;   // feature sums (f0/f3: see TRICK below)
    T gx, gy;                   // sum dx*c, dy*c with c = 1/r2 - 1/r4 (x2 at the end)
    T s0, s1;                   // unmasked alignment sums (non-TRICK fast path; EXACT: s[0..1])
};

// EXACT pair: Appendix B of SURVEY.md, one rounding per operation, NumPy's order.
__device__ __forceinline__ bool pair_exact(const double4 xi, const double4 xj, const bool diag, const double R,
                                           const double R2, const bool cen, Acc<double> &A, double &r2o) {
    const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
    double dvx = __dsub_rn(xi.z, xj.z), dvy = __dsub_rn(xi.w, xj.w);
    double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (diag) r2 = CUDART_INF;
    r2o = r2;
    const bool a = r2 < R2;
    const double r4 = __dmul_rn(r2, r2);
    const double q4x = __ddiv_rn(dx, r4), q2x = __ddiv_rn(dx, r2);
    const double q4y = __ddiv_rn(dy, r4), q2y = __ddiv_rn(dy, r2);
    if (a) {
        A.f0 = __dadd_rn(A.f0, dvx); A.f1 = __dadd_rn(A.f1, q4x); A.f2 = __dadd_rn(A.f2, q2x);
        A.f3 = __dadd_rn(A.f3, dvy); A.f4 = __dadd_rn(A.f4, q4y); A.f5 = __dadd_rn(A.f5, q2y);
    }
    double gx = __dadd_rn(__dmul_rn(-2.0, q4x), __dmul_rn(2.0, q2x));
    double gy = __dadd_rn(__dmul_rn(-2.0, q4y), __dmul_rn(2.0, q2y));
    if (r2 > R) { gx = 0.0; gy = 0.0; }          // unsquared radius, strict > (flocking_relative.py:225)
    if (!cen && !a) { dvx = 0.0; dvy = 0.0; gx = 0.0; gy = 0.0; }
    A.s0 = __dadd_rn(A.s0, dvx); A.s1 = __dadd_rn(A.s1, dvy);
    A.gx = __dadd_rn(A.gx, gx);  A.gy = __dadd_rn(A.gy, gy);
    return a;
}

// Fast pair.  TRICK (fp32): velocities are centred, so the unmasked alignment sum is N*v_i - sum_j v_j
// (O(N), done outside) and the masked one is deg*v_i - sum_j a*v_j: no dv subtraction per pair.
template <typename T, bool TRICK>
__device__ __forceinline__ bool pair_fast(const typename Tr<T>::V4 xi, const typename Tr<T>::V4 xj, const T R2,
                                          const T Rm, Acc<T> &A, T &r2o) {
    const T dx = xi.x - xj.x, dy = xi.y - xj.y;
    T r2 = Tr<T>::fma_(dx, dx, dy * dy);
    r2 = Tr<T>::max_(r2, Tr<T>::tiny());          // diagonal: dx = dy = 0 -> every product below is 0
    r2o = r2;
    const T inv2 = Tr<T>::rcp(r2);
    const bool a = r2 < R2;
    const T w2 = a ? inv2 : T(0);
    const T w4 = w2 * w2;
    A.f1 = Tr<T>::fma_(dx, w4, A.f1); A.f2 = Tr<T>::fma_(dx, w2, A.f2);
    A.f4 = Tr<T>::fma_(dy, w4, A.f4); A.f5 = Tr<T>::fma_(dy, w2, A.f5);
    if constexpr (TRICK) {
        const T af = a ? T(1) : T(0);
        A.f0 = Tr<T>::fma_(af, xj.z, A.f0); A.f3 = Tr<T>::fma_(af, xj.w, A.f3);
    } else {
        const T dvx = xi.z - xj.z, dvy = xi.w - xj.w;
        A.s0 += dvx; A.s1 += dvy;
        if (a) { A.f0 += dvx; A.f3 += dvy; }
    }
    T c = Tr<T>::fma_(-inv2, inv2, inv2);         // 1/r2 - 1/r4
    c = (r2 <= Rm) ? c : T(0);                    // flocking_relative.py:225 (and :205-206 folded into Rm)
    A.gx = Tr<T>::fma_(dx, c, A.gx); A.gy = Tr<T>::fma_(dy, c, A.gy);
    return a;
}

template <typename T, int LANES, bool EXACT, bool KNN>
__global__ void __launch_bounds__(512) flock_env_kernel(const KArgs a) {
    using V4 = typename Tr<T>::V4;
    constexpr bool CENTRE = std::is_same<T, float>::value;   // fp32: centre in fp64, then convert
    constexpr bool TRICK = CENTRE;
    static_assert(!EXACT || (std::is_same<T, double>::value && LANES == 1), "EXACT is fp64, one thread per row");

    extern __shared__ __align__(16) unsigned char smem[];
    V4 *xs = reinterpret_cast<V4 *>(smem + a.off_xs);
    double4 *stage = reinterpret_cast<double4 *>(smem + a.off_stage);
    unsigned *masks = reinterpret_cast<unsigned *>(smem + a.off_masks);
    T *invdeg = reinterpret_cast<T *>(smem + a.off_invdeg);
    double *stats = reinterpret_cast<double *>(smem + a.off_stats);   // per env: [svx, svy] (TRICK)

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int N = a.N;
    const int grp = blockIdx.x / a.tiles, tile = blockIdx.x - grp * a.tiles;
    const int env0 = grp * a.G;
    const int genv = min(a.G, a.E - env0);
    const int nag = genv * N;
    const bool writer = (tile == 0);

    // ---- prologue: load, integrate, publish x', stage --------------------------------------------
    for (int idx = tid; idx < nag; idx += nthr) {
        const size_t grow = (size_t)env0 * N + idx;
        double4 x = ld_state(a.x_in + grow * 4);
        if (a.u_in) {
            const double2 u = __ldg(reinterpret_cast<const double2 *>(a.u_in) + grow);
            x = integrate_exact(x, u, a.dt, a.ascale);
        }
        if (a.x_out && writer) st_state(a.x_out + grow * 4, x);
        if constexpr (CENTRE) stage[idx] = x;
        else xs[idx] = x;
    }
    __syncthreads();

    // ---- per-env statistics: reward (+ centring in fp32 mode); one warp per env -------------------
    T *reward = reinterpret_cast<T *>(a.reward);
    for (int g = warp; g < genv; g += nwarps) {
        if constexpr (EXACT) {
            if (lane == 0 && reward && writer) {      // NumPy's two-pass variance, sequential sums
                const double4 *xe = xs + g * N;
                double sx = 0.0, sy = 0.0;
                for (int j = 0; j < N; ++j) { sx = __dadd_rn(sx, xe[j].z); sy = __dadd_rn(sy, xe[j].w); }
                const double mx = __ddiv_rn(sx, (double)N), my = __ddiv_rn(sy, (double)N);
                double qx = 0.0, qy = 0.0;
                for (int j = 0; j < N; ++j) {
                    const double ex = __dsub_rn(xe[j].z, mx), ey = __dsub_rn(xe[j].w, my);
                    qx = __dadd_rn(qx, __dmul_rn(ex, ex)); qy = __dadd_rn(qy, __dmul_rn(ey, ey));
                }
                reward[env0 + g] = __dmul_rn(-1.0, __dadd_rn(__ddiv_rn(qx, (double)N), __ddiv_rn(qy, (double)N)));
            }
        } else if constexpr (CENTRE) {
            const double4 *se = stage + g * N;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int j = lane; j < N; j += 32) { const double4 v = se[j]; s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w; }
            const double invN = 1.0 / (double)N;
            const double m0 = warp_sum(s0) * invN, m1 = warp_sum(s1) * invN;
            const double m2 = warp_sum(s2) * invN, m3 = warp_sum(s3) * invN;
            double qx = 0, qy = 0;
            float svx = 0.f, svy = 0.f;
            for (int j = lane; j < N; j += 32) {
                const double4 v = se[j];
                const double ex = v.z - m2, ey = v.w - m3;
                const float4 c = make_float4((float)(v.x - m0), (float)(v.y - m1), (float)ex, (float)ey);
                xs[g * N + j] = c;
                qx = fma(ex, ex, qx); qy = fma(ey, ey, qy);
                svx += c.z; svy += c.w;
            }
            qx = warp_sum(qx); qy = warp_sum(qy); svx = warp_sum(svx); svy = warp_sum(svy);
            if (lane == 0) {
                stats[g * 2 + 0] = (double)svx; stats[g * 2 + 1] = (double)svy;
                if (reward && writer) reward[env0 + g] = (T)(-(qx + qy) * invN);
            }
        } else {
            if (reward && writer) {
                const double4 *xe = xs + g * N;
                double s2 = 0, s3 = 0;
                for (int j = lane; j < N; j += 32) { s2 += xe[j].z; s3 += xe[j].w; }
                const double invN = 1.0 / (double)N;
                const double m2 = warp_sum(s2) * invN, m3 = warp_sum(s3) * invN;
                double qx = 0, qy = 0;
                for (int j = lane; j < N; j += 32) {
                    const double ex = xe[j].z - m2, ey = xe[j].w - m3;
                    qx = fma(ex, ex, qx); qy = fma(ey, ey, qy);
                }
                qx = warp_sum(qx); qy = warp_sum(qy);
                if (lane == 0) reward[env0 + g] = (T)(-(qx + qy) * invN);
            }
        }
    }
    if constexpr (CENTRE) __syncthreads();   // xs / stats complete; stage (aliases masks) is dead from here

    // ---- pair loop -------------------------------------------------------------------------------
    const int nrows_env = a.row_hi - a.row_lo;
    const int rl = tid / LANES, sub = tid % LANES;
    const int fr = tile * a.RT + rl;                       // flat row over the group's rows
    const bool valid = (rl < a.RT) && (fr < genv * nrows_env);
    const int g = valid ? fr / nrows_env : 0;
    const int i = valid ? a.row_lo + (fr - g * nrows_env) : 0;
    const V4 *xe = xs + g * N;
    const V4 xi = xe[i];
    const T R2 = (T)a.R2;
    const T Rm = std::is_same<T, float>::value ? (T)a.Rm_f : (T)a.Rm;
    const bool cen = a.centralized != 0;

    Acc<T> A;
    A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = T(0);
    TopK<T> top;
    if constexpr (KNN) top.init(i);
    int deg = 0;
    const int nk = valid ? (N - sub + LANES - 1) / LANES : 0;   // pairs handled by this lane
    const int kdiag = (i % LANES == sub) ? i / LANES : -1;      // position of the diagonal, if ours
    unsigned *mrow = masks + (size_t)(rl * LANES + sub) * a.W;

    for (int w = 0; w * 32 < nk; ++w) {
        unsigned bits = 0;
        const int kbase = w * 32;
        const int kend = min(32, nk - kbase);
        auto body = [&](const int kk) {
            const int j = sub + LANES * (kbase + kk);
            const V4 xj = xe[j];
            T r2;
            bool adj;
            if constexpr (EXACT) adj = pair_exact(xi, xj, j == i, a.R, a.R2, cen, A, r2);
            else adj = pair_fast<T, TRICK>(xi, xj, R2, Rm, A, r2);
            bits |= adj ? (1u << kk) : 0u;
            if constexpr (KNN) {
                if constexpr (!EXACT) r2 = (j == i) ? Tr<T>::inf() : r2;
                top.insert(r2, j);
            }
        };
        if (kend == 32) {
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) body(kk);
        } else {
            for (int kk = 0; kk < kend; ++kk) body(kk);
        }
        if (kdiag >= kbase && kdiag < kbase + 32) bits &= ~(1u << (kdiag - kbase));   // fast path set it
        mrow[w] = bits;
        deg += __popc(bits);
    }

    // ---- reduce across the LANES threads of a row ----------------------------------------------------
    if constexpr (LANES > 1) {
#pragma unroll
        for (int m = LANES / 2; m > 0; m >>= 1) {
            A.f0 += shfl_xor(A.f0, m); A.f1 += shfl_xor(A.f1, m); A.f2 += shfl_xor(A.f2, m);
            A.f3 += shfl_xor(A.f3, m); A.f4 += shfl_xor(A.f4, m); A.f5 += shfl_xor(A.f5, m);
            A.gx += shfl_xor(A.gx, m); A.gy += shfl_xor(A.gy, m);
            if constexpr (!TRICK) { A.s0 += shfl_xor(A.s0, m); A.s1 += shfl_xor(A.s1, m); }
            deg += shfl_xor(deg, m);
        }
    }

    // ---- row epilogue: features, action, 1/deg -------------------------------------------------------
    const size_t grow = (size_t)(env0 + g) * N + i;
    if (valid && sub == 0) {
        T f0 = A.f0, f3 = A.f3, sdx, sdy;
        if constexpr (TRICK) {
            // the fast path's diagonal (r2 clamped to tiny) always passes r2 < R2, so A.f0 holds v_i once
            f0 = Tr<T>::fma_((T)(deg + 1), xi.z, -A.f0);
            f3 = Tr<T>::fma_((T)(deg + 1), xi.w, -A.f3);
            sdx = cen ? Tr<T>::fma_((T)N, xi.z, -(T)stats[g * 2 + 0]) : f0;
            sdy = cen ? Tr<T>::fma_((T)N, xi.w, -(T)stats[g * 2 + 1]) : f3;
        } else if constexpr (EXACT) {
            sdx = A.s0; sdy = A.s1;                  // pair_exact already applied the adjacency mask
        } else {
            sdx = cen ? A.s0 : f0; sdy = cen ? A.s1 : f3;
        }
        if (a.values) {
            T *v = reinterpret_cast<T *>(a.values) + grow * NFE;
            v[0] = f0; v[1] = A.f1; v[2] = A.f2; v[3] = f3; v[4] = A.f4; v[5] = A.f5;
        }
        if (a.u_next) {
            double ux, uy;
            if constexpr (EXACT) {
                ux = __dsub_rn(-A.gx, sdx);          // (-S.gx) - S.dvx      flocking_relative.py:209
                uy = __dsub_rn(-sdy, A.gy);          // (-S.dvy) - S.gy
                ux = fmin(fmax(ux, -10.0), 10.0); uy = fmin(fmax(uy, -10.0), 10.0);   // :210
                ux = __ddiv_rn(ux, a.ascale); uy = __ddiv_rn(uy, a.ascale);           // :211
            } else {
                T tx = -T(2) * A.gx - sdx, ty = -sdy - T(2) * A.gy;
                tx = Tr<T>::min_(Tr<T>::max_(tx, T(-10)), T(10));
                ty = Tr<T>::min_(Tr<T>::max_(ty, T(-10)), T(10));
                ux = (double)tx / a.ascale; uy = (double)ty / a.ascale;
            }
            reinterpret_cast<double2 *>(a.u_next)[grow] = make_double2(ux, uy);
        }
        T idg;
        if constexpr (EXACT) idg = __ddiv_rn(1.0, (double)max(deg, 1));               // :121-122
        else idg = T(1) / (T)max(deg, 1);
        invdeg[rl] = idg;
    }

    // ---- kNN observation: merge the LANES partial lists, write obs[i, 4m:4m+4] = x_i - x_nbr(m) ---------
    if constexpr (KNN) {
        int nb[KNN_K];
#pragma unroll
        for (int m = 0; m < KNN_K; ++m) {
            if constexpr (LANES > 1) {
                T bd = top.d[0]; int bi = top.id[0];
#pragma unroll
                for (int s = LANES / 2; s > 0; s >>= 1) {
                    const T od = shfl_xor(bd, s); const int oi = shfl_xor(bi, s);
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                }
                if (bi == top.id[0] && bd == top.d[0]) top.pop_front(i);
                nb[m] = bi;
            } else {
                nb[m] = top.id[m];
            }
        }
        if (valid && sub == 0 && a.obs) {
            T *o = reinterpret_cast<T *>(a.obs) + grow * (NXS * KNN_K);
#pragma unroll
            for (int m = 0; m < KNN_K; ++m) {
                const V4 xn = xe[nb[m]];
                if constexpr (EXACT) {
                    o[m * 4 + 0] = __dsub_rn(xi.x, xn.x); o[m * 4 + 1] = __dsub_rn(xi.y, xn.y);
                    o[m * 4 + 2] = __dsub_rn(xi.z, xn.z); o[m * 4 + 3] = __dsub_rn(xi.w, xn.w);
                } else {
                    o[m * 4 + 0] = xi.x - xn.x; o[m * 4 + 1] = xi.y - xn.y;
                    o[m * 4 + 2] = xi.z - xn.z; o[m * 4 + 3] = xi.w - xn.w;
                }
            }
        }
    }

    // ---- dense adj_mean rows: expand the bit masks, coalesced streaming stores --------------------------
    if (a.network == nullptr) return;
    __syncthreads();
    const int rows_cta = min(a.RT, genv * nrows_env - tile * a.RT);    // rows this CTA owns (>= 1)
    // flat row fr0 = tile*RT of the group -> (env, row).  With G > 1 there is one tile and all rows.
    T *net = reinterpret_cast<T *>(a.network);
    auto bit_of = [&](const int r, const int j) -> bool {
        const int s = j % LANES, k = j / LANES;
        return (masks[(size_t)(r * LANES + s) * a.W + (k >> 5)] >> (k & 31)) & 1u;
    };
    {
        // rows [row_lo + tile*RT, +rows_cta) of env0.. are consecutive in memory (with G > 1 there is one
        // tile and row_lo = 0): the CTA's output is ONE contiguous run of rows_cta*N elements.
        const size_t base = ((size_t)env0 * N + (size_t)a.row_lo + (size_t)tile * a.RT) * N;
        T *out = net + base;
        const int total = rows_cta * N;
        constexpr int VEC = 16 / (int)sizeof(T);
        const int head = min((int)((VEC - (base % VEC)) % VEC), total);
        for (int f = tid; f < head; f += nthr) {
            const int r = f / N, j = f - r * N;
            st_stream(out + f, bit_of(r, j) ? invdeg[r] : T(0));
        }
        const int nvec = (total - head) / VEC;
        if (nvec > 0) {
            const int stride = nthr * VEC;
            const int dr = stride / N, dj = stride - dr * N;
            int f = head + tid * VEC;
            int r = f / N, j = f - r * N;
            for (int q = tid; q < nvec; q += nthr) {
                T v[VEC];
                if (LANES == 1 && j + VEC <= N && ((j & 31) + VEC <= 32)) {
                    const unsigned nib = masks[(size_t)r * a.W + (j >> 5)] >> (j & 31);
                    const T val = invdeg[r];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[e] = ((nib >> e) & 1u) ? val : T(0);
                } else {
                    int rr = r, jj = j;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        v[e] = bit_of(rr, jj) ? invdeg[rr] : T(0);
                        if (++jj == N) { jj = 0; ++rr; }
                    }
                }
                if constexpr (VEC == 4) st_stream4(out + f, v[0], v[1], v[2], v[3]);
                else st_stream2(out + f, v[0], v[1]);
                f += stride; r += dr; j += dj;
                if (j >= N) { j -= N; ++r; }
            }
        }
        for (int f = head + nvec * VEC + tid; f < total; f += nthr) {
            const int r = f / N, j = f - r * N;
            st_stream(out + f, bit_of(r, j) ? invdeg[r] : T(0));
        }
    }
}

}  // namespace flock
