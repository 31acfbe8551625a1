// flock_env_kernel.cuh -- the env-resident fused flocking step (sm_100a), "filter then refine".
//
// One CTA owns either G whole environments (small N) or one row tile of one environment (large N);
// either way the whole state x[N,4] of its environment(s) is staged in shared memory, so each ordered
// pair (i,j) is visited exactly once per step.  Per env and step, in ONE launch:
//
//   zero-fill : the CTA's slice of the dense adj_mean output is cleared first with coalesced 16-byte
//               stores (fire and forget: the HBM write overlaps everything below)
//   prologue  : x' = integrate(x_in, u_in)                          flocking_relative.py:96-105
//               reward = -(var vx' + var vy')                       flocking_relative.py:145-147
//   phase A   : FILTER over all N^2 ordered pairs: r2 and one compare per pair, candidate bits
//               (r2 <= max(R, R^2)) kept in shared memory.          flocking_relative.py:113-117
//               fp32 output mode: packed f32x2 math on centroid-centred fp32 positions with a
//               conservative threshold; fp64 modes: fp64 r2 (exact mode: NumPy's unfused r2).
//               Flocking-v0: a running top-8 by r2 rides along.     flocking.py:20-24
//   phase B   : REFINE only the candidates (~deg of N) in fp64: exact adjacency bit, the 6 masked
//               feature sums, the potential gradient.               flocking_relative.py:117-129,214-226
//               EXACT: NumPy's operation order, separately rounded, IEEE divide, ascending j.
//   epilogue  : values, u_next (clip, /10), obs; the few non-zero adj_mean entries (1/deg) are scattered
//               into the zero-filled rows.                          flocking_relative.py:120-122,209-211
//
// Non-neighbour pairs contribute exact zeros to every masked sum of the reference, so skipping them in
// phase B changes nothing (not even in EXACT mode, up to the sign of a zero).  The centralized alignment
// term sum_j (v_i - v_j) runs over ALL j: EXACT accumulates it sequentially in phase A; the fast modes use
// N*(v_i - mean) - sum_j (v_j - mean) in fp64 (O(N)).
//
// LANES threads cooperate on one row (pairs interleaved across them, warp-shuffle reduction) so small
// batches still fill 148 SMs.
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
    int W;        // 32-bit candidate words per (row, lane)
    int NP2;      // pair slots per env in the packed fp32 position array: (N + 1) / 2
    double R, R2, dt, ascale;
    double Rm;    // fast modes: gradient mask as ONE compare r2 <= Rm (centralized: R; decentralized: also r2 < R2)
    double thrC;  // candidate threshold max(R, R2)
    unsigned off_xs, off_pf, off_masks, off_stats;
    unsigned off_zero;   // ZERO_BYTES of zeros: source of the bulk (TMA) zero-fill of the dense adjacency
    unsigned off_sort;   // LANES == 1: hist[SORT_BINS] + perm[RT] for the candidate-count row sort
    int tune;            // diagnostics: bit0 skip fill, bit1 skip scatter, bit2 plain-store fill
};

constexpr int ZERO_BYTES = 8192;
constexpr int SORT_BINS = 64;

// ---- bulk async copy (TMA, 1-D) shared -> global, used to clear the dense adjacency slice in the background ----
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, unsigned bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

constexpr int STATS_PER_ENV = 6;   // mean vx, mean vy, residual sum vx, residual sum vy, thrA, (pad)
constexpr int TOPK = KNN_K + 1;    // the diagonal (r2 = 0) rides along and is dropped at the end

template <typename T> struct TopK {
    T d[TOPK];
    int id[TOPK];
    __device__ __forceinline__ void init(int self) {
#pragma unroll
        for (int m = 0; m < TOPK; ++m) { d[m] = Tr<T>::inf(); id[m] = self; }
    }
    // strict < everywhere: among equal r2 the lower index (seen first) stays in front
    __device__ __forceinline__ void insert(T r2, int j) {
        if (r2 < d[TOPK - 1]) {
            d[TOPK - 1] = r2; id[TOPK - 1] = j;
#pragma unroll
            for (int p = TOPK - 1; p > 0; --p) {
                if (d[p] < d[p - 1]) {
                    const T td = d[p]; d[p] = d[p - 1]; d[p - 1] = td;
                    const int ti = id[p]; id[p] = id[p - 1]; id[p - 1] = ti;
                }
            }
        }
    }
    __device__ __forceinline__ void pop_front(int self) {
#pragma unroll
        for (int p = 0; p < TOPK - 1; ++p) { d[p] = d[p + 1]; id[p] = id[p + 1]; }
        d[TOPK - 1] = Tr<T>::inf(); id[TOPK - 1] = self;
    }
};

struct Acc {
    double f0, f1, f2, f3, f4, f5;   // masked feature sums
    double gx, gy;                   // EXACT: sum of grad; fast: sum d*(1/r2 - 1/r4) (x2 at the end)
    double s0, s1;                   // EXACT alignment sums (centralized: all j, phase A; else masked, phase B)
};

// EXACT refine of one candidate pair: Appendix B of SURVEY.md, one rounding per operation, NumPy's order.
__device__ __forceinline__ bool refine_exact(const double4 xi, const double4 xj, const double R, const double R2,
                                             const bool cen, Acc &A) {
    const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
    const double dvx = __dsub_rn(xi.z, xj.z), dvy = __dsub_rn(xi.w, xj.w);
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const bool a = r2 < R2;
    const double r4 = __dmul_rn(r2, r2);
    const double q4x = __ddiv_rn(dx, r4), q2x = __ddiv_rn(dx, r2);
    const double q4y = __ddiv_rn(dy, r4), q2y = __ddiv_rn(dy, r2);
    if (a) {
        A.f0 = __dadd_rn(A.f0, dvx); A.f1 = __dadd_rn(A.f1, q4x); A.f2 = __dadd_rn(A.f2, q2x);
        A.f3 = __dadd_rn(A.f3, dvy); A.f4 = __dadd_rn(A.f4, q4y); A.f5 = __dadd_rn(A.f5, q2y);
        if (!cen) { A.s0 = __dadd_rn(A.s0, dvx); A.s1 = __dadd_rn(A.s1, dvy); }      // :205-206
    }
    if (!(r2 > R) && (cen || a)) {                 // unsquared radius, strict > (flocking_relative.py:225)
        A.gx = __dadd_rn(A.gx, __dadd_rn(__dmul_rn(-2.0, q4x), __dmul_rn(2.0, q2x)));   // :224
        A.gy = __dadd_rn(A.gy, __dadd_rn(__dmul_rn(-2.0, q4y), __dmul_rn(2.0, q2y)));
    }
    return a;
}

// Fast refine (fp64, fused): one reciprocal instead of four divides.
__device__ __forceinline__ bool refine_fast(const double4 xi, const double4 xj, const double R2, const double Rm,
                                            Acc &A) {
    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
    const double r2 = fma(dx, dx, dy * dy);
    const double inv2 = Tr<double>::rcp(r2);
    const bool a = r2 < R2;
    if (a) {
        const double inv4 = inv2 * inv2;
        A.f0 += xi.z - xj.z; A.f3 += xi.w - xj.w;
        A.f1 = fma(dx, inv4, A.f1); A.f2 = fma(dx, inv2, A.f2);
        A.f4 = fma(dy, inv4, A.f4); A.f5 = fma(dy, inv2, A.f5);
    }
    if (r2 <= Rm) {
        const double c = fma(-inv2, inv2, inv2);   // 1/r2 - 1/r4
        A.gx = fma(dx, c, A.gx); A.gy = fma(dy, c, A.gy);
    }
    return a;
}

template <typename TO, int LANES, bool EXACT, bool KNN>
__global__ void __launch_bounds__(512) flock_env_kernel(const KArgs a) {
    constexpr bool PA32 = std::is_same<TO, float>::value;   // phase A on packed fp32 positions
    using TK = typename std::conditional<PA32, float, double>::type;
    static_assert(!EXACT || (!PA32 && LANES == 1), "EXACT is fp64, one thread per row");

    extern __shared__ __align__(16) unsigned char smem[];
    double4 *xs = reinterpret_cast<double4 *>(smem + a.off_xs);
    float4 *pf = reinterpret_cast<float4 *>(smem + a.off_pf);       // (px_2m, px_2m+1, py_2m, py_2m+1), centred
    unsigned *masks = reinterpret_cast<unsigned *>(smem + a.off_masks);
    double *stats = reinterpret_cast<double *>(smem + a.off_stats);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int N = a.N;
    const int grp = blockIdx.x / a.tiles, tile = blockIdx.x - grp * a.tiles;
    const int env0 = grp * a.G;
    const int genv = min(a.G, a.E - env0);
    const int nag = genv * N;
    const bool writer = (tile == 0);
    const int nrows_env = a.row_hi - a.row_lo;
    const int rows_cta = min(a.RT, genv * nrows_env - tile * a.RT);     // rows this CTA owns (>= 1)
    TO *net = reinterpret_cast<TO *>(a.network);
    // rows [row_lo + tile*RT, +rows_cta) of env0.. are consecutive in memory (with G > 1 there is one tile and
    // row_lo = 0): the CTA's slice of the dense output is ONE contiguous run of rows_cta*N elements.
    const size_t net_base = ((size_t)env0 * N + (size_t)a.row_lo + (size_t)tile * a.RT) * N;

    // ---- zero source for the bulk fill + sort histogram ------------------------------------------------
    {
        float4 *zb = reinterpret_cast<float4 *>(smem + a.off_zero);
        for (int q = tid; q < ZERO_BYTES / 16; q += nthr) zb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LANES == 1) for (int q = tid; q < SORT_BINS; q += nthr) reinterpret_cast<int *>(smem + a.off_sort)[q] = 0;
        fence_async_smem();      // make the generic-proxy zeros visible to the async proxy (TMA) before the barrier
    }

    // ---- prologue: load, integrate, publish x', stage -------------------------------------------------
    for (int idx = tid; idx < nag; idx += nthr) {
        const size_t grow = (size_t)env0 * N + idx;
        double4 x = ld_state(a.x_in + grow * 4);
        if (a.u_in) {
            const double2 u = __ldg(reinterpret_cast<const double2 *>(a.u_in) + grow);
            x = integrate_exact(x, u, a.dt, a.ascale);
        }
        if (a.x_out && writer) st_state(a.x_out + grow * 4, x);
        xs[idx] = x;
    }
    __syncthreads();

    // ---- zero-fill of the dense adj_mean slice: ONE thread queues 1-D bulk (TMA) stores of the shared zero block;
    // the copy engine streams them to HBM while every warp computes.  Unaligned head/tail elements: plain stores.
    if (net) {
        TO *out = net + net_base;
        const int total = rows_cta * N;
        constexpr int VEC = 16 / (int)sizeof(TO);
        const int head = min((int)((VEC - (net_base % VEC)) % VEC), total);
        const int nvec = (total - head) / VEC;
        if (tid < head) out[tid] = TO(0);
        if (tid < total - (head + nvec * VEC)) out[head + nvec * VEC + tid] = TO(0);
        if (a.tune & 4) {
            float4 *o4 = reinterpret_cast<float4 *>(out + head);
            for (int q = tid; q < nvec; q += nthr) o4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (tid == 0 && !(a.tune & 1)) {
            char *dst = reinterpret_cast<char *>(out + head);
            const size_t bytes = (size_t)nvec * 16;
            for (size_t off = 0; off < bytes; off += ZERO_BYTES)
                bulk_store_s2g(dst + off, smem + a.off_zero, (unsigned)min((size_t)ZERO_BYTES, bytes - off));
            bulk_commit();
        }
    }

    // ---- per-env statistics, one warp per env: reward, velocity mean (alignment trick), fp32 positions ----
    TO *reward = reinterpret_cast<TO *>(a.reward);
    for (int g = warp; g < genv; g += nwarps) {
        const double4 *xe = xs + g * N;
        if constexpr (EXACT) {
            if (lane == 0 && reward && writer) {      // NumPy's two-pass variance, sequential sums
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
        } else {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int j = lane; j < N; j += 32) { const double4 v = xe[j]; s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w; }
            const double invN = 1.0 / (double)N;
            const double m0 = warp_sum(s0) * invN, m1 = warp_sum(s1) * invN;
            const double m2 = warp_sum(s2) * invN, m3 = warp_sum(s3) * invN;
            double qx = 0, qy = 0, rx = 0, ry = 0;
            float pmax = 0.f;
            for (int j = lane; j < N; j += 32) {
                const double4 v = xe[j];
                const double ex = v.z - m2, ey = v.w - m3;
                qx = fma(ex, ex, qx); qy = fma(ey, ey, qy);
                rx += ex; ry += ey;
                if constexpr (PA32) {
                    const float cx = (float)(v.x - m0), cy = (float)(v.y - m1);
                    pmax = fmaxf(pmax, fmaxf(fabsf(cx), fabsf(cy)));
                    float *p = reinterpret_cast<float *>(pf + (size_t)g * a.NP2 + (j >> 1));
                    p[j & 1] = cx; p[2 + (j & 1)] = cy;
                }
            }
            qx = warp_sum(qx); qy = warp_sum(qy); rx = warp_sum(rx); ry = warp_sum(ry);
            if constexpr (PA32) {
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, m));
            }
            if (lane == 0) {
                double *st = stats + g * STATS_PER_ENV;
                st[0] = m2; st[1] = m3; st[2] = rx; st[3] = ry;
                if constexpr (PA32) {
                    // conservative fp32 candidate threshold: every pair with true r2 <= thrC passes.  Coordinates
                    // are rounded to within pmax*2^-24 each; |d(r2)| <= 2(|dx|+|dy|)*2*that + a few ulp of r2.
                    const float tc = (float)a.thrC;
                    st[4] = (double)(tc * (1.0f + 2e-6f) + 16.0f * sqrtf(tc) * pmax * 1.1920929e-7f + 1e-30f);
                    if (N & 1) {       // pad the odd slot with an agent at infinity: r2 = inf, never a candidate
                        float *p = reinterpret_cast<float *>(pf + (size_t)g * a.NP2 + (N >> 1));
                        p[1] = CUDART_INF_F; p[3] = CUDART_INF_F;
                    }
                }
                if (reward && writer) reward[env0 + g] = (TO)(-(qx + qy) * invN);
            }
        }
    }
    if constexpr (!EXACT) __syncthreads();

    // ---- row assignment -----------------------------------------------------------------------------------
    const int rl0 = tid / LANES, sub = tid % LANES;
    const int fr0 = tile * a.RT + rl0;                     // flat row over the group's rows
    const bool valid0 = (rl0 < a.RT) && (fr0 < genv * nrows_env);
    const int g0 = valid0 ? fr0 / nrows_env : 0;
    const int i0 = valid0 ? a.row_lo + (fr0 - g0 * nrows_env) : 0;
    const bool cen = a.centralized != 0;

    Acc A;
    A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = 0.0;

    // ---- phase A: filter ----------------------------------------------------------------------------------
    // Each lane owns bit positions k = 0.. of its words; j_of(k) maps them to agents.
    //   fp32 packed : pair slot m = sub + LANES*(k>>1), agent j = 2m + (k&1)
    //   fp64        : j = sub + LANES*k
    int cnt = 0;                                           // candidates found by this lane
    {
        const int g = g0, i = i0;
        const double4 *xe = xs + g * N;
        const double4 xi = xe[i];
        unsigned *mrow = masks + (size_t)(rl0 * LANES + sub) * a.W;
        TopK<TK> top;
        if constexpr (KNN) top.init(i);
        if constexpr (PA32) {
            const int nslots = valid0 ? (a.NP2 - sub + LANES - 1) / LANES : 0;
            const float4 *pe = pf + (size_t)g * a.NP2;
            const float *pif = reinterpret_cast<const float *>(pe + (i >> 1));
            const float pxi = pif[i & 1], pyi = pif[2 + (i & 1)];
            const float2 nx = make_float2(-pxi, -pxi), ny = make_float2(-pyi, -pyi);
            const float thrA = (float)stats[g * STATS_PER_ENV + 4];
            for (int w = 0; w * 16 < nslots; ++w) {
                unsigned bits = 0;
                const int tbase = w * 16;
                const int tend = min(16, nslots - tbase);
                auto body = [&](const int tt) {
                    const int m = sub + LANES * (tbase + tt);
                    const float4 q = pe[m];
                    const float2 dx = __fadd2_rn(make_float2(q.x, q.y), nx);
                    const float2 dy = __fadd2_rn(make_float2(q.z, q.w), ny);
                    const float2 r2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                    bits |= (r2.x <= thrA) ? (1u << (2 * tt)) : 0u;
                    bits |= (r2.y <= thrA) ? (2u << (2 * tt)) : 0u;
                    if constexpr (KNN) { top.insert(r2.x, 2 * m); top.insert(r2.y, 2 * m + 1); }
                };
                if (tend == 16) {
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) body(tt);
                } else {
                    for (int tt = 0; tt < tend; ++tt) body(tt);
                }
                mrow[w] = bits;
                cnt += __popc(bits);
            }
        } else {
            const int nk = valid0 ? (N - sub + LANES - 1) / LANES : 0;
            const double thrC = a.thrC;
            for (int w = 0; w * 32 < nk; ++w) {
                unsigned bits = 0;
                const int kbase = w * 32;
                const int kend = min(32, nk - kbase);
                auto body = [&](const int kk) {
                    const int j = sub + LANES * (kbase + kk);
                    const double4 xj = xe[j];
                    double r2;
                    bool c;
                    if constexpr (EXACT) {
                        const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
                        r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        c = (r2 < a.R2) || !(r2 > a.R);
                        if (cen) {   // sum_j (v_i - v_j) over ALL j, sequential (flocking_relative.py:204-208)
                            A.s0 = __dadd_rn(A.s0, __dsub_rn(xi.z, xj.z));
                            A.s1 = __dadd_rn(A.s1, __dsub_rn(xi.w, xj.w));
                        }
                    } else {
                        const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                        r2 = fma(dx, dx, dy * dy);
                        c = r2 <= thrC;
                    }
                    bits |= c ? (1u << kk) : 0u;
                    if constexpr (KNN) top.insert(r2, j);
                };
                if (kend == 32) {
#pragma unroll 8
                    for (int kk = 0; kk < 32; ++kk) body(kk);
                } else {
                    for (int kk = 0; kk < kend; ++kk) body(kk);
                }
                mrow[w] = bits;
                cnt += __popc(bits);
            }
        }

        // ---- kNN observation (independent of phase B): merge the LANES partial lists, drop the diagonal,
        //      obs[i, 4m:4m+4] = x_i - x_nbr(m)                                              flocking.py:20-24
        if constexpr (KNN) {
            int nb[TOPK];
#pragma unroll
            for (int m = 0; m < TOPK; ++m) {
                if constexpr (LANES > 1) {
                    TK bd = top.d[0]; int bi = top.id[0];
#pragma unroll
                    for (int s = LANES / 2; s > 0; s >>= 1) {
                        const TK od = shfl_xor(bd, s); const int oi = shfl_xor(bi, s);
                        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    }
                    if (bi == top.id[0] && bd == top.d[0]) top.pop_front(i);
                    nb[m] = bi;
                } else {
                    nb[m] = top.id[m];
                }
            }
            if (valid0 && sub == 0 && a.obs) {
                TO *o = reinterpret_cast<TO *>(a.obs) + ((size_t)(env0 + g) * N + i) * (NXS * KNN_K);
                int m = 0;
                bool dropped = false;
#pragma unroll
                for (int q = 0; q < TOPK; ++q) {
                    if (!dropped && nb[q] == i) { dropped = true; continue; }     // the diagonal (or an unused slot)
                    if (m < KNN_K) {
                        const double4 xn = xe[nb[q]];
                        if constexpr (EXACT) {
                            o[m * 4 + 0] = __dsub_rn(xi.x, xn.x); o[m * 4 + 1] = __dsub_rn(xi.y, xn.y);
                            o[m * 4 + 2] = __dsub_rn(xi.z, xn.z); o[m * 4 + 3] = __dsub_rn(xi.w, xn.w);
                        } else {
                            o[m * 4 + 0] = (TO)(xi.x - xn.x); o[m * 4 + 1] = (TO)(xi.y - xn.y);
                            o[m * 4 + 2] = (TO)(xi.z - xn.z); o[m * 4 + 3] = (TO)(xi.w - xn.w);
                        }
                        ++m;
                    }
                }
            }
        }
    }

    // ---- balance phase B: counting sort of the CTA's rows by candidate count (LANES == 1) -------------------
    // Neighbour counts vary a lot from row to row; lanes of a warp would idle until its busiest row is done.
    // After the sort every warp works on rows with (nearly) equal counts.  EXACT keeps the identity mapping only
    // when the centralized alignment sums (held in this thread's registers) must stay with their row.
    int rl = rl0;
    bool valid = valid0;
    constexpr bool SORT = (LANES == 1) && !EXACT;
    if constexpr (SORT) {
        int *hist = reinterpret_cast<int *>(smem + a.off_sort);
        int *perm = hist + SORT_BINS;
        const int bin = min(cnt, SORT_BINS - 1);
        int rank = 0;
        if (valid0) rank = atomicAdd(&hist[bin], 1);
        __syncthreads();
        if (warp == 0) {                                    // exclusive scan of the 64 bins, two per lane
            const int h0 = hist[2 * lane], h1 = hist[2 * lane + 1];
            int v = h0 + h1;
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, m); if (lane >= m) v += t; }
            const int excl = v - (h0 + h1);
            hist[2 * lane] = excl; hist[2 * lane + 1] = excl + h0;
        }
        __syncthreads();
        if (valid0) perm[hist[bin] + rank] = rl0;
        __syncthreads();
        valid = tid < rows_cta;
        rl = valid ? perm[rows_cta - 1 - tid] : 0;          // heaviest rows first
    }
    const int fr = tile * a.RT + rl;
    const int g = valid ? fr / nrows_env : 0;
    const int i = valid ? a.row_lo + (fr - g * nrows_env) : 0;
    const double4 *xe = xs + g * N;
    const double4 xi = xe[i];
    unsigned *mrow = masks + (size_t)(rl * LANES + sub) * a.W;
    int nk;
    if constexpr (PA32) nk = valid ? 2 * ((a.NP2 - sub + LANES - 1) / LANES) : 0;
    else nk = valid ? (N - sub + LANES - 1) / LANES : 0;
    const int nw = (nk + 31) >> 5;
    auto j_of = [&](const int k) -> int {
        if constexpr (PA32) return 2 * (sub + LANES * (k >> 1)) + (k & 1);
        else return sub + LANES * k;
    };

    // ---- phase B: refine the candidates in fp64, ascending j per lane; rewrite the words with the exact bits ----
    // ONE flat loop over this lane's candidates (not a loop per word): lanes advance through their words
    // independently, so a warp iterates max-over-lanes(total candidates) times, not sum-over-words(max).
    int deg = 0;
    {
        int w = 0;
        bool more = nw > 0;
        unsigned cand = more ? mrow[0] : 0u, ex = 0;
        for (;;) {
            while (more && cand == 0) {                     // word exhausted: publish its exact bits, fetch the next
                mrow[w] = ex; deg += __popc(ex); ex = 0;
                ++w; more = w < nw;
                if (more) cand = mrow[w];
            }
            if (!more) break;
            const int k = __ffs(cand) - 1;
            cand &= cand - 1;
            const int j = j_of(w * 32 + k);
            if (j != i) {                                   // diagonal: r2 = +inf in the reference (:115)
                const double4 xj = xe[j];
                bool adj;
                if constexpr (EXACT) adj = refine_exact(xi, xj, a.R, a.R2, cen, A);
                else adj = refine_fast(xi, xj, a.R2, a.Rm, A);
                ex |= adj ? (1u << k) : 0u;
            }
        }
    }

    // ---- reduce across the LANES threads of a row ----------------------------------------------------------
    if constexpr (LANES > 1) {
#pragma unroll
        for (int m = LANES / 2; m > 0; m >>= 1) {
            A.f0 += shfl_xor(A.f0, m); A.f1 += shfl_xor(A.f1, m); A.f2 += shfl_xor(A.f2, m);
            A.f3 += shfl_xor(A.f3, m); A.f4 += shfl_xor(A.f4, m); A.f5 += shfl_xor(A.f5, m);
            A.gx += shfl_xor(A.gx, m); A.gy += shfl_xor(A.gy, m);
            deg += shfl_xor(deg, m);
        }
    }

    // ---- row epilogue: features, action, 1/deg ---------------------------------------------------------------
    const size_t grow = (size_t)(env0 + g) * N + i;
    TO idg;
    if constexpr (EXACT) idg = __ddiv_rn(1.0, (double)max(deg, 1));               // :121-122
    else if constexpr (PA32) idg = __frcp_rn((float)max(deg, 1));
    else idg = Tr<double>::rcp((double)max(deg, 1));
    if (valid && sub == 0) {
        if (a.values) {
            TO *v = reinterpret_cast<TO *>(a.values) + grow * NFE;
            v[0] = (TO)A.f0; v[1] = (TO)A.f1; v[2] = (TO)A.f2; v[3] = (TO)A.f3; v[4] = (TO)A.f4; v[5] = (TO)A.f5;
        }
        if (a.u_next) {
            double ux, uy;
            if constexpr (EXACT) {
                ux = __dsub_rn(-A.gx, A.s0);             // (-S.gx) - S.dvx      flocking_relative.py:209
                uy = __dsub_rn(-A.s1, A.gy);             // (-S.dvy) - S.gy
                ux = fmin(fmax(ux, -10.0), 10.0); uy = fmin(fmax(uy, -10.0), 10.0);   // :210
                ux = __ddiv_rn(ux, a.ascale); uy = __ddiv_rn(uy, a.ascale);           // :211
            } else {
                const double *st = stats + g * STATS_PER_ENV;
                // centralized: sum_j (v_i - v_j) = N*(v_i - mean) - sum_j (v_j - mean); decentralized: masked sum = f0/f3
                const double sdx = cen ? fma((double)N, xi.z - st[0], -st[2]) : A.f0;
                const double sdy = cen ? fma((double)N, xi.w - st[1], -st[3]) : A.f3;
                const double ia = Tr<double>::rcp(a.ascale);     // 1/10 to ~1 ulp; one extra rounding vs a true divide
                ux = fmin(fmax(-2.0 * A.gx - sdx, -10.0), 10.0) * ia;
                uy = fmin(fmax(-sdy - 2.0 * A.gy, -10.0), 10.0) * ia;
            }
            reinterpret_cast<double2 *>(a.u_next)[grow] = make_double2(ux, uy);
        }
    }

    // ---- scatter the non-zero adj_mean entries (1/deg) into the zero-filled rows -------------------------------
    // Thread 0 waits for the bulk zero stores of this CTA to complete, orders them (async proxy) before the
    // generic-proxy stores below, then the barrier releases everybody.  Each CTA only touches its own slice.
    if (net == nullptr) return;
    if (tid == 0) { bulk_wait_all(); fence_async_all(); }
    __syncthreads();
    if (valid && !(a.tune & 2)) {
        TO *orow = net + grow * N;
        int w = 0;
        bool more = nw > 0;
        unsigned ex = more ? mrow[0] : 0u;
        for (;;) {
            while (more && ex == 0) { ++w; more = w < nw; if (more) ex = mrow[w]; }
            if (!more) break;
            const int k = __ffs(ex) - 1;
            ex &= ex - 1;
            orow[j_of(w * 32 + k)] = idg;
        }
    }
}

}  // namespace flock
