// flock_group_kernel.cuh -- the batched flocking step for small environments (N <= 512), sm_100a.
//
// Same algorithm and the same per-row arithmetic as flock_env_kernel.cuh ("filter then refine", one thread per row),
// restructured around WHOLE environments:
//
//   * A CTA owns G complete envs (G*N rows, one thread per row; G is chosen so that G*N fills the CTA: 5 envs of 100
//     agents in 512 threads).  Nothing is staged twice, no window crosses an env, every index is a compile-time
//     expression when N is a template constant (NC = 100 / 50; NC = 0: run-time N).
//   * The kernel runs `steps` consecutive steps of the driver loop {u = controller(); step(u)} without
//     leaving the SM: the state lives in shared memory (two copies, by step parity), the row owner integrates its own
//     agent with the action it has just computed and publishes it for the next step, so ONE block barrier separates two
//     steps.  Envs never interact, so no CTA ever waits for another: flock_rollout_* is one launch, and the tail of the
//     grid is paid once per rollout instead of once per step.  Every step still writes all of its outputs (values, the
//     dense adjacency, reward, obs) -- the stores of step t drain while step t+1 filters.
//   * steps = 1 is the per-step API (flock_step_*): state and action come from global memory, x_out / u_next go back.
//
//   per step and row i (flocking_relative.py):
//     stats     reward = -(var vx + var vy), sum_j v_j (alignment identity)                     :145-147, :204-208
//     filter    all N pairs, packed f32x2 sign test on positions relative to the env origin      :113-117
//     refine    the ~deg candidates in fp64: exact adjacency bit, 6 feature sums, gradient       :117-129, :214-226
//     epilogue  values, u = clip(.)/10, next state = integrate(x_i, u)                           :96-105, :209-211
//     rows      the warp's 32 rows of adj_mean from the exact bit masks, 16-byte streaming stores :120-122
//     kNN       Flocking-v0: 7 nearest by r2, re-ranked with the reference's own keys              flocking.py:20-24
#pragma once
#include "flock_env_kernel.cuh"

namespace flock {

struct GArgs {
    const double *x_in;       // [E][N][4] state before the first step
    const double *u_in;       // [E][N][2] action of the first step (NULL: no integration in the first step)
    double *x_out;            // state after the last step (nullable)
    double *u_out;            // controller() on the state after the last step (nullable)
    void *values, *network, *reward, *rewards_all, *obs;
    AuxOut aux;
    int E, N, G, steps;
    int n_main, G_tail;       // CTAs [0, n_main) own G envs each; the CTAs after them own G_tail (< G) envs each -- the last,
                              // partial round of the grid is spread over all SMs instead of leaving most of them idle
    int NP2;                  // pair slots per env in the packed fp32 position array
    int W;                    // 32-bit candidate words per row
    int centralized;
    float thr_a, thr_b, thr_c;   // per-row filter threshold: thr_a + thr_b * (max(|px|, |py|) + thr_c)
    float inv_n;
    double thrC, R2, Rm, dt, inv_ascale, ulim;
    StepVariant var;
    unsigned off_xs, off_pf, off_masks, off_stats, off_idg, off_obs;
    unsigned off_bar;                // the split step barrier: one mbarrier
    unsigned off_masks2, off_hint;   // Flocking-v0: kNN candidate mask (same layout as masks), last step's neighbours [T][8] u16
    unsigned xs_bytes, pf_bytes, stats_bytes;    // size of ONE copy of the parity-duplicated arrays
    int obs_direct;           // Flocking-v0: every lane stores its own observation row (7 x 16 bytes) instead of staging the
                              // warp's rows in shared memory for a fully coalesced copy (no staging area: more CTAs per SM)
};

// mask words are word-major with a fixed stride GMS (immediate offsets everywhere): a template parameter, 256 for CTAs of up
// to 256 threads, else 512
constexpr int GMAXT = 512;

// The CTA's agents in shared memory: positions and velocities in two arrays of 16-byte elements.  The refine loop loads
// RANDOM agents (each lane its own candidate): 16-byte elements spread over 8 bank groups, a 32-byte struct over 4 -- the
// step is bound by the L1 / shared-memory pipe, and those loads were half of its wavefronts.
struct EnvState {
    const double2 *p, *v;
    __device__ __forceinline__ double4 operator[](const int j) const {
        const double2 a = p[j], b = v[j];
        return make_double4(a.x, a.y, b.x, b.y);
    }
    __device__ __forceinline__ EnvState operator+(const int k) const { return EnvState{p + k, v + k}; }
};

// ---- the step barrier, split into arrive and wait (an mbarrier, one arrival per warp) -------------------------------------
// A warp ARRIVES as soon as its rows have published their next state (and it has finished reading the current one) and
// only WAITS at the top of the next step: the expansion of its 32 dense rows -- a quarter of the step's instructions,
// touching only the warp's own mask words -- runs while the slower warps of the CTA finish their refine loops, instead of
// every warp idling at a block barrier behind the slowest one (the largest stall reason of the round-2 capture).
__device__ __forceinline__ void gbar_init(unsigned long long *bar, const unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gbar_arrive(unsigned long long *bar) {       // release: the warp's shared-memory writes
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void gbar_wait(unsigned long long *bar, const unsigned parity) {
    // try_wait suspends the warp in hardware for up to the time hint before it answers "not yet": a waiting warp must not
    // spend issue slots polling (without the hint the polling loop was 12 % of the executed instructions of a step)
#ifndef FLOCK_GBAR_SLEEP
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "GBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra GBAR_DONE_%=;\n\t"
        "bra GBAR_WAIT_%=;\n"
        "GBAR_DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity), "r"(0x989680u) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "GBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GBAR_DONE_%=;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra GBAR_WAIT_%=;\n"
        "GBAR_DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity), "n"(FLOCK_GBAR_SLEEP) : "memory");
#endif
}

// ---- refine of one candidate for a row owned by ONE thread (this kernel) ------------------------------------------------
// Same arithmetic as refine_fast_pred2 (flock_env_kernel.cuh), but the row's mask words belong to this thread alone, so the
// exact adjacency word is built in a register (adjw |= low when adjacent) and stored when the walk leaves the word -- no
// shared-memory atomic per rejected candidate, no branch around it.
template <bool NEWTON1>
__device__ __forceinline__ void refine_group_pred(const double4 xi, const unsigned xaddr, const unsigned vaddr, const unsigned low,
                                                  const double R2, const double Rm, Acc &A, int &deg, unsigned &adjw) {
#define FLOCK_GROUP_REFINE_ASM(SECOND_NEWTON_STEP)                                                                  \
    asm volatile(                                                                                                   \
        "{\n\t"                                                                                                     \
        ".reg .pred pv, pa, pg;\n\t"                                                                                \
        ".reg .f64 xjx, xjy, xjz, xjw, dx, dy, r2, y, e, ya, y4, yg, dvx, dvy, c, ma;\n\t"                          \
        "setp.ne.u32 pv, %10, 0;\n\t"                                                                               \
        "ld.shared.v2.f64 {xjx, xjy}, [%11];\n\t"                                                                   \
        "ld.shared.v2.f64 {xjz, xjw}, [%12];\n\t"                                                                   \
        "sub.f64 dx, %13, xjx;\n\t"                                                                                 \
        "sub.f64 dy, %14, xjy;\n\t"                                                                                 \
        "mul.f64 r2, dy, dy;\n\t"                                                                                   \
        "fma.rn.f64 r2, dx, dx, r2;\n\t"                                                                            \
        "setp.lt.and.f64 pa, r2, %17, pv;\n\t"  /* adjacency: r2 < R^2 (squared radius) */                          \
        "setp.le.and.f64 pg, r2, %18, pv;\n\t"  /* gradient mask: r2 <= Rm */                                       \
        "rcp.approx.ftz.f64 y, r2;\n\t"                                                                             \
        "neg.f64 e, r2;\n\t"                                                                                        \
        "fma.rn.f64 c, e, y, 0d3FF0000000000000;\n\t"                                                               \
        "fma.rn.f64 y, y, c, y;\n\t"                                                                                \
        SECOND_NEWTON_STEP                                                                                          \
        "selp.f64 ya, y, 0d0000000000000000, pa;\n\t" /* the masks as selects: y * 1.0 and y * 0.0 (y finite) */    \
        "selp.f64 yg, y, 0d0000000000000000, pg;\n\t"                                                               \
        "mul.f64 y4, ya, ya;\n\t"                                                                                   \
        "sub.f64 dvx, %15, xjz;\n\t"                                                                                \
        "sub.f64 dvy, %16, xjw;\n\t"                                                                                \
        "selp.f64 ma, 0d3FF0000000000000, 0d0000000000000000, pa;\n\t"                                              \
        "fma.rn.f64 %0, dvx, ma, %0;\n\t"     /* + dvx * 1.0 / + dvx * 0.0: exact (dvx finite) */                    \
        "fma.rn.f64 %1, dx, y4, %1;\n\t"                                                                            \
        "fma.rn.f64 %2, dx, ya, %2;\n\t"                                                                            \
        "fma.rn.f64 %3, dvy, ma, %3;\n\t"                                                                           \
        "fma.rn.f64 %4, dy, y4, %4;\n\t"                                                                            \
        "fma.rn.f64 %5, dy, ya, %5;\n\t"                                                                            \
        "neg.f64 c, yg;\n\t"                                                                                        \
        "fma.rn.f64 c, c, yg, yg;\n\t" /* 1/r2 - 1/r4 */                                                            \
        "fma.rn.f64 %6, dx, c, %6;\n\t"                                                                             \
        "fma.rn.f64 %7, dy, c, %7;\n\t"                                                                             \
        "@pa add.s32 %8, %8, 1;\n\t"                                                                                \
        "@pa or.b32 %9, %9, %10;\n\t"                                                                               \
        "}"                                                                                                         \
        : "+d"(A.f0), "+d"(A.f1), "+d"(A.f2), "+d"(A.f3), "+d"(A.f4), "+d"(A.f5), "+d"(A.gx), "+d"(A.gy), "+r"(deg),  \
          "+r"(adjw)                                                                                               \
        : "r"(low), "r"(xaddr), "r"(vaddr), "d"(xi.x), "d"(xi.y), "d"(xi.z), "d"(xi.w), "d"(R2), "d"(Rm)            \
        : "memory")
    if constexpr (NEWTON1) { FLOCK_GROUP_REFINE_ASM(""); }
    else { FLOCK_GROUP_REFINE_ASM(FLOCK_NEWTON2); }
#undef FLOCK_GROUP_REFINE_ASM
}

// ---- in-warp expansion of 32 consecutive rows of adj_mean, N a compile-time constant ---------------------------------
// Row r of the warp lives N elements after row r-1: with N known every store address is dst + immediate, the mask word
// and 1/deg loads use immediate offsets too, so a row costs two LDS, the bit -> value selects and one 16-byte store.
template <typename TO, int NC, int STRIDE>
__device__ __forceinline__ void expand_rows_const(TO *wout, const unsigned *wm, const TO *idgs, const int nrows, const int lane) {
    constexpr int VMAX = 16 / (int)sizeof(TO);
    constexpr int V = (NC % VMAX == 0) ? VMAX : ((VMAX == 4 && NC % 2 == 0) ? 2 : ((VMAX == 2 && NC % 2 == 0) ? 2 : 1));
    constexpr int nv = NC / V;                          // groups per row
    if constexpr (nv <= 32) {
        constexpr int rpi = 32 / nv;                    // rows per iteration
        const int sr = lane / nv;                       // this lane's sub-row within an iteration
        const int cv = lane - sr * nv;
        const bool on = sr < rpi;
        const int c = on ? cv * V : 0;
        const int sh = c & 31;
        const unsigned *mp = wm + (c >> 5) * STRIDE + (on ? sr : 0);
        const TO *ip = idgs + (on ? sr : 0);
        TO *dst = wout + (on ? sr : 0) * NC + c;
        const int last = nrows - sr;                    // this lane stores while r < last
        if (nrows == 32 && rpi == 1) {
            // four rows per shared-memory access: the mask words of consecutive rows (same word index) and their 1/deg are
            // contiguous, so one LDS.128 each feeds four stores (the step is bound by the L1 / shared-memory pipe)
#pragma unroll
            for (int r4 = 0; r4 < 32; r4 += 4) {
                const uint4 mw = *reinterpret_cast<const uint4 *>(mp + r4);
                TO v0, v1, v2, v3;
                if constexpr (sizeof(TO) == 4) {
                    const float4 ig = *reinterpret_cast<const float4 *>(ip + r4);
                    v0 = ig.x; v1 = ig.y; v2 = ig.z; v3 = ig.w;
                } else {
                    const double2 ia = *reinterpret_cast<const double2 *>(ip + r4), ib = *reinterpret_cast<const double2 *>(ip + r4 + 2);
                    v0 = ia.x; v1 = ia.y; v2 = ib.x; v3 = ib.y;
                }
                store_group<TO, V>(dst + (r4 + 0) * NC, mw.x >> sh, v0, on);
                store_group<TO, V>(dst + (r4 + 1) * NC, mw.y >> sh, v1, on);
                store_group<TO, V>(dst + (r4 + 2) * NC, mw.z >> sh, v2, on);
                store_group<TO, V>(dst + (r4 + 3) * NC, mw.w >> sh, v3, on);
            }
        } else if (nrows == 32) {
#pragma unroll
            for (int r = 0; r < 32; r += rpi) {
                const bool go = on && (rpi == 1 || r < last);
                store_group<TO, V>(dst + r * NC, mp[r] >> sh, ip[r], go);
            }
        } else {
            for (int r = 0; r < nrows; r += rpi) {
                const bool go = on && r < last;
                store_group<TO, V>(dst + r * NC, go ? (mp[r] >> sh) : 0u, go ? ip[r] : TO(0), go);
            }
        }
    } else {
        const int sh = (lane * V) & 31;
        const unsigned *mp0 = wm + ((lane * V) >> 5) * STRIDE;
        for (int r = 0; r < nrows; ++r) {
            const TO v = idgs[r];
            const unsigned *mp = mp0 + r;
            TO *dst = wout + (size_t)r * NC + lane * V;
#pragma unroll 4
            for (int cv = lane; cv < nv; cv += 32) {
                store_group<TO, V>(dst, *mp >> sh, v, true);
                mp += V * STRIDE; dst += 32 * V;
            }
        }
    }
}

// generic N: flock_env_kernel's expand_rows with this kernel's mask stride
template <typename TO, int V, int GMS>
__device__ __forceinline__ void expand_rows_g(TO *wout, const unsigned *wm, const TO *idgs, const int N, const int nrows,
                                              const int lane) {
    const int nv = N / V;
    const int passes = (nv + 31) >> 5;
    const int rpi = (nv >= 32) ? 1 : 32 / nv;
    const int sr = (nv >= 32) ? 0 : lane / nv;
    const int cv0 = (nv >= 32) ? lane : lane - sr * nv;
    const bool lane_on = (nv >= 32) || (sr < rpi);
    if (passes == 1) {
        const bool on = lane_on && cv0 < nv;
        const int c = on ? cv0 * V : 0;
        const int sh = c & 31;
        const unsigned *mp = wm + (c >> 5) * GMS + (on ? sr : 0);
        const TO *ip = idgs + (on ? sr : 0);
        TO *dst = wout + (size_t)(on ? sr : 0) * N + c;
        const size_t dstep = (size_t)rpi * N;
        const int last = nrows - sr;
        for (int r = 0; r < nrows; r += rpi) {
            const bool go = on && r < last;
            store_group<TO, V>(dst, go ? (mp[r] >> sh) : 0u, go ? ip[r] : TO(0), go);
            dst += dstep;
        }
    } else {
        const int sh = (lane * V) & 31;
        const unsigned *mp0 = wm + ((lane * V) >> 5) * GMS;
        for (int r = 0; r < nrows; ++r) {
            const TO v = idgs[r];
            const unsigned *mp = mp0 + r;
            TO *dst = wout + (size_t)r * N + lane * V;
#pragma unroll 4
            for (int cv = lane; cv < nv; cv += 32) {
                store_group<TO, V>(dst, *mp >> sh, v, true);
                mp += V * GMS; dst += 32 * V;
            }
        }
    }
}

template <typename TO, bool KNN, int NC, int GMS>
#ifndef FLOCK_GROUP_MINB
#define FLOCK_GROUP_MINB 2
#endif
__global__ void __launch_bounds__(GMAXT, FLOCK_GROUP_MINB) flock_group_kernel(const GArgs a) {
    constexpr bool PA32 = std::is_same<TO, float>::value;
    using TK = typename std::conditional<PA32, float, double>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = NC ? NC : a.N;
    const int NP2 = NC ? (NC + 1) / 2 : a.NP2;
    const int tid = threadIdx.x, lane = tid & 31;
    const int bid = blockIdx.x;
    const int e0 = (bid < a.n_main) ? bid * a.G : a.n_main * a.G + (bid - a.n_main) * a.G_tail;
    const int genv = min((bid < a.n_main) ? a.G : a.G_tail, a.E - e0);
    const int nrows = genv * N;                            // rows (agents) of this CTA
    const bool valid = tid < nrows;
    const int r = valid ? tid : 0;                         // idle threads shadow row 0 and store nothing
    int g;
    if constexpr (NC != 0) g = r / NC;
    else g = (int)(((float)r + 0.5f) * a.inv_n);           // exact: r < 2^20
    const int i = r - g * N;
    const int env = e0 + g;
    const size_t grow = (size_t)e0 * N + r;                // global row (agent) index
    const bool cen = a.centralized != 0;
    const int steps = a.steps;
    unsigned *mrow = reinterpret_cast<unsigned *>(smem + a.off_masks) + tid;     // word w of this thread at mrow[w * GMS]
    TO *idg_s = reinterpret_cast<TO *>(smem + a.off_idg);
    TO *net = reinterpret_cast<TO *>(a.network);
    const double R2r = a.R2, Rmr = a.Rm;
    unsigned long long &step_bar = *reinterpret_cast<unsigned long long *>(smem + a.off_bar);   // (arrive: end of a step's epilogue)
    if (tid == 0) gbar_init(&step_bar, blockDim.x >> 5);

    // ---- step-0 staging: state and action from global memory ---------------------------------------------------------------
    for (int q = tid; q < 2 * genv; q += blockDim.x)       // ready counters of both parities
        reinterpret_cast<double *>(smem + a.off_stats + (q >= genv ? a.stats_bytes : 0u))[(q % genv) * STATS_PER_ENV + 2] = 0.0;
    pdl_wait();                                            // (programmatic dependent launch: the previous step's stores are visible)
    if constexpr (KNN) {
        // last step's neighbours, if the caller kept them (FlockAux.knn_hint); else "self": a zero radius -> the first step
        // searches all N agents
        uint4 hv = make_uint4((unsigned)i | ((unsigned)i << 16), (unsigned)i | ((unsigned)i << 16), (unsigned)i | ((unsigned)i << 16), (unsigned)i);
        if (a.aux.knn_hint != nullptr && valid) hv = reinterpret_cast<const uint4 *>(a.aux.knn_hint)[grow];
        *reinterpret_cast<uint4 *>(smem + a.off_hint + (size_t)tid * 16) = hv;
    }
    if (valid) {
        double4 x = ld_state(a.x_in + grow * 4);
        const double2 ref = __ldg(reinterpret_cast<const double2 *>(a.x_in + (size_t)env * N * 4));   // env origin: agent 0 before the step
        if (a.u_in) {
            const double2 u = __ldg(reinterpret_cast<const double2 *>(a.u_in) + grow);
            const bool plain = (a.var.u_clip == 0.0) && (a.var.n_leaders == 0) && (a.var.state_scale == 1.0);
            x = plain ? integrate_exact(x, u, a.dt, a.var.step_scale) : integrate_variant(x, u, i, a.dt, a.var);
        }
        reinterpret_cast<double2 *>(smem + a.off_xs)[r] = make_double2(x.x, x.y);
        reinterpret_cast<double2 *>(smem + a.off_xs + a.xs_bytes / 2)[r] = make_double2(x.z, x.w);
        if constexpr (PA32) {
            float *p = reinterpret_cast<float *>(reinterpret_cast<float4 *>(smem + a.off_pf) + g * NP2 + (i >> 1));
            p[i & 1] = (float)(x.x - ref.x); p[2 + (i & 1)] = (float)(x.y - ref.y);
            if ((N & 1) && i == N - 1) { p[1] = CUDART_INF_F; p[3] = CUDART_INF_F; }     // odd N: pad slot at infinity
        }
    }

#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        const unsigned par = (unsigned)(t & 1);
        const bool last_step = (t == steps - 1);
        const EnvState xs{reinterpret_cast<const double2 *>(smem + a.off_xs + par * a.xs_bytes),
                          reinterpret_cast<const double2 *>(smem + a.off_xs + par * a.xs_bytes + a.xs_bytes / 2)};
        const float4 *pf = reinterpret_cast<const float4 *>(smem + a.off_pf + par * a.pf_bytes);
        double *stats = reinterpret_cast<double *>(smem + a.off_stats + par * a.stats_bytes);
        // the state of step t is complete, and nobody still reads the copy this step overwrites (every warp arrived after
        // its last read of step t-1's state)
        if (t == 0) __syncthreads();
        else gbar_wait(&step_bar, (unsigned)((t - 1) & 1));

        // ---- per-env velocity sums and reward: one warp per env, published through a step counter -----------------------
        {
            const int warp = tid >> 5, nwarps = blockDim.x >> 5;
            for (int ge = warp; ge < genv; ge += nwarps) {
                const EnvState xe = xs + ge * N;
                const double rx = xe[0].z, ry = xe[0].w;    // sums about agent 0's velocity (a member of the flock)
                double sx = 0, sy = 0, qx = 0, qy = 0;
                for (int j = lane; j < N; j += 32) {
                    const double ex = xe[j].z - rx, ey = xe[j].w - ry;
                    sx += ex; sy += ey; qx = fma(ex, ex, qx); qy = fma(ey, ey, qy);
                }
                sx = warp_sum(sx); sy = warp_sum(sy);
                const double q = warp_sum(qx + qy);
                if (lane == 0) {
                    const double invN = 1.0 / (double)N;
                    const TO rew = (TO)(-(q - (sx * sx + sy * sy) * invN) * invN);          // -(var vx + var vy)
                    if (a.rewards_all) reinterpret_cast<TO *>(a.rewards_all)[(size_t)t * a.E + e0 + ge] = rew;
                    if (a.reward && (last_step || !a.rewards_all)) reinterpret_cast<TO *>(a.reward)[e0 + ge] = rew;
                    double *st = stats + ge * STATS_PER_ENV;
                    st[0] = fma((double)N, rx, sx); st[1] = fma((double)N, ry, sy);          // sum_j v_j
                    __threadfence_block();
                    reinterpret_cast<volatile double *>(st)[2] = (double)(t + 1);            // ready
                }
            }
        }

        if (tid - lane >= nrows) {                         // a warp without rows (a CTA of the tail round) only keeps the barrier
            if (!last_step && lane == 0) gbar_arrive(&step_bar);
            continue;
        }
        const EnvState xe = xs + g * N;
        const double4 xi = xe[i];
        if (last_step && a.x_out && valid) st_state(a.x_out + grow * 4, xi);

        // ---- phase A: filter over all N pairs; candidate bits in shared memory, bit k of the row <-> agent k ------------------
        // Flocking-v0 (KNN): a second mask collects the agents within tau, an UPPER BOUND on the row's 7th-nearest distance
        // derived from last step's neighbours (the 7 nearest of this step are among the agents at least as close as the
        // farthest of ANY 7 agents); neighbourhoods change slowly, so that mask holds 7-10 agents instead of N.  The hint
        // only sets the radius: stale, missing or garbage hints cost time (a mask with < 7 agents falls back to all N),
        // never correctness.
        // cnt: candidates of the row WITHOUT the diagonal (r2 = 0 is always a candidate; +inf in the reference, :115 -- its bit is
        // cleared before the word is stored); idle: iterations the refine walk spends on empty words (every empty word after
        // word 0 and before the row's last non-empty word costs it one)
        int cnt = 0, cntk = 0, idle = 0, empties = 0;
        const int diag_w = i >> 5;
        const unsigned diag_m = ~(1u << (i & 31));
        auto count_word = [&](const int w, const unsigned bits) {
            cnt += __popc(bits);
            idle = bits ? empties : idle;
            empties += (w >= 1 && bits == 0u) ? 1 : 0;
        };
        int nw;
        [[maybe_unused]] unsigned *mkrow = reinterpret_cast<unsigned *>(smem + a.off_masks2) + tid;
        [[maybe_unused]] unsigned short *hint = reinterpret_cast<unsigned short *>(smem + a.off_hint) + tid * 8;
        [[maybe_unused]] double tau = 0.0;
        [[maybe_unused]] int hj[KNN_K];                      // last step's neighbours and their distances NOW (the reference's
        [[maybe_unused]] double hk[KNN_K];                   // unfused form, the ordering key of the observation)
        if constexpr (KNN) {
            const uint4 hv4 = *reinterpret_cast<const uint4 *>(hint);      // 8 x u16 in one 16-byte load (no bank conflicts)
            const unsigned hw[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
#pragma unroll
            for (int m = 0; m < KNN_K; ++m) {
                hj[m] = min((int)((hw[m >> 1] >> ((m & 1) * 16)) & 0xffffu), N - 1);
                const double4 xh = xe[hj[m]];
                const double dx = __dsub_rn(xi.x, xh.x), dy = __dsub_rn(xi.y, xh.y);
                hk[m] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                tau = fmax(tau, hk[m]);
            }
        }
        if constexpr (PA32) {
            const int nslots = NP2;
            nw = (2 * nslots + 31) >> 5;
            const float4 *pe = pf + g * NP2;
            const float *pif = reinterpret_cast<const float *>(pe + (i >> 1));
            const float pxi = pif[i & 1], pyi = pif[2 + (i & 1)];
            const float2 nx = make_float2(-pxi, -pxi), ny = make_float2(-pyi, -pyi);
            // conservative per-row thresholds (cover the fp32 rounding of both positions and of the fused form)
            const float pm = fmaxf(fabsf(pxi), fabsf(pyi));
            const float thrA = fmaf(a.thr_b, pm + a.thr_c, a.thr_a);
            const float2 nthr2 = make_float2(-thrA, -thrA);
            [[maybe_unused]] float2 nthrk2 = nthr2;
            if constexpr (KNN) {
                const float tk = __double2float_ru(tau), rk = sqrtf(tk);
                const float thrK = tk * (1.0f + 3e-6f) + 16.0f * rk * (pm + 2.0f * rk) * 5.9604645e-8f;
                nthrk2 = make_float2(-thrK, -thrK);
            }
            // s = dx^2 + (dy^2 - thr): the candidate test is the SIGN of s; one funnel shift per pair moves it into the word
            // (first pair ends up highest: one BREV per word)
            auto word = [&](const int w, const int tend, auto FULL) {
                unsigned bits = 0;
                [[maybe_unused]] unsigned bitsk = 0;
                const int tbase = w * 16;
                auto body = [&](const int tt) {
                    const float4 q = pe[tbase + tt];
                    const float2 dx = __fadd2_rn(make_float2(q.x, q.y), nx);
                    const float2 dy = __fadd2_rn(make_float2(q.z, q.w), ny);
                    const float2 s = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, nthr2));
                    bits = __funnelshift_l(__float_as_uint(s.x), bits, 1);
                    bits = __funnelshift_l(__float_as_uint(s.y), bits, 1);
                    if constexpr (KNN) {
                        const float2 sk = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, nthrk2));
                        bitsk = __funnelshift_l(__float_as_uint(sk.x), bitsk, 1);
                        bitsk = __funnelshift_l(__float_as_uint(sk.y), bitsk, 1);
                    }
                };
                constexpr int CT = decltype(FULL)::value;        // > 0: compile-time slot count of this word
                if constexpr (CT > 0) {
#pragma unroll
                    for (int tt = 0; tt < CT; ++tt) body(tt);
                    bits = __brev(bits) >> (32 - 2 * CT);
                    if constexpr (KNN) bitsk = __brev(bitsk) >> (32 - 2 * CT);
                } else {
                    for (int tt = 0; tt < tend; ++tt) body(tt);
                    bits = __brev(bits) >> (32 - 2 * tend);          // 0 < tend <= 16
                    if constexpr (KNN) bitsk = __brev(bitsk) >> (32 - 2 * tend);
                }
                bits &= (w == diag_w) ? diag_m : 0xffffffffu;
                mrow[w * GMS] = bits;
                count_word(w, bits);
                if constexpr (KNN) { mkrow[w * GMS] = bitsk; cntk += __popc(bitsk); }
            };
            if constexpr (NC != 0) {
                constexpr int NS = (NC + 1) / 2, FW = NS / 16, TAIL = NS - FW * 16;
#pragma unroll
                for (int w = 0; w < FW; ++w) word(w, 16, std::integral_constant<int, 16>{});
                if constexpr (TAIL > 0) word(FW, TAIL, std::integral_constant<int, TAIL>{});
            } else {
                for (int w = 0; w * 16 < nslots; ++w) {
                    const int tend = min(16, nslots - w * 16);
                    if (tend == 16) word(w, 16, std::integral_constant<int, 16>{});
                    else word(w, tend, std::integral_constant<int, 0>{});
                }
            }
        } else {
            nw = (N + 31) >> 5;
            const double thrC = a.thrC;
            [[maybe_unused]] const double thrK = tau * (1.0 + 1e-12);   // (fused r2 here, unfused keys in tau)
            for (int w = 0; w * 32 < N; ++w) {
                unsigned bits = 0;
                [[maybe_unused]] unsigned bitsk = 0;
                const int kbase = w * 32;
                const int kend = min(32, N - kbase);
                auto body = [&](const int kk) {
                    const int j = kbase + kk;
                    const double4 xj = xe[j];
                    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                    const double r2 = fma(dx, dx, dy * dy);
                    bits |= (r2 <= thrC) ? (1u << kk) : 0u;
                    if constexpr (KNN) bitsk |= (r2 <= thrK) ? (1u << kk) : 0u;
                };
                if (kend == 32) {
#pragma unroll 8
                    for (int kk = 0; kk < 32; ++kk) body(kk);
                } else {
                    for (int kk = 0; kk < kend; ++kk) body(kk);
                }
                bits &= (w == diag_w) ? diag_m : 0xffffffffu;
                mrow[w * GMS] = bits;
                count_word(w, bits);
                if constexpr (KNN) { mkrow[w * GMS] = bitsk; cntk += __popc(bitsk); }
            }
        }

        // ---- kNN observation: obs[i, 4m:4m+4] = x_i - x_nbr(m), staged per warp and stored coalesced      flocking.py:20-24
        if constexpr (KNN) {
            constexpr bool KPACK = PA32;
            // Selection keys.  fp32 mode: the fp32 bits of r2 (non-negative, so they order as unsigned integers) with the low
            // ceil(log2 N) mantissa bits replaced by the agent index -- a branch-free min/max network; fp64 mode: fused fp64 r2.
            // Both keep TOPK = 9 survivors (the row itself + 7 + one slack slot); knn_rerank orders them exactly.
            TopK<double> top;
            unsigned kp[TOPK];
            unsigned kidx = 1;
            if constexpr (KPACK) {
                while ((int)kidx < N) kidx <<= 1;
                kidx -= 1;
            }
            auto consider = [&](const int j) {
                if constexpr (KPACK) {
                    const float *pj = reinterpret_cast<const float *>(pf + g * NP2 + (j >> 1));
                    const float *pif = reinterpret_cast<const float *>(pf + g * NP2 + (i >> 1));
                    const float dx = pj[j & 1] - pif[i & 1], dy = pj[2 + (j & 1)] - pif[2 + (i & 1)];
                    unsigned k = (__float_as_uint(fmaf(dx, dx, dy * dy)) & ~kidx) | (unsigned)j;
#pragma unroll
                    for (int m = 0; m < TOPK; ++m) { const unsigned lo = min(kp[m], k); k = max(kp[m], k); kp[m] = lo; }
                } else {
                    const double4 xj = xe[j];
                    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                    top.insert(fma(dx, dx, dy * dy), j);
                }
            };
            // The mask holds every agent with r2 <= tau (the filter has no false negatives; the row itself is skipped).  If the 7th neighbour found in it is
            // itself within tau, no agent outside the mask can be closer: done.  Otherwise (no usable hint: first step,
            // teleported agents, N <= 8) the warp searches all N agents.
            int nb[TOPK];
            // Fast path (nearly every row of an evolving flock): the mask holds exactly the row itself, last step's 7
            // neighbours and at most two newcomers.  Then those <= 9 agents are the survivors -- no selection, only the exact
            // ordering.  The check is a count: clearing the row's and the 7 hints' bits must remove exactly 8 set bits (so the
            // hints are distinct, in range and not the row itself) and leave 0..2.
            bool fast = false;
            if (N > KNN_K + 1) {
                auto mkbit = [&](const int j, const bool set) {
                    unsigned *wp = mkrow + (j >> 5) * GMS;
                    *wp = set ? (*wp | (1u << (j & 31))) : (*wp & ~(1u << (j & 31)));
                };
                mkbit(i, false);
#pragma unroll
                for (int m = 0; m < KNN_K; ++m) mkbit(hj[m], false);
                int rem = 0, e0n = i, e1n = i;               // unused newcomer slots hold the row itself (sorts last)
                for (int w = 0; w < nw; ++w) {
                    const unsigned b = mkrow[w * GMS];
                    const unsigned b2 = b & (b - 1);
                    const int j0 = w * 32 + __ffs(b) - 1, j1 = w * 32 + __ffs(b2) - 1;
                    if (rem == 0) { e0n = b ? j0 : e0n; e1n = b2 ? j1 : e1n; }
                    else if (rem == 1) { e1n = b ? j0 : e1n; }
                    rem += __popc(b);
                }
                fast = __all_sync(__activemask(), !valid || (rem == cntk - (KNN_K + 1) && rem <= TOPK - KNN_K));
                if (fast) {
                    double key[TOPK];
#pragma unroll
                    for (int m = 0; m < KNN_K; ++m) { nb[m] = hj[m]; key[m] = hk[m]; }
                    nb[KNN_K] = e0n; nb[KNN_K + 1] = e1n;
#pragma unroll
                    for (int m = KNN_K; m < TOPK; ++m) {
                        const double4 xj = xe[nb[m]];
                        const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
                        key[m] = (nb[m] == i) ? CUDART_INF : __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    }
                    // last step's order with the newcomers appended is nearly sorted: odd-even transposition passes until the
                    // warp is in order (TOPK passes always suffice; the bound also ends rows with non-finite keys)
                    auto cx = [&](const int q) {
                        const bool sw = key[q + 1] < key[q] || (key[q + 1] == key[q] && nb[q + 1] < nb[q]);
                        const double ka = sw ? key[q + 1] : key[q], kb = sw ? key[q] : key[q + 1];
                        const int ia = sw ? nb[q + 1] : nb[q], ib = sw ? nb[q] : nb[q + 1];
                        key[q] = ka; key[q + 1] = kb; nb[q] = ia; nb[q + 1] = ib;
                    };
#pragma unroll 1
                    for (int pass = 0; pass < (TOPK + 1) / 2; ++pass) {
                        bool sorted = true;
#pragma unroll
                        for (int m = 0; m + 1 < TOPK; ++m)
                            sorted = sorted && (key[m] < key[m + 1] || (key[m] == key[m + 1] && nb[m] <= nb[m + 1]));
                        if (__all_sync(__activemask(), sorted)) break;
#pragma unroll
                        for (int q = 0; q + 1 < TOPK; q += 2) cx(q);
#pragma unroll
                        for (int q = 1; q + 1 < TOPK; q += 2) cx(q);
                    }
                } else {                                     // put the bits back: the search below scans the whole mask
                    mkbit(i, true);
#pragma unroll
                    for (int m = 0; m < KNN_K; ++m) mkbit(hj[m], true);
                    cntk = 0;
                    for (int w = 0; w < nw; ++w) cntk += __popc(mkrow[w * GMS]);
                }
            }
            bool use_mask = __all_sync(__activemask(), !valid || cntk >= KNN_K + 1) && N > KNN_K + 1;
#pragma unroll 1
            for (int attempt = 0; attempt < 2 && !fast; ++attempt) {
                if constexpr (KPACK) {
#pragma unroll
                    for (int m = 0; m < TOPK; ++m) kp[m] = (0xFFFFFFFFu & ~kidx) | (unsigned)i;
                } else {
                    top.init(i);
                }
                if (use_mask) {
                    // one flat loop over the lane's mask bits (the advance to the next word is part of the body, as in the
                    // refine loop): a warp iterates max-over-lanes(candidates) times, not the sum of the per-word maxima
                    unsigned b = mkrow[0];
                    int w = 0, left = valid ? cntk : 0;
                    while (left > 0) {
                        if (b == 0) { ++w; b = mkrow[w * GMS]; }
                        if (b != 0) {
                            const int j = w * 32 + __ffs(b) - 1;
                            b &= b - 1;
                            --left;
                            if (j != i) consider(j);
                        }
                    }
                } else {
                    for (int j = 0; j < N; ++j) if (j != i) consider(j);
                }
                __syncwarp();
#pragma unroll
                for (int m = 0; m < TOPK; ++m) {
                    if constexpr (KPACK) nb[m] = (int)(kp[m] & kidx);
                    else nb[m] = top.id[m];
                }
                const double key7 = knn_rerank(xi, i, xe, nb);
                if (!use_mask || __all_sync(__activemask(), !valid || key7 <= tau)) break;
                use_mask = false;
            }
            {                                                // next step's radius comes from these
                uint4 hv;
                hv.x = (unsigned)nb[0] | ((unsigned)nb[1] << 16); hv.y = (unsigned)nb[2] | ((unsigned)nb[3] << 16);
                hv.z = (unsigned)nb[4] | ((unsigned)nb[5] << 16); hv.w = (unsigned)nb[6];
                *reinterpret_cast<uint4 *>(hint) = hv;
                if (a.aux.knn_hint != nullptr && last_step && valid) reinterpret_cast<uint4 *>(a.aux.knn_hint)[grow] = hv;
            }
            if (a.obs && a.obs_direct) {
                if (valid) {
                    TO *o = reinterpret_cast<TO *>(a.obs) + grow * (NXS * KNN_K);
#pragma unroll
                    for (int m = 0; m < KNN_K; ++m) {
                        const double4 xn = xe[nb[m]];
                        if constexpr (PA32) {
                            __stcs(reinterpret_cast<float4 *>(o) + m, make_float4((float)(xi.x - xn.x), (float)(xi.y - xn.y),
                                                                                  (float)(xi.z - xn.z), (float)(xi.w - xn.w)));
                        } else {
                            __stcs(reinterpret_cast<double2 *>(o) + 2 * m, make_double2(xi.x - xn.x, xi.y - xn.y));
                            __stcs(reinterpret_cast<double2 *>(o) + 2 * m + 1, make_double2(xi.z - xn.z, xi.w - xn.w));
                        }
                    }
                }
            } else if (a.obs) {
                TO *const obs_w = reinterpret_cast<TO *>(smem + a.off_obs) + (size_t)(tid - lane) * (NXS * KNN_K);
                TO *o = obs_w + lane * (NXS * KNN_K);
#pragma unroll
                for (int m = 0; m < KNN_K; ++m) {            // the diagonal and unused slots sort last: a zero block (N <= 7)
                    const double4 xn = xe[nb[m]];
                    o[m * 4 + 0] = (TO)(xi.x - xn.x); o[m * 4 + 1] = (TO)(xi.y - xn.y);
                    o[m * 4 + 2] = (TO)(xi.z - xn.z); o[m * 4 + 3] = (TO)(xi.w - xn.w);
                }
                __syncwarp();
                const int wr0 = tid - lane;                  // first row of this warp within the CTA
                const int nrw = min(32, nrows - wr0);
                if (nrw > 0) {
                    const int nvec = nrw * (NXS * KNN_K) * (int)sizeof(TO) / 16;
                    const float4 *src = reinterpret_cast<const float4 *>(obs_w);
                    float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<TO *>(a.obs) + ((size_t)e0 * N + wr0) * (NXS * KNN_K));
                    for (int v = lane; v < nvec; v += 32) __stcs(dst + v, src[v]);
                }
                __syncwarp();
            }
        }

        // ---- phase B: refine the candidates in fp64, ascending j; the words end up holding the exact adjacency bits -------
        Acc A;
        A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = 0.0;
        int deg = 0;
        {
            // The walk: `cand` holds the unvisited candidates of the current word, `xbase` the shared-memory address of that
            // word's first agent, `adjw` the exact adjacency bits found in it (stored when the walk leaves the word: the row's
            // words belong to this thread alone).  The whole warp runs max-over-lanes(cnt + idle) iterations of straight-line
            // code -- a counted, warp-uniform loop: no vote, no divergent exit; a lane that has run out idles on the
            // word's first agent with everything selected off (that agent may be the row itself: r2 = 0 gives a non-finite
            // reciprocal, which only ever feeds selects).
            const unsigned xe_s = (unsigned)__cvta_generic_to_shared(xe.p);
            const unsigned voff = a.xs_bytes / 2;             // the velocity array follows the position array
            unsigned waddr = (unsigned)__cvta_generic_to_shared(mrow);
            const unsigned wlast = waddr + (unsigned)(nw - 1) * (GMS * 4);
            unsigned cand = valid ? mrow[0] : 0u, nxt = mrow[GMS];   // (the mask region has one spare word row)
            unsigned adjw = 0;
            unsigned xbase = xe_s;
            const int niter = __reduce_max_sync(0xffffffffu, valid ? cnt + idle : 0);
            auto walk_step = [&]() {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "setp.eq.u32 p, %0, 0;\n\t"
                             "setp.lt.and.u32 p, %2, %5, p;\n\t"   // (the walk ends in the row's last word)
                             "@p st.shared.u32 [%2], %4;\n\t"
                             "@p mov.u32 %4, 0;\n\t"
                             "@p mov.u32 %0, %1;\n\t"
                             "@p add.u32 %2, %2, %6;\n\t"
                             "@p add.u32 %3, %3, 512;\n\t"
                             "@p ld.shared.u32 %1, [%2+%6];\n\t}"
                             : "+r"(cand), "+r"(nxt), "+r"(waddr), "+r"(xbase), "+r"(adjw) : "r"(wlast), "n"(GMS * 4) : "memory");
                unsigned low, xaddr;        // lowest candidate of the word and its agent (a lane without one: the word's first agent)
                asm volatile("{\n\t.reg .b32 t;\n\t"
                             "neg.s32 t, %0;\n\t"
                             "and.b32 %1, %0, t;\n\t"
                             "xor.b32 %0, %0, %1;\n\t"
                             "or.b32 t, %1, 1;\n\t"
                             "bfind.u32 t, t;\n\t"
                             "mad.lo.u32 %2, t, 16, %3;\n\t}"
                             : "+r"(cand), "=r"(low), "=r"(xaddr) : "r"(xbase));
                refine_group_pred<PA32>(xi, xaddr, xaddr + voff, low, R2r, Rmr, A, deg, adjw);
            };
            // two candidates per iteration, in order: the loads and reciprocals of the pair overlap (the sums stay sequential, so
            // the result does not change), an odd count ends with a single step
            // (Flocking-v0 keeps one per iteration: its register budget is spent on the kNN state, the pair only adds spills)
            int it = 0;
            if constexpr (!KNN) {
#pragma unroll 1
                for (; it + 2 <= niter; it += 2) { walk_step(); walk_step(); }
            }
#pragma unroll 1
            for (; it < niter; ++it) walk_step();
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(waddr), "r"(adjw) : "memory");    // the last word visited
        }
        __syncwarp();

        // ---- row epilogue: features, action, next state ----------------------------------------------------------------------
        const TO idg = inv_deg(deg, TO(0));
        idg_s[tid] = idg;
        double ux, uy;
        {
            double sdx = A.f0, sdy = A.f3;                   // decentralized: the masked sums (:205-206)
            if (cen) {                                       // centralized: sum_j (v_i - v_j) = N*v_i - sum_j v_j, about agent 0
                const volatile double *st = stats + g * STATS_PER_ENV;
                const double want = (double)(t + 1);
                while (st[2] < want) { }                     // published right after the barrier by the stats warp
                __threadfence_block();
                const double rx = xe[0].z, ry = xe[0].w;
                sdx = fma((double)N, xi.z - rx, -(st[0] - (double)N * rx));
                sdy = fma((double)N, xi.w - ry, -(st[1] - (double)N * ry));
            }
            const double ia = a.inv_ascale, ul = a.ulim;
            const double rawx = -2.0 * A.gx - sdx, rawy = -sdy - 2.0 * A.gy;
            ux = fmin(fmax(rawx * ia, -ul), ul);
            uy = fmin(fmax(rawy * ia, -ul), ul);
            if (a.aux.nonfinite != nullptr) {                // fmin/fmax swallow NaN: test the unclamped sums
                const double chk = ((A.f1 + A.f2) + (A.f4 + A.f5)) + (rawx + rawy);
                if (valid && !(fabs(chk) < CUDART_INF)) a.aux.nonfinite[env] = 1;
            }
        }
        if (valid) {
            if (a.values) {
                TO *v = reinterpret_cast<TO *>(a.values) + grow * NFE;
                if constexpr (PA32) {
                    float2 *v2 = reinterpret_cast<float2 *>(v);
                    v2[0] = make_float2((float)A.f0, (float)A.f1); v2[1] = make_float2((float)A.f2, (float)A.f3);
                    v2[2] = make_float2((float)A.f4, (float)A.f5);
                } else {
                    double2 *v2 = reinterpret_cast<double2 *>(v);
                    v2[0] = make_double2(A.f0, A.f1); v2[1] = make_double2(A.f2, A.f3); v2[2] = make_double2(A.f4, A.f5);
                }
            }
            if (!last_step) {
                // the row owner advances its own agent: next state and packed position into the other parity's copy
                const double2 u = make_double2(ux, uy);
                const bool plain = (a.var.u_clip == 0.0) && (a.var.n_leaders == 0) && (a.var.state_scale == 1.0);
                const double4 xn = plain ? integrate_exact(xi, u, a.dt, a.var.step_scale) : integrate_variant(xi, u, i, a.dt, a.var);
                reinterpret_cast<double2 *>(smem + a.off_xs + (par ^ 1u) * a.xs_bytes)[r] = make_double2(xn.x, xn.y);
                reinterpret_cast<double2 *>(smem + a.off_xs + (par ^ 1u) * a.xs_bytes + a.xs_bytes / 2)[r] = make_double2(xn.z, xn.w);
                if constexpr (PA32) {
                    const double ox = xe[0].x, oy = xe[0].y;   // env origin: agent 0 before the step
                    float *p = reinterpret_cast<float *>(reinterpret_cast<float4 *>(smem + a.off_pf + (par ^ 1u) * a.pf_bytes) + g * NP2 + (i >> 1));
                    p[i & 1] = (float)(xn.x - ox); p[2 + (i & 1)] = (float)(xn.y - oy);
                    if ((N & 1) && i == N - 1) { p[1] = CUDART_INF_F; p[3] = CUDART_INF_F; }
                }
            } else if (a.u_out) {
                reinterpret_cast<double2 *>(a.u_out)[grow] = make_double2(ux, uy);
            }
        }

        // ---- optional outputs (FlockAux): exact adjacency bits, degree, get_stats() ------------------------------------------------
        if (a.aux.adj_bits != nullptr || a.aux.degree != nullptr || a.aux.min_dist != nullptr || a.aux.vel_diff != nullptr) {
            if (valid) {
                if (a.aux.degree) a.aux.degree[grow] = deg;
                if (a.aux.adj_bits) for (int w = 0; w < a.aux.WB; ++w) a.aux.adj_bits[grow * a.aux.WB + w] = mrow[w * GMS];
                if (a.aux.min_dist) {
                    // min_j sqrt(r2) (flocking_relative.py:141-142): every neighbour is closer than every non-neighbour,
                    // so scan the adjacency bits; a row without neighbours scans all agents
                    double m2 = CUDART_INF;
                    if (deg > 0) {
                        for (int w = 0; w < nw; ++w) {
                            unsigned b = mrow[w * GMS];
                            while (b) {
                                const int j = w * 32 + __ffs(b) - 1; b &= b - 1;
                                const double dx = xi.x - xe[j].x, dy = xi.y - xe[j].y;
                                m2 = fmin(m2, fma(dx, dx, dy * dy));
                            }
                        }
                    } else {
                        for (int j = 0; j < N; ++j) {
                            if (j == i) continue;
                            const double dx = xi.x - xe[j].x, dy = xi.y - xe[j].y;
                            m2 = fmin(m2, fma(dx, dx, dy * dy));
                        }
                    }
                    reinterpret_cast<TO *>(a.aux.min_dist)[grow] = (TO)sqrt(m2);
                }
                if (a.aux.vel_diff) {
                    const volatile double *st = stats + g * STATS_PER_ENV;
                    const double want = (double)(t + 1);
                    while (st[2] < want) { }
                    __threadfence_block();
                    const double invN = 1.0 / (double)N;
                    const double ex = xi.z - st[0] * invN, ey = xi.w - st[1] * invN;      // :138-140
                    reinterpret_cast<TO *>(a.aux.vel_diff)[grow] = (TO)sqrt(fma(ex, ex, ey * ey));
                }
            }
        }

        // ---- dense adj_mean rows: every warp expands its own 32 rows from their exact bit masks -------------------------------
        __syncwarp();                                        // the lanes' next state, mask words and 1/deg (shared memory) are complete
        if (!last_step && lane == 0) gbar_arrive(&step_bar);
        if (net != nullptr) {
            const int wr0 = tid - lane;
            const int nrw = min(32, nrows - wr0);
            if (nrw > 0) {
                TO *wout = net + ((size_t)e0 * N + wr0) * N;
                const unsigned *wm = reinterpret_cast<const unsigned *>(smem + a.off_masks) + wr0;
                const TO *idgs = idg_s + wr0;
                if constexpr (NC != 0) {
                    expand_rows_const<TO, NC, GMS>(wout, wm, idgs, nrw, lane);
                } else {
                    constexpr int VMAX = 16 / (int)sizeof(TO);
                    if (N % VMAX == 0) expand_rows_g<TO, VMAX, GMS>(wout, wm, idgs, N, nrw, lane);
                    else if (VMAX == 4 && N % 2 == 0) expand_rows_g<TO, 2, GMS>(wout, wm, idgs, N, nrw, lane);
                    else expand_rows_g<TO, 1, GMS>(wout, wm, idgs, N, nrw, lane);
                }
            }
            __syncwarp();                                    // the next step rewrites the words
        }
    }
    pdl_launch();
}

}  // namespace flock
