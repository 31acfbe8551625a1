// flock_swarm_kernel.cuh -- the flocking step for swarms too large to stage in shared memory (sm_100a).
//
// Same algorithm as flock_env_kernel.cuh ("filter then refine"), restructured for N up to 2^16+ agents per env:
//
//   flock_prep_kernel  (O(N)) : x' = integrate(x_in, u_in) for ALL agents -> x_out; packed fp32 positions relative
//                               to the env origin -> scratch; per-block partial velocity sums -> scratch (every row CTA
//                               adds them up in block order: deterministic, shard independent, no tickets or counters to
//                               clear between launches).                        flocking_relative.py:96-105,145-147
//   flock_swarm_kernel (rows) : a CTA owns 256/LANES rows of one env.  The packed fp32 positions of ALL agents stream
//                               through shared memory in tiles, double buffered with 1-D TMA bulk loads completing on
//                               mbarriers (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes).
//                               Per tile: phase A (packed f32x2 filter, conservative threshold, two rows per thread)
//                               appends the few candidates to per-thread lists; after the last tile phase B refines
//                               them in fp64 (x_j read through L2).  The dense adj_mean rows (256 KB each) are zero-filled
//                               by dedicated FILL WARPS of the CTA with direct streaming stores while the other warps
//                               filter (the store queue never blocks a filtering warp); one block barrier later the few
//                               non-zeros are stored over them.                  :111-134,194-226, flocking.py:20-24
//
// The row arithmetic does not depend on how rows are sharded over CTAs or GPUs.  Exact-order (NumPy bit-exact) mode
// is not offered here: its sequential all-j sums belong to the env-resident kernel (N up to a few thousand).
#pragma once
#include "flock_env_kernel.cuh"

namespace flock {

struct SArgs {
    const double *x_in;     // prep: state before the step
    const double *u_in;     // prep: action (nullable: no integration)
    double *x_cur;          // prep writes, pair kernel reads: x' [E][N][4]
    double *u_next;
    void *values;
    void *network;
    void *reward;
    void *obs;
    float4 *pf;             // scratch: packed positions [E][NP2]
    double *partials;       // scratch: [E][PB][4]  per prep block: sum (v - ref), sum (v - ref)^2
    unsigned *needy_cnt;    // scratch: [E] rows whose kNN search needs a full rescan (cleared by the prep kernel)
    unsigned *needy_rows;   // scratch: [E][NEEDY_CAP]
    double *fix_d;          // scratch: [E][NEEDY_CAP][KFIX_SPLIT][TOPK] partial kNN lists of the fix-up kernel (r2) ...
    int *fix_i;             //          ... and their agent indices
    unsigned *fix_ticket;   // scratch: [E][NEEDY_CAP] arrival tickets (cleared by the prep kernel and by their last user)
    unsigned *ovf;          // scratch: [E][rows][32][SW_LCAP2] second tier of the candidate lists (touched only on overflow)
    int E, N, NP2, PB;
    int row_lo, row_hi;
    int RT;                 // rows per CTA (SW_R * 256 / LANES)
    int TS;                 // pair slots per tile
    int centralized;
    double R, R2, Rm, thrC, dt, ascale, inv_ascale;
    StepVariant var;
    unsigned off_tiles, off_lists, off_bar, off_tot;
    int fill_warps;         // warps per CTA that only zero-fill the CTA's dense rows (0 when network == NULL)
};

constexpr int PREP_THREADS = 256;
constexpr int NEEDY_CAP = 2048;     // rows per env the kNN fix-up kernel can take; beyond that the row kernel rescans itself
constexpr int KFIX_SPLIT = 32;      // CTAs that share a queued row in the kNN fix-up kernel
constexpr int SW_R = 2;         // rows per thread in the filter
constexpr int SW_LCAP = 16;     // candidates per (row, lane) in shared memory ...
constexpr int SW_LCAP2 = 112;   // ... and in the global second tier (dense swarms): 128 per (row, lane) before the full rescan
constexpr int SW_THREADS = 256; // filtering threads per CTA (fill warps come on top)
constexpr int SW_FILL_WARPS = 2;
}  // namespace flock
#ifndef FLOCK_ARGS_ONLY
namespace flock {

// ---- mbarrier / TMA load helpers -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_g2s(void *sdst, const void *gsrc, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- O(N) preparation ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PREP_THREADS) flock_prep_kernel(const SArgs a) {
    const int env = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    const int N = a.N;
    const size_t e0 = (size_t)env * N;
    pdl_wait();
    pdl_launch();               // the row kernel may start now: its fill warps zero the dense rows while this kernel runs
    const double4 ref = ld_state(a.x_in + e0 * 4);          // origin / reference velocity: agent 0 before the step
    const int j = b * PREP_THREADS + tid;
    double sx = 0, sy = 0, qx = 0, qy = 0;
    if (j < N) {
        double4 x = ld_state(a.x_in + (e0 + j) * 4);
        if (a.u_in) x = integrate_variant(x, __ldg(reinterpret_cast<const double2 *>(a.u_in) + e0 + j), j, a.dt, a.var);
        st_state(a.x_cur + (e0 + j) * 4, x);
        float *p = reinterpret_cast<float *>(a.pf + (size_t)env * a.NP2 + (j >> 1));
        p[j & 1] = (float)(x.x - ref.x); p[2 + (j & 1)] = (float)(x.y - ref.y);
        if ((N & 1) && j == N - 1) { p[1] = CUDART_INF_F; p[3] = CUDART_INF_F; }
        const double ex = x.z - ref.z, ey = x.w - ref.w;
        sx = ex; sy = ey; qx = ex * ex; qy = ey * ey;
    }
    // fixed-shape block reduction; the row CTAs add the per-block partials in block order
    __shared__ double red[PREP_THREADS / 32][4];
    sx = warp_sum(sx); sy = warp_sum(sy); qx = warp_sum(qx); qy = warp_sum(qy);
    if ((tid & 31) == 0) { red[tid >> 5][0] = sx; red[tid >> 5][1] = sy; red[tid >> 5][2] = qx; red[tid >> 5][3] = qy; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (int w = 0; w < PREP_THREADS / 32; ++w) { t0 += red[w][0]; t1 += red[w][1]; t2 += red[w][2]; t3 += red[w][3]; }
        double *pp = a.partials + ((size_t)env * a.PB + b) * 4;
        pp[0] = t0; pp[1] = t1; pp[2] = t2; pp[3] = t3;
        if (b == 0) a.needy_cnt[env] = 0u;                   // the kNN fix-up queue of this step starts empty
    }
    if (b == 0) for (int q = tid; q < NEEDY_CAP; q += PREP_THREADS) a.fix_ticket[(size_t)env * NEEDY_CAP + q] = 0u;
}

// ---- the row kernel --------------------------------------------------------------------------------------------------
// A CTA owns SW_R * 256/LANES consecutive rows of one env: thread (rl, sub) works for rows SW_R*rl .. SW_R*rl + SW_R-1.
//   tile loop : every packed slot loaded from shared memory (one LDS.128) is tested against the thread's SW_R rows --
//               per row 2 FADD2 + 2 FFMA2 + 2 funnel shifts, the candidate test is the sign of dx^2 + (dy^2 - thr) --
//               so the loop is ALU bound with SW_R independent chains per load (with one row per thread it stalled on
//               the LDS latency at 4 warps per scheduler: 47 % issue utilisation).  The rare candidates (a few per row
//               and tile) are appended, ascending j, to the thread's own list in shared memory: no atomics, no order
//               that depends on timing.
//   refine    : after the last tile each thread refines the candidates of its lists in fp64 (x_j through L2), row by
//               row; the LANES partial sums of a row are combined with xor shuffles.  Per-lane ascending j plus a fixed
//               shuffle pattern: the result does not depend on how rows are sharded over CTAs or GPUs.
//   overflow  : a (row, lane) list holds SW_LCAP candidates in shared memory and spills into a second tier of SW_LCAP2 in
//               global scratch (a swarm several times denser than the reference's: same code path, a few extra L2 accesses);
//               only a row with more than SW_LCAP + SW_LCAP2 candidates on one lane is recomputed by scanning all agents.
template <typename TO, int LANES, bool KNN>
__global__ void __launch_bounds__(SW_THREADS + 32 * SW_FILL_WARPS, 2) flock_swarm_kernel(const SArgs a) {
    static_assert(LANES >= 8 && LANES <= 32, "swarm rows are split over 8..32 lanes");
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *tiles = reinterpret_cast<float4 *>(smem + a.off_tiles);          // 2 stages x TS slots
    unsigned *lists = reinterpret_cast<unsigned *>(smem + a.off_lists);      // [thread][SW_R][SW_LCAP] candidate indices
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + a.off_bar);   // 2 mbarriers

    const int tid = threadIdx.x, lane = tid & 31;
    const int env = blockIdx.y;
    const int N = a.N, NP2 = a.NP2;
    const int rl = tid / LANES, sub = tid % LANES;
    const int RTC = SW_R * (256 / LANES);                                    // rows per CTA
    const int ibase = a.row_lo + blockIdx.x * RTC + rl * SW_R;               // first row of this thread
    const size_t e0 = (size_t)env * N;
    const double *xc = a.x_cur + e0 * 4;
    const float4 *pfe = a.pf + (size_t)env * NP2;
    TO *net = reinterpret_cast<TO *>(a.network);
    const int ntiles = (NP2 + a.TS - 1) / a.TS;

    // ---- fill warps: zero the CTA's rows of the dense adjacency (one contiguous run of RTC * N elements) with coalesced
    //      16-byte streaming stores while the other warps filter; they meet again at the block barrier before the non-zero
    //      entries are stored.  (A filtering warp that also stores stalls on the store queue.)  They do not wait for the
    //      prep kernel: nothing they touch depends on it, so the fill starts while prep still runs.
    if (tid >= SW_THREADS) {
        if (net) {
            const int r0 = a.row_lo + blockIdx.x * RTC, r1 = min(r0 + RTC, a.row_hi);
            if (r0 < r1) {
                const size_t base = (e0 + r0) * (size_t)N;
                const size_t total = (size_t)(r1 - r0) * N;
                constexpr unsigned VEC = 16 / (unsigned)sizeof(TO);
                const size_t headn = min((size_t)((VEC - (unsigned)(base & (VEC - 1))) & (VEC - 1)), total);
                const size_t nvec = (total - headn) / VEC;
                TO *out = net + base;
                const int ft = tid - SW_THREADS, nft = blockDim.x - SW_THREADS;
                if ((size_t)ft < headn) out[ft] = TO(0);
                if ((size_t)ft < total - (headn + nvec * VEC)) out[headn + nvec * VEC + ft] = TO(0);
                float4 *zdst = reinterpret_cast<float4 *>(out + headn);
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                size_t v = ft;
                for (; v + 7 * (size_t)nft < nvec; v += 8 * (size_t)nft) {     // 8 stores in flight per lane
#pragma unroll
                    for (int q = 0; q < 8; ++q) __stcs(zdst + v + (size_t)q * nft, z4);
                }
                for (; v < nvec; v += nft) __stcs(zdst + v, z4);
            }
        }
        __syncthreads();                                          // (B) the zero stores precede the non-zero stores
        return;
    }

    // ---- setup (the 256 filtering threads): barriers, the env's totals, first two tiles in flight ---------------------------
    double *tot = reinterpret_cast<double *>(smem + a.off_tot);   // (ref vx, ref vy, sum vx, sum vy, -, ref px, ref py, -)
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    fence_async_smem();
    pdl_wait();                                                   // the prep kernel's outputs are visible from here on
    auto issue_tile = [&](const int t) {                                     // one thread: queue tile t into stage t&1
        const int s0 = t * a.TS, ns = min(a.TS, NP2 - s0);
        mbar_expect_tx(&bars[t & 1], (unsigned)ns * 16u);
        bulk_load_g2s(tiles + (size_t)(t & 1) * a.TS, pfe + s0, (unsigned)ns * 16u, &bars[t & 1]);
    };
    if (tid < 32) {
        // warp 0: the env's velocity sums = the prep blocks' partials added in block order (lane-strided, then a fixed
        // shuffle tree): the same bits in every CTA, on every rank, whatever the shard
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        const double *pp = a.partials + (size_t)env * a.PB * 4;
        for (int q = lane; q < a.PB; q += 32) {
            const double2 p01 = *reinterpret_cast<const double2 *>(pp + (size_t)q * 4);
            const double2 p23 = *reinterpret_cast<const double2 *>(pp + (size_t)q * 4 + 2);
            t0 += p01.x; t1 += p01.y; t2 += p23.x; t3 += p23.y;
        }
        t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2); t3 = warp_sum(t3);
        if (lane == 0) {
            const double4 ref = ld_state(a.x_in + e0 * 4);        // origin / reference velocity: agent 0 before the step
            const double invN = 1.0 / (double)N;
            const double rew = -((t2 - t0 * t0 * invN) + (t3 - t1 * t1 * invN)) * invN;      // -(var vx + var vy)
            tot[0] = ref.z; tot[1] = ref.w; tot[2] = t0; tot[3] = t1; tot[4] = rew; tot[5] = ref.x; tot[6] = ref.y;
            if (a.reward && blockIdx.x == 0) reinterpret_cast<TO *>(a.reward)[env] = (TO)rew;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SW_THREADS) : "memory");  // barriers initialised, totals published
    if (tid == 0) { issue_tile(0); if (ntiles > 1) issue_tile(1); }

    // ---- per-row filter state ----------------------------------------------------------------------------------------
    // Flocking-v0: the 7 nearest neighbours are searched among the candidates, so the candidate radius is inflated
    // (x4 in r2: ~36 candidates per row at the reference density); rows that still find fewer than 7 rescan all agents.
    const float tc = (float)a.thrC * (KNN ? 4.0f : 1.0f), rc = sqrtf(tc);
    bool valid[SW_R];
    float2 nx[SW_R], ny[SW_R], nthr[SW_R];
    int cn[SW_R];                                            // candidates of this thread per row (may exceed SW_LCAP)
    unsigned *lst[SW_R], *ovl[SW_R];
#pragma unroll
    for (int r = 0; r < SW_R; ++r) {
        const int i = ibase + r;
        valid[r] = (i < a.row_hi) && (i < a.row_lo + (int)(blockIdx.x + 1) * RTC);
        const double2 pi_ = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)(valid[r] ? i : a.row_lo) * 4));
        const float pxi = (float)(pi_.x - tot[5]), pyi = (float)(pi_.y - tot[6]);   // same rounding as the packed array
        nx[r] = make_float2(-pxi, -pxi); ny[r] = make_float2(-pyi, -pyi);
        const float thrA = tc * (1.0f + 3e-6f) + 16.0f * rc * (fmaxf(fabsf(pxi), fabsf(pyi)) + 2.0f * rc) * 5.9604645e-8f;
        nthr[r] = make_float2(-thrA, -thrA);
        cn[r] = 0;
        lst[r] = lists + ((size_t)tid * SW_R + r) * SW_LCAP;
        // second tier: [env][row - row_lo][32 lanes][SW_LCAP2]
        ovl[r] = a.ovf + ((((size_t)env * (a.row_hi - a.row_lo) + (size_t)(valid[r] ? i - a.row_lo : 0)) * 32 + sub) * SW_LCAP2);
    }

    // ---- stream the tiles: filter only ---------------------------------------------------------------------------------
    for (int t = 0; t < ntiles; ++t) {
        const int s0 = t * a.TS, ns = min(a.TS, NP2 - s0);
        mbar_wait(&bars[t & 1], (unsigned)((t >> 1) & 1));
        const float4 *pe = tiles + (size_t)(t & 1) * a.TS;
        // slots sub, sub+LANES, ... ; bit 2q / 2q+1 of a word = the two agents of the word's q-th slot
        const int nsl = (ns - sub + LANES - 1) / LANES;
        for (int w = 0; w * 16 < nsl; ++w) {
            unsigned bits[SW_R];
#pragma unroll
            for (int r = 0; r < SW_R; ++r) bits[r] = 0;
            const int qb = w * 16, qe = min(16, nsl - qb);
            // s = dx^2 + (dy^2 - thrA): the candidate test is the SIGN of s (two packed FFMA2, no compare); one funnel
            // shift per pair moves it into the word (first pair ends up highest: one BREV per word)
            auto body = [&](const int q) {
                const float4 v = pe[sub + LANES * (qb + q)];
#pragma unroll
                for (int r = 0; r < SW_R; ++r) {
                    const float2 dx = __fadd2_rn(make_float2(v.x, v.y), nx[r]);
                    const float2 dy = __fadd2_rn(make_float2(v.z, v.w), ny[r]);
                    const float2 s = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, nthr[r]));
                    bits[r] = __funnelshift_l(__float_as_uint(s.x), bits[r], 1);
                    bits[r] = __funnelshift_l(__float_as_uint(s.y), bits[r], 1);
                }
            };
            if (qe == 16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) body(q);
#pragma unroll
                for (int r = 0; r < SW_R; ++r) bits[r] = __brev(bits[r]);
            } else {
                for (int q = 0; q < qe; ++q) body(q);
#pragma unroll
                for (int r = 0; r < SW_R; ++r) bits[r] = __brev(bits[r]) >> (32 - 2 * qe);   // 0 < qe < 16
            }
            // the rare candidates go to the thread's lists, ascending j
#pragma unroll
            for (int r = 0; r < SW_R; ++r) {
                unsigned c = valid[r] ? bits[r] : 0u;
                while (c) {
                    const int k = __ffs(c) - 1;
                    c &= c - 1;
                    const int j = 2 * (s0 + sub + LANES * (qb + (k >> 1))) + (k & 1);
                    if (j == ibase + r || j >= N) continue;
                    if (cn[r] < SW_LCAP) lst[r][cn[r]] = (unsigned)j;
                    else if (cn[r] < SW_LCAP + SW_LCAP2) ovl[r][cn[r] - SW_LCAP] = (unsigned)j;
                    ++cn[r];
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SW_THREADS) : "memory");   // the filtering threads are done with stage t&1
        if (tid == 0 && t + 2 < ntiles) issue_tile(t + 2);
    }

    // ---- refine, row by row (the lists are private to the thread: no barrier needed) ------------------------------------
    const bool cen = a.centralized != 0;
    TO idgs[SW_R];
    bool overs[SW_R];
#pragma unroll 1
    for (int r = 0; r < SW_R; ++r) {
        const int i = valid[r] ? ibase + r : a.row_lo;
        const size_t grow = e0 + i;
        const double4 xi = ld_state(xc + (size_t)i * 4);
        Acc A;
        A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = 0.0;
        TopK<double> top;                                    // kNN by exact fp64 r2 over the candidates
        if constexpr (KNN) top.init(i);
        int deg = 0;
        // a list that overflowed anywhere in the row: the whole row is recomputed from all agents (same on every lane)
        bool over = valid[r] && cn[r] > SW_LCAP + SW_LCAP2;
#pragma unroll
        for (int m = LANES / 2; m > 0; m >>= 1) over |= (__shfl_xor_sync(0xffffffffu, over ? 1 : 0, m) != 0);
        const int ncand = valid[r] ? min(cn[r], SW_LCAP + SW_LCAP2) : 0;
        auto slot = [&](const int c) -> unsigned * { return c < SW_LCAP ? lst[r] + c : ovl[r] + (c - SW_LCAP); };
        if (!over) {
            for (int c = 0; c < ncand; ++c) {
                const int j = (int)*slot(c);
                const double4 xj = ld_state(xc + (size_t)j * 4);
                if constexpr (KNN) {
                    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                    top.insert(fma(dx, dx, dy * dy), j);
                }
                if (refine_fast<sizeof(TO) == 4>(xi, xj, a.R2, a.Rm, A)) { *slot(c) = (unsigned)j | 0x80000000u; ++deg; }
            }
        } else if (valid[r]) {
            for (int j = sub; j < N; j += LANES) {
                if (j == i) continue;
                const double4 xj = ld_state(xc + (size_t)j * 4);
                const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                const double r2 = fma(dx, dx, dy * dy);
                if constexpr (KNN) top.insert(r2, j);
                if (r2 <= a.thrC && refine_fast<sizeof(TO) == 4>(xi, xj, a.R2, a.Rm, A)) ++deg;
            }
        }
        // reduce across the LANES threads of the row
#pragma unroll
        for (int m = LANES / 2; m > 0; m >>= 1) {
            A.f0 += shfl_xor(A.f0, m); A.f1 += shfl_xor(A.f1, m); A.f2 += shfl_xor(A.f2, m);
            A.f3 += shfl_xor(A.f3, m); A.f4 += shfl_xor(A.f4, m); A.f5 += shfl_xor(A.f5, m);
            A.gx += shfl_xor(A.gx, m); A.gy += shfl_xor(A.gy, m);
            deg += shfl_xor(deg, m);
        }
        const TO idg = inv_deg(deg, TO(0));

        if (valid[r] && sub == 0) {
            if (a.values) {
                TO *v = reinterpret_cast<TO *>(a.values) + grow * NFE;
                v[0] = (TO)A.f0; v[1] = (TO)A.f1; v[2] = (TO)A.f2; v[3] = (TO)A.f3; v[4] = (TO)A.f4; v[5] = (TO)A.f5;
            }
            if (a.u_next) {
                double sdx = A.f0, sdy = A.f3;
                if (cen) {                                   // sum_j (v_i - v_j) = N*(v_i - ref) - sum_j (v_j - ref)
                    sdx = fma((double)N, xi.z - tot[0], -tot[2]);
                    sdy = fma((double)N, xi.w - tot[1], -tot[3]);
                }
                double ux = fmin(fmax(-2.0 * A.gx - sdx, -10.0), 10.0) * a.inv_ascale;
                double uy = fmin(fmax(-sdy - 2.0 * A.gy, -10.0), 10.0) * a.inv_ascale;
                if (a.var.u_clip > 0.0) {
                    ux = fmin(fmax(ux, -a.var.u_clip), a.var.u_clip); uy = fmin(fmax(uy, -a.var.u_clip), a.var.u_clip);
                }
                reinterpret_cast<double2 *>(a.u_next)[grow] = make_double2(ux, uy);
            }
        }

        // ---- kNN observation ------------------------------------------------------------------------------------------
        if constexpr (KNN) {
            int nb[TOPK];
            auto merge = [&](int *out) {                     // every lane of the row ends with the row's sorted ids
#pragma unroll
                for (int m = 0; m < TOPK; ++m) {
                    double bd = top.d[0]; int bi = top.id[0];
#pragma unroll
                    for (int s2 = LANES / 2; s2 > 0; s2 >>= 1) {
                        const double od = shfl_xor(bd, s2); const int oi = shfl_xor(bi, s2);
                        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    }
                    if (bi == top.id[0] && bd == top.d[0]) top.pop_front(i);
                    out[m] = bi;
                }
            };
            merge(nb);
            // Fewer than 7 neighbours inside the inflated radius (flock edge, stray agents): the row is queued for
            // flock_knn_fix_kernel, which rescans ALL agents with a whole CTA per row after this kernel (so the few such
            // rows do not stretch this kernel's tail).  If the queue is full the row rescans here.
            bool need = valid[r] && !over && (nb[KNN_K - 1] == i);
            if (need && a.obs) {
                unsigned slot = 0;
                if (sub == 0) slot = atomicAdd(a.needy_cnt + env, 1u);
                slot = __shfl_sync(__activemask(), slot, (lane / LANES) * LANES);
                if (slot < (unsigned)NEEDY_CAP) {
                    if (sub == 0) a.needy_rows[(size_t)env * NEEDY_CAP + slot] = (unsigned)i;
                    need = false;
                }
            }
            if (__any_sync(0xffffffffu, need)) {
                top.init(i);
                if (need) {
                    for (int j = sub; j < N; j += LANES) {
                        if (j == i) continue;
                        const double2 pj = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)j * 4));
                        const double dx = xi.x - pj.x, dy = xi.y - pj.y;
                        top.insert(fma(dx, dx, dy * dy), j);
                    }
                }
                int nb2[TOPK];
                merge(nb2);
                if (need) {
#pragma unroll
                    for (int m = 0; m < TOPK; ++m) nb[m] = nb2[m];
                }
            }
            // exact order / membership among the survivors, with the reference's own keys (unfused r2, ties to the lower index)
            knn_rerank(xi, i, reinterpret_cast<const double4 *>(xc), nb);
            if (valid[r] && sub == 0 && a.obs) {
                TO *o = reinterpret_cast<TO *>(a.obs) + grow * (NXS * KNN_K);
#pragma unroll
                for (int m = 0; m < KNN_K; ++m) {             // the diagonal is never a candidate here; unused slots hold i
                    const double4 xn = ld_state(xc + (size_t)min(nb[m], N - 1) * 4);
                    o[m * 4 + 0] = (TO)(xi.x - xn.x); o[m * 4 + 1] = (TO)(xi.y - xn.y);
                    o[m * 4 + 2] = (TO)(xi.z - xn.z); o[m * 4 + 3] = (TO)(xi.w - xn.w);
                }
            }
        }

        idgs[r] = idg; overs[r] = over;
    }

    // ---- the non-zero adj_mean entries, stored over the zero-filled rows: only this waits for the fill warps (everything
    //      above -- filter, refine, features, actions, observations -- overlapped with the fill) --------------------------------
    __syncthreads();                                         // (B) the fill warps have issued the CTA's zero stores
    if (net != nullptr) {
#pragma unroll 1
        for (int r = 0; r < SW_R; ++r) {
            if (!valid[r]) continue;
            const int i = ibase + r;
            TO *orow = net + (e0 + i) * (size_t)N;
            const TO idg = idgs[r];
            if (!overs[r]) {
                const int ncand = min(cn[r], SW_LCAP + SW_LCAP2);
                for (int c = 0; c < ncand; ++c) {
                    const unsigned e = c < SW_LCAP ? lst[r][c] : ovl[r][c - SW_LCAP];
                    if (e & 0x80000000u) orow[e & 0x7fffffffu] = idg;
                }
            } else {                                         // rare: recompute the row's adjacency
                const double2 pi_ = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)i * 4));
                for (int j = sub; j < N; j += LANES) {
                    if (j == i) continue;
                    const double2 pj = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)j * 4));
                    const double dx = pi_.x - pj.x, dy = pi_.y - pj.y;
                    if (fma(dx, dx, dy * dy) < a.R2) orow[j] = idg;
                }
            }
        }
    }
}

// ---- kNN fix-up: the queued rows rescan all agents (exact fp64 r2, ties to the lower index) ----------------------------
// A queued row costs a full scan of the swarm, and the kernel sits on the critical path of the step (the all-gather and the
// next prep wait for it).  One CTA per row ran the whole scan on ONE SM: ~65 us for a single queued row (the sorted insert
// diverges, so a warp pays the full bubble whenever any lane inserts).  Now KFIX_SPLIT CTAs share a row: each scans its slice of
// the agents, leaves its TOPK best in scratch, and the last one to arrive (a ticket per row) merges the KFIX_SPLIT lists with
// one warp and writes the observation.  Deterministic: the merge orders by (r2, index) whatever the arrival order.
constexpr int KFIX_THREADS = 256;
constexpr int KFIX_SLOTS = 9;       // rows in flight: the grid is KFIX_SPLIT * KFIX_SLOTS CTAs (~2 per SM)
template <typename TO>
__global__ void __launch_bounds__(KFIX_THREADS) flock_knn_fix_kernel(const SArgs a) {
    const int env = blockIdx.y;
    pdl_wait();
    const unsigned cnt = min(a.needy_cnt[env], (unsigned)NEEDY_CAP);
    constexpr int NW = KFIX_THREADS / 32;
    __shared__ double sd[NW * TOPK];
    __shared__ int si[NW * TOPK];
    __shared__ unsigned s_last;
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x % KFIX_SPLIT, slot = blockIdx.x / KFIX_SPLIT;
    const double *xc = a.x_cur + (size_t)env * N * 4;
    const int per = (N + KFIX_SPLIT - 1) / KFIX_SPLIT, j_lo = split * per, j_hi = min(N, j_lo + per);
    auto warp_merge = [&](TopK<double> &top, const int i, double (&od)[TOPK], int (&oi)[TOPK]) {   // the warp's TOPK best, in order
#pragma unroll
        for (int m = 0; m < TOPK; ++m) {
            double bd = top.d[0]; int bi = top.id[0];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const double xd = shfl_xor(bd, s); const int xi_ = shfl_xor(bi, s);
                if (xd < bd || (xd == bd && xi_ < bi)) { bd = xd; bi = xi_; }
            }
            if (bi == top.id[0] && bd == top.d[0]) top.pop_front(i);
            od[m] = bd; oi[m] = bi;
        }
    };
    for (unsigned q = slot; q < cnt; q += KFIX_SLOTS) {
        const int i = (int)a.needy_rows[(size_t)env * NEEDY_CAP + q];
        const double4 xi = ld_state(xc + (size_t)i * 4);
        TopK<double> top;
        top.init(i);
        // eight independent loads in flight per thread (the insert's early-out depends on the running 9th-best distance)
        for (int j0 = j_lo + tid; j0 < j_hi; j0 += KFIX_THREADS * 8) {
            double2 pj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * KFIX_THREADS;
                pj[u] = (j < j_hi) ? __ldg(reinterpret_cast<const double2 *>(xc + (size_t)j * 4)) : make_double2(CUDART_INF, CUDART_INF);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * KFIX_THREADS;
                const double dx = xi.x - pj[u].x, dy = xi.y - pj[u].y;
                if (j != i && j < j_hi) top.insert(fma(dx, dx, dy * dy), j);
            }
        }
        double md[TOPK]; int mi[TOPK];
        warp_merge(top, i, md, mi);
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < TOPK; ++m) { sd[warp * TOPK + m] = md[m]; si[warp * TOPK + m] = mi[m]; }
        }
        __syncthreads();
        double *pd = a.fix_d + (((size_t)env * NEEDY_CAP + q) * KFIX_SPLIT) * TOPK;
        int *pi = a.fix_i + (((size_t)env * NEEDY_CAP + q) * KFIX_SPLIT) * TOPK;
        if (warp == 0) {                                     // the CTA's TOPK best -> scratch
            TopK<double> fin;
            fin.init(i);
            for (int e = lane; e < NW * TOPK; e += 32) if (si[e] != i) fin.insert(sd[e], si[e]);
            warp_merge(fin, i, md, mi);
            if (lane == 0) {
#pragma unroll
                for (int m = 0; m < TOPK; ++m) { pd[split * TOPK + m] = md[m]; pi[split * TOPK + m] = mi[m]; }
                __threadfence();
                s_last = (atomicAdd(a.fix_ticket + (size_t)env * NEEDY_CAP + q, 1u) == (unsigned)(KFIX_SPLIT - 1)) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (s_last && warp == 0) {                           // last CTA of the row: merge the KFIX_SPLIT lists, write the row
            __threadfence();
            TopK<double> fin;
            fin.init(i);
            for (int e = lane; e < KFIX_SPLIT * TOPK; e += 32) {
                const double d = __ldcg(pd + e); const int id = __ldcg(pi + e);
                if (id != i) fin.insert(d, id);
            }
            warp_merge(fin, i, md, mi);
            knn_rerank(xi, i, reinterpret_cast<const double4 *>(xc), mi);
            if (lane < KNN_K && a.obs) {
                TO *o = reinterpret_cast<TO *>(a.obs) + ((size_t)env * N + i) * (NXS * KNN_K);
                int nbm = mi[0];
#pragma unroll
                for (int m = 1; m < KNN_K; ++m) nbm = (lane == m) ? mi[m] : nbm;
                const double4 xn = ld_state(xc + (size_t)nbm * 4);
                o[lane * 4 + 0] = (TO)(xi.x - xn.x); o[lane * 4 + 1] = (TO)(xi.y - xn.y);
                o[lane * 4 + 2] = (TO)(xi.z - xn.z); o[lane * 4 + 3] = (TO)(xi.w - xn.w);
            }
            if (lane == 0) a.fix_ticket[(size_t)env * NEEDY_CAP + q] = 0u;
        }
        __syncthreads();
    }
}

}  // namespace flock
#endif  // FLOCK_ARGS_ONLY
