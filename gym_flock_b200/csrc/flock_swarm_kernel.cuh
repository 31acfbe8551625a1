// flock_swarm_kernel.cuh -- the flocking step for swarms too large to stage in shared memory (sm_100a).
//
// Same algorithm as flock_env_kernel.cuh ("filter then refine"), restructured for N up to 2^16+ agents per env:
//
//   flock_prep_kernel  (O(N)) : x' = integrate(x_in, u_in) for ALL agents -> x_out; packed fp32 positions relative
//                               to the env origin -> scratch; deterministic two-stage reduction of the velocity sums
//                               (alignment identity, reward) -> scratch.       flocking_relative.py:96-105,145-147
//   flock_swarm_kernel (rows) : a CTA owns 256/LANES rows of one env.  The packed fp32 positions of ALL agents stream
//                               through shared memory in tiles, double buffered with 1-D TMA bulk loads completing on
//                               mbarriers (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes).
//                               Per tile: phase A (packed f32x2 filter, conservative threshold) then phase B (fp64
//                               refine of the few candidates, x_j read through L2).  Neighbours are collected in a
//                               per-row list; the dense adj_mean rows are cleared by per-warp TMA bulk stores issued
//                               at kernel start and the non-zeros scattered at the end.   :111-134,194-226, flocking.py:20-24
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
    double *partials;       // scratch: [E][PB][4]
    double *totals;         // scratch: [E][8]  (ref vx, ref vy, sum vx, sum vy, reward, ref px, ref py, pad)
    unsigned *counters;     // scratch: [E] tickets of the two-stage reduction (zero between launches)
    unsigned *needy_cnt;    // scratch: [E] rows whose kNN search needs a full rescan (cleared every step)
    unsigned *needy_rows;   // scratch: [E][NEEDY_CAP]
    int E, N, NP2, PB;
    int row_lo, row_hi;
    int RT;                 // rows per CTA (256 / LANES)
    int TS;                 // pair slots per tile
    int WT;                 // 32-bit candidate words per (thread, tile)
    int CAP;                // neighbour-list capacity per row
    int centralized;
    double R, R2, Rm, thrC, dt, ascale, inv_ascale;
    StepVariant var;
    unsigned off_tiles, off_masks, off_lists, off_cnt, off_zero, off_bar;
    unsigned zero_bytes;
};

constexpr int PREP_THREADS = 256;
constexpr int NEEDY_CAP = 8192;     // rows per env the kNN fix-up kernel can take; beyond that the row kernel rescans itself

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
__global__ void __launch_bounds__(PREP_THREADS) flock_prep_kernel(const SArgs a, const int want_reward_f32) {
    const int env = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    const int N = a.N;
    const size_t e0 = (size_t)env * N;
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
    // stage 1: fixed-shape block reduction
    __shared__ double red[PREP_THREADS / 32][4];
    sx = warp_sum(sx); sy = warp_sum(sy); qx = warp_sum(qx); qy = warp_sum(qy);
    if ((tid & 31) == 0) { red[tid >> 5][0] = sx; red[tid >> 5][1] = sy; red[tid >> 5][2] = qx; red[tid >> 5][3] = qy; }
    __syncthreads();
    __shared__ bool last;
    if (tid == 0) {
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (int w = 0; w < PREP_THREADS / 32; ++w) { t0 += red[w][0]; t1 += red[w][1]; t2 += red[w][2]; t3 += red[w][3]; }
        double *pp = a.partials + ((size_t)env * a.PB + b) * 4;
        pp[0] = t0; pp[1] = t1; pp[2] = t2; pp[3] = t3;
        __threadfence();
        last = (atomicAdd(a.counters + env, 1u) == (unsigned)(gridDim.x - 1));
    }
    __syncthreads();
    // stage 2: the last block to arrive adds the partials in block order (deterministic, shard independent)
    if (last && tid == 0) {
        __threadfence();
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        const volatile double *pp = a.partials + (size_t)env * a.PB * 4;
        for (int q = 0; q < (int)gridDim.x; ++q) { t0 += pp[q * 4 + 0]; t1 += pp[q * 4 + 1]; t2 += pp[q * 4 + 2]; t3 += pp[q * 4 + 3]; }
        double *tt = a.totals + (size_t)env * 8;
        const double invN = 1.0 / (double)N;
        const double rew = -((t2 - t0 * t0 * invN) + (t3 - t1 * t1 * invN)) * invN;      // -(var vx + var vy)
        tt[0] = ref.z; tt[1] = ref.w; tt[2] = t0; tt[3] = t1; tt[4] = rew; tt[5] = ref.x; tt[6] = ref.y;
        if (a.reward) {
            if (want_reward_f32) reinterpret_cast<float *>(a.reward)[env] = (float)rew;
            else reinterpret_cast<double *>(a.reward)[env] = rew;
        }
        a.counters[env] = 0u;                                // ready for the next launch
    }
}

// ---- the row kernel --------------------------------------------------------------------------------------------------
template <typename TO, int LANES, bool KNN>
__global__ void __launch_bounds__(256, 2) flock_swarm_kernel(const SArgs a) {
    static_assert(LANES >= 8 && LANES <= 32, "swarm rows are split over 8..32 lanes");
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *tiles = reinterpret_cast<float4 *>(smem + a.off_tiles);          // 2 stages x TS slots
    unsigned *masks = reinterpret_cast<unsigned *>(smem + a.off_masks);      // [thread][WT]
    unsigned *lists = reinterpret_cast<unsigned *>(smem + a.off_lists);      // [row][CAP] neighbour indices
    int *lcnt = reinterpret_cast<int *>(smem + a.off_cnt);                   // [row] neighbours found (may exceed CAP)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + a.off_bar);   // 2 mbarriers

    const int tid = threadIdx.x, lane = tid & 31;
    const int env = blockIdx.y;
    const int N = a.N, NP2 = a.NP2;
    const int rl = tid / LANES, sub = tid % LANES;
    const int i = a.row_lo + blockIdx.x * a.RT + rl;
    const bool valid = (rl < a.RT) && (i < a.row_hi);
    const size_t e0 = (size_t)env * N;
    const double *xc = a.x_cur + e0 * 4;
    const float4 *pfe = a.pf + (size_t)env * NP2;
    TO *net = reinterpret_cast<TO *>(a.network);
    const size_t grow = e0 + (valid ? i : a.row_lo);
    const int ntiles = (NP2 + a.TS - 1) / a.TS;

    // ---- setup: zero block, list counters, barriers; first two tiles in flight -----------------------------------------
    {
        for (int q = tid; q < a.RT; q += blockDim.x) lcnt[q] = 0;
        if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
        fence_async_smem();
    }
    __syncthreads();
    auto issue_tile = [&](const int t) {                                     // one thread: queue tile t into stage t&1
        const int s0 = t * a.TS, ns = min(a.TS, NP2 - s0);
        mbar_expect_tx(&bars[t & 1], (unsigned)ns * 16u);
        bulk_load_g2s(tiles + (size_t)(t & 1) * a.TS, pfe + s0, (unsigned)ns * 16u, &bars[t & 1]);
    };
    if (tid == 0) { issue_tile(0); if (ntiles > 1) issue_tile(1); }

    // ---- zero-fill of the warp's rows of the dense adjacency: plain coalesced 16-byte streaming stores, one slice per
    //      tile of the main loop, so the writes drain to HBM while the warp filters.  (1-D TMA bulk stores from a shared
    //      zero block were measured slower: the bulk-store queue throttles at ~3.6-4.8 TB/s, direct stores reach 6.3.)
    float4 *zdst = nullptr;                                  // first 16-byte vector of the warp's region
    size_t znvec = 0;
    if (net) {
        const int wr0 = a.row_lo + blockIdx.x * a.RT + (tid - lane) / LANES;    // first row of this warp
        const int wr1 = min(min(wr0 + 32 / LANES, a.row_lo + (int)(blockIdx.x + 1) * a.RT), a.row_hi);
        if (wr0 < wr1) {
            const size_t base = (e0 + wr0) * (size_t)N;
            const size_t total = (size_t)(wr1 - wr0) * N;
            constexpr unsigned VEC = 16 / (unsigned)sizeof(TO);
            const size_t headn = min((size_t)((VEC - (unsigned)(base & (VEC - 1))) & (VEC - 1)), total);
            znvec = (total - headn) / VEC;
            TO *out = net + base;
            if ((size_t)lane < headn) out[lane] = TO(0);
            if ((size_t)lane < total - (headn + znvec * VEC)) out[headn + znvec * VEC + lane] = TO(0);
            zdst = reinterpret_cast<float4 *>(out + headn);
        }
    }
    const size_t zper = ((znvec + (size_t)ntiles - 1) / (size_t)ntiles + 255) & ~(size_t)255;   // vectors per tile (x256)
    auto zero_slice = [&](const int t) {
        const size_t lo = (size_t)t * zper, hi = min(lo + zper, znvec);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        size_t v = lo + lane;
        for (; v + 7 * 32 < hi; v += 8 * 32) {               // 8 stores in flight per lane, 512 contiguous bytes per warp store
#pragma unroll
            for (int q = 0; q < 8; ++q) __stcs(zdst + v + q * 32, z4);
        }
        for (; v < hi; v += 32) __stcs(zdst + v, z4);
    };

    // ---- row state ------------------------------------------------------------------------------------------------------
    const double4 xi = ld_state(xc + (size_t)(valid ? i : a.row_lo) * 4);
    const double *tot = a.totals + (size_t)env * 8;
    const float pxi = (float)(xi.x - tot[5]), pyi = (float)(xi.y - tot[6]);     // same rounding as the packed array
    const float2 nx = make_float2(-pxi, -pxi), ny = make_float2(-pyi, -pyi);
    // Flocking-v0: the 7 nearest neighbours are searched among the candidates, so the candidate radius is inflated
    // (x4 in r2: ~36 candidates per row at the reference density); rows that still find fewer than 7 rescan all agents.
    const float tc = (float)a.thrC * (KNN ? 4.0f : 1.0f), rc = sqrtf(tc);
    const float thrA = tc * (1.0f + 3e-6f) + 16.0f * rc * (fmaxf(fabsf(pxi), fabsf(pyi)) + 2.0f * rc) * 5.9604645e-8f;
    const float2 nthr2 = make_float2(-thrA, -thrA);
    unsigned *mrow = masks + (size_t)tid * a.WT;
    unsigned *lrow = lists + (size_t)rl * a.CAP;
    Acc A;
    A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = 0.0;
    TopK<double> top;                                       // kNN by exact fp64 r2 over the candidates
    if constexpr (KNN) top.init(i);
    const bool cen = a.centralized != 0;

    // ---- stream the tiles ------------------------------------------------------------------------------------------------
    for (int t = 0; t < ntiles; ++t) {
        const int s0 = t * a.TS, ns = min(a.TS, NP2 - s0);
        if (znvec) zero_slice(t);                            // queued before the wait: drains while the tile is filtered
        mbar_wait(&bars[t & 1], (unsigned)((t >> 1) & 1));
        const float4 *pe = tiles + (size_t)(t & 1) * a.TS;
        // phase A on this tile: slots sub, sub+LANES, ... ; bit 2q / 2q+1 of the lane's words = the slot's two agents
        const int nsl = valid ? (ns - sub + LANES - 1) / LANES : 0;
        int cnt = 0;
        for (int w = 0; w * 16 < nsl; ++w) {
            unsigned bits = 0;
            const int qb = w * 16, qe = min(16, nsl - qb);
            // s = dx^2 + (dy^2 - thrA): the candidate test is the SIGN of s (two packed FFMA2, no compare); one funnel
            // shift per pair moves it into the word (first pair ends up highest: one BREV per word)
            auto body = [&](const int q) {
                const float4 v = pe[sub + LANES * (qb + q)];
                const float2 dx = __fadd2_rn(make_float2(v.x, v.y), nx);
                const float2 dy = __fadd2_rn(make_float2(v.z, v.w), ny);
                const float2 s = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, nthr2));
                bits = __funnelshift_l(__float_as_uint(s.x), bits, 1);
                bits = __funnelshift_l(__float_as_uint(s.y), bits, 1);
            };
            if (qe == 16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) body(q);
                bits = __brev(bits);
            } else {
                for (int q = 0; q < qe; ++q) body(q);
                bits = __brev(bits) >> (32 - 2 * qe);            // 0 < qe < 16
            }
            mrow[w] = bits;
            cnt += __popc(bits);
        }
        // phase B on this tile's candidates (x_j through L2); neighbours go to the row's list
        {
            int w = -1;
            unsigned cand = 0;
            for (int c = 0; c < cnt; ++c) {
                while (cand == 0) { ++w; cand = mrow[w]; }
                const int k = __ffs(cand) - 1;
                cand &= cand - 1;
                const int kk = w * 32 + k;
                const int j = 2 * (s0 + sub + LANES * (kk >> 1)) + (kk & 1);
                if (j == i || j >= N) continue;
                const double4 xj = ld_state(xc + (size_t)j * 4);
                if constexpr (KNN) {
                    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                    top.insert(fma(dx, dx, dy * dy), j);
                }
                if (refine_fast<sizeof(TO) == 4>(xi, xj, a.R2, a.Rm, A)) {
                    const int pos = atomicAdd(&lcnt[rl], 1);
                    if (pos < a.CAP) lrow[pos] = (unsigned)j;
                }
            }
        }
        __syncthreads();                                     // everyone is done with stage t&1
        if (tid == 0 && t + 2 < ntiles) issue_tile(t + 2);
    }

    // ---- reduce across the LANES threads of a row ----------------------------------------------------------------------
#pragma unroll
    for (int m = LANES / 2; m > 0; m >>= 1) {
        A.f0 += shfl_xor(A.f0, m); A.f1 += shfl_xor(A.f1, m); A.f2 += shfl_xor(A.f2, m);
        A.f3 += shfl_xor(A.f3, m); A.f4 += shfl_xor(A.f4, m); A.f5 += shfl_xor(A.f5, m);
        A.gx += shfl_xor(A.gx, m); A.gy += shfl_xor(A.gy, m);
    }
    const int deg = valid ? lcnt[rl] : 0;                    // complete: the last tile ended with a block barrier
    TO idg;
    if constexpr (sizeof(TO) == 4) idg = __frcp_rn((float)max(deg, 1));
    else idg = Tr<double>::rcp((double)max(deg, 1));

    if (valid && sub == 0) {
        if (a.values) {
            TO *v = reinterpret_cast<TO *>(a.values) + grow * NFE;
            v[0] = (TO)A.f0; v[1] = (TO)A.f1; v[2] = (TO)A.f2; v[3] = (TO)A.f3; v[4] = (TO)A.f4; v[5] = (TO)A.f5;
        }
        if (a.u_next) {
            double sdx = A.f0, sdy = A.f3;
            if (cen) {                                       // sum_j (v_i - v_j) = N*(v_i - ref) - sum_j (v_j - ref)
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

    // ---- kNN observation ------------------------------------------------------------------------------------------------
    if constexpr (KNN) {
        int nb[TOPK];
        auto merge = [&](int *out) {                         // every lane of the row ends with the row's sorted ids
#pragma unroll
            for (int m = 0; m < TOPK; ++m) {
                double bd = top.d[0]; int bi = top.id[0];
#pragma unroll
                for (int s = LANES / 2; s > 0; s >>= 1) {
                    const double od = shfl_xor(bd, s); const int oi = shfl_xor(bi, s);
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
        bool need = valid && (nb[KNN_K - 1] == i);
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
        if (valid && sub == 0 && a.obs) {
            TO *o = reinterpret_cast<TO *>(a.obs) + grow * (NXS * KNN_K);
#pragma unroll
            for (int m = 0; m < KNN_K; ++m) {                 // the diagonal is never a candidate here; unused slots hold i
                const double4 xn = ld_state(xc + (size_t)min(nb[m], N - 1) * 4);
                o[m * 4 + 0] = (TO)(xi.x - xn.x); o[m * 4 + 1] = (TO)(xi.y - xn.y);
                o[m * 4 + 2] = (TO)(xi.z - xn.z); o[m * 4 + 3] = (TO)(xi.w - xn.w);
            }
        }
    }

    // ---- scatter the non-zero adj_mean entries -------------------------------------------------------------------------
    if (net == nullptr) return;
    __syncwarp();                                            // the warp's zero stores precede its non-zero stores
    if (valid) {
        TO *orow = net + grow * N;
        if (deg <= a.CAP) {
            for (int c = sub; c < deg; c += LANES) orow[lrow[c]] = idg;
        } else {                                             // rare: denser than the list; recompute the row's adjacency
            for (int j = sub; j < N; j += LANES) {
                if (j == i) continue;
                const double2 pj = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)j * 4));
                const double dx = xi.x - pj.x, dy = xi.y - pj.y;
                if (fma(dx, dx, dy * dy) < a.R2) orow[j] = idg;
            }
        }
    }
}

// ---- kNN fix-up: one CTA per queued row rescans all agents (exact fp64 r2, ties to the lower index) -----------------
template <typename TO>
__global__ void __launch_bounds__(256) flock_knn_fix_kernel(const SArgs a) {
    const int env = blockIdx.y;
    const unsigned cnt = min(a.needy_cnt[env], (unsigned)NEEDY_CAP);
    __shared__ double sd[8 * TOPK];
    __shared__ int si[8 * TOPK];
    const int N = a.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *xc = a.x_cur + (size_t)env * N * 4;
    for (unsigned q = blockIdx.x; q < cnt; q += gridDim.x) {
        const int i = (int)a.needy_rows[(size_t)env * NEEDY_CAP + q];
        const double4 xi = ld_state(xc + (size_t)i * 4);
        TopK<double> top;
        top.init(i);
        for (int j = tid; j < N; j += 256) {
            if (j == i) continue;
            const double2 pj = __ldg(reinterpret_cast<const double2 *>(xc + (size_t)j * 4));
            const double dx = xi.x - pj.x, dy = xi.y - pj.y;
            top.insert(fma(dx, dx, dy * dy), j);
        }
#pragma unroll
        for (int m = 0; m < TOPK; ++m) {                     // warp merge: the warp's TOPK best, in order
            double bd = top.d[0]; int bi = top.id[0];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const double od = shfl_xor(bd, s); const int oi = shfl_xor(bi, s);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (bi == top.id[0] && bd == top.d[0]) top.pop_front(i);
            if (lane == 0) { sd[warp * TOPK + m] = bd; si[warp * TOPK + m] = bi; }
        }
        __syncthreads();
        if (tid == 0 && a.obs) {                             // 64 entries: one thread finishes the merge
            TopK<double> fin;
            fin.init(i);
            for (int e = 0; e < 8 * TOPK; ++e) if (si[e] != i) fin.insert(sd[e], si[e]);
            // equal r2: insert keeps the earlier entry in front; entries arrive in warp order, not index order
            TO *o = reinterpret_cast<TO *>(a.obs) + ((size_t)env * N + i) * (NXS * KNN_K);
            for (int m = 0; m < KNN_K; ++m) {
                const double4 xn = ld_state(xc + (size_t)fin.id[m] * 4);
                o[m * 4 + 0] = (TO)(xi.x - xn.x); o[m * 4 + 1] = (TO)(xi.y - xn.y);
                o[m * 4 + 2] = (TO)(xi.z - xn.z); o[m * 4 + 3] = (TO)(xi.w - xn.w);
            }
        }
        __syncthreads();
    }
}

}  // namespace flock
