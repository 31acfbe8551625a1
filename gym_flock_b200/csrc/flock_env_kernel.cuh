// flock_env_kernel.cuh -- the env-resident fused flocking step (sm_100a), "filter then refine".
//
// A CTA owns a WINDOW of consecutive rows (agents i) of the flattened [E x rows] space -- small envs are packed
// back to back so every lane has a row; large envs are cut into row tiles -- and stages the full state x[N,4] of
// the environments its rows belong to in shared memory.  Each ordered pair (i,j) is visited exactly once per
// step.  After ONE block barrier (state staged) every warp runs on its own:
//
//   prologue  : x' = integrate(x_in, u_in)                          flocking_relative.py:96-105
//               reward = -(var vx' + var vy')                       flocking_relative.py:145-147
//   phase A   : FILTER over all N^2 ordered pairs, candidate bits (r2 <= max(R, R^2)) kept in shared memory, word-major.
//               fp32 output mode: packed f32x2 math on fp32 positions relative to a per-env origin; the test is the
//               SIGN of dx^2 + (dy^2 - thr) with a conservative per-row threshold, shifted into the word (no compare).
//               fp64 modes: fp64 r2 (exact mode: NumPy's unfused r2).         flocking_relative.py:113-117
//               Flocking-v0: a running top-8 by r2 rides along.     flocking.py:20-24
//   phase B   : REFINE only the candidates (~deg of N) in fp64, branch-free in the fast modes: exact adjacency bit
//               (cleared in place when the candidate is no neighbour), the 6 masked feature sums, the potential
//               gradient.                                           flocking_relative.py:117-129,214-226
//               EXACT: NumPy's operation order, separately rounded, IEEE divide, ascending j.
//   epilogue  : values, u_next (clip, /10), obs (staged per warp, stored coalesced).   flocking_relative.py:120-122,209-211
//   dense rows: one thread per row and full rows: every warp expands its 32 rows of adj_mean from the exact bit masks
//               with coalesced 16-byte streaming stores -- every byte written exactly once (partial-sector writes,
//               not the store path, are what made "zero-fill + scatter" slow).  Otherwise (sharded rows, several
//               lanes per row): zero-fill with the same stores up front, the non-zeros stored over them at the end.
//
// Non-neighbour pairs contribute exact zeros to every masked sum of the reference, so skipping them in
// phase B changes nothing (not even in EXACT mode, up to the sign of a zero).  The centralized alignment
// term sum_j (v_i - v_j) runs over ALL j: EXACT accumulates it sequentially in phase A; the fast modes use
// N*(v_i - ref) - sum_j (v_j - ref) in fp64 (O(N), sums gathered while staging).
//
// LANES threads cooperate on one row (pairs interleaved across them, warp-shuffle reduction) so small
// batches still fill 148 SMs.
#pragma once
#include <type_traits>
#ifndef FLOCK_MAXT
#define FLOCK_MAXT 256
#endif
#ifndef FLOCK_MINB
#define FLOCK_MINB 4   // resident 256-thread CTAs per SM the fast kernels are compiled for (register cap)
#endif
#include "flock_common.cuh"

namespace flock {

struct AuxOut {               // optional extra outputs (FlockAux of the C ABI); all nullable
    int *nonfinite;           // [E] set to 1 when a feature / action of the env is NaN or Inf
    unsigned *adj_bits;       // [E][N][WB] exact adjacency bits: bit (j & 31) of word j >> 5 (one thread per row only)
    int *degree;              // [E][N]
    void *min_dist;           // [E][N] TO  get_stats()['min_dists']   flocking_relative.py:141-142
    void *vel_diff;           // [E][N] TO  get_stats()['vel_diffs']   flocking_relative.py:138-140
    void *knn_hint;           // [E][N][8] u16 in/out: the row's 7 nearest of the previous step (a search-radius hint only)
    int WB;                   // words per row: ceil(N / 32)
};

struct KArgs {
    AuxOut aux;
    const double *x_in;
    const double *u_in;
    double *x_out;
    double *u_next;
    void *values;
    void *network;
    void *reward;
    void *obs;
    int E, N;
    int RT;       // rows per CTA window
    int tiles;    // 0: flat windows over E*nre rows; > 0: row tiles per env (windows never cross envs)
    int GMAX;     // max envs a window touches (staging capacity)
    int row_lo, row_hi;
    int centralized;
    int W;        // 32-bit candidate words per (row, lane)
    int NP2;      // pair slots per env in the packed fp32 position array: (N + 1) / 2
    double R, R2, dt, ascale;
    double inv_ascale;   // 1/ascale (fast modes multiply; EXACT divides)
    double ulim;         // fast modes: the action clamp after scaling, min(10/ascale, u_clip)
    StepVariant var;     // flocking variants (leader / stochastic); defaults reproduce FlockingRelative
    double Rm;    // fast modes: gradient mask as ONE compare r2 <= Rm (centralized: R; decentralized: also r2 < R2)
    double thrC;  // candidate threshold max(R, R2)
    unsigned off_xs, off_pf, off_masks, off_stats, off_idg, off_obs;
    int out_mode; // dense adjacency output: 0 = zero-fill up front + the non-zeros stored over it; 1 = in-warp row expansion
    int tune;     // diagnostics: bit0 skip fill, bit1 skip scatter, bit2 skip refine, bit3 skip filter (bit4: host, out_mode 0),
                  // bit6 scatter the non-zeros as 8-byte stores instead of whole sectors
    unsigned off_zero, zero_bytes;   // out_mode 0, exact order, full rows: a block of zeros in shared memory that the warps'
                                     // row runs are filled from with 1-D TMA bulk stores (0: plain streaming stores)
};

// ---- 1-D TMA bulk stores shared -> global (the async proxy: the issuing warp never waits in the LSU's store queue) -------
__device__ __forceinline__ void bulk_store(void *dst_global, const unsigned src_shared, const unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(__cvta_generic_to_global(dst_global)), "r"(src_shared), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }   // writes complete
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

constexpr int MS = FLOCK_MAXT;     // mask words are stored word-major with a fixed stride: word w of thread t at masks[w * MS + t]

// one agent's state from shared memory by 32-bit shared address (two LDS.128, no generic-address arithmetic)
__device__ __forceinline__ double4 lds_state(const unsigned saddr) {
    double4 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%4];\n\tld.shared.v2.f64 {%2, %3}, [%4+16];"
                 : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "r"(saddr));
    return v;
}

__device__ __forceinline__ int bfind(const unsigned v) {        // index of the highest set bit (FLO)
    int k;
    asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(v));
    return k;
}

// shared-memory double accumulate (a handful per env and CTA; CAS loop is fine)
__device__ __forceinline__ void smem_add(double *p, double v) { atomicAdd(p, v); }

constexpr int STATS_PER_ENV = 4;   // sum vx, sum vy over the env (fp64), ready flag, pad
constexpr int TOPK = KNN_K + 2;    // the diagonal (r2 = 0) rides along, plus one slack slot: the fast modes select with
                                   // approximate keys and re-rank the TOPK survivors with the reference's own keys

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

// Final ranking of the kNN candidates with the REFERENCE's keys: r2 = dx*dx + dy*dy, every operation rounded separately
// (flocking_relative.py:114), ties to the lower index.  `id` holds TOPK agent indices found with approximate keys (fp32 /
// fused fp64); the row's own index (the diagonal, or an unused slot) sorts last.  The approximate order is almost always
// right, so check first and run the insertion network only when some lane of the warp is out of order.
// Returns the exact r2 of the KNN_K-th neighbour (+inf when the row has fewer).
template <typename XE>
__device__ __forceinline__ double knn_rerank(const double4 xi, const int i, const XE &xe, int (&id)[TOPK]) {
    double key[TOPK];
#pragma unroll
    for (int m = 0; m < TOPK; ++m) {
        const double4 xj = xe[id[m]];
        const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
        key[m] = (id[m] == i) ? CUDART_INF : __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    }
    // the row itself (r2 = 0) leads the approximate ranking when the caller let the diagonal ride along: rotate it to the end
    // (it sorts last) instead of paying for the network on every row
    const bool lead = (id[0] == i);
#pragma unroll
    for (int m = 0; m + 1 < TOPK; ++m) { id[m] = lead ? id[m + 1] : id[m]; key[m] = lead ? key[m + 1] : key[m]; }
    id[TOPK - 1] = lead ? i : id[TOPK - 1];
    key[TOPK - 1] = lead ? CUDART_INF : key[TOPK - 1];
    bool sorted = true;
#pragma unroll
    for (int m = 0; m + 1 < TOPK; ++m)
        sorted = sorted && (key[m] < key[m + 1] || (key[m] == key[m + 1] && id[m] <= id[m + 1]));
    if (__any_sync(__activemask(), !sorted)) {
#pragma unroll
        for (int p = 1; p < TOPK; ++p) {
#pragma unroll
            for (int q = p; q > 0; --q) {
                const bool sw = key[q] < key[q - 1] || (key[q] == key[q - 1] && id[q] < id[q - 1]);
                const double ka = sw ? key[q] : key[q - 1], kb = sw ? key[q - 1] : key[q];
                const int ia = sw ? id[q] : id[q - 1], ib = sw ? id[q - 1] : id[q];
                key[q - 1] = ka; key[q] = kb; id[q - 1] = ia; id[q] = ib;
            }
        }
    }
    return key[KNN_K - 1];
}

// 1/deg in the output type, shared by every kernel so that row shards, lanes and kernels agree bit for bit.
__device__ __forceinline__ float inv_deg(const int deg, float) {
    const float d = (float)max(deg, 1);
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    const float e = fmaf(-d, y, 1.0f);         // one Newton step on the MUFU seed
    return fmaf(y, e, y);
}
__device__ __forceinline__ double inv_deg(const int deg, double) { return Tr<double>::rcp((double)max(deg, 1)); }

struct Acc {
    double f0, f1, f2, f3, f4, f5;   // masked feature sums
    double gx, gy;                   // EXACT: sum of grad; fast: sum d*(1/r2 - 1/r4) (x2 at the end)
    double s0, s1;                   // EXACT alignment sums (centralized: all j, phase A; else masked, phase B)
};

// EXACT refine of one candidate pair: Appendix B of SURVEY.md, one rounding per operation, NumPy's order.
// novel: the pair involves an obstacle (FlockingObstacle), its velocity difference is zero (flocking_obstacle.py:81-82)
__device__ __forceinline__ bool refine_exact(const double4 xi, const double4 xj, const double R, const double R2,
                                             const bool cen, Acc &A, const bool novel = false) {
    const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
    const double dvx = novel ? 0.0 : __dsub_rn(xi.z, xj.z), dvy = novel ? 0.0 : __dsub_rn(xi.w, xj.w);
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

// Fast refine (fp64, fused): one reciprocal instead of four divides.  Candidates are a superset (conservative fp32
// filter, possibly an inflated kNN radius), so pairs beyond both radii leave early.
// NEWTON1 (fp32 output mode): one Newton step on the 2^-23 seed (relative error ~2^-46, far below float rounding).
template <bool NEWTON1 = false>
__device__ __forceinline__ bool refine_fast(const double4 xi, const double4 xj, const double R2, const double Rm,
                                            Acc &A) {
    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
    const double r2 = fma(dx, dx, dy * dy);
    const bool a = r2 < R2;
    const bool g = r2 <= Rm;
    if (a || g) {
        const double inv2 = NEWTON1 ? Tr<double>::rcp1(r2) : Tr<double>::rcp(r2);
        if (a) {
            const double inv4 = inv2 * inv2;
            A.f0 += xi.z - xj.z; A.f3 += xi.w - xj.w;
            A.f1 = fma(dx, inv4, A.f1); A.f2 = fma(dx, inv2, A.f2);
            A.f4 = fma(dy, inv4, A.f4); A.f5 = fma(dy, inv2, A.f5);
        }
        if (g) {
            const double c = fma(-inv2, inv2, inv2);   // 1/r2 - 1/r4
            A.gx = fma(dx, c, A.gx); A.gy = fma(dy, c, A.gy);
        }
    }
    return a;
}

// Fast refine of one candidate as straight-line, branch-free code (the lanes of a warp disagree on almost every
// condition, and every divergent region costs its BSSY/BSYNC pair and re-issued instructions).  Same arithmetic as
// refine_fast -- one reciprocal (seed + Newton) instead of four divides, fused multiply-adds, fp64 throughout -- with the
// two masks applied as exact factors 1.0 / 0.0 (x * 1.0 and + 0.0 are exact, so the sums are bit-identical to the
// branchy form).  The caller passes a valid agent address even when the lane idles (low == 0): pv masks everything.
//   !adjacent  : the candidate's bit is cleared in its mask word (shared-memory RED at waddr)
#define FLOCK_REFINE_PRED_ASM(SECOND_NEWTON_STEP, LD_VEL, EXTRA_BODY, EXTRA_IN)                                   \
    asm volatile(                                                                                                   \
        "{\n\t"                                                                                                     \
        ".reg .pred pv, pa, pg, pr;\n\t"                                                                            \
        ".reg .b32 nl, hi;\n\t"                                                                                     \
        ".reg .f64 xjx, xjy, xjz, xjw, dx, dy, r2, y, e, ya, y4, yg, dvx, dvy, c;\n\t"                      \
        "setp.ne.u32 pv, %11, 0;\n\t"                                                                            \
        "ld.shared.v2.f64 {xjx, xjy}, [%12];\n\t"                                                              \
        LD_VEL                                                                                                      \
        "sub.f64 dx, %14, xjx;\n\t"                                                                              \
        "sub.f64 dy, %15, xjy;\n\t"                                                                              \
        "mul.f64 r2, dy, dy;\n\t"                                                                                   \
        "fma.rn.f64 r2, dx, dx, r2;\n\t"                                                                            \
        EXTRA_BODY                                                                                                  \
        "setp.lt.and.f64 pa, r2, %18, pv;\n\t"  /* adjacency: r2 < R^2 (squared radius) */                        \
        "setp.le.and.f64 pg, r2, %19, pv;\n\t"  /* gradient mask: r2 <= Rm */                                     \
        "not.pred pr, pa;\n\t"                                                                                      \
        "and.pred pr, pr, pv;\n\t"              /* a candidate that is not a neighbour */                         \
        "rcp.approx.ftz.f64 y, r2;\n\t"                                                                             \
        "neg.f64 e, r2;\n\t"                                                                                        \
        "fma.rn.f64 c, e, y, 0d3FF0000000000000;\n\t"                                                               \
        "fma.rn.f64 y, y, c, y;\n\t"                                                                                \
        SECOND_NEWTON_STEP                                                                                          \
        "selp.f64 ya, y, 0d0000000000000000, pa;\n\t" /* the masks as selects: y * 1.0 and y * 0.0 (y finite) */    \
        "selp.f64 yg, y, 0d0000000000000000, pg;\n\t"                                                               \
        "mul.f64 y4, ya, ya;\n\t"                                                                                   \
        "sub.f64 dvx, %16, xjz;\n\t"                                                                             \
        "sub.f64 dvy, %17, xjw;\n\t"                                                                             \
        "@pa add.f64 %0, %0, dvx;\n\t"         /* = fma(dvx, 1.0, f0); + dvx * 0.0 otherwise */                  \
        "fma.rn.f64 %1, dx, y4, %1;\n\t"                                                                      \
        "fma.rn.f64 %2, dx, ya, %2;\n\t"                                                                      \
        "@pa add.f64 %3, %3, dvy;\n\t"                                                                         \
        "fma.rn.f64 %4, dy, y4, %4;\n\t"                                                                      \
        "fma.rn.f64 %5, dy, ya, %5;\n\t"                                                                      \
        "neg.f64 c, yg;\n\t"                                                                                        \
        "fma.rn.f64 c, c, yg, yg;\n\t" /* 1/r2 - 1/r4 */                                                            \
        "fma.rn.f64 %6, dx, c, %6;\n\t"                                                                       \
        "fma.rn.f64 %7, dy, c, %7;\n\t"                                                                       \
        "@pa add.s32 %8, %8, 1;\n\t"                                                                        \
        "@pv add.s32 %9, %9, -1;\n\t"                                                                     \
        "not.b32 nl, %11;\n\t"                                                                                   \
        "@pr red.shared.and.b32 [%13], nl;\n\t"                                                                \
        "}"                                                                                                         \
        : "+d"(A.f0), "+d"(A.f1), "+d"(A.f2), "+d"(A.f3), "+d"(A.f4), "+d"(A.f5), "+d"(A.gx), "+d"(A.gy), "+r"(deg),  \
          "+r"(left), "=f"(r2f) /* %10: the pair's r2 as an fp32 key, written by EXTRA_BODY only */                  \
        : "r"(low), "r"(xaddr), "r"(waddr), "d"(xi.x), "d"(xi.y), "d"(xi.z), "d"(xi.w), "d"(R2), "d"(Rm) EXTRA_IN   \
        : "memory")

#define FLOCK_NEWTON2 "fma.rn.f64 c, e, y, 0d3FF0000000000000;\n\tfma.rn.f64 y, y, c, y;\n\t"
#define FLOCK_NOTHING

template <bool NEWTON1>
__device__ __forceinline__ void refine_fast_pred(const double4 xi, const unsigned xaddr, const unsigned low, const unsigned waddr,
                                                 const double R2, const double Rm, Acc &A, int &deg, int &left) {
    [[maybe_unused]] float r2f;
    if constexpr (NEWTON1) {            // fp32 output mode: seed (2^-23) + one Newton step
        FLOCK_REFINE_PRED_ASM("", "ld.shared.v2.f64 {xjz, xjw}, [%12+16];\n\t", "", FLOCK_NOTHING);
    } else {                            // fp64 output mode: two Newton steps (full double accuracy)
        FLOCK_REFINE_PRED_ASM(FLOCK_NEWTON2, "ld.shared.v2.f64 {xjz, xjw}, [%12+16];\n\t", "", FLOCK_NOTHING);
    }
}

// the same with positions and velocities in two separate shared-memory arrays (flock_group_kernel): xaddr -> (px, py),
// vaddr -> (vx, vy)
#define FLOCK_VADDR_IN , "r"(vaddr)
template <bool NEWTON1>
__device__ __forceinline__ void refine_fast_pred2(const double4 xi, const unsigned xaddr, const unsigned vaddr, const unsigned low,
                                                  const unsigned waddr, const double R2, const double Rm, Acc &A, int &deg, int &left) {
    [[maybe_unused]] float r2f;
    if constexpr (NEWTON1) {
        FLOCK_REFINE_PRED_ASM("", "ld.shared.v2.f64 {xjz, xjw}, [%20];\n\t", "", FLOCK_VADDR_IN);
    } else {
        FLOCK_REFINE_PRED_ASM(FLOCK_NEWTON2, "ld.shared.v2.f64 {xjz, xjw}, [%20];\n\t", "", FLOCK_VADDR_IN);
    }
}

// ... and returning the pair's r2 as an fp32 selection key for Flocking-v0's kNN (0xFFFFFFFF when the lane idles): the refine
// loop already has the exact r2 of every candidate, so the kNN search rides along instead of running its own pass
#define FLOCK_R2KEY_BODY "cvt.rn.f32.f64 %10, r2;\n\t"
template <bool NEWTON1>
__device__ __forceinline__ float refine_fast_pred2_key(const double4 xi, const unsigned xaddr, const unsigned vaddr, const unsigned low,
                                                       const unsigned waddr, const double R2, const double Rm, Acc &A, int &deg,
                                                       int &left) {
    float r2f;
    if constexpr (NEWTON1) {
        FLOCK_REFINE_PRED_ASM("", "ld.shared.v2.f64 {xjz, xjw}, [%20];\n\t", FLOCK_R2KEY_BODY, FLOCK_VADDR_IN);
    } else {
        FLOCK_REFINE_PRED_ASM(FLOCK_NEWTON2, "ld.shared.v2.f64 {xjz, xjw}, [%20];\n\t", FLOCK_R2KEY_BODY, FLOCK_VADDR_IN);
    }
    return r2f;
}

// ---- in-warp expansion of the warp's rows of adj_mean (small rows) -----------------------------------------------
// The warp's rows are consecutive rows of the [E*N, N] output.  A row is cut into nv = N/V groups of V elements
// (16 bytes when V is maximal); a lane owns one group column (and, when rows are short, one of RPI rows handled per
// iteration), so its mask word and shift are loop invariants.  Mask words are stored word-major ([word][thread]) and
// 1/deg of every row sits in shared memory, so in the unrolled loop both loads use immediate offsets: per row two LDS,
// the bit -> value selects and one predicated, coalesced streaming store.  No partial sectors, nothing to wait for.
template <typename TO, int V>
__device__ __forceinline__ void store_group(TO *dst, const unsigned bits, const TO v, const bool on) {
    const unsigned p = on ? 1u : 0u;
    if constexpr (sizeof(TO) == 4 && V == 4) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                     :: "l"(dst), "f"((bits & 1u) ? v : 0.f), "f"((bits & 2u) ? v : 0.f), "f"((bits & 4u) ? v : 0.f),
                        "f"((bits & 8u) ? v : 0.f), "r"(p) : "memory");
    } else if constexpr (sizeof(TO) == 4 && V == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.cs.v2.f32 [%0], {%1, %2};\n\t}"
                     :: "l"(dst), "f"((bits & 1u) ? v : 0.f), "f"((bits & 2u) ? v : 0.f), "r"(p) : "memory");
    } else if constexpr (sizeof(TO) == 8 && V == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.cs.v2.f64 [%0], {%1, %2};\n\t}"
                     :: "l"(dst), "d"((bits & 1u) ? v : 0.0), "d"((bits & 2u) ? v : 0.0), "r"(p) : "memory");
    } else {
        if (on) __stcs(dst, (bits & 1u) ? v : TO(0));
    }
}

// wm: mask word 0 of the warp's first row (word q of row r at wm[q * MS + r]); idgs: 1/deg of the warp's rows
template <typename TO, int V>
__device__ __forceinline__ void expand_rows(TO *wout, const unsigned *wm, const TO *idgs, const int N,
                                            const int nrows, const int lane) {
    const int nv = N / V;                               // groups per row
    if (nv > 16 && nv <= 32 && nrows == 32) {           // a full warp of rows, one row per iteration: fixed trip count
        const bool on = lane < nv;
        const int c = on ? lane * V : 0;
        const int sh = c & 31;
        const unsigned *mp = wm + (c >> 5) * MS;
        TO *dst = wout + c;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            store_group<TO, V>(dst, mp[r] >> sh, idgs[r], on);
            dst += N;
        }
        return;
    }
    const int passes = (nv + 31) >> 5;                  // > 1 when a row has more than 32 groups
    const int rpi = (nv >= 32) ? 1 : 32 / nv;           // rows per iteration when rows are short
    const int sr = (nv >= 32) ? 0 : lane / nv;          // this lane's sub-row within an iteration
    const int cv0 = (nv >= 32) ? lane : lane - sr * nv; // this lane's group column (first pass)
    const bool lane_on = (nv >= 32) || (sr < rpi);
    if (passes == 1) {                                  // short rows (several per iteration) or a partial warp
        const bool on = lane_on && cv0 < nv;
        const int c = on ? cv0 * V : 0;
        const int sh = c & 31;
        const unsigned *mp = wm + (c >> 5) * MS + (on ? sr : 0);
        const TO *ip = idgs + (on ? sr : 0);
        TO *dst = wout + (size_t)(on ? sr : 0) * N + c;
        const size_t dstep = (size_t)rpi * N;
        const int last = nrows - sr;                    // this lane stores while r < last
        for (int r = 0; r < nrows; r += rpi) {
            const bool go = on && r < last;
            store_group<TO, V>(dst, go ? (mp[r] >> sh) : 0u, go ? ip[r] : TO(0), go);
            dst += dstep;
        }
    } else {
        // long rows (more than 32 groups): the lane's groups are cv = lane + 32 t, so its bit offset within a mask word is
        // a constant and its word advances by V per step: per group one LDS, the selects and one coalesced 16-byte store
        // (a warp-wide store covers 512 contiguous bytes of the row).  Every byte of the row is written exactly once.
        const int sh = (lane * V) & 31;
        const unsigned *mp0 = wm + ((lane * V) >> 5) * MS;
        for (int r = 0; r < nrows; ++r) {
            const TO v = idgs[r];
            const unsigned *mp = mp0 + r;
            TO *dst = wout + (size_t)r * N + lane * V;
#pragma unroll 4
            for (int cv = lane; cv < nv; cv += 32) {
                store_group<TO, V>(dst, *mp >> sh, v, true);
                mp += V * MS; dst += 32 * V;
            }
        }
    }
}

// ---- one window of rows: [r0, r1) of the flat [E x rows] space and the envs [e0, e0 + nenv) it touches ---------------
struct Win { int r0, r1, e0, nenv; };

__device__ __forceinline__ Win window_of(const KArgs &a, const int w) {
    const int nre = a.row_hi - a.row_lo;
    Win W;
    if (a.tiles > 0) {
        const int e = w / a.tiles, t = w - e * a.tiles;
        W.r0 = e * nre + t * a.RT;
        W.r1 = min(W.r0 + a.RT, (e + 1) * nre);
    } else {
        W.r0 = w * a.RT;
        W.r1 = min(W.r0 + a.RT, a.E * nre);
    }
    W.e0 = W.r0 / nre;
    W.nenv = (W.r1 - 1) / nre - W.e0 + 1;
    return W;
}

// the locals every phase of a window uses (unused ones cost nothing)
#define FLOCK_WIN_LOCALS                                                                                         \
    [[maybe_unused]] constexpr bool PA32 = std::is_same<TO, float>::value; /* phase A on packed fp32 */                \
    [[maybe_unused]] double4 *xs = reinterpret_cast<double4 *>(smem + a.off_xs);                                  \
    [[maybe_unused]] float4 *pf = reinterpret_cast<float4 *>(smem + a.off_pf); /* (px_2m, px_2m+1, py_2m, py_2m+1) - env origin */ \
    [[maybe_unused]] double *stats = reinterpret_cast<double *>(smem + a.off_stats);                              \
    [[maybe_unused]] const int tid = threadIdx.x, nthr = blockDim.x;                                              \
    [[maybe_unused]] const int lane = tid & 31;                                                                   \
    [[maybe_unused]] const int N = a.N;                                                                           \
    [[maybe_unused]] const int nre = a.row_hi - a.row_lo; /* rows per env handled by this call */                 \
    [[maybe_unused]] const int r0 = win.r0, r1 = win.r1, e0 = win.e0, nenv = win.nenv;                            \
    [[maybe_unused]] const float inv_n = 1.0f / (float)N, inv_nre = 1.0f / (float)nre;                            \
    [[maybe_unused]] TO *net = reinterpret_cast<TO *>(a.network);                                                 \
    [[maybe_unused]] TO *reward = reinterpret_cast<TO *>(a.reward);

// ---- phase 0 (all threads; the caller follows it with a block barrier): load, integrate, publish x', stage ---------
template <typename TO, bool EXACT>
__device__ __forceinline__ void stage_window(const KArgs &a, unsigned char *smem, const Win win) {
    FLOCK_WIN_LOCALS

    // ---- stats flags ----------------------------------------------------------------------------------------------------
    for (int q = tid; q < nenv; q += nthr) stats[q * STATS_PER_ENV + 2] = 0.0;
    if constexpr (EXACT) {      // the zero block the bulk stores read (generic writes -> async-proxy reads: proxy fence)
        if (a.zero_bytes) {
            float4 *z = reinterpret_cast<float4 *>(smem + a.off_zero);
            for (unsigned q = (unsigned)tid; q < a.zero_bytes / 16u; q += (unsigned)nthr) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            fence_proxy_async();
        }
    }

    // ---- prologue: load, integrate, publish x', stage (one flat loop over the agents of the staged envs) -----------
    {
        const int nag = nenv * N;
        const size_t a0 = (size_t)e0 * N;
        // the plain step (FlockingRelative / Flocking) keeps the loop body short; the variants take the general path
        const bool plain = (a.var.u_clip == 0.0) && (a.var.n_leaders == 0) && (a.var.n_obstacles == 0) && (a.var.state_scale == 1.0);
        struct Ag { double4 x; double2 u, ref; int ge, j; };
        // all loads first (state, action, the env's origin) ...
        auto fetch = [&](const int idx, Ag &g) {
            const size_t ga = a0 + idx;
            g.ge = (nenv > 1) ? (int)(((float)idx + 0.5f) * inv_n) : 0;              // exact: idx < 2^20
            g.j = idx - g.ge * N;
            g.x = ld_state(a.x_in + ga * 4);
            g.u = make_double2(0.0, 0.0);
            if (a.u_in) g.u = __ldg(reinterpret_cast<const double2 *>(a.u_in) + ga);
            g.ref = make_double2(0.0, 0.0);
            if constexpr (PA32) g.ref = __ldg(reinterpret_cast<const double2 *>(a.x_in + (a0 + (size_t)g.ge * N) * 4));
        };
        // ... then the arithmetic and the stores
        auto finish = [&](const int idx, Ag &g) {
            const size_t ga = a0 + idx;
            double4 x = g.x;
            const int ge = g.ge, j = g.j;
            if (a.u_in) {
                if (plain) x = integrate_exact(x, g.u, a.dt, a.var.step_scale);
                else x = integrate_variant(x, g.u, j, a.dt, a.var);
            }
            // every agent's new state is published exactly once: by the CTA that owns local row (j mod rows-per-env)
            // of its env (with full rows that is simply the CTA owning agent j's row)
            const int frj = (e0 + ge) * nre + ((nre == N) ? j : j % nre);
            if (a.x_out && frj >= r0 && frj < r1) st_state(a.x_out + ga * 4, x);
            xs[idx] = x;
            if constexpr (PA32) {
                float *p = reinterpret_cast<float *>(pf + (size_t)ge * a.NP2 + (j >> 1));   // origin: agent 0 before the step
                p[j & 1] = (float)(x.x - g.ref.x); p[2 + (j & 1)] = (float)(x.y - g.ref.y);
                if ((N & 1) && j == N - 1) { p[1] = CUDART_INF_F; p[3] = CUDART_INF_F; }   // odd N: pad slot at infinity
            }
        };
        // two agents per thread and iteration: both loads are in flight before the first use
        for (int idx = tid; idx < nag; idx += 2 * nthr) {
            const int idx1 = idx + nthr;
            const bool two = idx1 < nag;
            Ag g0, g1;
            fetch(idx, g0);
            if (two) fetch(idx1, g1);
            finish(idx, g0);
            if (two) finish(idx1, g1);
        }
    }
}

// ---- per-env velocity sums (alignment trick) and rewards: one warp per env, published through a flag ----------------
template <typename TO, bool EXACT>
__device__ __forceinline__ void window_stats(const KArgs &a, unsigned char *smem, const Win win) {
    FLOCK_WIN_LOCALS
    {
        const int warp = tid >> 5, nwarps = nthr >> 5;
        for (int ge = warp; ge < nenv; ge += nwarps) {
            const int env = e0 + ge;
            const bool owner = reward && (r0 <= env * nre) && (env * nre < r1);     // CTA holding the env's first row
            const double4 *xe = xs + ge * N;
            if constexpr (EXACT) {
                if (lane == 0 && (owner || a.aux.vel_diff != nullptr)) {   // NumPy's two-pass variance with sequential sums (:145-147)
                    double sx = 0.0, sy = 0.0;
                    for (int j = 0; j < N; ++j) { sx = __dadd_rn(sx, xe[j].z); sy = __dadd_rn(sy, xe[j].w); }
                    const double mx = __ddiv_rn(sx, (double)N), my = __ddiv_rn(sy, (double)N);
                    if (owner) {
                        double qx = 0.0, qy = 0.0;
                        for (int j = 0; j < N; ++j) {
                            const double ex = __dsub_rn(xe[j].z, mx), ey = __dsub_rn(xe[j].w, my);
                            qx = __dadd_rn(qx, __dmul_rn(ex, ex)); qy = __dadd_rn(qy, __dmul_rn(ey, ey));
                        }
                        reward[env] = __dmul_rn(-1.0, __dadd_rn(__ddiv_rn(qx, (double)N), __ddiv_rn(qy, (double)N)));
                    }
                    double *st = stats + ge * STATS_PER_ENV;      // np.mean(v, axis=0) for get_stats() (:138)
                    st[0] = mx; st[1] = my;
                    __threadfence_block();
                    reinterpret_cast<volatile double *>(st)[2] = 1.0;
                }
            } else {
                // sums about agent 0's velocity (a member of the flock: no cancellation problem)
                const double rx = xe[0].z, ry = xe[0].w;
                double sx = 0, sy = 0, qx = 0, qy = 0;
                for (int j = lane; j < N; j += 32) {
                    const double ex = xe[j].z - rx, ey = xe[j].w - ry;
                    sx += ex; sy += ey; qx = fma(ex, ex, qx); qy = fma(ey, ey, qy);
                }
                sx = warp_sum(sx); sy = warp_sum(sy);
                if (owner) {
                    const double q = warp_sum(qx + qy);
                    const double invN = 1.0 / (double)N;
                    if (lane == 0) reward[env] = (TO)(-(q - (sx * sx + sy * sy) * invN) * invN);   // -(var x + var y)
                }
                if (lane == 0) {
                    double *st = stats + ge * STATS_PER_ENV;
                    st[0] = fma((double)N, rx, sx); st[1] = fma((double)N, ry, sy);      // sum_j v_j
                    __threadfence_block();
                    reinterpret_cast<volatile double *>(st)[2] = 1.0;                     // ready
                }
            }
        }
    }

}

// ---- the rows of a window: filter, (kNN,) refine, row epilogue; leaves the exact adjacency bits in `masks` and
//      (small full rows, out_mode 1) 1/deg in `idg_s` for expand_window(); out_mode 0 scatters into the zero-filled rows.
//      From here on warps are independent.
template <typename TO, int LANES, bool EXACT, bool KNN>
__device__ __forceinline__ void compute_rows(const KArgs &a, unsigned char *smem, unsigned *masks, TO *idg_s, const Win win) {
    FLOCK_WIN_LOCALS
    using TK = typename std::conditional<PA32, float, double>::type;
    // ---- row assignment -----------------------------------------------------------------------------------
    const int rl = tid / LANES, sub = tid % LANES;
    const int fr = r0 + rl;                                 // flat row
    const bool valid = fr < r1;
    int g = 0;
    if (valid && nenv > 1) g = (int)(((float)(fr - e0 * nre) + 0.5f) * inv_nre);   // env of this row relative to e0 (exact)
    const int env = e0 + g;
    const int i = valid ? a.row_lo + (fr - env * nre) : 0;
    const bool cen = a.centralized != 0;
    const int nobs = a.var.n_obstacles;                     // FlockingObstacle: agents [0, nobs)
    const double4 *xe = xs + g * N;
    const double4 xi = xe[i];
    unsigned *mrow = masks + tid;                           // word-major: word w of this thread at mrow[w * MS]
    const size_t grow = (size_t)env * N + i;                // global row (agent) index

    // ---- out_mode 0: zero-fill of the warp's rows of the dense adjacency, up front -------------------------------------
    // The rows of a warp are consecutive flat rows; with full rows (nre == N) they are ONE contiguous run in memory,
    // otherwise one run per env they belong to.  Plain coalesced 16-byte streaming stores (a warp store covers 512
    // contiguous bytes); they drain to HBM while the warp computes, and the few non-zero entries are stored over them at
    // the end.  (1-D TMA bulk stores from a shared zero block throttle at ~3.6 TB/s; direct stores reach 6.3.)
    if (net && a.out_mode == 0 && !(a.tune & 1)) {
        const int wr0 = r0 + (tid - lane) / LANES;                           // first flat row of this warp
        const int wr1 = min(wr0 + 32 / LANES, r1);
        int s0 = wr0;
        while (s0 < wr1) {                                                   // runs (1 unless rows are sharded)
            const int senv = s0 / nre;
            const int s1 = (nre == N) ? wr1 : min(wr1, (senv + 1) * nre);
            const size_t base = ((size_t)senv * N + a.row_lo + (s0 - senv * nre)) * N;
            const unsigned total = (unsigned)(s1 - s0) * (unsigned)N;
            constexpr unsigned VEC = 16 / (unsigned)sizeof(TO);
            const unsigned headn = min((unsigned)((VEC - (unsigned)(base & (VEC - 1))) & (VEC - 1)), total);
            const unsigned nvec = (total - headn) / VEC;
            TO *out = net + base;
            if ((unsigned)lane < headn) out[lane] = TO(0);
            if ((unsigned)lane < total - (headn + nvec * VEC)) out[headn + nvec * VEC + lane] = TO(0);
            float4 *d4 = reinterpret_cast<float4 *>(out + headn);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned v = (unsigned)lane;
            if constexpr (EXACT) {
                // Exact order: the compute phase is long and LDS-heavy, and a warp that issues its run as STG sits in the
                // store queue's throttle -- with it every shared-memory load of the SM (measured at 1024 x 1000: step =
                // compute 1.09 ms + fill 1.41 ms, no overlap, wherever the stores were issued).  Bulk stores from a block
                // of zeros leave the LSU alone; the warp waits for them only before it stores the non-zeros.
                if (a.zero_bytes) {
                    if (lane == 0) {
                        char *dst = reinterpret_cast<char *>(d4);
                        const unsigned zs = (unsigned)__cvta_generic_to_shared(smem + a.off_zero);
                        for (size_t left = (size_t)nvec * 16u; left > 0;) {
                            const unsigned nb = (unsigned)min(left, (size_t)a.zero_bytes);
                            bulk_store(dst, zs, nb);
                            dst += nb; left -= nb;
                        }
                        bulk_commit();
                    }
                    s0 = s1;
                    continue;
                }
            }
            for (; v + 7u * 32u < nvec; v += 8u * 32u) {
#pragma unroll
                for (unsigned q = 0; q < 8u; ++q) __stcs(d4 + v + q * 32u, z4);
            }
            for (; v < nvec; v += 32u) __stcs(d4 + v, z4);
            s0 = s1;
        }
    }

    Acc A;
    A.f0 = A.f1 = A.f2 = A.f3 = A.f4 = A.f5 = A.gx = A.gy = A.s0 = A.s1 = 0.0;
    TopK<TK> top;
    // fp32 mode, one thread per row: the running top-8 is a branch-free min/max network on packed keys -- the fp32 bits of
    // r2 (non-negative, so they order as unsigned integers) with the low ceil(log2 N) mantissa bits replaced by the agent
    // index.  16 VIMNMX per pair, no divergence; ties and near-ties (r2 equal to 2^-(23 - bits) relative) resolve to the
    // lower index, the order NumPy's argsort gives equal keys.  (The compare-and-bubble insert cost ~74 instr/pair.)
    constexpr bool KPACK = KNN && PA32 && (LANES == 1);
    unsigned kp[TOPK];
    unsigned kidx = 1;                                      // index mask: next power of two >= N, minus 1
    if constexpr (KPACK) {
        while ((int)kidx < N) kidx <<= 1;
        kidx -= 1;
#pragma unroll
        for (int m = 0; m < TOPK; ++m) kp[m] = (0xFFFFFFFFu & ~kidx) | (unsigned)i;     // unused slot: "self", at +inf
    } else if constexpr (KNN) {
        top.init(i);
    }
    auto kp_insert = [&](unsigned k) {
#pragma unroll
        for (int m = 0; m < TOPK; ++m) { const unsigned lo = min(kp[m], k); k = max(kp[m], k); kp[m] = lo; }
    };

    // ---- phase A: filter ----------------------------------------------------------------------------------
    // Each lane owns bit positions k = 0.. of its words; j_of(k) maps them to agents.
    //   fp32 packed : pair slot m = sub + LANES*(k>>1), agent j = 2m + (k&1)
    //   fp64        : j = sub + LANES*k
    int nk;                                                 // bit positions used by this lane
    int cnt = 0;                                            // candidates found by this lane
    if constexpr (PA32) {
        const int nslots = valid ? (a.NP2 - sub + LANES - 1) / LANES : 0;
        nk = 2 * nslots;
        const float4 *pe = pf + (size_t)g * a.NP2;
        const float *pif = reinterpret_cast<const float *>(pe + (i >> 1));
        const float pxi = pif[i & 1], pyi = pif[2 + (i & 1)];
        const float2 nx = make_float2(-pxi, -pxi), ny = make_float2(-pyi, -pyi);
        // Conservative per-row fp32 threshold: every pair with true r2 <= thrC passes.  For such a pair
        // |p_j|inf <= |p_i|inf + rc (rc = sqrt(thrC)); each coordinate is rounded to <= |p|*2^-24, so
        // |d(dx)| <= (2|p_i|inf + 3rc)*2^-24 and |d(r2)| <= 4 rc |d(dx)| + a few ulp(r2).
        const float tc = (float)a.thrC, rc = sqrtf(tc);
        const float thrA = tc * (1.0f + 3e-6f) + 16.0f * rc * (fmaxf(fabsf(pxi), fabsf(pyi)) + 2.0f * rc) * 5.9604645e-8f;
        const float2 nthr2 = make_float2(-thrA, -thrA);
        for (int w = 0; w * 16 < ((a.tune & 8) ? 0 : nslots); ++w) {
            unsigned bits = 0;
            const int tbase = w * 16;
            const int tend = min(16, nslots - tbase);
            if constexpr (KNN) {
                auto body = [&](const int tt) {
                    const int m = sub + LANES * (tbase + tt);
                    const float4 q = pe[m];
                    const float2 dx = __fadd2_rn(make_float2(q.x, q.y), nx);
                    const float2 dy = __fadd2_rn(make_float2(q.z, q.w), ny);
                    const float2 r2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                    bits |= (r2.x < thrA) ? (1u << (2 * tt)) : 0u;
                    bits |= (r2.y < thrA) ? (2u << (2 * tt)) : 0u;
                    if constexpr (KPACK) {
                        const unsigned jw = (unsigned)(2 * tbase);                       // LANES == 1: m = tbase + tt
                        kp_insert(((__float_as_uint(r2.x) & ~kidx) | jw) + (unsigned)(2 * tt));
                        kp_insert(((__float_as_uint(r2.y) & ~kidx) | jw) + (unsigned)(2 * tt + 1));
                    } else {
                        top.insert(r2.x, 2 * m); top.insert(r2.y, 2 * m + 1);
                    }
                };
                if (tend == 16) {
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) body(tt);
                } else {
                    for (int tt = 0; tt < tend; ++tt) body(tt);
                }
            } else {
                // s = dx^2 + (dy^2 - thrA): the candidate test is the SIGN of s (two packed FFMA2, no compare); one
                // funnel shift per pair moves it into the word (first pair ends up highest: one BREV per word)
                auto body = [&](const int tt) {
                    const float4 q = pe[sub + LANES * (tbase + tt)];
                    const float2 dx = __fadd2_rn(make_float2(q.x, q.y), nx);
                    const float2 dy = __fadd2_rn(make_float2(q.z, q.w), ny);
                    const float2 s = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, nthr2));
                    bits = __funnelshift_l(__float_as_uint(s.x), bits, 1);
                    bits = __funnelshift_l(__float_as_uint(s.y), bits, 1);
                };
                if (tend == 16) {
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) body(tt);
                    bits = __brev(bits);
                } else {
                    for (int tt = 0; tt < tend; ++tt) body(tt);
                    bits = __brev(bits) >> (32 - 2 * tend);          // 0 < tend < 16
                }
            }
            mrow[w * MS] = bits;
            cnt += __popc(bits);
        }
    } else {
        nk = valid ? (N - sub + LANES - 1) / LANES : 0;
        const double thrC = a.thrC;
        [[maybe_unused]] const double thrX = (a.R2 <= a.R) ? a.R : __longlong_as_double(__double_as_longlong(a.R2) - 1);   // (R^2 > 0)
        // EXACT, centralized: the alignment sums run over ALL j in order (flocking_relative.py:204-208).  The obstacle test
        // (rows / columns of obstacles contribute exact zeros) is per pair only when the env has obstacles: three loop variants
        // picked by one warp-uniform branch.  SUMS 0: no sums (decentralized / fast modes), 1: every pair, 2: per-pair test.
        auto words = [&](auto SUMS) {
            [[maybe_unused]] constexpr int SM = decltype(SUMS)::value;
            for (int w = 0; w * 32 < nk; ++w) {
                unsigned bits = 0;
                const int kbase = w * 32;
                const int kend = min(32, nk - kbase);
                auto body = [&](const int kk) {
                    const int j = sub + LANES * (kbase + kk);
                    const double4 xj = xe[j];
                    double r2;
                    [[maybe_unused]] bool c;
                    if constexpr (EXACT) {
                        const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
                        r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        // candidate = (r2 < R^2) || !(r2 > R), the union of the two masks refine_exact applies: one compare,
                        // !(r2 > thrX) with thrX = R if R^2 <= R, else the double just below R^2 (NaN stays a candidate);
                        // the bit is one predicated OR
                        asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@!p or.b32 %0, %0, %3;\n\t}"
                            : "+r"(bits) : "d"(r2), "d"(thrX), "r"(1u << kk));
                        if (SM == 1 || (SM == 2 && !(i < nobs || j < nobs))) {
                            A.s0 = __dadd_rn(A.s0, __dsub_rn(xi.z, xj.z));
                            A.s1 = __dadd_rn(A.s1, __dsub_rn(xi.w, xj.w));
                        }
                    } else {
                        const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                        r2 = fma(dx, dx, dy * dy);
                        c = r2 <= thrC;
                        bits |= c ? (1u << kk) : 0u;
                    }
                    if constexpr (KNN) top.insert(r2, j);
                };
                if (kend == 32) {
                    if constexpr (EXACT) {
#pragma unroll
                        for (int kk = 0; kk < 32; ++kk) body(kk);
                    } else {
#pragma unroll 8
                        for (int kk = 0; kk < 32; ++kk) body(kk);
                    }
                } else {
                    for (int kk = 0; kk < kend; ++kk) body(kk);
                }
                mrow[w * MS] = bits;
                cnt += __popc(bits);
            }
        };
        if constexpr (EXACT) {
            if (!cen) words(std::integral_constant<int, 0>{});
            else if (nobs == 0) words(std::integral_constant<int, 1>{});
            else words(std::integral_constant<int, 2>{});
        } else {
            words(std::integral_constant<int, 0>{});
        }
    }
    const int nw = (nk + 31) >> 5;
    auto j_of = [&](const int k) -> int {
        if constexpr (PA32) return 2 * (sub + LANES * (k >> 1)) + (k & 1);
        else return sub + LANES * k;
    };

    // ---- kNN observation (independent of phase B): merge the LANES partial lists, drop the diagonal,
    //      obs[i, 4m:4m+4] = x_i - x_nbr(m)                                                 flocking.py:20-24
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
            } else if constexpr (KPACK) {
                const int id = (int)(kp[m] & kidx);
                nb[m] = (id < N) ? id : i;                   // (odd N: the pad slot at +inf is not an agent)
            } else {
                nb[m] = top.id[m];
            }
        }
        if constexpr (!EXACT) knn_rerank(xi, i, xe, nb);     // exact order / membership among the TOPK survivors
        // One thread per row and full rows: the warp's 32 observation rows are contiguous in memory (32 x 28 values).  A
        // lane writing its own row directly issues 28 scalar stores 112 bytes apart from its neighbours' -- partial sectors
        // that made Flocking-v0 4x slower than FlockingRelative-v0; stage the block in shared memory and copy it out with
        // coalesced 16-byte stores instead.
        const bool staged = (LANES == 1) && (nre == N) && a.obs != nullptr;
        TO *const obs_w = reinterpret_cast<TO *>(smem + a.off_obs) + (size_t)(tid - lane) * (NXS * KNN_K);
        if (valid && sub == 0 && a.obs) {
            TO *o = staged ? obs_w + lane * (NXS * KNN_K) : reinterpret_cast<TO *>(a.obs) + grow * (NXS * KNN_K);
            int m = 0;
            bool dropped = false;
#pragma unroll
            for (int q = 0; q < TOPK; ++q) {
                if (!dropped && nb[q] == i) { dropped = true; continue; }     // the diagonal (or an unused slot)
                if (m < KNN_K) {
                    const double4 xn = xe[nb[q]];
                    typename Tr<TO>::V4 v;
                    if constexpr (EXACT) {
                        v.x = __dsub_rn(xi.x, xn.x); v.y = __dsub_rn(xi.y, xn.y);
                        v.z = __dsub_rn(xi.z, xn.z); v.w = __dsub_rn(xi.w, xn.w);
                    } else {
                        v.x = (TO)(xi.x - xn.x); v.y = (TO)(xi.y - xn.y);
                        v.z = (TO)(xi.z - xn.z); v.w = (TO)(xi.w - xn.w);
                    }
                    o[m * 4 + 0] = v.x; o[m * 4 + 1] = v.y; o[m * 4 + 2] = v.z; o[m * 4 + 3] = v.w;
                    ++m;
                }
            }
        }
        if (staged) {
            __syncwarp();
            const int wr0 = r0 + (tid - lane);                    // first flat row of this warp == its global row
            const int nrw = min(32, r1 - wr0);
            if (nrw > 0) {
                const int nvec = nrw * (NXS * KNN_K) * (int)sizeof(TO) / 16;
                const float4 *src = reinterpret_cast<const float4 *>(obs_w);
                float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<TO *>(a.obs) + (size_t)wr0 * (NXS * KNN_K));
                for (int v = lane; v < nvec; v += 32) __stcs(dst + v, src[v]);
            }
        }
    }

    // ---- phase B: refine the candidates in fp64, ascending j per lane; rewrite the words with the exact bits ----
    // One flat loop per lane; the advance to the next word is straight-line code inside the loop body, so a warp
    // iterates ~max-over-lanes(total candidates) times.
    int deg = 0;
    const double R2r = a.R2, Rmr = a.Rm;
    if (!(a.tune & 4)) {
        // the diagonal (r2 = 0: always a candidate; +inf in the reference, :115) is removed up front
        {
            int kd = -1;
            if constexpr (PA32) { if (valid && ((i >> 1) % LANES) == sub) kd = 2 * ((i >> 1) / LANES) + (i & 1); }
            else { if (valid && (i % LANES) == sub) kd = i / LANES; }
            if (kd >= 0) { mrow[(kd >> 5) * MS] &= ~(1u << (kd & 31)); --cnt; }
        }
        // One flat loop per lane over its candidates, ascending j.  Advancing to the next mask word is predicated
        // straight-line code (some lane of the warp advances in almost every iteration) with the following word
        // prefetched; a candidate that turns out not to be a neighbour has its bit cleared in place (shared-memory
        // RED), so after the loop the words hold the exact adjacency bits.
        const unsigned xe_s = (unsigned)__cvta_generic_to_shared(xe);   // 32-bit shared addresses: stay in registers
        unsigned waddr = (unsigned)__cvta_generic_to_shared(mrow);       // current mask word of this thread
        const unsigned xdummy = xe_s + ((i == 0 && N > 1) ? 32u : 0u);
        unsigned cand = 0, nxt = 0;
        if (nw > 0) { cand = mrow[0]; nxt = mrow[MS]; }     // (the mask region has one spare word row)
        int jb = 0;                                         // bit position of the current word's bit 0
        int left = cnt;                                     // candidates still to refine: they are known to exist
        while (left > 0) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "setp.eq.u32 p, %0, 0;\n\t"
                         "@p mov.u32 %0, %1;\n\t"
                         "@p add.u32 %2, %2, %4;\n\t"
                         "@p add.s32 %3, %3, 32;\n\t"
                         "@p ld.shared.u32 %1, [%2+%4];\n\t}"
                         : "+r"(cand), "+r"(nxt), "+r"(waddr), "+r"(jb) : "n"(MS * 4) : "memory");
            const unsigned low = cand & (0u - cand);        // lowest candidate bit (0: an empty word, the lane idles)
            cand ^= low;
            const int k = bfind(low);
            int j;
            if constexpr (LANES == 1) j = jb + k;           // one thread per row: bit position == agent index
            else j = j_of(jb + k);
            // (an idle lane reads a harmless agent: not itself, so r2 > 0 and every masked term stays finite)
            const unsigned xaddr = low ? xe_s + (unsigned)j * 32u : xdummy;
            if constexpr (EXACT) {
                if (low != 0) {
                    const double4 xj = lds_state(xaddr);
                    const bool adj = refine_exact(xi, xj, a.R, a.R2, cen, A, i < nobs || j < nobs);
                    deg += adj ? 1 : 0;
                    if (!adj) asm volatile("red.shared.and.b32 [%0], %1;" :: "r"(waddr), "r"(~low) : "memory");
                    --left;
                }
            } else {
                refine_fast_pred<PA32>(xi, xaddr, low, waddr, R2r, Rmr, A, deg, left);
            }
        }
    }
    __syncwarp();                                           // reconverge here: the epilogue runs with all 32 lanes

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

    // ---- FlockingObstacle, fast modes: velocity differences with an obstacle row or column are zero --------------------
    // (flocking_obstacle.py:81-82).  The refine loop summed them like any other pair; take the (at most nobs) terms out
    // again with the same adjacency test.  EXACT handled them in place.
    double osx = 0.0, osy = 0.0;                            // sum over the obstacles of (v_j - v_0)   (centralized alignment)
    if constexpr (!EXACT) {
        if (nobs > 0 && valid && sub == 0) {
            if (i < nobs) { A.f0 = 0.0; A.f3 = 0.0; }
            for (int j = 0; j < nobs; ++j) {
                const double4 xj = xe[j];
                osx += xj.z - xe[0].z; osy += xj.w - xe[0].w;
                if (j == i || i < nobs) continue;
                const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                if (fma(dx, dx, dy * dy) < R2r) { A.f0 -= xi.z - xj.z; A.f3 -= xi.w - xj.w; }
            }
        }
    }

    // ---- row epilogue: features, action, 1/deg ---------------------------------------------------------------
    TO idg;
    if constexpr (EXACT) idg = __ddiv_rn(1.0, (double)max(deg, 1));               // :121-122
    else idg = inv_deg(deg, TO(0));
    if (valid && sub == 0) {
        if (a.values) {
            TO *v = reinterpret_cast<TO *>(a.values) + grow * NFE;
            if constexpr (PA32) {                       // 24 B per row, 8-byte aligned: three float2 stores
                float2 *v2 = reinterpret_cast<float2 *>(v);
                v2[0] = make_float2((float)A.f0, (float)A.f1); v2[1] = make_float2((float)A.f2, (float)A.f3);
                v2[2] = make_float2((float)A.f4, (float)A.f5);
            } else {
                double2 *v2 = reinterpret_cast<double2 *>(v);   // 48 B per row, 16-byte aligned
                v2[0] = make_double2(A.f0, A.f1); v2[1] = make_double2(A.f2, A.f3); v2[2] = make_double2(A.f4, A.f5);
            }
        }
        if (a.u_next) {
            double ux, uy;
            if constexpr (EXACT) {
                ux = __dsub_rn(-A.gx, A.s0);             // (-S.gx) - S.dvx      flocking_relative.py:209
                uy = __dsub_rn(-A.s1, A.gy);             // (-S.dvy) - S.gy
                ux = fmin(fmax(ux, -10.0), 10.0); uy = fmin(fmax(uy, -10.0), 10.0);   // :210
                ux = __ddiv_rn(ux, a.ascale); uy = __ddiv_rn(uy, a.ascale);           // :211
            } else {
                double sdx = A.f0, sdy = A.f3;           // decentralized: the masked sums (:205-206)
                if (cen) {                               // centralized: sum_j (v_i - v_j) = N*v_i - sum_j v_j, about agent 0
                    const volatile double *st = stats + g * STATS_PER_ENV;
                    while (st[2] == 0.0) { }             // published long ago by the stats warp (never waits on us)
                    __threadfence_block();
                    const double rx = xe[0].z, ry = xe[0].w;
                    sdx = fma((double)N, xi.z - rx, -(st[0] - (double)N * rx));
                    sdy = fma((double)N, xi.w - ry, -(st[1] - (double)N * ry));
                    if (nobs > 0) {                          // only the N - nobs ordinary agents count, and only for them
                        sdx = (i < nobs) ? 0.0 : fma((double)(N - nobs), xi.z - rx, -((st[0] - (double)N * rx) - osx));
                        sdy = (i < nobs) ? 0.0 : fma((double)(N - nobs), xi.w - ry, -((st[1] - (double)N * ry) - osy));
                    }
                }
                // clip(v, +-10) / ascale, then the variant's clip (flocking_stoch.py:43), as ONE clamp of v * (1/ascale)
                // to +-ulim = min(10 / ascale, u_clip) (host): same values, saturated or not; one rounding vs a true divide
                const double ia = a.inv_ascale, ul = a.ulim;
                ux = fmin(fmax((-2.0 * A.gx - sdx) * ia, -ul), ul);
                uy = fmin(fmax((-sdy - 2.0 * A.gy) * ia, -ul), ul);
            }
            if constexpr (EXACT) {
                if (a.var.u_clip > 0.0) {                // FlockingStochastic.controller (flocking_stoch.py:43)
                    ux = fmin(fmax(ux, -a.var.u_clip), a.var.u_clip); uy = fmin(fmax(uy, -a.var.u_clip), a.var.u_clip);
                }
            }
            reinterpret_cast<double2 *>(a.u_next)[grow] = make_double2(ux, uy);
        }
    }

    // ---- optional outputs (FlockAux): non-finite guard, degree, exact adjacency bits, get_stats() -------------------------
    if (a.aux.nonfinite != nullptr || a.aux.degree != nullptr || a.aux.adj_bits != nullptr || a.aux.min_dist != nullptr ||
        a.aux.vel_diff != nullptr) {
        if (a.aux.nonfinite != nullptr && valid && sub == 0) {
            const double chk = ((A.f0 + A.f1) + (A.f2 + A.f3)) + ((A.f4 + A.f5) + (A.gx + A.gy)) + ((xi.x + xi.y) + (xi.z + xi.w));
            if (!(fabs(chk) < CUDART_INF)) a.aux.nonfinite[env] = 1;
        }
        if (a.aux.degree != nullptr && valid && sub == 0) a.aux.degree[grow] = deg;
        if constexpr (LANES == 1) {
            if (a.aux.adj_bits != nullptr && valid)
                for (int w = 0; w < a.aux.WB; ++w) a.aux.adj_bits[grow * a.aux.WB + w] = (w < nw) ? mrow[w * MS] : 0u;
        }
        if (a.aux.min_dist != nullptr) {
            // min_j sqrt(r2) (flocking_relative.py:141-142; r2 is symmetric, so the column minimum is the row minimum).  Every
            // neighbour is closer than every non-neighbour: scan the exact adjacency bits; a row without neighbours scans all.
            auto r2_of = [&](const int j) -> double {
                const double4 xj = xe[j];
                if constexpr (EXACT) {
                    const double dx = __dsub_rn(xi.x, xj.x), dy = __dsub_rn(xi.y, xj.y);
                    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                } else {
                    const double dx = xi.x - xj.x, dy = xi.y - xj.y;
                    return fma(dx, dx, dy * dy);
                }
            };
            double m2 = CUDART_INF;
            if (valid) {
                for (int w = 0; w < nw; ++w) {
                    unsigned b = mrow[w * MS];
                    while (b) { const int k = __ffs(b) - 1; b &= b - 1; m2 = fmin(m2, r2_of(j_of(w * 32 + k))); }
                }
            }
            if constexpr (LANES > 1) {
#pragma unroll
                for (int m = LANES / 2; m > 0; m >>= 1) m2 = fmin(m2, shfl_xor(m2, m));
            }
            if (__any_sync(0xffffffffu, valid && deg == 0)) {
                if (valid && deg == 0) for (int j = sub; j < N; j += LANES) if (j != i) m2 = fmin(m2, r2_of(j));
                if constexpr (LANES > 1) {
#pragma unroll
                    for (int m = LANES / 2; m > 0; m >>= 1) m2 = fmin(m2, shfl_xor(m2, m));
                }
            }
            if (valid && sub == 0) {
                if constexpr (EXACT) reinterpret_cast<TO *>(a.aux.min_dist)[grow] = __dsqrt_rn(m2);
                else reinterpret_cast<TO *>(a.aux.min_dist)[grow] = (TO)sqrt(m2);
            }
        }
        if (a.aux.vel_diff != nullptr && valid && sub == 0) {
            const volatile double *st = stats + g * STATS_PER_ENV;
            while (st[2] == 0.0) { }
            __threadfence_block();
            if constexpr (EXACT) {                        // sqrt(sum((v - mean)^2, axis=1)) (:138-140), NumPy's rounding
                const double ex = __dsub_rn(xi.z, st[0]), ey = __dsub_rn(xi.w, st[1]);
                reinterpret_cast<TO *>(a.aux.vel_diff)[grow] = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
            } else {
                const double invN = 1.0 / (double)N;
                const double ex = xi.z - st[0] * invN, ey = xi.w - st[1] * invN;
                reinterpret_cast<TO *>(a.aux.vel_diff)[grow] = (TO)sqrt(fma(ex, ex, ey * ey));
            }
        }
    }

    // ---- dense adj_mean rows ---------------------------------------------------------------------------------------
    if (net == nullptr) return;
    if (a.out_mode == 1) {       // small full rows (LANES == 1; host guarantees it): expand_window() does the rest
        if constexpr (LANES == 1) idg_s[tid] = idg;
        return;
    }
    // out_mode 0: store the few non-zero entries (1/deg) over the warp's zero-filled rows.  A warp only ever touches its
    // own rows; __syncwarp orders the lanes' zero stores before the non-zero stores (bulk stores: lane 0 waits for their
    // completion first; the proxy fence orders the async-proxy writes before the generic stores that follow).
    if constexpr (EXACT) {
        if (a.zero_bytes && lane == 0 && !(a.tune & 1)) { bulk_wait_all(); fence_proxy_async(); }
    }
    __syncwarp();
    if (valid && !(a.tune & 2)) {
        TO *orow = net + grow * N;
        int w = 0;
        unsigned ex = nw > 0 ? mrow[0] : 0u;
        // fp64 rows that start on a 32-byte sector (one thread per row): a non-zero is stored together with the other three
        // entries of its sector, as ONE 32-byte store.  An 8-byte store into a row that has already left the L2 is a
        // read-modify-write of the sector in HBM (ECC); a whole sector is just a write.
        bool whole = false;
        if constexpr (sizeof(TO) == 8 && LANES == 1) whole = ((reinterpret_cast<uintptr_t>(orow) & 31u) == 0u) && !(a.tune & 64);
        while (w < nw) {
            if (ex == 0) { ++w; ex = (w < nw) ? mrow[w * MS] : 0u; }
            if (ex != 0) {
                const int k = __ffs(ex) - 1;
                if constexpr (sizeof(TO) == 8 && LANES == 1) {
                    const int k4 = k & ~3, j4 = w * 32 + k4;             // LANES == 1: j_of(k) = k
                    if (whole && j4 + 4 <= N) {
                        const unsigned q = ex >> k4;                     // (bits below k in this sector are already clear)
                        ex &= ~(0xFu << k4);
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};"
                                     :: "l"(orow + j4), "d"((q & 1u) ? (double)idg : 0.0), "d"((q & 2u) ? (double)idg : 0.0),
                                        "d"((q & 4u) ? (double)idg : 0.0), "d"((q & 8u) ? (double)idg : 0.0) : "memory");
                        continue;
                    }
                }
                ex &= ex - 1;
                orow[j_of(w * 32 + k)] = idg;
            }
        }
    }
}


// ---- small rows: every warp expands its own 32 rows of the window [r0, r1) from their exact bit masks (LANES == 1) -----
template <typename TO>
__device__ __forceinline__ void expand_window(TO *network, const int N, const int tune, const unsigned *masks, const TO *idg_s,
                                              const int r0, const int r1) {
    const int tid = threadIdx.x, lane = tid & 31;
    __syncwarp();                                    // the lanes' mask words and 1/deg (shared memory) are complete
    const int wr0 = r0 + (tid - lane);               // first flat row of this warp == its global row (nre == N)
    const int nrows = min(32, r1 - wr0);
    if (nrows > 0 && !(tune & 2)) {
        TO *wout = network + (size_t)wr0 * N;
        const unsigned *wm = masks + (tid - lane);
        const TO *idgs = idg_s + (tid - lane);
        constexpr int VMAX = 16 / (int)sizeof(TO);
        if (N % VMAX == 0) expand_rows<TO, VMAX>(wout, wm, idgs, N, nrows, lane);
        else if (VMAX == 4 && N % 2 == 0) expand_rows<TO, 2>(wout, wm, idgs, N, nrows, lane);
        else expand_rows<TO, 1>(wout, wm, idgs, N, nrows, lane);
    }
}

// Resident-CTA targets: the fast kernels are latency bound (divergent refine loops), so occupancy matters more than
// registers: 256-thread bound, FLOCK_MINB CTAs per SM.  EXACT keeps its registers.
// One CTA = one window.
#ifndef FLOCK_EXACT_MINB
#define FLOCK_EXACT_MINB 2
#endif
template <typename TO, int LANES, bool EXACT, bool KNN>
__global__ void __launch_bounds__(FLOCK_MAXT, EXACT ? FLOCK_EXACT_MINB : FLOCK_MINB) flock_env_kernel(const KArgs a) {
    static_assert(!EXACT || (!std::is_same<TO, float>::value && LANES == 1), "EXACT is fp64, one thread per row");
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *masks = reinterpret_cast<unsigned *>(smem + a.off_masks);
    TO *idg_s = reinterpret_cast<TO *>(smem + a.off_idg);
    const Win win = window_of(a, blockIdx.x);
    pdl_wait();
    stage_window<TO, EXACT>(a, smem, win);
    __syncthreads();
    window_stats<TO, EXACT>(a, smem, win);
    compute_rows<TO, LANES, EXACT, KNN>(a, smem, masks, idg_s, win);
    if constexpr (LANES == 1) {
        if (a.network != nullptr && a.out_mode == 1)
            expand_window<TO>(reinterpret_cast<TO *>(a.network), a.N, a.tune, masks, idg_s, win.r0, win.r1);
    }
}

}  // namespace flock
