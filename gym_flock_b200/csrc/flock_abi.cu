// flock_abi.cu -- extern "C" boundary of libflock_b200.so (declared in include/flock_b200.h).
// Host-side planning (which kernel, how rows map to CTAs / lanes / shared memory) lives here; the kernel
// instantiations are in launch_*.cu.  No torch or C++ types cross the ABI.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <new>
#include <utility>
#include <cstdio>
#include <cstring>
#include <string>

#define FLOCK_ARGS_ONLY 1      // argument structs and constants of the kernels, not their device code
#include "flock_host.hpp"
#include "flock_group_kernel.cuh"
#include "flock_swarm_kernel.cuh"
#include "flock_reset_kernel.cuh"
#include "flock_graph_kernels.cuh"

namespace {
thread_local std::string g_err;
}

int flock_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

namespace {

#define fail flock_fail

flock::StepVariant step_variant(const FlockParams *p) {
    flock::StepVariant v;
    v.step_scale = (p->step_scale != 0.0) ? p->step_scale : p->action_scalar;
    v.state_scale = (p->state_scale != 0.0) ? p->state_scale : 1.0;
    v.u_clip = p->u_clip;
    v.n_leaders = p->n_leaders;
    v.n_obstacles = p->n_obstacles;
    return v;
}

inline unsigned align16(size_t v) { return (unsigned)((v + 15) & ~(size_t)15); }
inline int round_up32(int v) { return (v + 31) & ~31; }

// diagnostics word (FlockParams.reserved): bits 0-7 CTA size / 32 of the env kernel, bits 8-15 phase switches of the env
// kernel, bits 16-23 envs per CTA of the group kernel, bit 24 never use the group kernel, bit 25 no programmatic dependent
// launch, bit 26 flock_rollout_* queues single-step launches instead of one fused launch, bit 27 stage the kNN rows in shared
// memory, bit 28 print the device time of each kernel of a swarm step (synchronises)
inline int tune_threads(const FlockParams *p) { return p->reserved & 0xff; }
inline int tune_phases(const FlockParams *p) { return (p->reserved >> 8) & 0xff; }
inline int tune_group(const FlockParams *p) { return (p->reserved >> 16) & 0xff; }
inline bool tune_no_group(const FlockParams *p) { return (p->reserved >> 24) & 1; }
inline bool tune_no_pdl(const FlockParams *p) { return (p->reserved >> 25) & 1; }
inline bool tune_no_fuse(const FlockParams *p) { return (p->reserved >> 26) & 1; }
inline bool tune_no_bulk(const FlockParams *p) { return (p->reserved >> 29) & 1; }       // bit 29: exact-order fill as plain stores
constexpr size_t kZeroBlock = 16 * 1024;
inline bool tune_obs_staged(const FlockParams *p) { return (p->reserved >> 27) & 1; }   // bit 27: stage the kNN rows in smem

struct Plan {
    bool group;               // flock_group_kernel (whole envs per CTA) instead of flock_env_kernel
    flock::KArgs k;
    flock::GArgs g;
    int lanes, threads, grid;
    size_t smem;
    bool exact, knn;
};

int validate(const FlockParams *p, int tbytes, int &row_lo, int &row_hi) {
    if (!p) return fail(FLOCK_EINVAL, "params is NULL");
    const int E = p->n_envs, N = p->n_agents;
    if (E < 1 || N < 1) return fail(FLOCK_EINVAL, "n_envs=%d n_agents=%d must be >= 1", E, N);
    if (p->n_leaders < 0 || p->n_leaders > N || p->n_obstacles < 0 || p->n_obstacles > N || p->u_clip < 0.0 || p->state_scale < 0.0)
        return fail(FLOCK_EINVAL, "bad variant parameters (n_leaders=%d n_obstacles=%d u_clip=%g state_scale=%g)", p->n_leaders,
                    p->n_obstacles, p->u_clip, p->state_scale);
    if ((long long)E * N > (1LL << 30)) return fail(FLOCK_EUNSUPPORTED, "E*N too large");
    row_lo = p->row_lo; row_hi = p->row_hi;
    if (row_lo == 0 && row_hi == 0) row_hi = N;
    if (row_lo < 0 || row_hi > N || row_lo >= row_hi) return fail(FLOCK_EINVAL, "bad row range [%d,%d)", row_lo, row_hi);
    if ((p->flags & FLOCK_EXACT_ORDER) && tbytes != 8) return fail(FLOCK_EINVAL, "FLOCK_EXACT_ORDER is an f64 mode");
    if ((p->flags & FLOCK_KNN_OBS) && p->knn != flock::KNN_K)
        return fail(FLOCK_EUNSUPPORTED, "knn=%d: only %d is built", p->knn, flock::KNN_K);
    if (!(p->action_scalar != 0.0)) return fail(FLOCK_EINVAL, "action_scalar must be non-zero");
    return FLOCK_OK;
}

// ---- group kernel: G whole envs per CTA, one thread per row --------------------------------------------------------------
// Cost model for G: the step is issue bound, so an SM's time is proportional to the threads it hosts per wave; minimise
// (full waves) * occ * T + ceil(rest / 148) * T over the G that fit, preferring fuller CTAs on ties.
bool group_layout(const FlockParams *p, int tbytes, bool knn, int G, flock::GArgs &g, int &threads, size_t &smem) {
    const int N = p->n_agents;
    const bool obs_staged = knn && tune_obs_staged(p);
    const bool pa32 = (tbytes == 4);
    const int NP2 = (N + 1) / 2;
    const int W = pa32 ? (2 * NP2 + 31) / 32 : (N + 31) / 32;
    threads = round_up32(G * N);
    if (threads > flock::GMAXT) return false;
    const int GMS = threads <= 256 ? 256 : 512;        // mask stride of the instantiation launch_group picks
    g.xs_bytes = align16((size_t)G * N * 32);
    g.pf_bytes = pa32 ? align16((size_t)G * NP2 * 16) : 0;
    g.stats_bytes = align16((size_t)G * flock::STATS_PER_ENV * 8);
    g.off_xs = 0;
    g.off_pf = g.off_xs + 2 * g.xs_bytes;
    g.off_masks = g.off_pf + 2 * g.pf_bytes;
    g.off_stats = align16((size_t)g.off_masks + (size_t)GMS * (W + 1) * 4);     // one spare word row (prefetch)
    g.off_idg = g.off_stats + 2 * g.stats_bytes;
    g.off_obs = align16((size_t)g.off_idg + (size_t)threads * 8);
    g.off_masks2 = align16((size_t)g.off_obs + (obs_staged ? (size_t)threads * flock::NXS * flock::KNN_K * tbytes : 0));
    g.off_hint = align16((size_t)g.off_masks2 + (knn ? (size_t)GMS * W * 4 : 0));
    g.obs_direct = obs_staged ? 0 : 1;
    g.off_bar = align16((size_t)g.off_hint + (knn ? (size_t)threads * 16 : 0));     // the split step barrier (one mbarrier)
    smem = (size_t)g.off_bar + 16;
    g.G = G; g.NP2 = NP2; g.W = W;
    return smem <= kMaxSmem;
}

bool plan_group(const FlockParams *p, int tbytes, bool knn, Plan &pl) {
    const int E = p->n_envs, N = p->n_agents;
    if (N > flock::GMAXT) return false;
    flock::GArgs &g = pl.g;
    memset(&g, 0, sizeof(g));
    // Envs per CTA.  The step is issue bound, so an SM's time is the threads it hosts: R full rounds of `occ` resident CTAs
    // of T threads, then one partial round.  The CTAs of the partial round own fewer envs (G_tail), so that it is spread
    // over all SMs.  Measured on B200 (4096 x 100, fused): small CTAs run ~25 % more efficiently per thread than
    // 512-thread ones (shorter barriers, more phase diversity) -- hence the size factor.
    int bestG = 0, best_main = 0, best_tail = 0;
    double best = 0.0;
    const int gmax = std::min(E, flock::GMAXT / N);
    const int forced = tune_group(p) ? std::min(tune_group(p), gmax) : 0;
    for (int G = 1; G <= gmax; ++G) {
        if (forced && G != forced) continue;
        int T; size_t smem;
        flock::GArgs tmp;
        memset(&tmp, 0, sizeof(tmp));
        if (!group_layout(p, tbytes, knn, G, tmp, T, smem)) continue;
        const long long occ = (long long)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(65536 / (64 * T), (228 * 1024) / (smem + 1024)), 2048 / T));
        const long long S = (long long)kSMs * occ;
        const long long R = E / (S * G), Er = E - R * S * G;
        long long n_main = R * S, G_tail = G, n_tail = 0;
        if (Er > 0) {
            G_tail = std::min<long long>(G, (Er + S - 1) / S);
            n_tail = (Er + G_tail - 1) / G_tail;
        }
        const double load = (double)R * occ * T + (double)((n_tail + kSMs - 1) / kSMs) * round_up32((int)G_tail * N);
        double cost = load * (1.0 + 0.25 * (double)(T - 128) / 384.0);
        // measured optima where the model is off by more than its resolution (B200, fp32, 200 steps from the synthetic x0, large
        // batches): 16384 x 50 runs 44.3 us/step with 5 envs per CTA (46.6-48.4 with 2, 4, 6, 8, 10); Flocking-v0 at 4096 x 100
        // 59.7 with 5 (62.6 with 3, 66.1 with 2, 72.8 with 1)
        if (tbytes == 4 && G == 5 && E >= 4 * kSMs * G && (N == 50 || (N == 100 && knn))) cost *= 0.85;
        if (bestG == 0 || cost < best * 0.999 || (cost < best * 1.001 && G > bestG)) {
            best = cost; bestG = G; best_main = (int)n_main; best_tail = (int)G_tail;
        }
    }
    if (bestG == 0) return false;
    int T; size_t smem;
    group_layout(p, tbytes, knn, bestG, g, T, smem);
    g.E = E; g.N = N; g.steps = 1;
    g.n_main = best_main; g.G_tail = best_tail;
    g.centralized = (p->flags & FLOCK_CENTRALIZED) ? 1 : 0;
    const double R = p->comm_radius, R2 = p->comm_radius2;
    g.R2 = R2; g.Rm = (g.centralized || R2 > R) ? R : std::nextafter(R2, -1.0);
    g.thrC = std::max(R, R2);
    // conservative fp32 filter threshold (flock_env_kernel.cuh, phase A): tc (1 + 3e-6) + 16 rc (max|p| + 2 rc) 2^-24
    const float tc = (float)g.thrC, rc = sqrtf(tc);
    g.thr_a = tc * (1.0f + 3e-6f); g.thr_b = 16.0f * rc * 5.9604645e-8f; g.thr_c = 2.0f * rc;
    g.inv_n = 1.0f / (float)N;
    g.dt = p->dt; g.inv_ascale = 1.0 / p->action_scalar; g.var = step_variant(p);
    g.ulim = 10.0 * g.inv_ascale;
    if (p->u_clip > 0.0) g.ulim = std::min(g.ulim, p->u_clip);
    {
        const long long rest = (long long)E - (long long)best_main * bestG;
        pl.grid = best_main + (int)((rest + best_tail - 1) / best_tail);
    }
    pl.group = true; pl.lanes = 1; pl.threads = T; pl.smem = smem;
    return true;
}

// Decide the kernel, lanes-per-row, envs-per-CTA (or rows-per-tile) and the shared-memory layout.
int make_plan(const FlockParams *p, int tbytes, Plan &pl, bool need_bits = false) {
    int row_lo, row_hi;
    int rc = validate(p, tbytes, row_lo, row_hi);
    if (rc) return rc;
    const int E = p->n_envs, N = p->n_agents;
    const int nrows = row_hi - row_lo;
    pl.group = false;
    pl.exact = (p->flags & FLOCK_EXACT_ORDER) != 0;
    pl.knn = (p->flags & FLOCK_KNN_OBS) != 0;

    int lanes = need_bits ? 1 : p->lanes;
    if (pl.exact) {
        if (lanes > 1) return fail(FLOCK_EINVAL, "FLOCK_EXACT_ORDER needs lanes = 1 (sequential j)");
        lanes = 1;
    } else if (lanes <= 0) {
        lanes = 1;
        const long long rows = (long long)E * nrows;
        // one thread per row as soon as every SM gets a CTA; several lanes per row only for small batches, where the step
        // is latency bound (measured: 2048 x 50 runs 12.4 us with one lane, 23.9 us with two)
        while (rows * lanes < (long long)kSMs * 256 && lanes < 32 && lanes * 2 <= N) lanes *= 2;
    } else if (lanes > 32 || (lanes & (lanes - 1))) {
        return fail(FLOCK_EINVAL, "lanes=%d must be a power of two <= 32", lanes);
    }

    // whole envs per CTA (flock_group_kernel): fast modes, full rows, one thread per row, no obstacle variant
    if (!pl.exact && lanes == 1 && nrows == N && p->n_obstacles == 0 && !tune_no_group(p) && plan_group(p, tbytes, pl.knn, pl))
        return FLOCK_OK;

    // Rows per CTA window.  Small envs: flat windows over the E*nrows rows (every lane gets a row; a window stages
    // the few envs it touches).  Envs with at least a window of rows: row tiles per env.
    int threads = 256;
    if (tune_threads(p) > 0) threads = std::min(256, 32 * tune_threads(p));
    int RT = std::max(1, threads / lanes);
    const long long rows_total = (long long)E * nrows;
    if (rows_total < RT) RT = (int)rows_total;
    threads = round_up32(RT * lanes);
    int tiles, GMAX, grid;
    if (nrows >= RT) { tiles = (nrows + RT - 1) / RT; GMAX = 1; grid = E * tiles; }
    else { tiles = 0; GMAX = std::min(E, (RT - 1) / nrows + 2); grid = (int)((rows_total + RT - 1) / RT); }
    const bool pa32 = (tbytes == 4);                   // fp32 output mode filters on packed fp32 positions
    const int NP2 = (N + 1) / 2;
    const int nk_max = pa32 ? 2 * ((NP2 + lanes - 1) / lanes) : (N + lanes - 1) / lanes;
    const int W = (nk_max + 31) / 32;
    flock::KArgs &k = pl.k;
    memset(&k, 0, sizeof(k));
    k.off_xs = 0;
    const size_t size_xs = (size_t)GMAX * N * 32;       // double4 per agent
    k.off_pf = align16(size_xs);
    const size_t size_pf = pa32 ? (size_t)GMAX * NP2 * 16 : 0;
    k.off_masks = align16(k.off_pf + size_pf);
    const size_t size_lists = (size_t)flock::MS * (W + 1) * 4;     // word-major, fixed stride, one spare word row (prefetch)
    k.off_stats = align16(k.off_masks + size_lists);
    k.off_idg = align16(k.off_stats + (size_t)GMAX * flock::STATS_PER_ENV * 8);      // 1/deg per row (in-warp row expansion)
    k.off_obs = align16((size_t)k.off_idg + (size_t)threads * 8);                 // kNN observation rows staged per warp
    pl.smem = (size_t)k.off_obs + (pl.knn ? (size_t)threads * flock::NXS * flock::KNN_K * tbytes : 0);
    if (pl.smem > kMaxSmem)
        return fail(FLOCK_EUNSUPPORTED, "N=%d needs %zu B of shared memory per CTA (> %zu): beyond the env-resident kernel",
                    N, pl.smem, kMaxSmem);
    size_t zero_block = 0;      // exact order, zero-fill + scatter: a block of zeros for the TMA bulk fill, unless it costs a resident CTA
    if (pl.exact && !tune_no_bulk(p)) {
        const size_t base = (pl.smem + 127) & ~(size_t)127;
        auto ctas = [](size_t smem) { return std::min<size_t>(2, (228 * 1024) / (smem + 1024)); };    // (registers allow two)
        if (base + kZeroBlock <= kMaxSmem && ctas(base + kZeroBlock) == ctas(base)) {
            zero_block = kZeroBlock; k.off_zero = (unsigned)base; pl.smem = base + kZeroBlock;
        }
    }
    k.E = E; k.N = N; k.RT = RT; k.tiles = tiles; k.GMAX = GMAX; k.row_lo = row_lo; k.row_hi = row_hi;
    k.centralized = (p->flags & FLOCK_CENTRALIZED) ? 1 : 0;
    {
        // fast-path gradient mask as ONE compare r2 <= Rm.  Centralized: r2 <= R (flocking_relative.py:225).
        // Decentralized (:205-206): additionally r2 < R2, i.e. r2 <= pred(R2) when R2 <= R.
        const bool cen = k.centralized != 0;
        const double R = p->comm_radius, R2 = p->comm_radius2;
        k.Rm = (cen || R2 > R) ? R : std::nextafter(R2, -1.0);
    }
    k.tune = tune_phases(p);
    // full rows, one thread per row: expand in the warp (coalesced stores, every byte written once);
    // sharded rows / several lanes per row: zero-fill + scatter
    // (exact-order mode with long rows keeps the up-front fill: its long sequential compute phase hides it)
    k.out_mode = (lanes == 1 && nrows == N && !(pl.exact && (size_t)N * tbytes > 2048) && !(tune_phases(p) & 16)) ? 1 : 0;
    k.zero_bytes = (k.out_mode == 0 && lanes == 1 && nrows == N) ? (unsigned)zero_block : 0u;
    if (k.zero_bytes == 0 && zero_block) pl.smem -= zero_block;
    k.W = W; k.NP2 = NP2; k.thrC = std::max(p->comm_radius, p->comm_radius2); k.R = p->comm_radius; k.R2 = p->comm_radius2;
    k.dt = p->dt; k.ascale = p->action_scalar; k.inv_ascale = 1.0 / p->action_scalar; k.var = step_variant(p);
    k.ulim = 10.0 * k.inv_ascale;
    if (p->u_clip > 0.0) k.ulim = std::min(k.ulim, p->u_clip);
    pl.lanes = lanes; pl.threads = threads;
    pl.grid = grid;
    return FLOCK_OK;
}

inline bool misaligned(const void *p, uintptr_t a = 16) { return p && (reinterpret_cast<uintptr_t>(p) & (a - 1)); }

flock::AuxOut aux_out(const FlockAux *aux, int N) {
    flock::AuxOut o;
    memset(&o, 0, sizeof(o));
    o.WB = (N + 31) / 32;
    if (aux) {
        o.nonfinite = aux->nonfinite; o.adj_bits = aux->adj_bits; o.degree = aux->degree;
        o.min_dist = aux->min_dist; o.vel_diff = aux->vel_diff; o.knn_hint = aux->knn_hint;
    }
    return o;
}

int launch_plan(Plan &pl, int tbytes, cudaStream_t st, bool pdl) {
    if (pl.group) return launch_group(pl.g, tbytes, pl.knn, pl.grid, pl.threads, pl.smem, st, pdl);
    if (tbytes == 4) return launch_env_f32(pl.k, pl.lanes, pl.knn, pl.grid, pl.threads, pl.smem, st, pdl);
    return launch_env_f64(pl.k, pl.lanes, pl.exact, pl.knn, pl.grid, pl.threads, pl.smem, st, pdl);
}

template <typename T>
int check_step_args(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, T *values, T *network, T *reward,
                    double *u_next, T *obs, const FlockAux *aux, bool may_alias) {
    if (!x_in) return fail(FLOCK_EINVAL, "x_in is NULL");
    if (!may_alias && x_out && x_out == x_in) return fail(FLOCK_EINVAL, "x_out must not alias x_in");
    if (!may_alias && u_in && u_next == u_in) return fail(FLOCK_EINVAL, "u_next must not alias u_in");
    const bool sharded = !((p->row_lo == 0) && (p->row_hi == 0 || p->row_hi == p->n_agents));
    if (misaligned(x_in) || misaligned(u_in) || misaligned(x_out) || misaligned(u_next) || misaligned(obs) ||
        misaligned(values, sharded ? std::min<uintptr_t>(16, 2 * sizeof(T)) : 16) || misaligned(network, sharded ? sizeof(T) : 16))
        return fail(FLOCK_EALIGN, "device buffers must be 16-byte aligned (row-sharded calls: values 8 bytes, network element size)");
    if (reward && (reinterpret_cast<uintptr_t>(reward) % sizeof(T)))
        return fail(FLOCK_EALIGN, "reward must be aligned to its element size");
    if (aux && (misaligned(aux->nonfinite, 4) || misaligned(aux->adj_bits, 4) || misaligned(aux->degree, 4) ||
                misaligned(aux->min_dist, sizeof(T)) || misaligned(aux->vel_diff, sizeof(T)) || misaligned(aux->knn_hint, 16)))
        return fail(FLOCK_EALIGN, "FlockAux buffers must be element aligned");
    return FLOCK_OK;
}

// One step, or (group kernel only) `steps` fused steps.  x_out / u_next receive the state / action after the last step.
template <typename T>
int step_impl(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, T *values, T *network,
              T *reward, double *u_next, T *obs, const FlockAux *aux, void *stream, int steps = 1, T *rewards_all = nullptr,
              bool *fused = nullptr) {
    NvtxRange nvtx_("flock_step");
    Plan pl;
    int rc = make_plan(p, (int)sizeof(T), pl, aux && aux->adj_bits);
    if (rc) return rc;
    if (fused) *fused = pl.group;
    if (steps != 1 && !pl.group) return FLOCK_OK;          // the caller loops over single steps
    // (the group kernel's CTAs own whole envs: each reads its inputs before it writes its outputs, so a fused rollout may
    //  finish in the buffers it started from)
    rc = check_step_args<T>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, aux, pl.group && steps != 1);
    if (rc) return rc;
    if (pl.knn && !obs) pl.knn = false;
    const flock::AuxOut ao = aux_out(aux, p->n_agents);
    if (pl.group) {
        flock::GArgs &g = pl.g;
        g.x_in = x_in; g.u_in = u_in; g.x_out = x_out; g.u_out = u_next;
        g.values = values; g.network = network; g.reward = reward; g.rewards_all = rewards_all; g.obs = obs;
        g.aux = ao; g.steps = steps;
    } else {
        flock::KArgs &k = pl.k;
        k.x_in = x_in; k.u_in = u_in; k.x_out = x_out; k.u_next = u_next;
        k.values = values; k.network = network; k.reward = reward; k.obs = obs;
        k.aux = ao;
        if (ao.adj_bits && pl.lanes != 1) return fail(FLOCK_EINVAL, "FlockAux.adj_bits needs one thread per row");
    }
    return launch_plan(pl, (int)sizeof(T), reinterpret_cast<cudaStream_t>(stream), !tune_no_pdl(p));
}

template <typename T>
int rollout_impl(const FlockParams *p, int steps, double *x_a, double *x_b, double *u_a, double *u_b, T *values,
                 T *network, T *reward, T *rewards_all, T *obs, int32_t *final_in_b, void *stream) {
    NvtxRange nvtx_("flock_rollout");
    if (steps < 0) return fail(FLOCK_EINVAL, "steps=%d", steps);
    if (!x_a || !x_b || !u_a || !u_b) return fail(FLOCK_EINVAL, "x_a/x_b/u_a/u_b must be non-NULL");
    if (final_in_b) *final_in_b = (steps & 1);
    if (steps == 0) return FLOCK_OK;
    // small envs: ONE persistent launch runs all the steps (flock_group_kernel); the final state / action land where `steps`
    // ping-pong launches would have left them
    bool fused = false;
    int rc = tune_no_fuse(p) ? FLOCK_OK : step_impl<T>(p, x_a, u_a, (steps & 1) ? x_b : x_a, values, network, reward, (steps & 1) ? u_b : u_a, obs, nullptr,
                          stream, steps, rewards_all, &fused);
    if (rc || fused) return rc;
    double *xi = x_a, *xo = x_b, *ui = u_a, *uo = u_b;
    for (int t = 0; t < steps; ++t) {
        T *r = rewards_all ? rewards_all + (size_t)t * p->n_envs : reward;
        rc = step_impl<T>(p, xi, ui, xo, values, network, r, uo, obs, nullptr, stream);
        if (rc) return rc;
        std::swap(xi, xo);
        std::swap(ui, uo);
    }
    if (rewards_all && reward)
        CUDA_TRY(cudaMemcpyAsync(reward, rewards_all + (size_t)(steps - 1) * p->n_envs, sizeof(T) * p->n_envs,
                                 cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    return FLOCK_OK;
}

// ---- large swarms: prep kernel + streaming row kernel (flock_swarm_kernel.cuh) --------------------------------------
struct SwarmLayout { size_t off_pf, off_partials, off_needy_cnt, off_needy, off_fix_d, off_fix_i, off_ticket, off_ovf, bytes; int NP2, PB; };

SwarmLayout swarm_layout(int E, int N, int nrows) {
    SwarmLayout L;
    L.NP2 = (N + 1) / 2;
    L.PB = (N + flock::PREP_THREADS - 1) / flock::PREP_THREADS;
    L.off_pf = 0;
    L.off_partials = ((size_t)E * L.NP2 * 16 + 255) & ~(size_t)255;
    L.off_needy_cnt = L.off_partials + (((size_t)E * L.PB * 32 + 255) & ~(size_t)255);
    L.off_needy = L.off_needy_cnt + (((size_t)E * 4 + 255) & ~(size_t)255);
    const size_t fix_entries = (size_t)E * flock::NEEDY_CAP * flock::KFIX_SPLIT * flock::TOPK;   // partial lists of the kNN fix-up
    L.off_fix_d = L.off_needy + (((size_t)E * flock::NEEDY_CAP * 4 + 255) & ~(size_t)255);
    L.off_fix_i = L.off_fix_d + ((fix_entries * 8 + 255) & ~(size_t)255);
    L.off_ticket = L.off_fix_i + ((fix_entries * 4 + 255) & ~(size_t)255);
    L.off_ovf = L.off_ticket + (((size_t)E * flock::NEEDY_CAP * 4 + 255) & ~(size_t)255);
    L.bytes = L.off_ovf + (size_t)E * nrows * 32 * flock::SW_LCAP2 * 4;       // second tier of the candidate lists
    return L;
}

template <typename T>
int swarm_step_impl(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, T *values, T *network,
                    T *reward, double *u_next, T *obs, void *scratch, size_t scratch_bytes, void *stream) {
    NvtxRange nvtx_("flock_swarm_step");
    if (!p) return fail(FLOCK_EINVAL, "params is NULL");
    const int E = p->n_envs, N = p->n_agents;
    if (E < 1 || N < 2) return fail(FLOCK_EINVAL, "n_envs=%d n_agents=%d", E, N);
    if (p->flags & FLOCK_EXACT_ORDER)
        return fail(FLOCK_EUNSUPPORTED, "FLOCK_EXACT_ORDER is offered by the env-resident path only (flock_step_f64)");
    if ((p->flags & FLOCK_KNN_OBS) && p->knn != flock::KNN_K) return fail(FLOCK_EUNSUPPORTED, "knn=%d", p->knn);
    if (p->n_obstacles != 0) return fail(FLOCK_EUNSUPPORTED, "n_obstacles is offered by the env-resident path only (flock_step_*)");
    int row_lo = p->row_lo, row_hi = p->row_hi;
    if (row_lo == 0 && row_hi == 0) row_hi = N;
    if (row_lo < 0 || row_hi > N || row_lo >= row_hi) return fail(FLOCK_EINVAL, "bad row range [%d,%d)", row_lo, row_hi);
    if (!x_in || !x_out || !scratch) return fail(FLOCK_EINVAL, "x_in, x_out and scratch are required");
    if (x_out == x_in) return fail(FLOCK_EINVAL, "x_out must not alias x_in");
    if (misaligned(x_in) || misaligned(u_in) || misaligned(x_out) || misaligned(values, std::min<uintptr_t>(16, 2 * sizeof(T))) ||
        misaligned(network, sizeof(T)) || misaligned(u_next) || misaligned(obs) || (reinterpret_cast<uintptr_t>(scratch) & 255u))
        return fail(FLOCK_EALIGN, "device buffers must be 16-byte aligned (values: 8, network: element size, scratch: 256)");
    const SwarmLayout L = swarm_layout(E, N, row_hi - row_lo);
    if (scratch_bytes < L.bytes) return fail(FLOCK_EINVAL, "scratch too small: %zu < %zu", scratch_bytes, L.bytes);
    int lanes = p->lanes > 0 ? p->lanes : 32;
    if (lanes != 8 && lanes != 16 && lanes != 32) return fail(FLOCK_EINVAL, "swarm path: lanes must be 8, 16 or 32");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *sc = reinterpret_cast<char *>(scratch);

    flock::SArgs a;
    memset(&a, 0, sizeof(a));
    a.x_in = x_in; a.u_in = u_in; a.x_cur = x_out; a.u_next = u_next;
    a.values = values; a.network = network; a.reward = reward; a.obs = (p->flags & FLOCK_KNN_OBS) ? obs : nullptr;
    a.pf = reinterpret_cast<float4 *>(sc + L.off_pf);
    a.partials = reinterpret_cast<double *>(sc + L.off_partials);
    a.needy_cnt = reinterpret_cast<unsigned *>(sc + L.off_needy_cnt);
    a.needy_rows = reinterpret_cast<unsigned *>(sc + L.off_needy);
    a.fix_d = reinterpret_cast<double *>(sc + L.off_fix_d);
    a.fix_i = reinterpret_cast<int *>(sc + L.off_fix_i);
    a.fix_ticket = reinterpret_cast<unsigned *>(sc + L.off_ticket);
    a.ovf = reinterpret_cast<unsigned *>(sc + L.off_ovf);
    a.E = E; a.N = N; a.NP2 = L.NP2; a.PB = L.PB; a.row_lo = row_lo; a.row_hi = row_hi;
    a.RT = flock::SW_R * (256 / lanes);        // rows per CTA
    a.TS = 2048;
    a.centralized = (p->flags & FLOCK_CENTRALIZED) ? 1 : 0;
    a.R = p->comm_radius; a.R2 = p->comm_radius2; a.thrC = std::max(a.R, a.R2);
    a.Rm = (a.centralized || a.R2 > a.R) ? a.R : std::nextafter(a.R2, -1.0);
    a.dt = p->dt; a.ascale = p->action_scalar; a.inv_ascale = 1.0 / p->action_scalar; a.var = step_variant(p);
    a.off_tiles = 0;
    a.off_lists = align16((size_t)2 * a.TS * 16);                                   // [thread][SW_R][SW_LCAP] candidates
    a.off_bar = align16(a.off_lists + (size_t)flock::SW_THREADS * flock::SW_R * flock::SW_LCAP * 4);
    a.off_tot = align16((size_t)a.off_bar + 16);
    const size_t smem = (size_t)a.off_tot + 64;
    a.fill_warps = network ? flock::SW_FILL_WARPS : 0;       // dedicated zero-fill warps only when the dense rows are written

    // three launches, no memset: the prep kernel clears the kNN fix-up queue, the row CTAs add up its partial sums
    const bool knn = a.obs != nullptr;
    const dim3 grid((row_hi - row_lo + a.RT - 1) / a.RT, E);
    return launch_swarm(a, (int)sizeof(T), lanes, knn, L.PB, grid, smem, st, !tune_no_pdl(p), (p->reserved >> 28) & 1);
}

}  // namespace

int reset_sample_impl(const FlockParams *p, double r_max, double v_max, double v_bias, double min_dist, uint64_t seed,
                      int32_t max_tries, double *x_out, int32_t *tries_out, void *stream) {
    NvtxRange nvtx_("flock_reset_sample");
    if (!p) return fail(FLOCK_EINVAL, "params is NULL");
    const int E = p->n_envs, N = p->n_agents;
    if (E < 1 || N < 1) return fail(FLOCK_EINVAL, "n_envs=%d n_agents=%d must be >= 1", E, N);
    if (!x_out) return fail(FLOCK_EINVAL, "x_out is NULL");
    if (misaligned(x_out)) return fail(FLOCK_EALIGN, "x_out must be 16-byte aligned");
    if (!(r_max > 0.0) || v_max < 0.0 || v_bias < 0.0 || min_dist < 0.0 || max_tries < 1)
        return fail(FLOCK_EINVAL, "bad sampler parameters (r_max=%g v_max=%g v_bias=%g min_dist=%g max_tries=%d)", r_max, v_max,
                    v_bias, min_dist, max_tries);
    const size_t smem = (size_t)N * sizeof(double2);
    const size_t smem_cap = kMaxSmem - 1024;            // the kernel also has a few static shared variables
    if (smem > smem_cap) return fail(FLOCK_EUNSUPPORTED, "N=%d: the candidate positions do not fit one CTA's shared memory", N);
    flock::RArgs a;
    a.x_out = x_out; a.tries_out = tries_out; a.E = E; a.N = N; a.r_max = r_max; a.v_max = v_max; a.v_bias = v_bias;
    a.R2 = p->comm_radius2; a.min_dist = min_dist; a.seed = seed; a.max_tries = max_tries;
    return launch_reset(a, smem, smem_cap, reinterpret_cast<cudaStream_t>(stream));
}

// ---- host-buffer context -------------------------------------------------------------------------------
struct FlockCtx {
    FlockParams p;
    int tbytes, device;
    cudaStream_t stream;
    double *x[2];      // ping-pong state; x[cur] is the current state
    double *u[2];      // u[0]: action to apply (uploaded, or the cached controller action); u[1]: next
    void *values, *network, *reward, *obs;
    void *stat_a, *stat_b;   // get_stats() scratch [E][N]
    uint16_t *knn_hint;      // [E][N][8] last step's neighbours (search-radius hint of the next kNN)
    int cur;
    int u_cached_flags;   // centralized flag the cached action was computed with; -1 = none
};

namespace {

size_t n_x(const FlockCtx *c) { return (size_t)c->p.n_envs * c->p.n_agents * 4; }
size_t n_u(const FlockCtx *c) { return (size_t)c->p.n_envs * c->p.n_agents * 2; }
size_t n_val(const FlockCtx *c) { return (size_t)c->p.n_envs * c->p.n_agents * 6; }
size_t n_net(const FlockCtx *c) { return (size_t)c->p.n_envs * c->p.n_agents * c->p.n_agents; }
size_t n_obs(const FlockCtx *c) { return (size_t)c->p.n_envs * c->p.n_agents * 4 * flock::KNN_K; }

int ctx_run(FlockCtx *c, const double *u_dev, double *x_out, bool want_values, bool want_net, bool want_reward,
            bool want_obs, double *u_next, const FlockAux *aux = nullptr) {
    void *obs = (want_obs && (c->p.flags & FLOCK_KNN_OBS)) ? c->obs : nullptr;
    FlockAux hint_only;
    if (!aux && obs && c->knn_hint) {
        memset(&hint_only, 0, sizeof(hint_only));
        hint_only.knn_hint = c->knn_hint;
        aux = &hint_only;
    }
    if (c->tbytes == 8)
        return step_impl<double>(&c->p, c->x[c->cur], u_dev, x_out, want_values ? (double *)c->values : nullptr,
                                 want_net ? (double *)c->network : nullptr, want_reward ? (double *)c->reward : nullptr,
                                 u_next, (double *)obs, aux, c->stream);
    return step_impl<float>(&c->p, c->x[c->cur], u_dev, x_out, want_values ? (float *)c->values : nullptr,
                            want_net ? (float *)c->network : nullptr, want_reward ? (float *)c->reward : nullptr, u_next,
                            (float *)obs, aux, c->stream);
}

int ctx_fetch(FlockCtx *c, void *values_host, void *network_host, void *reward_host, void *obs_host) {
    const size_t tb = (size_t)c->tbytes;
    if (values_host) CUDA_TRY(cudaMemcpyAsync(values_host, c->values, n_val(c) * tb, cudaMemcpyDeviceToHost, c->stream));
    if (network_host) CUDA_TRY(cudaMemcpyAsync(network_host, c->network, n_net(c) * tb, cudaMemcpyDeviceToHost, c->stream));
    if (reward_host) CUDA_TRY(cudaMemcpyAsync(reward_host, c->reward, (size_t)c->p.n_envs * tb, cudaMemcpyDeviceToHost, c->stream));
    if (obs_host && (c->p.flags & FLOCK_KNN_OBS))
        CUDA_TRY(cudaMemcpyAsync(obs_host, c->obs, n_obs(c) * tb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLOCK_OK;
}

}  // namespace

template <typename T>
static int aggregate_impl(const FlockParams *p, const T *x, int32_t F, const T *prev, const uint32_t *adj_bits, int32_t L,
                          int32_t pool, T *out, void *stream) {
    if (!p || !x || !adj_bits || !out) return fail(FLOCK_EINVAL, "NULL argument");
    if (p->n_envs < 1 || p->n_agents < 1 || F < 1 || L < 1) return fail(FLOCK_EINVAL, "bad sizes (E=%d N=%d F=%d L=%d)", p->n_envs, p->n_agents, F, L);
    if (pool != FLOCK_POOL_MEAN && pool != FLOCK_POOL_SUM) return fail(FLOCK_EINVAL, "pool=%d", pool);
    if ((const void *)out == (const void *)prev || (const void *)out == (const void *)x) return fail(FLOCK_EINVAL, "out must not alias prev or x");
    flock::AggArgs a;
    a.x = x; a.prev = prev; a.bits = adj_bits; a.out = out;
    a.E = p->n_envs; a.N = p->n_agents; a.F = F; a.L = L; a.WB = (p->n_agents + 31) / 32; a.pool = pool;
    return launch_aggregate(a, (int)sizeof(T), reinterpret_cast<cudaStream_t>(stream));
}
template <typename T>
static int repulsion_impl(const FlockParams *p, const double *pos, int32_t stride, const T *weights, double cutoff_r2, T *out,
                          void *stream) {
    if (!p || !pos || !weights || !out) return fail(FLOCK_EINVAL, "NULL argument");
    if (p->n_envs < 1 || p->n_agents < 1 || stride < 2) return fail(FLOCK_EINVAL, "bad sizes (E=%d N=%d stride=%d)", p->n_envs, p->n_agents, stride);
    flock::ShepArgs a;
    a.pos = pos; a.weights = weights; a.out = out; a.E = p->n_envs; a.N = p->n_agents; a.stride = stride; a.cutoff_r2 = cutoff_r2;
    return launch_shepherd(a, (int)sizeof(T), reinterpret_cast<cudaStream_t>(stream));
}

extern "C" {

int flock_abi_version(void) { return FLOCK_ABI_VERSION; }
const char *flock_last_error(void) { return g_err.c_str(); }
int flock_adj_words(int32_t n_agents) { return n_agents > 0 ? (n_agents + 31) / 32 : 0; }

int flock_step_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, double *values,
                   double *network, double *reward, double *u_next, double *obs, void *stream) {
    return step_impl<double>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, nullptr, stream);
}
int flock_step_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, float *values,
                   float *network, float *reward, double *u_next, float *obs, void *stream) {
    return step_impl<float>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, nullptr, stream);
}
int flock_step_aux_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, double *values,
                       double *network, double *reward, double *u_next, double *obs, const FlockAux *aux, void *stream) {
    return step_impl<double>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, aux, stream);
}
int flock_step_aux_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, float *values,
                       float *network, float *reward, double *u_next, float *obs, const FlockAux *aux, void *stream) {
    return step_impl<float>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, aux, stream);
}
int flock_rollout_f64(const FlockParams *p, int32_t steps, double *x_a, double *x_b, double *u_a, double *u_b,
                      double *values, double *network, double *reward, double *rewards_all, double *obs,
                      int32_t *final_in_b, void *stream) {
    return rollout_impl<double>(p, steps, x_a, x_b, u_a, u_b, values, network, reward, rewards_all, obs, final_in_b, stream);
}
int flock_rollout_f32(const FlockParams *p, int32_t steps, double *x_a, double *x_b, double *u_a, double *u_b,
                      float *values, float *network, float *reward, float *rewards_all, float *obs,
                      int32_t *final_in_b, void *stream) {
    return rollout_impl<float>(p, steps, x_a, x_b, u_a, u_b, values, network, reward, rewards_all, obs, final_in_b, stream);
}

int flock_reset_sample(const FlockParams *p, double r_max, double v_max, double v_bias, double min_dist, uint64_t seed,
                       int32_t max_tries, double *x_out, int32_t *tries_out, void *stream) {
    return reset_sample_impl(p, r_max, v_max, v_bias, min_dist, seed, max_tries, x_out, tries_out, stream);
}

size_t flock_swarm_scratch_bytes(const FlockParams *p) {
    if (!p || p->n_envs < 1 || p->n_agents < 1) return 0;
    int lo = p->row_lo, hi = p->row_hi;
    if (lo == 0 && hi == 0) hi = p->n_agents;
    if (lo < 0 || hi > p->n_agents || lo >= hi) return 0;
    return swarm_layout(p->n_envs, p->n_agents, hi - lo).bytes;
}
int flock_swarm_step_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, double *values,
                         double *network, double *reward, double *u_next, double *obs, void *scratch,
                         size_t scratch_bytes, void *stream) {
    return swarm_step_impl<double>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, scratch, scratch_bytes, stream);
}
int flock_swarm_step_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, float *values,
                         float *network, float *reward, double *u_next, float *obs, void *scratch, size_t scratch_bytes,
                         void *stream) {
    return swarm_step_impl<float>(p, x_in, u_in, x_out, values, network, reward, u_next, obs, scratch, scratch_bytes, stream);
}

int flock_aggregate_f64(const FlockParams *p, const double *x, int32_t F, const double *prev, const uint32_t *adj_bits,
                        int32_t L, int32_t pool, double *out, void *stream) {
    return aggregate_impl<double>(p, x, F, prev, adj_bits, L, pool, out, stream);
}
int flock_aggregate_f32(const FlockParams *p, const float *x, int32_t F, const float *prev, const uint32_t *adj_bits,
                        int32_t L, int32_t pool, float *out, void *stream) {
    return aggregate_impl<float>(p, x, F, prev, adj_bits, L, pool, out, stream);
}
int flock_repulsion_f64(const FlockParams *p, const double *pos, int32_t stride, const double *weights, double cutoff_r2,
                        double *out, void *stream) {
    return repulsion_impl<double>(p, pos, stride, weights, cutoff_r2, out, stream);
}
int flock_repulsion_f32(const FlockParams *p, const double *pos, int32_t stride, const float *weights, double cutoff_r2,
                        float *out, void *stream) {
    return repulsion_impl<float>(p, pos, stride, weights, cutoff_r2, out, stream);
}

void *flock_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void flock_host_free(void *p) { if (p) cudaFreeHost(p); }

int flock_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int flock_ctx_create(FlockCtx **out, const FlockParams *p, int32_t dtype_bytes, int32_t device) {
    if (!out || !p) return fail(FLOCK_EINVAL, "NULL argument");
    if (dtype_bytes != 4 && dtype_bytes != 8) return fail(FLOCK_EINVAL, "dtype_bytes must be 4 or 8");
    Plan pl;
    int rc = make_plan(p, dtype_bytes, pl);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    FlockCtx *c = new (std::nothrow) FlockCtx();
    if (!c) return fail(FLOCK_ENOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->p = *p; c->tbytes = dtype_bytes; c->device = device; c->cur = 0; c->u_cached_flags = -1;
    const size_t tb = (size_t)dtype_bytes;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc((void **)&c->x[i], n_x(c) * 8);
        if (e == cudaSuccess) e = cudaMalloc((void **)&c->u[i], n_u(c) * 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->u[i], 0, n_u(c) * 8, c->stream);
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->values, n_val(c) * tb);
    if (e == cudaSuccess) e = cudaMalloc(&c->network, n_net(c) * tb);
    if (e == cudaSuccess) e = cudaMalloc(&c->reward, (size_t)p->n_envs * tb);
    if (e == cudaSuccess) e = cudaMalloc(&c->obs, n_obs(c) * tb);
    if (e == cudaSuccess && (p->flags & FLOCK_KNN_OBS)) {
        const size_t hb = (size_t)p->n_envs * p->n_agents * 16;
        e = cudaMalloc((void **)&c->knn_hint, hb);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->knn_hint, 0, hb, c->stream);
    }
    if (e != cudaSuccess) {
        flock_ctx_destroy(c);
        return fail(e == cudaErrorMemoryAllocation ? FLOCK_ENOMEM : FLOCK_ECUDA, "ctx_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return FLOCK_OK;
}

void flock_ctx_destroy(FlockCtx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (int i = 0; i < 2; ++i) { cudaFree(c->x[i]); cudaFree(c->u[i]); }
    cudaFree(c->values); cudaFree(c->network); cudaFree(c->reward); cudaFree(c->obs);
    cudaFree(c->stat_a); cudaFree(c->stat_b); cudaFree(c->knn_hint);
    cudaGetLastError();
    delete c;
}

int flock_ctx_set_params(FlockCtx *c, const FlockParams *p) {
    if (!c || !p) return fail(FLOCK_EINVAL, "NULL argument");
    if (p->n_envs != c->p.n_envs || p->n_agents != c->p.n_agents) return fail(FLOCK_EINVAL, "E and N are fixed per context");
    Plan pl;
    int rc = make_plan(p, c->tbytes, pl);
    if (rc) return rc;
    c->p = *p;
    c->u_cached_flags = -1;
    return FLOCK_OK;
}

int flock_ctx_set_state(FlockCtx *c, const double *x_host, void *values_host, void *network_host, void *reward_host,
                        void *obs_host) {
    NvtxRange nvtx_("flock_ctx_set_state");
    if (!c || !x_host) return fail(FLOCK_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->x[c->cur], x_host, n_x(c) * 8, cudaMemcpyHostToDevice, c->stream));
    int rc = ctx_run(c, nullptr, nullptr, values_host != nullptr, network_host != nullptr, reward_host != nullptr,
                     obs_host != nullptr, c->u[0]);
    if (rc) return rc;
    c->u_cached_flags = c->p.flags & FLOCK_CENTRALIZED;
    return ctx_fetch(c, values_host, network_host, reward_host, obs_host);
}

int flock_ctx_upload_state(FlockCtx *c, const double *x_host) {
    if (!c || !x_host) return fail(FLOCK_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(c->x[c->cur], x_host, n_x(c) * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLOCK_OK;
}

int flock_ctx_step(FlockCtx *c, const double *u_host, void *values_host, void *network_host, void *reward_host,
                   void *obs_host) {
    NvtxRange nvtx_("flock_ctx_step");
    if (!c) return fail(FLOCK_EINVAL, "NULL context");
    CUDA_TRY(cudaSetDevice(c->device));
    if (u_host) {
        CUDA_TRY(cudaMemcpyAsync(c->u[0], u_host, n_u(c) * 8, cudaMemcpyHostToDevice, c->stream));
    } else if (c->u_cached_flags != (c->p.flags & FLOCK_CENTRALIZED)) {
        return fail(FLOCK_EINVAL, "no cached controller action for the current state / mode; call flock_ctx_controller or pass u");
    }
    int rc = ctx_run(c, c->u[0], c->x[c->cur ^ 1], values_host != nullptr, network_host != nullptr,
                     reward_host != nullptr, obs_host != nullptr, c->u[1]);
    if (rc) return rc;
    c->cur ^= 1;
    std::swap(c->u[0], c->u[1]);
    c->u_cached_flags = c->p.flags & FLOCK_CENTRALIZED;
    return ctx_fetch(c, values_host, network_host, reward_host, obs_host);
}

int flock_ctx_controller(FlockCtx *c, int32_t centralized, double *u_host) {
    NvtxRange nvtx_("flock_ctx_controller");
    if (!c || !u_host) return fail(FLOCK_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const int want = centralized ? FLOCK_CENTRALIZED : 0;
    if (c->u_cached_flags != want) {
        const int saved = c->p.flags;
        c->p.flags = (saved & ~FLOCK_CENTRALIZED) | want;
        int rc = ctx_run(c, nullptr, nullptr, false, false, false, false, c->u[0]);
        c->p.flags = saved;
        if (rc) return rc;
        c->u_cached_flags = ((saved & FLOCK_CENTRALIZED) == want) ? want : -1;   // cache only the env's own mode
    }
    CUDA_TRY(cudaMemcpyAsync(u_host, c->u[0], n_u(c) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLOCK_OK;
}

int flock_ctx_get_state(FlockCtx *c, double *x_host) {
    if (!c || !x_host) return fail(FLOCK_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(x_host, c->x[c->cur], n_x(c) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLOCK_OK;
}

int flock_ctx_get_stats(FlockCtx *c, void *vel_diffs_host, void *min_dists_host) {
    NvtxRange nvtx_("flock_ctx_get_stats");
    if (!c) return fail(FLOCK_EINVAL, "NULL context");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->p.n_envs * c->p.n_agents * c->tbytes;
    if (!c->stat_a) CUDA_TRY(cudaMalloc(&c->stat_a, bytes));
    if (!c->stat_b) CUDA_TRY(cudaMalloc(&c->stat_b, bytes));
    FlockAux aux;
    memset(&aux, 0, sizeof(aux));
    aux.vel_diff = vel_diffs_host ? c->stat_a : nullptr;
    aux.min_dist = min_dists_host ? c->stat_b : nullptr;
    // a pair pass on the current state with no other output (the cached action is left alone: u_next = NULL)
    int rc = ctx_run(c, nullptr, nullptr, false, false, false, false, nullptr, &aux);
    if (rc) return rc;
    if (vel_diffs_host) CUDA_TRY(cudaMemcpyAsync(vel_diffs_host, c->stat_a, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (min_dists_host) CUDA_TRY(cudaMemcpyAsync(min_dists_host, c->stat_b, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLOCK_OK;
}

}  // extern "C"
