/* flock_b200.h -- C ABI of libflock_b200.so: gym-flock's flocking step on B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of katetolstaya/gym-flock: the per-step update
 * of FlockingRelativeEnv / FlockingEnv.  Reference citations (file:line) are relative to
 * gym_flock/envs/flocking/ of the reference tree.
 *
 *   what the reference does per step                         where
 *   -------------------------------------------------------  ------------------------------
 *   step(u): Euler integration of x[N,4] with u*10           flocking_relative.py:91-105
 *   compute_helpers(): diff, r2, adjacency, 6 features        flocking_relative.py:111-134
 *   instant_cost(): -(var vx + var vy)                        flocking_relative.py:145-147
 *   controller(): Tanner potential gradient + alignment       flocking_relative.py:194-226
 *   get_observation(): 7-nearest-neighbour state differences  flocking.py:20-24
 *
 * The reference is Python/NumPy and has no FFI; the "binding" a maintainer adds is a ctypes stub
 * (see INTEGRATION.md).  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - State-side arrays are ALWAYS double: x[E][N][4] = (px,py,vx,vy), u[E][N][2] in [-1,1].
 *   - Observation-side arrays (values, network, reward, obs) are double for *_f64 and float for *_f32.
 *   - All device buffers are caller-owned, contiguous, 16-byte aligned (`reward`: element aligned; in row-sharded calls,
 *     where a caller may pass shard-sized buffers shifted back by row_lo rows, `values` needs 8 bytes and `network` only
 *     element alignment).  The library allocates nothing
 *     behind device-pointer entry points; they are asynchronous on `stream` (a cudaStream_t passed as
 *     void*), re-entrant, and never synchronise.
 *   - Every entry point returns 0 on success or a negative FLOCK_E* code; it never throws or exits.
 *     flock_last_error() returns a thread-local message for the last failure.
 */
#ifndef FLOCK_B200_H
#define FLOCK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOCK_ABI_VERSION 4

/* flags */
#define FLOCK_CENTRALIZED   0x1   /* controller(centralized=True), flocking_relative.py:200-206 */
#define FLOCK_EXACT_ORDER   0x2   /* f64 only: sequential-j, unfused IEEE arithmetic == NumPy bit for bit */
#define FLOCK_KNN_OBS       0x4   /* also write Flocking-v0's obs[N][4*knn] (flocking.py:20-24) */

/* error codes */
#define FLOCK_OK             0
#define FLOCK_EINVAL        -1   /* bad shape / flag / null pointer */
#define FLOCK_EALIGN        -2   /* pointer not 16-byte aligned */
#define FLOCK_EUNSUPPORTED  -3   /* size outside what the selected mode supports */
#define FLOCK_ECUDA         -4   /* CUDA runtime error (message in flock_last_error) */
#define FLOCK_ENOMEM        -5

typedef struct FlockParams {
    int32_t n_envs;         /* E: independent environments stepped in lockstep */
    int32_t n_agents;       /* N  (self.n_agents, flocking_relative.py:38) */
    double  comm_radius;    /* R  (:39) -- the UNSQUARED radius masks the potential gradient (:225) */
    double  comm_radius2;   /* R*R (:44) -- adjacency threshold, strict < (:117) */
    double  dt;             /* (:40) */
    double  action_scalar;  /* 10.0 (:64) */
    int32_t flags;          /* FLOCK_* */
    int32_t knn;            /* neighbours in obs; 7 (flocking.py:9); only with FLOCK_KNN_OBS */
    int32_t row_lo;         /* swarm sharding: this call computes rows [row_lo,row_hi) of every env; */
    int32_t row_hi;         /*   0,0 (or 0,N) = all rows.  x_in / u_in / x_out are always full [E][N][*].  Outputs are
                             *   addressed by GLOBAL row but only rows [row_lo,row_hi) are ever touched, so a caller
                             *   may pass shard-sized buffers shifted back by row_lo rows. */
    int32_t lanes;          /* threads cooperating on one row (power of two <= 32); 0 = choose */
    int32_t reserved;       /* 0.  (Diagnostics word of tools/ and bench.py A/B runs -- CTA size, phase switches, envs per
                             *   CTA; a non-zero value can switch whole phases off, so bindings must pass 0.) */
    /* ---- flocking variants through the same kernel; all zero = FlockingRelative / Flocking ---- */
    double  step_scale;     /* step(): a = u * step_scale.  0 = action_scalar (flocking_relative.py:96);
                             *   1.0 for FlockingLeader (self.u = u, flocking_leader.py:24); 6.0 for FlockingStochastic */
    double  state_scale;    /* FlockingStochastic: x *= scale before, x /= scale after the Euler step
                             *   (flocking_stoch.py:18,33); 0 or 1 = none */
    double  u_clip;         /* FlockingStochastic: clip the action before the step and the controller output to
                             *   +-u_clip (flocking_stoch.py:16,43); 0 = none */
    int32_t n_leaders;      /* FlockingLeader: agents [0, n_leaders) are not actuated (flocking_leader.py:11-12,27-31) */
    int32_t n_obstacles;    /* FlockingObstacle: agents [0, n_obstacles) are not actuated and every velocity difference with
                               one of them is zero (flocking_obstacle.py:19-20,33-50,81-82); env-resident path only */
} FlockParams;

int         flock_abi_version(void);
const char *flock_last_error(void);
int         flock_device_count(void);
/* page-locked host memory for the host-buffer API (so NumPy-side bindings can get full PCIe rate) */
void       *flock_host_alloc(size_t bytes);
void        flock_host_free(void *p);

/* One fused step on device buffers.
 *   x' = u_in ? integrate(x_in, u_in) : x_in          (flocking_relative.py:96-105)
 *   x_out (if non-NULL) <- x'   (ALL rows of every env, also in row-sharded calls; x_out may NOT alias x_in)
 *   then ONE pass over the ordered pairs of x' produces, for rows [row_lo,row_hi):
 *     values [E][N][6]      state_values   (:124-129)           (nullable)
 *     network[E][N][N]      adj_mat_mean   (:117-122,131-132)   (nullable: skip the dense write)
 *     reward [E]            instant_cost   (:145-147)           (nullable)
 *     u_next [E][N][2]      controller() on x', double          (nullable)
 *     obs    [E][N][4*knn]  FlockingEnv.get_observation()       (FLOCK_KNN_OBS; nullable)
 * Replaces: FlockingRelativeEnv.step + compute_helpers + instant_cost + the following controller().
 */
int flock_step_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out,
                   double *values, double *network, double *reward, double *u_next, double *obs,
                   void *stream);
int flock_step_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out,
                   float *values, float *network, float *reward, double *u_next, float *obs,
                   void *stream);

/* Optional extra outputs of one step (SURVEY.md 5 "aux", 8f #4).  Every pointer is nullable; pass aux = NULL for none.
 * Device buffers, caller-owned; observation-side types follow the entry point (double for *_f64, float for *_f32). */
typedef struct FlockAux {
    int32_t  *nonfinite;    /* [E]  set to 1 when a feature, state or action of the env is NaN / Inf (coincident agents give
                             *      0/0 in the reference, flocking_relative.py:124).  Never cleared by the library. */
    uint32_t *adj_bits;     /* [E][N][flock_adj_words(N)]  the exact adjacency as bits: bit (j & 31) of word (j >> 5) of row i
                             *      is r2[i][j] < comm_radius2 (:117).  The 'bitmask' adjacency mode: 32x smaller than the
                             *      dense fp32 network; with network = NULL the N^2 write is skipped altogether.  The call
                             *      then runs one thread per row (lanes is ignored). */
    int32_t  *degree;       /* [E][N]  neighbours per row (the divisor of adj_mean, before the 0 -> 1 fix, :120-121) */
    void     *min_dist;     /* [E][N]  get_stats()['min_dists']: min_j sqrt(r2[i][j]) (:141-142) */
    void     *vel_diff;     /* [E][N]  get_stats()['vel_diffs']: |v_i - mean_j v_j| (:138-140) */
    uint16_t *knn_hint;     /* [E][N][8] in/out, 16-byte aligned; FLOCK_KNN_OBS only.  The 7 nearest neighbours of each row at
                             *      the previous step (slot 7 unused).  Read before the step as a SEARCH-RADIUS hint -- the 7
                             *      nearest now are among the agents at least as close as the farthest of any 7 agents -- and
                             *      rewritten with this step's neighbours.  Contents never affect results, only speed: zeros,
                             *      stale or arbitrary values make the row search all agents.  Batched path (one thread per
                             *      row, n_agents <= 512) only; ignored elsewhere. */
} FlockAux;

/* words per row of FlockAux.adj_bits: ceil(n_agents / 32) */
int flock_adj_words(int32_t n_agents);

/* flock_step_* with the optional outputs above.  Replaces additionally: FlockingRelativeEnv.get_stats(). */
int flock_step_aux_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out,
                       double *values, double *network, double *reward, double *u_next, double *obs,
                       const FlockAux *aux, void *stream);
int flock_step_aux_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out,
                       float *values, float *network, float *reward, double *u_next, float *obs,
                       const FlockAux *aux, void *stream);

/* Graph-signal aggregation on the adjacency the step produced, from its BITS (FlockAux.adj_bits; never the dense [N,N]):
 *   x    [E][N][F]      the nodes' current features (e.g. state_values, F = 6)
 *   prev [E][N][F*L]    last step's aggregate, or NULL
 *   out  [E][N][F*L]    prev != NULL (delayed aggregation, gym_flock/envs/old/flocking_multi.py:180-196):
 *                         out[i] = [ x[i] | pool_{j in N(i)} prev[j][0 : F*(L-1)] ]
 *                       prev == NULL (K-hop of the current signal): out[i] = [ x[i] | (S x)[i] | ... | (S^(L-1) x)[i] ]
 *   pool  FLOCK_POOL_MEAN: np.nanmean over the connected senders (:36, :262-265; S = D^-1 A = state_network with
 *         mean_pooling; a node without neighbours receives 0), FLOCK_POOL_SUM: S = A.
 * f64 reproduces NumPy bit for bit (ascending-j sums, true divide).  out must not alias prev or x.  p: n_envs, n_agents. */
#define FLOCK_POOL_MEAN 0
#define FLOCK_POOL_SUM  1
int flock_aggregate_f64(const FlockParams *p, const double *x, int32_t F, const double *prev, const uint32_t *adj_bits,
                        int32_t L, int32_t pool, double *out, void *stream);
int flock_aggregate_f32(const FlockParams *p, const float *x, int32_t F, const float *prev, const uint32_t *adj_bits,
                        int32_t L, int32_t pool, float *out, void *stream);

/* The sheep controller of the shepherding env (gym_flock/envs/shepherding/shepherding.py:164-178): the same all-pairs
 * inverse-distance shape as the flocking potential.  out[e][i] = sum_j weights[j] * (p_i - p_j) / r2_ij over the agents with
 * r2 <= cutoff_r2 (2.0 in the reference; the diagonal is excluded).  pos[E][N][stride] doubles with x, y in columns 0, 1
 * (stride = 3 for the reference's (x, y, heading) state); weights[N] per SOURCE agent (force_weights, :50); out[E][N][2].
 * f64 follows the reference's operation order (bit-exact); the caller keeps rows n_shepherds.. (:178). */
int flock_repulsion_f64(const FlockParams *p, const double *pos, int32_t stride, const double *weights, double cutoff_r2,
                        double *out, void *stream);
int flock_repulsion_f32(const FlockParams *p, const double *pos, int32_t stride, const float *weights, double cutoff_r2,
                        float *out, void *stream);

/* `steps` fused steps with the internal controller (the driver loop {u=controller(); step(u)}),
 * without a per-step host round trip.  Batches of small envs (n_agents <= 512, full rows, no FLOCK_EXACT_ORDER) run as
 * ONE persistent launch that keeps the state in shared memory between steps (every step still writes all of its
 * outputs); everything else is `steps` launches queued back to back.  Both produce the bits of `steps` flock_step_* calls.
 *   x_a / x_b: ping-pong state buffers, x_a holds the start state; returns in *final_in_b whether the
 *              final state is in x_b (1) or x_a (0).   u_a / u_b: ping-pong actions; u_a must hold
 *              controller(x_a) on entry (as produced by a previous flock_step_* call's u_next).
 *   Observation-side outputs are overwritten every step (they hold the last step on return);
 *   rewards_all (nullable) is [steps][E] and receives every step's reward.
 */
int flock_rollout_f64(const FlockParams *p, int32_t steps, double *x_a, double *x_b, double *u_a, double *u_b,
                      double *values, double *network, double *reward, double *rewards_all, double *obs,
                      int32_t *final_in_b, void *stream);
int flock_rollout_f32(const FlockParams *p, int32_t steps, double *x_a, double *x_b, double *u_a, double *u_b,
                      float *values, float *network, float *reward, float *rewards_all, float *obs,
                      int32_t *final_in_b, void *stream);

/* Large swarms (N beyond what one CTA can stage, e.g. 65 536 agents): same semantics as flock_step_* (x_out is
 * required), computed by an O(N) preparation kernel plus a row kernel that streams all agents' positions through
 * shared memory with TMA bulk loads.  `scratch` is caller-owned device memory of flock_swarm_scratch_bytes(p) bytes
 * (p with the row range of the call: the second tier of the per-row candidate lists scales with the rows computed),
 * 256-byte aligned; its contents need not survive between calls.  FLOCK_EXACT_ORDER is not offered here;
 * lanes must be 8, 16 or 32 (0 = 32; use the same value on every rank for shard-independent bits).
 * Replaces the same reference lines as flock_step_* for a single giant swarm (BASELINE config 5). */
size_t flock_swarm_scratch_bytes(const FlockParams *p);
int flock_swarm_step_f64(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, double *values,
                         double *network, double *reward, double *u_next, double *obs, void *scratch,
                         size_t scratch_bytes, void *stream);
int flock_swarm_step_f32(const FlockParams *p, const double *x_in, const double *u_in, double *x_out, float *values,
                         float *network, float *reward, double *u_next, float *obs, void *scratch, size_t scratch_bytes,
                         void *stream);

/* Batched reset(): the reference's rejection sampler (flocking_relative.py:156-192) for every env at once, on the device.
 * Each env draws whole configurations -- positions uniform in the disk of radius sqrt(r_max) (:167-170), velocities
 * U(-v_max, v_max) + a common bias U(-v_bias, v_bias)^2 (:172-174) -- until the closest pair is >= min_dist apart (0.1 in
 * the reference, :160) and every agent has >= 2 neighbours with r2 < comm_radius2 (:164,183-184), at most max_tries times.
 * x_out[E][N][4] (device, 16-byte aligned) receives the accepted states; tries_out[E] (device int32, nullable) the number
 * of draws used, or -max_tries where no admissible configuration was found (x_out then holds the last draw).
 * Counter-based generator (Philox4x32-10) keyed by (seed, env): deterministic, independent of the batch size.  Parity
 * with the reference is statistical (same law, same acceptance test); NumPy's MT19937 stream is not reproduced.
 * Asynchronous on `stream`; the caller runs flock_step_* (u_in = NULL) afterwards for the observation. */
int flock_reset_sample(const FlockParams *p, double r_max, double v_max, double v_bias, double min_dist, uint64_t seed,
                       int32_t max_tries, double *x_out, int32_t *tries_out, void *stream);

/* ---- host-buffer context API (what a NumPy-only binding uses; no torch needed) ----------------
 * A context owns device buffers + pinned staging for (E, N); calls take HOST pointers, copy in,
 * launch, copy out and synchronise before returning.  dtype_bytes = 8 (f64) or 4 (f32).
 */
typedef struct FlockCtx FlockCtx;

int  flock_ctx_create(FlockCtx **out, const FlockParams *p, int32_t dtype_bytes, int32_t device);
void flock_ctx_destroy(FlockCtx *ctx);
int  flock_ctx_set_params(FlockCtx *ctx, const FlockParams *p);          /* E, N must not change */
/* x <- host state [E][N][4]; runs the pair pass on it (no integration): the reference's
 * `self.x = x0; self.compute_helpers()`.  Outputs nullable; observation-side outputs have the
 * context's dtype. */
int  flock_ctx_set_state(FlockCtx *ctx, const double *x_host,
                         void *values_host, void *network_host, void *reward_host, void *obs_host);
/* step(u): u_host [E][N][2] (NULL = use the cached controller action).  Outputs nullable. */
int  flock_ctx_step(FlockCtx *ctx, const double *u_host,
                    void *values_host, void *network_host, void *reward_host, void *obs_host);
/* controller(centralized): copies the controller action for the CURRENT state to u_host [E][N][2]. */
int  flock_ctx_controller(FlockCtx *ctx, int32_t centralized, double *u_host);
int  flock_ctx_get_state(FlockCtx *ctx, double *x_host);
/* x <- host state WITHOUT running the pair pass (the reference's in-place edit of env.x, e.g. flocking_leader.py:38-39):
 * the cached controller action and the observation-side outputs keep describing the previous state until the next step. */
int  flock_ctx_upload_state(FlockCtx *ctx, const double *x_host);
/* get_stats() for the current state (flocking_relative.py:136-143): vel_diffs[E][N], min_dists[E][N] in the context's
 * observation dtype; either may be NULL. */
int  flock_ctx_get_stats(FlockCtx *ctx, void *vel_diffs_host, void *min_dists_host);

#ifdef __cplusplus
}
#endif
#endif /* FLOCK_B200_H */
