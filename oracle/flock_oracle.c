/* TEST INFRASTRUCTURE ONLY -- scalar C restatement of gym-flock's flocking hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
 * library.  The product (gym_flock_b200) never links or calls it.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors for this path
 * (SURVEY.md 4, 8c).  This file is pinned against the reference code itself:
 * tests/test_oracle_vs_reference.py (runs where /root/reference exists) and
 * tests/test_oracle_golden.py (committed outputs of the reference, tests/golden/)
 * assert BIT equality of state, features, adjacency, actions, rewards and kNN
 * observations over 200-step rollouts.
 *
 * Exactness recipe: one row i at a time, j = 0..N-1 in order (NumPy's
 * np.sum(axis=1) over [N,N,6] accumulates sequentially in j), every product and sum
 * rounded separately (build with -ffp-contract=off, no -ffast-math), IEEE divide.
 *
 * Citations are to /root/reference/gym_flock/envs/flocking/flocking_relative.py
 * unless another file is named.  State layout: x[N][4] = (px, py, vx, vy), row major.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 4
#define NF 6
#define KNN_MAX 16

/* One row of compute_helpers(), :111-134: values_i[6]; adj_row / adj_mean_row [N] may be NULL; returns the degree. */
static double helpers_row(const double *x, int n, int i, double comm_radius2,
                          double *values_i, double *adj_row, double *adj_mean_row)
{
    const double pxi = x[i * NX + 0], pyi = x[i * NX + 1], vxi = x[i * NX + 2], vyi = x[i * NX + 3];
    double f[NF] = {0, 0, 0, 0, 0, 0};
    double deg = 0.0;
    for (int j = 0; j < n; ++j) {
        const double dx = pxi - x[j * NX + 0], dy = pyi - x[j * NX + 1];       /* :113 */
        const double dvx = vxi - x[j * NX + 2], dvy = vyi - x[j * NX + 3];
        double r2 = (dx * dx) + (dy * dy);                                     /* :114 */
        if (i == j) r2 = INFINITY;                                             /* :115 */
        const double a = (r2 < comm_radius2) ? 1.0 : 0.0;                      /* :117 */
        deg += a;                                                              /* :120 */
        const double r4 = r2 * r2;
        f[0] += dvx * a;                                                       /* :124-128 */
        f[1] += (dx / r4) * a;
        f[2] += (dx / r2) * a;
        f[3] += dvy * a;
        f[4] += (dy / r4) * a;
        f[5] += (dy / r2) * a;
        if (adj_row) adj_row[j] = a;
    }
    for (int k = 0; k < NF; ++k) values_i[k] = f[k];
    if (adj_mean_row) {
        const double d = (deg == 0.0) ? 1.0 : deg;                             /* :121 */
        for (int j = 0; j < n; ++j) {
            const double dx = pxi - x[j * NX + 0], dy = pyi - x[j * NX + 1];
            double r2 = (dx * dx) + (dy * dy);
            if (i == j) r2 = INFINITY;
            adj_mean_row[j] = ((r2 < comm_radius2) ? 1.0 : 0.0) / d;           /* :122 */
        }
    }
    return deg;
}

/* compute_helpers(), :111-134.  values[N][6]; adj / adj_mean [N][N] may be NULL. */
void flk_helpers_f64(const double *x, int n, double comm_radius2,
                     double *values, double *adj, double *adj_mean)
{
    for (int i = 0; i < n; ++i)
        helpers_row(x, n, i, comm_radius2, values + (size_t)i * NF, adj ? adj + (size_t)i * n : NULL,
                    adj_mean ? adj_mean + (size_t)i * n : NULL);
}

/* One row of controller() + potential_grad(), :194-226.  u_i[2] in [-1, 1]. */
static void controller_row(const double *x, int n, int i, double comm_radius, double comm_radius2,
                           int centralized, double *u_i)
{
    const double pxi = x[i * NX + 0], pyi = x[i * NX + 1], vxi = x[i * NX + 2], vyi = x[i * NX + 3];
    double s_dvx = 0, s_dvy = 0, s_gx = 0, s_gy = 0;
    for (int j = 0; j < n; ++j) {
        const double dx = pxi - x[j * NX + 0], dy = pyi - x[j * NX + 1];
        double dvx = vxi - x[j * NX + 2], dvy = vyi - x[j * NX + 3];
        double r2 = (dx * dx) + (dy * dy);
        if (i == j) r2 = INFINITY;
        const double r4 = r2 * r2;
        double gx = (-2.0 * (dx / r4)) + (2.0 * (dx / r2));                    /* :224 */
        double gy = (-2.0 * (dy / r4)) + (2.0 * (dy / r2));
        if (r2 > comm_radius) { gx = 0.0; gy = 0.0; }                          /* :225 unsquared R, strict > */
        if (!centralized) {                                                    /* :205-206 */
            const double a = (r2 < comm_radius2) ? 1.0 : 0.0;
            dvx *= a; dvy *= a; gx *= a; gy *= a;
        }
        s_dvx += dvx; s_dvy += dvy; s_gx += gx; s_gy += gy;                    /* :208 */
    }
    double ux = (-s_gx) - s_dvx;                                               /* :209 */
    double uy = (-s_dvy) - s_gy;
    ux = fmin(fmax(ux, -10.0), 10.0);                                          /* :210 */
    uy = fmin(fmax(uy, -10.0), 10.0);
    u_i[0] = ux / 10.0;                                                        /* :211 */
    u_i[1] = uy / 10.0;
}

/* controller() + potential_grad(), :194-226.  u[N][2] in [-1, 1]. */
void flk_controller_f64(const double *x, int n, double comm_radius, double comm_radius2,
                        int centralized, double *u)
{
    for (int i = 0; i < n; ++i) controller_row(x, n, i, comm_radius, comm_radius2, centralized, u + (size_t)i * 2);
}

/* step() integrator, :96-105, in place. */
void flk_integrate_f64(double *x, const double *u, int n, double dt, double action_scalar)
{
    for (int i = 0; i < n; ++i) {
        const double ax = u[i * 2 + 0] * action_scalar, ay = u[i * 2 + 1] * action_scalar;   /* :96 */
        const double px = x[i * NX + 0], py = x[i * NX + 1], vx = x[i * NX + 2], vy = x[i * NX + 3];
        x[i * NX + 0] = (px + (vx * dt)) + (((ax * dt) * dt) * 0.5);                         /* :99 */
        x[i * NX + 1] = (py + (vy * dt)) + (((ay * dt) * dt) * 0.5);                         /* :101 */
        x[i * NX + 2] = vx + (ax * dt);                                                      /* :103 */
        x[i * NX + 3] = vy + (ay * dt);                                                      /* :105 */
    }
}

/* instant_cost(), :145-147: -(var(vx) + var(vy)), population variance, NumPy's two-pass algorithm. */
double flk_reward_f64(const double *x, int n)
{
    double sx = 0, sy = 0;
    for (int i = 0; i < n; ++i) { sx += x[i * NX + 2]; sy += x[i * NX + 3]; }
    const double mx = sx / (double)n, my = sy / (double)n;
    double qx = 0, qy = 0;
    for (int i = 0; i < n; ++i) {
        const double ex = x[i * NX + 2] - mx, ey = x[i * NX + 3] - my;
        qx += ex * ex; qy += ey * ey;
    }
    return -1.0 * ((qx / (double)n) + (qy / (double)n));
}

/* FlockingEnv.get_observation(), flocking.py:20-24: the k smallest r2 of row i in ascending order
 * (ties: lowest index first; the reference's argsort leaves exact ties unspecified), block m =
 * x_i - x_nbr(m) over all 4 state components.  Requires n >= k (reference: IndexError otherwise). */
static void knn_row(const double *x, int n, int i, int k, double *obs_i, double *r2_sorted)
{
    double best[KNN_MAX + 2]; int idx[KNN_MAX + 2]; int cnt = 0;
    const int kk = r2_sorted ? k + 2 : k;            /* diagnostics: the k+2 smallest r2 (tie analysis in the tests) */
    for (int j = 0; j < n; ++j) {
        const double dx = x[i * NX + 0] - x[j * NX + 0], dy = x[i * NX + 1] - x[j * NX + 1];
        double r2 = (dx * dx) + (dy * dy);
        if (i == j) r2 = INFINITY;
        if (cnt < kk || r2 < best[cnt - 1]) {
            int p = (cnt < kk) ? cnt++ : kk - 1;
            while (p > 0 && r2 < best[p - 1]) { best[p] = best[p - 1]; idx[p] = idx[p - 1]; --p; }
            best[p] = r2; idx[p] = j;
        }
    }
    for (int m = 0; m < k; ++m)
        for (int c = 0; c < NX; ++c)
            obs_i[m * NX + c] = x[i * NX + c] - x[idx[m] * NX + c];
    if (r2_sorted) for (int m = 0; m < kk; ++m) r2_sorted[m] = (m < cnt) ? best[m] : INFINITY;
}

void flk_knn_obs_f64(const double *x, int n, int k, double *obs)
{
    if (k > KNN_MAX) k = KNN_MAX;
    for (int i = 0; i < n; ++i) knn_row(x, n, i, k, obs + (size_t)i * (NX * k), NULL);
}

/* A SUBSET of rows of one env (the giant-swarm tests: N = 65 536 is 4.3e9 ordered pairs, a row sample is seconds):
 * the same per-row functions as above for rows[0..nrows).  Outputs (NULL to skip): values[nrows][6], deg[nrows],
 * adj_mean[nrows][N], u[nrows][2], obs[nrows][4*knn], r2_sorted[nrows][knn+2] (the knn+2 smallest r2 of the row, ascending). */
void flk_rows_f64(const double *x, int n, const int *rows, int nrows, double comm_radius, double comm_radius2,
                  int centralized, int knn, double *values, double *deg, double *adj_mean, double *u, double *obs,
                  double *r2_sorted)
{
    if (knn > KNN_MAX) knn = KNN_MAX;
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < nrows; ++q) {
        const int i = rows[q];
        double v[NF];
        const double d = helpers_row(x, n, i, comm_radius2, values ? values + (size_t)q * NF : v, NULL,
                                     adj_mean ? adj_mean + (size_t)q * n : NULL);
        if (deg) deg[q] = d;
        if (u) controller_row(x, n, i, comm_radius, comm_radius2, centralized, u + (size_t)q * 2);
        if (obs && knn > 0)
            knn_row(x, n, i, knn, obs + (size_t)q * NX * knn, r2_sorted ? r2_sorted + (size_t)q * (knn + 2) : NULL);
    }
}

/* Driver loop of SURVEY.md 3.6 on E independent envs: steps x { u = controller(); step(u) }.
 * x[E][N][4] in/out.  Optional outputs (NULL to skip), all for the state after the last step:
 * values[E][N][6], adj_mean[E][N][N], u_last[E][N][2] (the action of the last step),
 * rewards[E][steps], obs[E][N][4*knn] (knn > 0).  nthreads <= 0: all cores. */
void flk_rollout_f64(double *x, int n_envs, int n, int steps, double comm_radius, double dt,
                     int centralized, int knn, double *values, double *adj_mean, double *u_last,
                     double *rewards, double *obs, int nthreads)
{
    const double r2c = comm_radius * comm_radius;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        double *u = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        double *v = (double *)malloc(sizeof(double) * NF * (size_t)n);
#pragma omp for schedule(dynamic, 1)
        for (int e = 0; e < n_envs; ++e) {
            double *xe = x + (size_t)e * n * NX;
            for (int t = 0; t < steps; ++t) {
                flk_controller_f64(xe, n, comm_radius, r2c, centralized, u);
                flk_integrate_f64(xe, u, n, dt, 10.0);
                const int last = (t == steps - 1);
                /* the reference recomputes the helpers every step (:107); keep the work, drop the copies */
                flk_helpers_f64(xe, n, r2c, (last && values) ? values + (size_t)e * n * NF : v, NULL,
                                (last && adj_mean) ? adj_mean + (size_t)e * n * n : NULL);
                const double r = flk_reward_f64(xe, n);
                if (rewards) rewards[(size_t)e * steps + t] = r;
            }
            if (steps == 0 && values)
                flk_helpers_f64(xe, n, r2c, values + (size_t)e * n * NF, NULL,
                                adj_mean ? adj_mean + (size_t)e * n * n : NULL);
            if (u_last && steps > 0) memcpy(u_last + (size_t)e * n * 2, u, sizeof(double) * 2 * (size_t)n);
            if (obs && knn > 0) flk_knn_obs_f64(xe, n, knn, obs + (size_t)e * n * NX * knn);
        }
        free(u); free(v);
    }
}

int flk_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
