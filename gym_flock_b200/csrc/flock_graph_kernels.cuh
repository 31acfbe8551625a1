// flock_graph_kernels.cuh -- the steps either side of the flocking step that share its all-pairs / adjacency structure
// (SURVEY.md 8f #3 and #4), sm_100a.
//
//   flock_aggregate_kernel : graph-signal aggregation a learner applies to (state_values, state_network), computed from
//                            the EXACT ADJACENCY BITS the step kernel emits (FlockAux.adj_bits) -- the dense [N,N] network
//                            is never materialised.           gym_flock/envs/old/flocking_multi.py:180-196, 214-265
//                              delayed : out[i] = [ x[i] | pool_{j in N(i)} prev[j][0 : F (L-1)] ]     (aggregate(), :191-195)
//                              K-hop   : out[i] = [ x[i] | (S x)[i] | ... | (S^(L-1) x)[i] ],  S = D^-1 A (mean) or A (sum)
//                            pool = np.nanmean over the connected senders (:36, :262-265): ascending-j sum, then a true divide
//                            by the neighbour count; rows without neighbours give 0 (:264).
//   flock_repulsion_kernel : the sheep controller of the shepherding env: sum_j w_j d / r2 over agents closer than a cutoff
//                            (r2 > 2 -> inf, diagonal inf).          gym_flock/envs/shepherding/shepherding.py:164-178
//                            f64: the reference's operation order, every operation rounded separately, ascending j.
#pragma once
#include "flock_common.cuh"

namespace flock {

struct AggArgs {
    const void *x;            // [E][N][F]
    const void *prev;         // [E][N][F*L] (delayed) or NULL (K-hop)
    const unsigned *bits;     // [E][N][WB]
    void *out;                // [E][N][F*L]
    int E, N, F, L, WB, pool;
};

struct ShepArgs {
    const double *pos;        // [E][N][stride]: x, y in columns 0, 1
    const void *weights;      // [N] weight of SOURCE agent j (force_weights, shepherding.py:50)
    void *out;                // [E][N][2]
    int E, N, stride;
    double cutoff_r2;         // 2.0 (:173)
};

}  // namespace flock
#ifndef FLOCK_ARGS_ONLY
namespace flock {

template <typename T> __device__ __forceinline__ T agg_finish(T sum, int cnt, int pool);
template <> __device__ __forceinline__ double agg_finish<double>(double sum, int cnt, int pool) {
    if (pool != 0) return sum;
    return cnt > 0 ? __ddiv_rn(sum, (double)cnt) : 0.0;        // nanmean: NaN (no neighbour) -> 0
}
template <> __device__ __forceinline__ float agg_finish<float>(float sum, int cnt, int pool) {
    if (pool != 0) return sum;
    return cnt > 0 ? __fdiv_rn(sum, (float)cnt) : 0.0f;
}
__device__ __forceinline__ double agg_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float agg_add(float a, float b) { return __fadd_rn(a, b); }

// one CTA per env; thread (i, c) of the hop pools column c of the source block over the neighbours of row i, ascending j
template <typename T>
__global__ void __launch_bounds__(256) flock_aggregate_kernel(const AggArgs a) {
    const int env = blockIdx.x, N = a.N, F = a.F, L = a.L, WB = a.WB;
    const int FL = F * L;
    const T *x = reinterpret_cast<const T *>(a.x) + (size_t)env * N * F;
    const T *prev = a.prev ? reinterpret_cast<const T *>(a.prev) + (size_t)env * N * FL : nullptr;
    T *out = reinterpret_cast<T *>(a.out) + (size_t)env * N * FL;
    const unsigned *bits = a.bits + (size_t)env * N * WB;
    pdl_wait();
    for (int idx = threadIdx.x; idx < N * F; idx += blockDim.x) {           // block 0: the node's own features
        const int i = idx / F, c = idx - i * F;
        out[(size_t)i * FL + c] = x[idx];
    }
    auto pool_block = [&](const T *src, const int src_stride, const int C, const int dst_col) {
        for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
            const int i = idx / C, c = idx - i * C;
            const unsigned *row = bits + (size_t)i * WB;
            T sum = T(0);
            int cnt = 0;
            for (int w = 0; w < WB; ++w) {
                unsigned b = row[w];
                cnt += __popc(b);
                while (b) {
                    const int j = w * 32 + __ffs(b) - 1;
                    b &= b - 1;
                    sum = agg_add(sum, src[(size_t)j * src_stride + c]);
                }
            }
            out[(size_t)i * FL + dst_col + c] = agg_finish<T>(sum, cnt, a.pool);
        }
    };
    if (prev != nullptr) {
        pool_block(prev, FL, F * (L - 1), F);                               // delayed: one hop of last step's aggregate
    } else {
        for (int k = 1; k < L; ++k) {                                       // K-hop: block k = S * block k-1
            __syncthreads();                                                // block k-1 (written by this CTA) is complete
            pool_block(out + (size_t)(k - 1) * F, FL, F, k * F);
        }
    }
}

// thread per row i, sequential j (N is a few tens)
template <typename T>
__global__ void __launch_bounds__(128) flock_repulsion_kernel(const ShepArgs a) {
    const int env = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const double *p = a.pos + (size_t)env * a.N * a.stride;
    const T *w = reinterpret_cast<const T *>(a.weights);
    const double xi = p[(size_t)i * a.stride], yi = p[(size_t)i * a.stride + 1];
    if constexpr (sizeof(T) == 8) {
        double sx = 0.0, sy = 0.0;
        for (int j = 0; j < a.N; ++j) {
            const double dx = __dsub_rn(xi, p[(size_t)j * a.stride]), dy = __dsub_rn(yi, p[(size_t)j * a.stride + 1]);   // :154-155
            double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));                                                 // :156
            if (r2 > a.cutoff_r2 || j == i) r2 = CUDART_INF;                                                             // :173-174
            sx = __dadd_rn(sx, __dmul_rn((double)w[j], __ddiv_rn(dx, r2)));                                              // :175-176
            sy = __dadd_rn(sy, __dmul_rn((double)w[j], __ddiv_rn(dy, r2)));
        }
        double *o = reinterpret_cast<double *>(a.out) + ((size_t)env * a.N + i) * 2;
        o[0] = sx; o[1] = sy;
    } else {
        float sx = 0.f, sy = 0.f;
        for (int j = 0; j < a.N; ++j) {
            const float dx = (float)(xi - p[(size_t)j * a.stride]), dy = (float)(yi - p[(size_t)j * a.stride + 1]);
            const float r2 = fmaf(dx, dx, dy * dy);
            const float inv = (r2 > (float)a.cutoff_r2 || j == i) ? 0.f : __fdividef(1.f, r2);
            sx = fmaf((float)w[j] * dx, inv, sx); sy = fmaf((float)w[j] * dy, inv, sy);
        }
        float *o = reinterpret_cast<float *>(a.out) + ((size_t)env * a.N + i) * 2;
        o[0] = sx; o[1] = sy;
    }
}

}  // namespace flock
#endif  // FLOCK_ARGS_ONLY
