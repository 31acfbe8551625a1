// launch_env_f64.cu -- instantiations of flock_env_kernel with double observation outputs (fast and exact-order modes).
#include "flock_host.hpp"
#include "flock_env_kernel.cuh"

namespace {
template <int LANES, bool EXACT, bool KNN>
int go(const flock::KArgs &k, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    auto kern = flock::flock_env_kernel<double, LANES, EXACT, KNN>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        configured_dev = dev;
    }
    return launch_pdl(kern, dim3(grid), dim3(threads), smem, st, pdl, k);
}
template <bool KNN>
int by_lanes(const flock::KArgs &k, int lanes, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    switch (lanes) {
        case 1: return go<1, false, KNN>(k, grid, threads, smem, st, pdl);
        case 2: return go<2, false, KNN>(k, grid, threads, smem, st, pdl);
        case 4: return go<4, false, KNN>(k, grid, threads, smem, st, pdl);
        case 8: return go<8, false, KNN>(k, grid, threads, smem, st, pdl);
        case 16: return go<16, false, KNN>(k, grid, threads, smem, st, pdl);
        case 32: return go<32, false, KNN>(k, grid, threads, smem, st, pdl);
    }
    return flock_fail(FLOCK_EINVAL, "lanes=%d", lanes);
}
}  // namespace

int launch_env_f64(const flock::KArgs &k, int lanes, bool exact, bool knn, int grid, int threads, size_t smem, cudaStream_t st,
                   bool pdl) {
    if (exact) return knn ? go<1, true, true>(k, grid, threads, smem, st, pdl) : go<1, true, false>(k, grid, threads, smem, st, pdl);
    return knn ? by_lanes<true>(k, lanes, grid, threads, smem, st, pdl) : by_lanes<false>(k, lanes, grid, threads, smem, st, pdl);
}
