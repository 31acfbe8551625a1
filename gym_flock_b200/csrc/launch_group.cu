// launch_group.cu -- instantiations of flock_group_kernel (whole envs per CTA, fused multi-step rollouts).
// N = 100 (the reference's default, flocking_relative.py:38) and N = 50 get compile-time sizes; any other N runs the NC = 0 build.
#include "flock_host.hpp"
#include "flock_group_kernel.cuh"

namespace {
template <typename TO, bool KNN, int NC, int GMS>
int go2(const flock::GArgs &g, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    auto kern = flock::flock_group_kernel<TO, KNN, NC, GMS>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        configured_dev = dev;
    }
    return launch_pdl(kern, dim3(grid), dim3(threads), smem, st, pdl, g);
}
template <typename TO, bool KNN, int NC>
int go(const flock::GArgs &g, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    // mask stride: 256 for CTAs of up to 256 threads (less shared memory, more CTAs per SM), else 512
    return threads <= 256 ? go2<TO, KNN, NC, 256>(g, grid, threads, smem, st, pdl) : go2<TO, KNN, NC, 512>(g, grid, threads, smem, st, pdl);
}
template <typename TO, bool KNN>
int by_n(const flock::GArgs &g, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    if (g.N == 100) return go<TO, KNN, 100>(g, grid, threads, smem, st, pdl);
    if (g.N == 50) return go<TO, KNN, 50>(g, grid, threads, smem, st, pdl);
    return go<TO, KNN, 0>(g, grid, threads, smem, st, pdl);
}
}  // namespace

int launch_group(const flock::GArgs &g, int tbytes, bool knn, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
    if (tbytes == 4) return knn ? by_n<float, true>(g, grid, threads, smem, st, pdl) : by_n<float, false>(g, grid, threads, smem, st, pdl);
    return knn ? by_n<double, true>(g, grid, threads, smem, st, pdl) : by_n<double, false>(g, grid, threads, smem, st, pdl);
}
