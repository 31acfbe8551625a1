// launch_swarm.cu -- the large-swarm kernels (prep + streaming row kernel + kNN fix-up) and the batched reset sampler.
#include <cstdio>
#include "flock_host.hpp"
#include "flock_swarm_kernel.cuh"
#include "flock_reset_kernel.cuh"

namespace {
template <typename TO, int LANES>
int swarm_go(const flock::SArgs &a, bool knn, dim3 grid, size_t smem, cudaStream_t st, bool pdl) {
    auto k0 = flock::flock_swarm_kernel<TO, LANES, false>;
    auto k1 = flock::flock_swarm_kernel<TO, LANES, true>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        configured_dev = dev;
    }
    const dim3 block(flock::SW_THREADS + 32 * a.fill_warps);
    return knn ? launch_pdl(k1, grid, block, smem, st, pdl, a) : launch_pdl(k0, grid, block, smem, st, pdl, a);
}
template <typename TO>
int swarm_all(const flock::SArgs &a, int lanes, bool knn, int PB, dim3 grid, size_t smem, cudaStream_t st, bool pdl, bool timing) {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (timing) { for (auto &e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], st); }
    int rc = launch_pdl(flock::flock_prep_kernel, dim3(PB, a.E), dim3(flock::PREP_THREADS), 0, st, pdl, a);
    if (rc) return rc;
    if (timing) cudaEventRecord(ev[1], st);
    switch (lanes) {
        case 8: rc = swarm_go<TO, 8>(a, knn, grid, smem, st, pdl); break;
        case 16: rc = swarm_go<TO, 16>(a, knn, grid, smem, st, pdl); break;
        default: rc = swarm_go<TO, 32>(a, knn, grid, smem, st, pdl); break;
    }
    if (timing) cudaEventRecord(ev[2], st);
    if (rc) return rc;
    // rows whose kNN search found fewer than 7 agents inside the inflated radius: one CTA per row rescans everything
    if (knn) rc = launch_pdl(flock::flock_knn_fix_kernel<TO>, dim3(flock::KFIX_SPLIT * flock::KFIX_SLOTS, a.E), dim3(flock::KFIX_THREADS), 0, st, pdl, a);
    if (timing) {                      // diagnostics (tools/shard_time.py): device time of each kernel of the step
        cudaEventRecord(ev[3], st);
        cudaEventSynchronize(ev[3]);
        float t01 = 0, t12 = 0, t23 = 0;
        cudaEventElapsedTime(&t01, ev[0], ev[1]); cudaEventElapsedTime(&t12, ev[1], ev[2]); cudaEventElapsedTime(&t23, ev[2], ev[3]);
        unsigned queued = 0;
        cudaMemcpy(&queued, a.needy_cnt, sizeof(queued), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[flock swarm step] prep %.1f us  row kernel %.1f us  knn fix-up %.1f us (%u rows queued)\n", t01 * 1e3,
                t12 * 1e3, t23 * 1e3, queued);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return rc;
}
}  // namespace

int launch_swarm(const flock::SArgs &a, int tbytes, int lanes, bool knn, int PB, dim3 grid, size_t smem, cudaStream_t st, bool pdl,
                 bool timing) {
    return tbytes == 4 ? swarm_all<float>(a, lanes, knn, PB, grid, smem, st, pdl, timing)
                       : swarm_all<double>(a, lanes, knn, PB, grid, smem, st, pdl, timing);
}

int launch_reset(const flock::RArgs &a, size_t smem, size_t smem_cap, cudaStream_t st) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        CUDA_TRY(cudaFuncSetAttribute(flock::flock_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured_dev = dev;
    }
    flock::flock_reset_kernel<<<a.E, flock::RESET_THREADS, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return FLOCK_OK;
}
