// launch_graph.cu -- graph-signal aggregation on the adjacency bits and the shepherding repulsion.
#include "flock_host.hpp"
#include "flock_graph_kernels.cuh"

int launch_aggregate(const flock::AggArgs &a, int tbytes, cudaStream_t st) {
    if (tbytes == 4) return launch_pdl(flock::flock_aggregate_kernel<float>, dim3(a.E), dim3(256), 0, st, true, a);
    return launch_pdl(flock::flock_aggregate_kernel<double>, dim3(a.E), dim3(256), 0, st, true, a);
}

int launch_shepherd(const flock::ShepArgs &a, int tbytes, cudaStream_t st) {
    const dim3 grid((a.N + 127) / 128, a.E);
    if (tbytes == 4) flock::flock_repulsion_kernel<float><<<grid, 128, 0, st>>>(a);
    else flock::flock_repulsion_kernel<double><<<grid, 128, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return FLOCK_OK;
}
