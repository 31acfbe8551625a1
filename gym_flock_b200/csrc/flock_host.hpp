// flock_host.hpp -- host-side plumbing shared by the translation units of libflock_b200.so.
// The kernels are templates in headers; their instantiations are split over several .cu files (launch_*.cu) so that the
// library builds in parallel.  Nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/flock_b200.h"

namespace flock {
struct KArgs;
struct GArgs;
struct SArgs;
struct RArgs;
struct AggArgs;
struct ShepArgs;
}  // namespace flock

// records the message for flock_last_error() and returns `code`
int flock_fail(int code, const char *fmt, ...);

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return flock_fail(FLOCK_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_));  \
    } while (0)

// NVTX range over one entry point of the C ABI (shows up as a named span in nsys / ncu timelines)
// (nvtx_ranges.cu: the NVTX header pulls in POSIX flock(), which cannot share a translation unit with namespace flock)
void flock_nvtx_push(const char *name);
void flock_nvtx_pop();
struct NvtxRange {
    explicit NvtxRange(const char *name) { flock_nvtx_push(name); }
    ~NvtxRange() { flock_nvtx_pop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

constexpr int kSMs = 148;
constexpr size_t kMaxSmem = 227 * 1024;

// launch with (optional) programmatic dependent launch: the next kernel of the stream may be scheduled while this one
// drains; every kernel of the library executes griddepcontrol.wait before its first global access
template <typename Kern, typename Args>
inline int launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Args &args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args));
    return FLOCK_OK;
}

// ---- launchers (one translation unit each) ---------------------------------------------------------------------------------
// env-resident kernel: tbytes 4 / 8 = observation type; lanes / exact / knn select the instantiation
int launch_env_f32(const flock::KArgs &k, int lanes, bool knn, int grid, int threads, size_t smem, cudaStream_t st, bool pdl);
int launch_env_f64(const flock::KArgs &k, int lanes, bool exact, bool knn, int grid, int threads, size_t smem, cudaStream_t st, bool pdl);
// group kernel (whole envs per CTA, `steps` fused steps)
int launch_group(const flock::GArgs &g, int tbytes, bool knn, int grid, int threads, size_t smem, cudaStream_t st, bool pdl);
// large swarms: prep + row kernel + kNN fix-up
int launch_swarm(const flock::SArgs &a, int tbytes, int lanes, bool knn, int PB, dim3 grid, size_t smem, cudaStream_t st, bool pdl,
                 bool timing);
int launch_reset(const flock::RArgs &a, size_t smem, size_t smem_cap, cudaStream_t st);
// consumers of the adjacency bits
int launch_aggregate(const flock::AggArgs &a, int tbytes, cudaStream_t st);
int launch_shepherd(const flock::ShepArgs &a, int tbytes, cudaStream_t st);
