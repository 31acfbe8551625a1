// nvtx_ranges.cu -- NVTX v3 (header-only: a no-op unless a profiler injects itself) behind two plain functions.
// Its own translation unit: the NVTX header includes POSIX <sys/file.h>, whose flock() collides with namespace flock.
#include <nvtx3/nvToolsExt.h>

void flock_nvtx_push(const char *name) { nvtxRangePushA(name); }
void flock_nvtx_pop() { nvtxRangePop(); }
