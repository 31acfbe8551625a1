// ubench_store.cu -- how fast can 148 SMs WRITE a dense [rows x 100] fp32 matrix (the C2 adjacency: 164 MB per step)?
//   direct : STG.128 streaming stores, a warp writes 32 consecutive rows (25 of 32 lanes active per row, as the kernel does)
//   flat   : the same bytes as one flat array, all 32 lanes active
//   bulk   : 1-D TMA bulk stores shared -> global (cp.async.bulk.global.shared::cta), one per CHUNK bytes, per warp
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_store.bin tools/ubench_store.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 100;

__global__ void __launch_bounds__(256) k_direct(float *out, int rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {
        float *dst = out + (size_t)r0 * N + lane * 4;
        if (lane < 25) {
#pragma unroll
            for (int r = 0; r < 32; ++r) __stcs(reinterpret_cast<float4 *>(dst + r * N), make_float4(0.f, 1.f, 0.f, 0.f));
        }
    }
}

__global__ void __launch_bounds__(256) k_flat(float4 *out, size_t nvec) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    // each warp walks 12.8 KB runs like k_direct does
    const size_t warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (size_t b = warp * 800; b < nvec; b += nwarps * 800) {
#pragma unroll
        for (int q = 0; q < 25; ++q) __stcs(out + b + q * 32 + lane, make_float4(0.f, 1.f, 0.f, 0.f));
    }
}

template <int CHUNK>
__global__ void __launch_bounds__(256) k_bulk(float *out, int rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned char *buf = smem + (size_t)wib * CHUNK;
    for (int q = lane * 16; q < CHUNK; q += 32 * 16) *reinterpret_cast<float4 *>(buf + q) = make_float4(0.f, 1.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t total = (size_t)rows * N * 4;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(buf);
    for (size_t r0 = (size_t)warp * 12800; r0 < total; r0 += (size_t)nwarps * 12800) {
        if (lane == 0) {
            for (int c = 0; c < 12800; c += CHUNK) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char *>(out) + r0 + c),
                             "r"(saddr), "r"(CHUNK) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> void timeit(const char *name, size_t bytes, F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 50;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%-40s %8.2f us / 164 MB   %7.1f GB/s   %s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const int rows = 4096 * 100;
    const size_t bytes = (size_t)rows * N * 4;
    float *out; cudaMalloc(&out, bytes);
    for (int ctas : {148 * 2, 148 * 4, 148 * 8, 1600}) {
        char nm[96];
        snprintf(nm, sizeof nm, "direct STG.128 rows (%d CTAs)", ctas);
        timeit(nm, bytes, [&] { k_direct<<<ctas, 256>>>(out, rows); });
        snprintf(nm, sizeof nm, "flat STG.128 (%d CTAs)", ctas);
        timeit(nm, bytes, [&] { k_flat<<<ctas, 256>>>(reinterpret_cast<float4 *>(out), bytes / 16); });
    }
    for (int ctas : {148 * 2, 148 * 4}) {
        char nm[96];
        cudaFuncSetAttribute(k_bulk<400>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12800);
        cudaFuncSetAttribute(k_bulk<3200>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12800);
        cudaFuncSetAttribute(k_bulk<12800>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 12800);
        snprintf(nm, sizeof nm, "TMA bulk store 400 B (%d CTAs)", ctas);
        timeit(nm, bytes, [&] { k_bulk<400><<<ctas, 256, 8 * 400>>>(out, rows); });
        snprintf(nm, sizeof nm, "TMA bulk store 3200 B (%d CTAs)", ctas);
        timeit(nm, bytes, [&] { k_bulk<3200><<<ctas, 256, 8 * 3200>>>(out, rows); });
        if (ctas <= 148 * 2) {
            snprintf(nm, sizeof nm, "TMA bulk store 12800 B (%d CTAs)", ctas);
            timeit(nm, bytes, [&] { k_bulk<12800><<<ctas, 256, 8 * 12800>>>(out, rows); });
        }
    }
    cudaFree(out);
    return 0;
}
