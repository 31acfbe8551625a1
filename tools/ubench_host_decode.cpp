// Host-side microbenchmark (no GPU): expanding adjacency bits + 1/deg into a dense [E,N,N] fp32 array with T threads.
// Result on the B200 box: profiles/r2_ubench_host_decode.txt.   g++ -O3 -march=native -pthread; usage: ./a.out [threads]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <thread>
#include <chrono>
#include <immintrin.h>
static void decode_rows(float *out, const uint32_t *bits, const float *idg, size_t r0, size_t r1, int N, int WB) {
    for (size_t r = r0; r < r1; ++r) {
        float *o = out + r * N;
        const uint32_t *b = bits + r * WB;
        const float v = idg[r];
        memset(o, 0, sizeof(float) * N);
        for (int w = 0; w < WB; ++w) {
            uint32_t m = b[w];
            while (m) { int k = __builtin_ctz(m); m &= m - 1; o[w * 32 + k] = v; }
        }
    }
}
int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : 16;
    const int E = 4096, N = 100, WB = 4;
    size_t rows = (size_t)E * N;
    std::vector<uint32_t> bits(rows * WB, 0);
    std::vector<float> idg(rows, 0.2f);
    for (size_t r = 0; r < rows; ++r) for (int k = 0; k < 6; ++k) { int j = rand() % N; bits[r * WB + j / 32] |= 1u << (j % 32); }
    float *out = (float *)aligned_alloc(4096, rows * N * sizeof(float));
    memset(out, 1, rows * N * sizeof(float));
    for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(decode_rows, out, bits.data(), idg.data(), rows * t / T, rows * (t + 1) / T, N, WB);
        for (auto &x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("T=%d  %.3f ms  %.1f GB/s\n", T, s * 1e3, rows * N * 4 / s / 1e9);
    }
    return 0;
}
