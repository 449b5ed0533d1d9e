// Debug harness: per-CTA role timeline of the persistent tensor-core conv (conv_tcp.cuh built with -DLC_TC_TIMING): globaltimer stamps of the transform,
// issuer and epilogue roles per tile, for the first, a middle and the last CTA, from one launch on cold inputs.
#define LC_TC_TIMING 1
#include "conv_tcp.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>
using namespace lc;
template <int C, int W>
void run() {
    const int B = 128;
    using K = tc::ConvTcpCfg<C, W>;
    size_t n = (size_t)B * W * W * C;
    const int NB = 12;
    float *x[NB], *y[NB], *w; unsigned long long* tm; int* err;
    for (int i = 0; i < NB; ++i) { cudaMalloc(&x[i], n * 4); cudaMalloc(&y[i], n * 4); cudaMemset(x[i], 0, n * 4); }
    cudaMalloc(&w, 9 * C * C * 4); cudaMalloc(&err, 4); cudaMemset(w, 0, 9 * C * C * 4); cudaMemset(err, 0, 4);
    const int grid = tc::conv_tcp_grid(B, C, W, 148);
    cudaMalloc(&tm, (size_t)grid * 64 * 8); cudaMemset(tm, 0, (size_t)grid * 64 * 8);
    for (int it = 0; it < NB; ++it) {       // rotating buffers: the last launch reads cold HBM
        tc::ConvTcArgs a{}; a.in = x[it]; a.wtc = w; a.out = y[it]; a.B = B; a.error_flag = err; a.timing = tm;
        tc::conv_tcp_launch<C, W, 0>(a, 148, 0); cudaDeviceSynchronize();
    }
    std::vector<unsigned long long> h((size_t)grid * 64);
    cudaMemcpy(h.data(), tm, h.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int b = 0; b < grid; ++b) { t0 = std::min(t0, h[b * 64]); t1 = std::max(t1, h[b * 64 + 60]); }
    printf("C=%d W=%d grid %d  span %.2f us (first CTA start -> last CTA end)\n", C, W, grid, (t1 - t0) / 1e3);
    const int show[3] = {0, grid / 2, grid - 1};
    for (int si = 0; si < 3; ++si) {
        const unsigned long long* q = h.data() + (size_t)show[si] * 64;
        auto T = [&](int s) { return q[s] ? (double)(q[s] - t0) / 1e3 : -1.0; };
        printf(" CTA %d: start %.2f  synced %.2f  affine ready %.2f  end %.2f\n", show[si], T(0), T(1), T(2), T(60));
        for (int k = 0; k <= K::TMAX; ++k)
            printf("   k=%d  chunk arrived %6.2f  staged %6.2f | mma issued %6.2f | acc seen %6.2f  stored %6.2f\n", k, T(10 + k), T(20 + k), T(30 + k), T(40 + k), T(50 + k));
    }
    {   // event-timed average over the rotating buffers (cold HBM: 12 x 2 x |X| > L2 for C = 16; back-to-back launches, no host sync in between)
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int reps = 240;
        cudaEventRecord(e0);
        for (int it = 0; it < reps; ++it) {
            tc::ConvTcArgs a{}; a.in = x[it % NB]; a.wtc = w; a.out = y[it % NB]; a.B = B; a.error_flag = err; a.timing = tm;
            tc::conv_tcp_launch<C, W, 0>(a, 148, 0);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(" average over %d back-to-back launches: %.2f us\n", reps, ms * 1e3 / reps);
    }
    int e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost); printf("err %d %s\n", e, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<16, 32>(); run<32, 16>(); return 0; }
