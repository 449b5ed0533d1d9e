// Debug harness: CUDA-event time per launch of the tensor-core conv in each prologue / epilogue configuration, 200 back-to-back launches
// (programmatic dependent launch, as in the step) rotating over NBUF buffer sets.  Build: see tools/build_variants.sh.
#include "conv_tcp.cuh"
#include <cstdio>
#include <vector>
using namespace lc;

template <int C, int W, bool PERSIST = false>
void run(int B) {
    using K = tc::ConvTcCfg<C, W>;
    constexpr int NBUF = 6;
    const size_t n = (size_t)B * W * W * C;
    float *x[NBUF], *y[NBUF], *o[NBUF], *w, *aff, *coef, *part, *gamma, *beta, *dg;
    unsigned int* counter; int* err;
    for (int i = 0; i < NBUF; ++i) { cudaMalloc(&x[i], n * 4); cudaMalloc(&y[i], n * 4); cudaMalloc(&o[i], n * 4); cudaMemset(x[i], 0, n * 4); cudaMemset(y[i], 0, n * 4); cudaMemset(o[i], 0, n * 4); }
    cudaMalloc(&w, 9 * C * C * 4); cudaMemset(w, 0, 9 * C * C * 4);
    cudaMalloc(&aff, 4 * C * 4); cudaMemset(aff, 0, 4 * C * 4);
    cudaMalloc(&coef, 3 * C * 4); cudaMemset(coef, 0, 3 * C * 4);
    cudaMalloc(&gamma, C * 4); cudaMalloc(&beta, C * 4); cudaMalloc(&dg, 2 * C * 4); cudaMemset(gamma, 0, C * 4); cudaMemset(beta, 0, C * 4);
    const int grid = PERSIST ? tc::conv_tcp_grid(B, C, W, 148) : (int)(((long long)B * K::PP + K::MROWS - 1) / K::MROWS);
    cudaMalloc(&part, (size_t)grid * 2 * C * 4 * 2); cudaMemset(part, 0, (size_t)grid * 2 * C * 4 * 2);
    cudaMalloc(&counter, 64); cudaMemset(counter, 0, 64); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto stat = [&](tc::ConvTcArgs& a, int defer) {
        a.stat.partial = part; a.stat.counter = counter; a.stat.gamma = gamma; a.stat.beta = beta; a.stat.scale = aff; a.stat.shift = aff + C;
        a.stat.mean = aff + 2 * C; a.stat.invstd = aff + 3 * C; a.stat.momentum = 0.1f; a.stat.eps = 1e-5f; a.stat.defer = defer;
    };
    auto lazy = [&](tc::ConvTcArgs& a) {
        a.pro_lazy.partial = part + (size_t)grid * 2 * C; a.pro_lazy.gamma = gamma; a.pro_lazy.beta = beta; a.pro_lazy.nparts = grid;
        a.pro_lazy.count = (float)B * W * W; a.pro_lazy.eps = 1e-5f;
    };
    auto bw = [&](tc::ConvTcArgs& a, int i, int reduce, int defer, int mask_out) {
        a.bw.y = y[(i + 1) % NBUF]; a.bw.mask_out = mask_out ? y[(i + 2) % NBUF] : nullptr; a.bw.scale = aff; a.bw.shift = aff + C; a.bw.mean = aff + 2 * C;
        a.bw.invstd = aff + 3 * C; a.bw.partial = reduce ? part : nullptr; a.bw.defer = defer; a.bw.counter = counter; a.bw.coef = coef; a.bw.dgamma = dg;
        a.bw.dbeta = dg + C;
    };
    auto blazy = [&](tc::ConvTcArgs& a) {
        a.pro_blazy.partial = part + (size_t)grid * 2 * C; a.pro_blazy.scale = aff; a.pro_blazy.mean = aff + 2 * C; a.pro_blazy.invstd = aff + 3 * C;
        a.pro_blazy.nparts = grid; a.pro_blazy.count = (float)B * W * W;
    };
    const char* names[] = {"fwd plain", "fwd + BN/ReLU prologue", "fwd + prologue + stats (last-CTA finalise)", "fwd + prologue + stats (deferred)",
                           "fwd + lazy prologue + stats (deferred)", "bwd: apply prologue (coef array)", "bwd: + mask(BN) + reduce (last-CTA)",
                           "bwd: + mask(BN) + reduce (deferred)", "bwd: lazy coef + addend + mask(out) + reduce (deferred)", "bwd: mask(BN) only, no reduce"};
    for (int v = 0; v < (PERSIST ? 5 : 10); ++v) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            for (int it = 0; it < 200; ++it) {
                const int i = it % NBUF;
                tc::ConvTcArgs a{};
                a.in = x[i]; a.wtc = w; a.out = o[i]; a.B = B; a.error_flag = err;
                if (v >= 1 && v <= 3) { a.pro_scale = aff; a.pro_shift = aff + C; }
                if (v == 2) stat(a, 0);
                if (v == 3 || v == 4) stat(a, 1);
                if (v == 4) lazy(a);
                if (v < 5) {
                    if constexpr (PERSIST) tc::conv_tcp_launch<C, W, 0>(a, 148, st); else tc::conv_tc_launch<C, W, 0>(a, st);
                    continue;
                }
                a.pro_y = y[i]; a.pro_coef = coef;
                if (v == 6) bw(a, i, 1, 0, 0);
                if (v == 7) bw(a, i, 1, 1, 0);
                if (v == 8) { bw(a, i, 1, 1, 1); blazy(a); a.addend = x[(i + 3) % NBUF]; }
                if (v == 9) bw(a, i, 0, 0, 0);
                tc::conv_tc_launch<C, W, 1>(a, st);
            }
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("C=%d W=%d  %-58s %7.2f us / launch\n", C, W, names[v], best * 1000.f / 200.f);
    }
    int e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost);
    printf("err %d  last cuda error: %s\n", e, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<16, 32>(128);
    printf("persistent kernel (conv_tcp.cuh):\n");
    run<16, 32, true>(128);
    run<32, 16>(128);
    printf("persistent kernel (conv_tcp.cuh):\n");
    run<32, 16, true>(128);
    run<64, 8>(128);
    return 0;
}
