// Debug harness: per-CTA phase timestamps of the tensor-core conv (build with -DLC_TC_TIMING).
#define LC_TC_TIMING 1
#include "../libcontinual_b200/csrc/conv_tc.cuh"
#include <cstdio>
#include <vector>
#include <algorithm>
using namespace lc;
int main(int argc, char** argv) {
    const int B = 128; constexpr int C = 16, W = 32;
    using K = tc::ConvTcCfg<C, W>;
    size_t n = (size_t)B * W * W * C;
    float *x, *y, *w; unsigned long long* tm; int* err;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&w, 9 * C * C * 4); cudaMalloc(&err, 4);
    cudaMemset(x, 0, n * 4); cudaMemset(w, 0, 9 * C * C * 4); cudaMemset(err, 0, 4);
    int grid = (int)(((long long)B * K::PP + K::MROWS - 1) / K::MROWS);
    cudaMalloc(&tm, (size_t)grid * 8 * 8); cudaMemset(tm, 0, (size_t)grid * 64);
    tc::ConvTcArgs a{}; a.in = x; a.wtc = w; a.out = y; a.B = B; a.error_flag = err; a.timing = tm;
    for (int it = 0; it < 3; ++it) { tc::conv_tc_launch<C, W>(a, 0); cudaDeviceSynchronize(); }
    std::vector<unsigned long long> h((size_t)grid * 8);
    cudaMemcpy(h.data(), tm, h.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int b = 0; b < grid; ++b) { t0 = std::min(t0, h[b * 8]); t1 = std::max(t1, h[b * 8 + 6]); }
    printf("grid %d  kernel span %.2f us (first CTA start -> last CTA end)\n", grid, (t1 - t0) / 1e3);
    const char* names[] = {"rowtab+issue cp.async", "wait loads", "transform+sync", "mma issue", "mma wait", "epilogue"};
    for (int ph = 0; ph < 6; ++ph) {
        std::vector<double> d;
        for (int b = 0; b < grid; ++b) d.push_back((double)(h[b * 8 + ph + 1] - h[b * 8 + ph]) / 1e3);
        std::sort(d.begin(), d.end());
        printf("  %-16s median %.2f  p90 %.2f  max %.2f us\n", names[ph], d[d.size() / 2], d[d.size() * 9 / 10], d.back());
    }
    { std::vector<double> d; for (int b = 0; b < grid; ++b) d.push_back((double)(h[b * 8 + 7] - h[b * 8]) / 1e3); std::sort(d.begin(), d.end());
      printf("  (setup+rowtab+sync within phase 0: median %.2f max %.2f us)\n", d[d.size() / 2], d.back()); }
    std::vector<double> st, en;
    for (int b = 0; b < grid; ++b) { st.push_back((h[b * 8] - t0) / 1e3); en.push_back((h[b * 8 + 6] - t0) / 1e3); }
    std::sort(st.begin(), st.end()); std::sort(en.begin(), en.end());
    printf("  CTA start: median %.2f p90 %.2f max %.2f us ; CTA end: p10 %.2f median %.2f max %.2f us\n", st[grid / 2], st[grid * 9 / 10], st.back(), en[grid / 10], en[grid / 2], en.back());
    int e; cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost); printf("err %d\n", e);
    return 0;
}
