// Debug harness: CUDA-event time per launch of the tensor-core weight gradient (wgrad3x3_tc_kernel) in each staging configuration, 200 back-to-back
// launches rotating over NBUF buffer sets (cold HBM).  Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc dir> tools/wgrad_variants.cu -o tools/_timing/wgrad_variants
// -I selects the header tree (the working tree's libcontinual_b200/csrc, or a snapshot of an older commit for an A/B run).
#include "wgrad_tc.cuh"
#include <cstdio>
#include <vector>
using namespace lc;

template <int C, int W>
void run(int B, int nsplit) {
    constexpr int NBUF = 8;
    const size_t n = (size_t)B * W * W * C;
    float *x[NBUF], *dy[NBUF], *y[NBUF], *part, *aff, *coef, *lpart, *dg;
    int* err;
    for (int i = 0; i < NBUF; ++i) {
        cudaMalloc(&x[i], n * 4); cudaMalloc(&dy[i], n * 4); cudaMalloc(&y[i], n * 4);
        cudaMemset(x[i], 0, n * 4); cudaMemset(dy[i], 0, n * 4); cudaMemset(y[i], 0, n * 4);
    }
    cudaMalloc(&part, (size_t)nsplit * 9 * C * C * 4);
    cudaMalloc(&aff, 4 * C * 4); cudaMemset(aff, 0, 4 * C * 4);
    cudaMalloc(&coef, 3 * C * 4); cudaMemset(coef, 0, 3 * C * 4);
    cudaMalloc(&dg, 2 * C * 4);
    const int nparts = 296;
    cudaMalloc(&lpart, (size_t)nparts * 2 * C * 4); cudaMemset(lpart, 0, (size_t)nparts * 2 * C * 4);
    cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"plain", "+ BN/ReLU input prologue", "+ dY apply (coef array)", "+ prologue + dY apply (lazy coef)"};
    for (int v = 0; v < 4; ++v) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            for (int it = 0; it < 200; ++it) {
                const int i = it % NBUF;
                tc::WgradTcArgs a{};
                a.in = x[i]; a.dy = dy[i]; a.partial = part; a.B = B; a.error_flag = err;
                if (v == 1 || v == 3) { a.pro_scale = aff; a.pro_shift = aff + C; }
                if (v >= 2) { a.dy_y = y[i]; a.dy_coef = coef; }
                if (v == 3) {
                    a.dy_blazy.partial = lpart; a.dy_blazy.scale = aff; a.dy_blazy.mean = aff + 2 * C; a.dy_blazy.invstd = aff + 3 * C;
                    a.dy_blazy.dgamma = dg; a.dy_blazy.dbeta = dg + C; a.dy_blazy.nparts = nparts; a.dy_blazy.count = (float)B * W * W; a.dy_blazy.write_grads = 1;
                }
                tc::wgrad_tc_launch<C, W>(a, nsplit, st);
            }
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("wgrad C=%d W=%d nsplit=%d  %-36s %7.2f us / launch\n", C, W, nsplit, names[v], best * 1000.f / 200.f);
    }
    int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    printf("err %d  last cuda error: %s\n", herr, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const int B = 128;
    run<16, 32>(B, 296);
    run<16, 32>(B, 148);
    run<32, 16>(B, 148);
    run<64, 8>(B, 50);
    run<64, 8>(B, 100);
    return 0;
}
