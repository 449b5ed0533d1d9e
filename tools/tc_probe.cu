// Probe: one tcgen05.mma (kind::tf32, K=8) with host-chosen descriptor parameters; prints max error vs the intended A*B^T.
#include "../libcontinual_b200/csrc/conv_tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace lc; using namespace lc::tc;

struct ProbeArgs { const float* a_img; const float* b_img; float* d_out; int a_bytes, b_bytes; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, idesc; int ncols; };

__global__ void probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* sA = sm; unsigned char* sB = sm + 65536;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 131072); uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < p.a_bytes / 4; i += 128) reinterpret_cast<float*>(sA)[i] = p.a_img[i];
    for (int i = tid; i < p.b_bytes / 4; i += 128) reinterpret_cast<float*>(sB)[i] = p.b_img[i];
    if (tid == 32) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(slot, 64);
    fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = *slot;
    if (tid == 0) {
        mma_tf32(tb, make_desc(smem_u32(sA), p.a_lbo, p.a_sbo), make_desc(smem_u32(sB), p.b_lbo, p.b_sbo), p.idesc, 0);
        mma_commit(bar);
    }
    mbar_wait(bar, 0); fence_after_sync();
    for (int c0 = 0; c0 < p.ncols; c0 += 16) {
        float v[16];
        tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) p.d_out[(size_t)tid * p.ncols + c0 + i] = v[i];
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

// logical matrices: A[M][8], B[N][8]
#include <cuda_bf16.h>
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__global__ void probe16_kernel(ProbeArgs p) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* sA = sm; unsigned char* sB = sm + 65536;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 131072); uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < p.a_bytes / 4; i += 128) reinterpret_cast<float*>(sA)[i] = p.a_img[i];
    for (int i = tid; i < p.b_bytes / 4; i += 128) reinterpret_cast<float*>(sB)[i] = p.b_img[i];
    if (tid == 32) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(slot, 64);
    fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = *slot;
    if (tid == 0) {
        mma_f16(tb, make_desc(smem_u32(sA), p.a_lbo, p.a_sbo), make_desc(smem_u32(sB), p.b_lbo, p.b_sbo), p.idesc, 0);
        mma_commit(bar);
    }
    mbar_wait(bar, 0); fence_after_sync();
    for (int c0 = 0; c0 < p.ncols; c0 += 16) {
        float v[16];
        tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) p.d_out[(size_t)tid * p.ncols + c0 + i] = v[i];
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}
// bf16, K = 16.  MN-major planar: block b (8 m's = 16 B) plane, K-row k at plane + 16 B * k; 8 K-rows per core matrix, LBO steps 8 K-rows
static void run16(const char* name, int M, int N, bool mn_major, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB, int planeA_rows, int planeB_rows) {
    std::vector<float> A(M * 16), Bm(N * 16);
    for (auto& v : A) v = (float)((rand() % 17) - 8);
    for (auto& v : Bm) v = (float)((rand() % 13) - 6);
    std::vector<__nv_bfloat16> aimg(32768, __float2bfloat16(0.f)), bimg(32768, __float2bfloat16(0.f));
    if (!mn_major) {
        for (int r = 0; r < M; ++r) for (int k = 0; k < 16; ++k) aimg[(k / 8) * planeA_rows * 8 + r * 8 + (k % 8)] = __float2bfloat16(A[r * 16 + k]);
        for (int r = 0; r < N; ++r) for (int k = 0; k < 16; ++k) bimg[(k / 8) * planeB_rows * 8 + r * 8 + (k % 8)] = __float2bfloat16(Bm[r * 16 + k]);
    } else {
        for (int r = 0; r < M; ++r) for (int k = 0; k < 16; ++k) aimg[(r / 8) * planeA_rows * 8 + k * 8 + (r % 8)] = __float2bfloat16(A[r * 16 + k]);
        for (int r = 0; r < N; ++r) for (int k = 0; k < 16; ++k) bimg[(r / 8) * planeB_rows * 8 + k * 8 + (r % 8)] = __float2bfloat16(Bm[r * 16 + k]);
    }
    float *da, *db, *dd; cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, aimg.data(), 65536, cudaMemcpyHostToDevice); cudaMemcpy(db, bimg.data(), 65536, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * 64 * 4);
    ProbeArgs p{da, db, dd, 65536, 65536, lboA, sboA, lboB, sboB, 0, N};
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    cudaFuncSetAttribute(probe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    probe16_kernel<<<1, 128, 131072 + 64>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * N); cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int m = 0; m < M; ++m) {
        const int lane = (M == 64) ? (m % 16) + 32 * (m / 16) : m;
        for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < 16; ++k) ref += (double)A[m * 16 + k] * Bm[n * 16 + k];
            maxerr = fmax(maxerr, fabs(ref - D[lane * N + n])); maxref = fmax(maxref, fabs(ref)); nz += D[lane * N + n] != 0.f;
        }
    }
    printf("%-46s M=%3d N=%2d  cuda=%s  max|err|=%g  (ref max %g, nonzero outputs %d/%d)\n", name, M, N, cudaGetErrorString(e), maxerr, maxref, nz, M * N);
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

static void run(const char* name, int M, int N, bool mn_major, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB, int planeA_rows, int planeB_rows) {
    std::vector<float> A(M * 8), Bm(N * 8);
    for (auto& v : A) v = (float)((rand() % 17) - 8);
    for (auto& v : Bm) v = (float)((rand() % 13) - 6);
    std::vector<float> aimg(16384, 0.f), bimg(16384, 0.f);
    if (!mn_major) {
        // K-major planar: chunk j (4 k's) plane, row r at plane + 16 B * r :  LBO = plane bytes, SBO = 128
        for (int r = 0; r < M; ++r) for (int k = 0; k < 8; ++k) aimg[(k / 4) * planeA_rows * 4 + r * 4 + (k % 4)] = A[r * 8 + k];
        for (int r = 0; r < N; ++r) for (int k = 0; k < 8; ++k) bimg[(k / 4) * planeB_rows * 4 + r * 4 + (k % 4)] = Bm[r * 8 + k];
    } else {
        // MN-major planar: block b (4 m's) plane, K-row k at plane + 16 B * k
        for (int r = 0; r < M; ++r) for (int k = 0; k < 8; ++k) aimg[(r / 4) * planeA_rows * 4 + k * 4 + (r % 4)] = A[r * 8 + k];
        for (int r = 0; r < N; ++r) for (int k = 0; k < 8; ++k) bimg[(r / 4) * planeB_rows * 4 + k * 4 + (r % 4)] = Bm[r * 8 + k];
    }
    float *da, *db, *dd; cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, aimg.data(), 65536, cudaMemcpyHostToDevice); cudaMemcpy(db, bimg.data(), 65536, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * 64 * 4);
    ProbeArgs p{da, db, dd, 65536, 65536, lboA, sboA, lboB, sboB, 0, N};
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    probe_kernel<<<1, 128, 131072 + 64>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * N); cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int m = 0; m < M; ++m) {
        const int lane = (M == 64) ? (m % 16) + 32 * (m / 16) : m;
        for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < 8; ++k) ref += (double)A[m * 8 + k] * Bm[n * 8 + k];
            maxerr = fmax(maxerr, fabs(ref - D[lane * N + n])); maxref = fmax(maxref, fabs(ref)); nz += D[lane * N + n] != 0.f;
        }
    }
    printf("%-46s M=%3d N=%2d  cuda=%s  max|err|=%g  (ref max %g, nonzero outputs %d/%d)\n", name, M, N, cudaGetErrorString(e), maxerr, maxref, nz, M * N);
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

int main() {
    // plane sizes in rows (of 16 B): A plane 200 rows, B plane 64 rows
    run("K-major  (LBO=plane, SBO=128)", 128, 16, false, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run("MN-major (LBO=128, SBO=plane)", 128, 16, true, 128, 200 * 16, 128, 64 * 16, 200, 64);
    run("MN-major (LBO=plane, SBO=128)", 128, 16, true, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run("MN-major M=64 (LBO=128, SBO=plane)", 64, 16, true, 128, 200 * 16, 128, 64 * 16, 200, 64);
    run("MN-major M=64 (LBO=plane, SBO=128)", 64, 16, true, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run("MN-major M=128 N=64 (LBO=128, SBO=plane)", 128, 64, true, 128, 200 * 16, 128, 64 * 16, 200, 64);
    run("MN-major M=128 N=64 (LBO=plane, SBO=128)", 128, 64, true, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run16("bf16 K-major  (LBO=plane, SBO=128)", 128, 16, false, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run16("bf16 MN-major (LBO=128, SBO=plane)", 128, 16, true, 128, 200 * 16, 128, 64 * 16, 200, 64);
    run16("bf16 MN-major (LBO=plane, SBO=128)", 128, 16, true, 200 * 16, 128, 64 * 16, 128, 200, 64);
    run16("bf16 MN-major M=64 (LBO=128, SBO=plane)", 64, 16, true, 128, 200 * 16, 128, 64 * 16, 200, 64);
    run16("bf16 MN-major M=64 (LBO=plane, SBO=128)", 64, 16, true, 200 * 16, 128, 64 * 16, 128, 200, 64);
    return 0;
}
