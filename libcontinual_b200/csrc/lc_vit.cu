// C entry points of the ViT-B/16 building blocks: tcgen05 GEMM (gemm_tc.cuh) and the row-wise kernels around it.
#include "../../include/lc_b200.h"
#include "gemm_tc.cuh"

#include <cstdio>

using namespace lc;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rank-3 bf16 tensor map {K (contiguous), rows, batch}, box {64, box_rows, 1}, 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap* m, const void* ptr, int K, int rows, int batch, long long ld, long long batch_stride, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return LC_ERR_CUDA;
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? batch_stride : (long long)rows * ld) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((gstr[0] % 16) != 0 || (gstr[1] % 16) != 0 || ((uintptr_t)ptr % 16) != 0) return LC_ERR_INVALID;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? LC_OK : LC_ERR_INVALID;
}

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmArgs& a, int batch, cudaStream_t st) {
    using K = tc::GemmCfg<BN>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc::gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    dim3 grid((a.N + BN - 1) / BN, (a.M + 127) / 128, batch);
    tc::gemm_bf16_kernel<BN><<<grid, K::NT, K::SMEM_BYTES, st>>>(ta, tb, a);
    return lc_launch_status();
}

}  // namespace

extern "C" {

int lc_gemm_bf16(const void* A, int lda, long long strideA, const void* B, int ldb, long long strideB, void* C, int ldc, long long strideC, int M, int N,
                 int K, int batch, const float* bias, const float* residual, int ldr, long long strideR, void* out2, int out_f32, float alpha,
                 int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(A && B && C && M >= 1 && N >= 1 && K >= 8 && K % 8 == 0 && batch >= 1 && lda >= K && ldb >= K && ldc >= N);
    // vectorised epilogue stores: rows of C (and of the residual) must start on 16-byte boundaries
    LC_CHECK_ARG(ldc % (out_f32 ? 4 : 8) == 0 && strideC % (out_f32 ? 4 : 8) == 0 && (residual == nullptr || (ldr % 4 == 0 && strideR % 4 == 0)));
    CUtensorMap ta, tb;
    const int bn = (N % 256 == 0 || N > 128) ? 256 : 128;
    int e = make_tmap(&ta, A, K, M, batch, lda, strideA, 128);
    if (e != LC_OK) return e;
    e = make_tmap(&tb, B, K, N, batch, ldb, strideB, bn);
    if (e != LC_OK) return e;
    tc::GemmArgs a{};
    a.out = C; a.bias = bias; a.residual = residual; a.out2 = out2; a.M = M; a.N = N; a.K = K; a.ldc = ldc; a.ldr = ldr;
    a.batch_stride_c = strideC; a.batch_stride_r = strideR; a.out_dtype = out_f32 ? tc::GEMM_OUT_F32 : tc::GEMM_OUT_BF16; a.alpha = alpha;
    a.error_flag = error_flag;
    return bn == 256 ? launch_gemm<256>(ta, tb, a, batch, (cudaStream_t)stream) : launch_gemm<128>(ta, tb, a, batch, (cudaStream_t)stream);
}

}  // extern "C"
