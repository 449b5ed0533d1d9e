// C entry points of the ViT-B/16 building blocks: tcgen05 GEMM (gemm_tc.cuh) and the row-wise kernels around it (vit_ops.cuh).
#include "../../include/lc_b200.h"
#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "vit_ops.cuh"
#include "lora_ops.cuh"
#include "prompt_ops.cuh"

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

// rank-4 bf16 tensor map {K (contiguous), rows, inner batch, outer batch}, box {64, box_rows, 1, 1}, 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap* m, const void* ptr, int K, int rows, int b_in, int b_out, long long ld, long long s_in, long long s_out, int box_rows,
              int* bcast) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return LC_ERR_CUDA;
    *bcast = 0;
    if (b_in > 1 && s_in == 0) { *bcast |= 1; b_in = 1; }       // operand shared across the batch: extent 1, the kernel passes coordinate 0
    if (b_out > 1 && s_out == 0) { *bcast |= 2; b_out = 1; }
    if (b_in <= 1) s_in = (long long)rows * ld;
    if (b_out <= 1) s_out = s_in * (b_in < 1 ? 1 : b_in);
    cuuint64_t gdim[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(b_in < 1 ? 1 : b_in), (cuuint64_t)(b_out < 1 ? 1 : b_out)};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)s_in * 2, (cuuint64_t)s_out * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if ((gstr[0] % 16) != 0 || (gstr[1] % 16) != 0 || (gstr[2] % 16) != 0 || ((uintptr_t)ptr % 16) != 0) return LC_ERR_INVALID;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? LC_OK : LC_ERR_INVALID;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int EPI>
int launch_gemm_epi(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmArgs& a, int batch, cudaStream_t st) {
    using K = tc::GemmCfg<BN, EPI>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc::gemm_bf16_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    const long long tiles = (long long)((a.N + BN - 1) / BN) * ((a.M + 127) / 128) * batch * (a.ksplit > 1 ? a.ksplit : 1);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    tc::gemm_bf16_kernel<BN, EPI><<<grid, K::NT, K::SMEM_BYTES, st>>>(ta, tb, a);
    return lc_launch_status();
}

// CTA-pair variant (gemm_tc.cuh CG = 2): 2-CTA clusters, 256 x 256 blocks, tb's box holds 128 rows
template <int EPI>
int launch_gemm_epi_cg2(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmArgs& a, int batch, cudaStream_t st) {
    using K = tc::GemmCfg<256, EPI, 2>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc::gemm_bf16_kernel<256, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    const long long blocks = (long long)((a.N + 255) / 256) * ((a.M + 255) / 256) * batch;
    const long long pairs = blocks < num_sms() / 2 ? blocks : num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(K::NT); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, tc::gemm_bf16_kernel<256, EPI, 2>, ta, tb, a) != cudaSuccess) return LC_ERR_CUDA;
    return lc_launch_status();
}
int launch_gemm_cg2(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmArgs& a, int batch, cudaStream_t st) {
    const int epi = (a.out_dtype == tc::GEMM_OUT_F32 ? tc::EPI_F32 : 0) | (a.residual != nullptr ? tc::EPI_RES : 0) | (a.out2 != nullptr ? tc::EPI_GELU2 : 0) |
                    (a.gelu_aux != nullptr ? tc::EPI_DGELU : 0);
    switch (epi) {
#define LC_EPI_CASE(E) case E: return launch_gemm_epi_cg2<E>(ta, tb, a, batch, st);
        LC_EPI_CASE(0) LC_EPI_CASE(1) LC_EPI_CASE(2) LC_EPI_CASE(3) LC_EPI_CASE(4) LC_EPI_CASE(5) LC_EPI_CASE(6) LC_EPI_CASE(7)
        LC_EPI_CASE(8) LC_EPI_CASE(9) LC_EPI_CASE(10) LC_EPI_CASE(11) LC_EPI_CASE(12) LC_EPI_CASE(13) LC_EPI_CASE(14) LC_EPI_CASE(15)
#undef LC_EPI_CASE
    }
    return LC_ERR_INVALID;
}
// 0: single-CTA kernel only; 1 (default): CTA pairs for the wide, tall GEMMs.  LC_GEMM_CG2=0 switches the pair variant off.
int gemm_cg2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LC_GEMM_CG2"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
    return v;
}

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmArgs& a, int batch, cudaStream_t st) {
    const int epi = (a.out_dtype == tc::GEMM_OUT_F32 ? tc::EPI_F32 : 0) | (a.residual != nullptr ? tc::EPI_RES : 0) | (a.out2 != nullptr ? tc::EPI_GELU2 : 0) |
                    (a.gelu_aux != nullptr ? tc::EPI_DGELU : 0);
    switch (epi) {
#define LC_EPI_CASE(E) case E: return launch_gemm_epi<BN, E>(ta, tb, a, batch, st);
        LC_EPI_CASE(0) LC_EPI_CASE(1) LC_EPI_CASE(2) LC_EPI_CASE(3) LC_EPI_CASE(4) LC_EPI_CASE(5) LC_EPI_CASE(6) LC_EPI_CASE(7)
        LC_EPI_CASE(8) LC_EPI_CASE(9) LC_EPI_CASE(10) LC_EPI_CASE(11) LC_EPI_CASE(12) LC_EPI_CASE(13) LC_EPI_CASE(14) LC_EPI_CASE(15)
#undef LC_EPI_CASE
    }
    return LC_ERR_INVALID;
}

int grid_for(long long n, int per_block) {
    long long b = (n + per_block - 1) / per_block;
    const long long cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

#ifdef LC_GEMM_TIMING
extern "C" int lc_debug_gemm_timing(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, tc::g_gemm_tstamp, sizeof(unsigned long long) * 148 * 32 * 8) == cudaSuccess ? 0 : -1;
}
#endif

#ifdef LC_ATTN_TIMING
extern "C" int lc_debug_attn_timing(unsigned long long* host_out, int clear) {
    if (clear) { static unsigned long long z[2048 * 16]; return cudaMemcpyToSymbol(tc::g_attn_tstamp, z, sizeof(z)) == cudaSuccess ? 0 : -1; }
    return cudaMemcpyFromSymbol(host_out, tc::g_attn_tstamp, sizeof(unsigned long long) * 2048 * 16) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" {

int lc_gemm_bf16_ex(const lc_gemm_desc* d, int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(d && d->A && d->B && d->C && d->M >= 1 && d->N >= 1 && d->K >= 8 && d->K % 8 == 0 && d->batch_in >= 1 && d->batch_out >= 1);
    LC_CHECK_ARG(d->lda >= d->K && d->ldb >= d->K && d->ldc >= d->N);
    const int cal = d->out_f32 ? 4 : 8;      // vectorised epilogue stores: rows of C / residual start on 16-byte boundaries
    LC_CHECK_ARG(d->ldc % cal == 0 && d->strideC_in % cal == 0 && d->strideC_out % cal == 0);
    LC_CHECK_ARG(d->residual == nullptr || (d->ldr % 4 == 0 && d->strideR_in % 4 == 0 && d->strideR_out % 4 == 0));
    LC_CHECK_ARG(d->gelu_mode >= 0 && d->gelu_mode <= 3 && ((d->gelu_mode & 2) == 0 || d->out2 != nullptr));
    CUtensorMap ta, tb;
    const int bn = d->N > 128 ? 256 : 128;
    // CTA pairs when a 256 x 256 block grid still fills the machine: wide panels, at least 74 block pairs' worth of rows
    const bool cg2 = gemm_cg2_enabled() && bn == 256 && d->ksplit <= 1 &&
                     (long long)((d->M + 255) / 256) * ((d->N + 255) / 256) * d->batch_in * d->batch_out >= 74;
    tc::GemmArgs a{};
    int e = make_tmap(&ta, d->A, d->K, d->M, d->batch_in, d->batch_out, d->lda, d->strideA_in, d->strideA_out, 128, &a.a_bcast);
    if (e != LC_OK) return e;
    e = make_tmap(&tb, d->B, d->K, d->N, d->batch_in, d->batch_out, d->ldb, d->strideB_in, d->strideB_out, cg2 ? 128 : bn, &a.b_bcast);
    if (e != LC_OK) return e;
    a.out = d->C; a.bias = d->bias; a.residual = d->residual; a.out2 = d->out2; a.gelu_aux = d->gelu_bwd_aux; a.M = d->M; a.N = d->N; a.K = d->K; a.ldc = (int)d->ldc; a.ldr = (int)d->ldr;
    a.c_stride_in = d->strideC_in; a.c_stride_out = d->strideC_out; a.r_stride_in = d->strideR_in; a.r_stride_out = d->strideR_out;
    a.batch_in = d->batch_in; a.batch_total = d->batch_in * d->batch_out; a.out_dtype = d->out_f32 ? tc::GEMM_OUT_F32 : tc::GEMM_OUT_BF16; a.alpha = d->alpha; a.error_flag = error_flag; a.gelu_mode = d->gelu_mode;
    if (d->ksplit > 1) {
        const int nkb = (d->K + 63) / 64, kper = (nkb + d->ksplit - 1) / d->ksplit;
        LC_CHECK_ARG(d->out_f32 && !d->bias && !d->residual && !d->out2 && !d->gelu_bwd_aux && (long long)(d->ksplit - 1) * kper < nkb && d->strideC_split % 4 == 0);
        a.ksplit = d->ksplit; a.c_stride_split = d->strideC_split;
    }
    const int batch = d->batch_in * d->batch_out;
    if (cg2) return launch_gemm_cg2(ta, tb, a, batch, (cudaStream_t)stream);
    return bn == 256 ? launch_gemm<256>(ta, tb, a, batch, (cudaStream_t)stream) : launch_gemm<128>(ta, tb, a, batch, (cudaStream_t)stream);
}

int lc_conv_gemm_bf16(const lc_conv_desc* d, int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(d && d->X && d->Wk && d->Y && d->N >= 1 && d->C >= 64 && d->C % 64 == 0 && d->Cout >= 1 && d->ks >= 1 && d->ks <= 7 && d->stride >= 1 && d->pad >= 0);
    LC_CHECK_ARG(d->Ho == (d->H + 2 * d->pad - d->ks) / d->stride + 1 && d->Wo == (d->W + 2 * d->pad - d->ks) / d->stride + 1 && d->Wo >= 1 && 128 % d->Wo == 0);
    const int hw = d->Ho * d->Wo;
    LC_CHECK_ARG(hw % 128 == 0 || 128 % hw == 0);
    LC_CHECK_ARG(d->ldc >= d->Cout && d->ldc % (d->out_f32 ? 4 : 8) == 0 && (d->residual == nullptr || d->ldr % 4 == 0) && ((uintptr_t)d->X % 16) == 0);
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return LC_ERR_CUDA;
    const int Ht = hw >= 128 ? 128 / d->Wo : d->Ho, Nt = hw >= 128 ? 1 : 128 / hw;
    // box extents are in tensor elements: Wt outputs at element stride s span (Wt - 1) * s + 1 inputs
    CUtensorMap ta, tb;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t gstr[3] = {(cuuint64_t)d->C * 2, (cuuint64_t)d->W * d->C * 2, (cuuint64_t)d->H * d->W * d->C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)((d->Wo - 1) * d->stride + 1), (cuuint32_t)((Ht - 1) * d->stride + 1), (cuuint32_t)Nt};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        LC_CHECK_ARG(box[1] <= 256 && box[2] <= 256);
        if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return LC_ERR_INVALID;
    }
    const int K = d->ks * d->ks * d->C;
    const int bn = d->Cout > 128 ? 256 : 128;
    tc::GemmArgs a{};
    int e = make_tmap(&tb, d->Wk, K, d->Cout, 1, 1, K, 0, 0, bn, &a.b_bcast);
    if (e != LC_OK) return e;
    a.b_bcast = 0;
    a.out = d->Y; a.bias = d->bias; a.residual = d->residual; a.M = d->N * hw; a.N = d->Cout; a.K = K; a.ldc = (int)d->ldc; a.ldr = (int)d->ldr;
    a.batch_in = 1; a.batch_total = 1; a.out_dtype = d->out_f32 ? tc::GEMM_OUT_F32 : tc::GEMM_OUT_BF16; a.alpha = 1.f; a.error_flag = error_flag;
    a.conv_ks = d->ks; a.conv_cchunks = d->C / 64; a.conv_stride = d->stride; a.conv_pad = d->pad; a.conv_hw = hw; a.conv_wo = d->Wo;
    return bn == 256 ? launch_gemm<256>(ta, tb, a, 1, (cudaStream_t)stream) : launch_gemm<128>(ta, tb, a, 1, (cudaStream_t)stream);
}

int lc_gemm_bf16(const void* A, int lda, long long strideA, const void* B, int ldb, long long strideB, void* C, int ldc, long long strideC, int M, int N,
                 int K, int batch, const float* bias, const float* residual, int ldr, long long strideR, void* out2, int out_f32, float alpha,
                 int* error_flag, lc_stream_t stream) {
    lc_gemm_desc d{};
    d.A = A; d.lda = lda; d.strideA_in = strideA; d.B = B; d.ldb = ldb; d.strideB_in = strideB; d.C = C; d.ldc = ldc; d.strideC_in = strideC;
    d.bias = bias; d.residual = residual; d.ldr = ldr; d.strideR_in = strideR; d.out2 = out2; d.M = M; d.N = N; d.K = K; d.batch_in = batch;
    d.batch_out = 1; d.out_f32 = out_f32; d.alpha = alpha;
    return lc_gemm_bf16_ex(&d, error_flag, stream);
}

int lc_attn_forward_prefix(const void* qkv_bf16, void* out_bf16, float* lse2, int batch, int T, int heads, const void* pk_bf16, const void* pv_bf16, int P,
                            int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(qkv_bf16 && out_bf16 && lse2 && batch >= 1 && T >= 1 && heads >= 1 && batch <= 65535 && heads <= 65535);
    LC_CHECK_ARG(P >= 0 && (P == 0 || (pk_bf16 && pv_bf16)) && T + P <= 256);
    const size_t smem = tc::attn_fwd_smem(T + P);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(tc::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LC_ERR_CUDA;
        attr_smem = smem;
    }
    tc::AttnFwdArgs a{reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), lse2, T, heads, error_flag,
                      P ? reinterpret_cast<const __nv_bfloat16*>(pk_bf16) : nullptr, P ? reinterpret_cast<const __nv_bfloat16*>(pv_bf16) : nullptr, P};
    tc::attn_fwd_kernel<<<dim3((T + 127) / 128, heads, batch), 256, smem, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_attn_forward(const void* qkv_bf16, void* out_bf16, float* lse2, int batch, int T, int heads, int* error_flag, lc_stream_t stream) {
    return lc_attn_forward_prefix(qkv_bf16, out_bf16, lse2, batch, T, heads, nullptr, nullptr, 0, error_flag, stream);
}

int lc_attn_backward_prefix(const void* qkv_bf16, const void* dout_bf16, const float* lse2, void* dqkv_bf16, int batch, int T, int heads, const void* pk_bf16,
                             const void* pv_bf16, float* dpk, float* dpv, int P, int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(qkv_bf16 && dout_bf16 && lse2 && dqkv_bf16 && batch >= 1 && T >= 1 && heads >= 1 && batch <= 65535);
    LC_CHECK_ARG(P >= 0 && (P == 0 || (pk_bf16 && pv_bf16 && dpk && dpv)) && T + P <= 256);
    const size_t smem = tc::attn_bwd_smem(T + P);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(tc::attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LC_ERR_CUDA;
        attr_smem = smem;
    }
    tc::AttnBwdArgs a{reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), reinterpret_cast<const __nv_bfloat16*>(dout_bf16), lse2,
                      reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), T, heads, error_flag, P ? reinterpret_cast<const __nv_bfloat16*>(pk_bf16) : nullptr,
                      P ? reinterpret_cast<const __nv_bfloat16*>(pv_bf16) : nullptr, dpk, dpv, P};
    tc::attn_bwd_kernel<<<dim3(heads, batch), 256, smem, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_attn_backward(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* rowdot, void* dqkv_bf16, int batch, int T,
                     int heads, int* error_flag, lc_stream_t stream) {
    (void)out_bf16; (void)rowdot;       // the softmax row term sum_j P_j dP_j is formed on chip (attn_tc.cuh); O and the scratch are not read
    return lc_attn_backward_prefix(qkv_bf16, dout_bf16, lse2, dqkv_bf16, batch, T, heads, nullptr, nullptr, nullptr, nullptr, 0, error_flag, stream);
}

int lc_vit_patchify(const float* img, void* out_bf16, int batch, lc_stream_t stream) {
    LC_CHECK_ARG(img && out_bf16 && batch >= 1);
    patchify_kernel<<<grid_for((long long)batch * 196 * 96, 256), 256, 0, (cudaStream_t)stream>>>(img, reinterpret_cast<__nv_bfloat16*>(out_bf16), batch);
    return lc_launch_status();
}
int lc_vit_set_rows(float* x, long long batch_stride, int batch, int row0, int nrows, const float* src, const float* add, int dim, lc_stream_t stream) {
    LC_CHECK_ARG(x && src && batch >= 1 && nrows >= 1 && dim % 4 == 0);
    set_rows_kernel<<<dim3(batch, nrows), 192, 0, (cudaStream_t)stream>>>(x, batch_stride, row0, src, add, dim);
    return lc_launch_status();
}
int lc_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps, long long rows, int dim, void* out_bf16, float* out_f32,
                         float* stat, lc_stream_t stream) {
    LC_CHECK_ARG(x && gamma && beta && rows >= 1 && dim == 768 && (out_bf16 || out_f32));
    layernorm_fwd_kernel<768><<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(x, gamma, beta, eps, rows, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                                       out_f32, stat);
    return lc_launch_status();
}
int lc_layernorm_backward(const float* dh, const float* dh_pool, int T, int n_active, const float* x, const float* gamma, float eps, long long rows, int dim,
                          const float* res, float* out_f32, void* out_bf16, lc_stream_t stream) {
    LC_CHECK_ARG((dh != nullptr) != (dh_pool != nullptr) && x && gamma && rows >= 1 && dim == 768 && (out_f32 || out_bf16));
    LC_CHECK_ARG(dh_pool == nullptr || (T >= 1 && n_active >= 1 && n_active <= T && rows % T == 0));
    layernorm_bwd_kernel<768><<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(dh, dh_pool, T, n_active, x, gamma, eps, rows, res, out_f32,
                                                                                       reinterpret_cast<__nv_bfloat16*>(out_bf16));
    return lc_launch_status();
}
int lc_sum_batch_rows(const float* x, long long batch_stride, int batch, int nrows, int dim, float* out, lc_stream_t stream) {
    LC_CHECK_ARG(x && out && batch >= 1 && nrows >= 1 && dim % 4 == 0);
    sum_batch_rows_kernel<<<nrows, 192, 0, (cudaStream_t)stream>>>(x, batch_stride, batch, dim, out);
    return lc_launch_status();
}
int lc_vit_pool_rows(const float* y, long long batch_stride, int batch, int r0, int nr, int dim, float* feat, lc_stream_t stream) {
    LC_CHECK_ARG(y && feat && batch >= 1 && nr >= 1 && dim % 4 == 0);
    pool_rows_kernel<<<batch, 192, 0, (cudaStream_t)stream>>>(y, batch_stride, r0, nr, dim, feat);
    return lc_launch_status();
}
int lc_linear_head(const float* feat, const float* W, const float* bias, int batch, int ncls, int dim, float* logits, int ld, lc_stream_t stream) {
    LC_CHECK_ARG(feat && W && logits && batch >= 1 && ncls >= 1 && ld >= ncls);
    linear_head_kernel<<<(batch * ncls + 3) / 4, 128, 0, (cudaStream_t)stream>>>(feat, W, bias, batch, ncls, dim, logits, ld);
    return lc_launch_status();
}
int lc_linear_head_backward(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int batch, int dim, float* dW, float* db,
                            float* dfeat, lc_stream_t stream) {
    LC_CHECK_ARG(dlogits && feat && W && dW && dfeat && ncls >= 1 && batch >= 1 && dim % 4 == 0 && ldl >= ncls);
    linear_head_bwd_kernel<<<ncls + batch, 192, 0, (cudaStream_t)stream>>>(dlogits, ldl, feat, W, ncls, batch, dim, dW, db, dfeat);
    return lc_launch_status();
}
int lc_l2p_backward(const float* dprompts, const int64_t* ids, int pool, int top_k, int length, int dim, float* dpool, const float* dkey_in, float coeff,
                    float* dkey_out, lc_stream_t stream) {
    LC_CHECK_ARG(dprompts && ids && dpool && dkey_in && dkey_out && pool >= 1 && top_k >= 1 && top_k <= pool && length >= 1 && dim % 4 == 0);
    l2p_backward_kernel<<<pool * length + pool, 192, 0, (cudaStream_t)stream>>>(dprompts, reinterpret_cast<const long long*>(ids), pool, top_k, length, dim, dpool,
                                                                             dkey_in, coeff, dkey_out);
    return lc_launch_status();
}
int lc_cast_bf16(const float* in, void* out_bf16, long long n, lc_stream_t stream) {
    LC_CHECK_ARG(in && out_bf16 && n >= 1);
    cast_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n);
    return lc_launch_status();
}

int lc_prompt_key_match(const float* query, const float* const* keys, float* const* dkeys, int nlayers, int batch, int pool, int dim, int task_id,
                        int64_t* idx, float* loss, lc_stream_t stream) {
    LC_CHECK_ARG(query && keys && idx && nlayers >= 1 && nlayers <= kPromptMaxLayers && batch >= 1 && pool >= 1 && dim == 768 && task_id < pool);
    LC_CHECK_ARG(task_id < 0 || (dkeys && loss));
    KeyMatchArgs a{};
    a.q = query; a.idx = reinterpret_cast<long long*>(idx); a.loss = loss; a.nl = nlayers; a.B = batch; a.pool = pool; a.D = dim; a.task_id = task_id;
    for (int l = 0; l < nlayers; ++l) { a.K[l] = keys[l]; a.dK[l] = dkeys ? dkeys[l] : nullptr; LC_CHECK_ARG(a.K[l] && (task_id < 0 || a.dK[l])); }
    prompt_key_match_kernel<768><<<1, 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_coda_prompt_forward(const float* query, const float* const* K, const float* const* A, const float* const* p, void* const* pk_bf16, void* const* pv_bf16,
                           int nlayers, int batch, int nk, int length, int dim, float* alpha, float* vnorm, lc_stream_t stream) {
    LC_CHECK_ARG(query && K && A && p && pk_bf16 && pv_bf16 && alpha && vnorm && nlayers >= 1 && nlayers <= kPromptMaxLayers && batch >= 1 && nk >= 1 &&
                 nk <= kCodaMaxK && length >= 2 && length % 2 == 0 && dim == 768);
    CodaArgs a{};
    a.q = query; a.alpha = alpha; a.vnorm = vnorm; a.B = batch; a.nk = nk; a.Lp = length; a.D = dim;
    for (int l = 0; l < nlayers; ++l) {
        a.K[l] = K[l]; a.A[l] = A[l]; a.p[l] = p[l]; a.pk[l] = reinterpret_cast<__nv_bfloat16*>(pk_bf16[l]); a.pv[l] = reinterpret_cast<__nv_bfloat16*>(pv_bf16[l]);
        LC_CHECK_ARG(a.K[l] && a.A[l] && a.p[l] && a.pk[l] && a.pv[l]);
    }
    coda_prompt_fwd_kernel<<<dim3(batch, nlayers), 192, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_coda_prompt_backward(const float* query, const float* const* K, const float* const* A, const float* const* p, const float* const* dpk, const float* const* dpv,
                            float* const* dK, float* const* dA, float* const* dp, int nlayers, int batch, int nk, int length, int dim, const float* alpha,
                            const float* vnorm, float* dalpha, lc_stream_t stream) {
    LC_CHECK_ARG(query && K && A && p && dpk && dpv && dK && dA && dp && alpha && vnorm && dalpha && nlayers >= 1 && nlayers <= kPromptMaxLayers && batch >= 1 &&
                 nk >= 1 && nk <= kCodaMaxK && length >= 2 && length % 2 == 0 && dim == 768);
    CodaBwdArgs a{};
    a.q = query; a.alpha = alpha; a.vnorm = vnorm; a.dalpha = dalpha; a.B = batch; a.nk = nk; a.Lp = length; a.D = dim;
    for (int l = 0; l < nlayers; ++l) {
        a.K[l] = K[l]; a.A[l] = A[l]; a.p[l] = p[l]; a.dpk[l] = dpk[l]; a.dpv[l] = dpv[l]; a.dK[l] = dK[l]; a.dA[l] = dA[l]; a.dp[l] = dp[l];
        LC_CHECK_ARG(a.K[l] && a.A[l] && a.p[l] && a.dpk[l] && a.dpv[l] && a.dK[l] && a.dA[l] && a.dp[l]);
    }
    coda_prompt_bwd_alpha_kernel<<<dim3(batch, nlayers), 192, 0, (cudaStream_t)stream>>>(a);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    coda_prompt_bwd_param_kernel<<<dim3(length + nk, nlayers), 192, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_gather_rows_bf16(const float* src, const int64_t* idx, long long idx_stride, int rows, int dim, int batch, void* out_bf16, lc_stream_t stream) {
    LC_CHECK_ARG(src && out_bf16 && rows >= 1 && rows <= 65535 && dim % 4 == 0 && batch >= 1);
    gather_rows_bf16_kernel<<<dim3(batch, rows), 192, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<const long long*>(idx), idx_stride, rows, dim,
                                                                                  reinterpret_cast<__nv_bfloat16*>(out_bf16));
    return lc_launch_status();
}
int lc_split_bf16(const float* x, void* hi_bf16, void* lo_bf16, long long n, lc_stream_t stream) {
    LC_CHECK_ARG(x && hi_bf16 && lo_bf16 && n >= 1);
    split_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(hi_bf16), reinterpret_cast<__nv_bfloat16*>(lo_bf16), n);
    return lc_launch_status();
}
int lc_gpm_project_tc(float* grad, const void* proj_hi_bf16, const void* proj_lo_bf16, int rows, int dim, void* g_hi_bf16, void* g_lo_bf16, int* error_flag,
                      lc_stream_t stream) {
    LC_CHECK_ARG(grad && proj_hi_bf16 && proj_lo_bf16 && g_hi_bf16 && g_lo_bf16 && rows >= 1 && dim >= 8 && dim % 8 == 0);
    int e = lc_split_bf16(grad, g_hi_bf16, g_lo_bf16, (long long)rows * dim, stream);
    if (e != LC_OK) return e;
    // grad <- grad - g M with M symmetric (M = U U^T): B operand [N = j][K = k] = M itself.  Three in-place passes, each grad <- grad - A_i B_i^T
    const void* As[3] = {g_hi_bf16, g_lo_bf16, g_hi_bf16};
    const void* Bs[3] = {proj_hi_bf16, proj_hi_bf16, proj_lo_bf16};
    for (int i = 0; i < 3; ++i) {
        lc_gemm_desc d{};
        d.A = As[i]; d.lda = dim; d.B = Bs[i]; d.ldb = dim; d.C = grad; d.ldc = dim; d.residual = grad; d.ldr = dim;
        d.M = rows; d.N = dim; d.K = dim; d.batch_in = 1; d.batch_out = 1; d.out_f32 = 1; d.alpha = -1.f;
        e = lc_gemm_bf16_ex(&d, error_flag, stream);
        if (e != LC_OK) return e;
    }
    return LC_OK;
}
int lc_transpose_bf16(const void* in_bf16, long long ld_in, long long rows, int cols, void* out_bf16, long long ld_out, lc_stream_t stream) {
    LC_CHECK_ARG(in_bf16 && out_bf16 && rows >= 1 && cols >= 1 && ld_in >= cols && ld_out >= rows);
    transpose_bf16_kernel<<<dim3((unsigned)((ld_out + 63) / 64), (cols + 63) / 64), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(in_bf16), ld_in, rows, cols, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld_out);
    return lc_launch_status();
}
int lc_lora_merge(const float* w, const float* A, const float* B, const float* scale, int slab_mask, int layers, int dim, int rank, void* wb_bf16,
                  void* wbt_bf16, float* w_out, lc_stream_t stream) {
    LC_CHECK_ARG(w && A && B && layers >= 1 && dim >= 32 && dim % 32 == 0 && rank >= 1 && rank <= kLoraMaxR && slab_mask >= 1 && slab_mask <= 7);
    LC_CHECK_ARG(wb_bf16 || wbt_bf16 || w_out);
    LoraMergeArgs a{};
    a.w = w; a.A = A; a.B = B; a.scale = scale; a.wb = reinterpret_cast<__nv_bfloat16*>(wb_bf16); a.wbt = reinterpret_cast<__nv_bfloat16*>(wbt_bf16);
    a.w_out = w_out; a.D = dim; a.R = rank; a.ns = 0;
    for (int s = 0; s < 3; ++s) if (slab_mask & (1 << s)) a.slab[a.ns++] = s;
    lora_merge_kernel<<<dim3(dim / 32, dim / 32, layers * a.ns), dim3(32, 8), 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
long long lc_lora_bgrad_partial_floats(int nslab, int dim, int rank, int nchunk) { return (long long)nslab * dim * rank * nchunk; }
int lc_rowouter_bf16(const void* x_bf16, long long ldx, int x0, int x_slab_stride, int nslab, int dim, const float* z, int ldz, int z0, int z_slab_stride, int rank,
                     long long rows, float* partial, int nchunk, float* out, int transposed, const float* scale_dev, lc_stream_t stream) {
    LC_CHECK_ARG(x_bf16 && z && partial && out && nslab >= 1 && dim >= 768 && dim % 768 == 0 && rank >= 1 && rank <= 16 && rows >= 1 && nchunk >= 1);
    LC_CHECK_ARG(ldx % 4 == 0 && x0 % 4 == 0 && x_slab_stride % 4 == 0 && z0 >= 0 && ldz >= z0 + (nslab - 1) * z_slab_stride + rank);
    const int rows_per = (int)((rows + nchunk - 1) / nchunk);
    const dim3 grid(nslab * (dim / 768), nchunk);
    const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
    if (rank <= 10) rowouter_partial_kernel<10><<<grid, 192, 0, (cudaStream_t)stream>>>(X, ldx, x0, x_slab_stride, dim, z, ldz, z0, z_slab_stride, rank, rows, rows_per, partial);
    else rowouter_partial_kernel<16><<<grid, 192, 0, (cudaStream_t)stream>>>(X, ldx, x0, x_slab_stride, dim, z, ldz, z0, z_slab_stride, rank, rows, rows_per, partial);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    const long long per_chunk = (long long)nslab * dim * rank;
    rowouter_reduce_kernel<<<grid_for(per_chunk, 256), 256, 0, (cudaStream_t)stream>>>(partial, per_chunk, nchunk, out, dim, rank, transposed, scale_dev);
    return lc_launch_status();
}
int lc_lora_bgrad_rows(const void* x_bf16, long long ldx, int x0, int x_slab_stride, int nslab, int dim, const float* z, int ldz, int rank, long long rows, float* partial,
                       int nchunk, float* out, lc_stream_t stream) {
    return lc_rowouter_bf16(x_bf16, ldx, x0, x_slab_stride, nslab, dim, z, ldz, 0, rank, rank, rows, partial, nchunk, out, 0, nullptr, stream);
}
int lc_coldot_accumulate(const float* g, int ldg, const float* z, int ldz, int z0, int cols, int rank, long long rows, const float* col_weight, float* partial, int nchunk,
                         float* dmag, lc_stream_t stream) {
    LC_CHECK_ARG(g && z && partial && dmag && cols >= 1 && cols <= 128 && rank >= 1 && cols % rank == 0 && rows >= 1 && nchunk >= 1 && ldg >= cols && ldz >= z0 + cols);
    const int rows_per = (int)((rows + nchunk - 1) / nchunk);
    coldot_partial_kernel<<<nchunk, 256, 0, (cudaStream_t)stream>>>(g, ldg, z, ldz, z0, cols, rows, rows_per, partial);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    coldot_finish_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(partial, nchunk, cols, rank, col_weight, dmag);
    return lc_launch_status();
}

}  // extern "C"
