// C entry points of the generic layer kernels (nn_ops.cuh) used by the ResNet18 and AlexNet_TRGP engines.  See include/lc_b200.h.
#include "../../include/lc_b200.h"
#include "nn_ops.cuh"

using namespace lc;
using namespace lc::nn;

namespace {
constexpr int kColsumBlocks = 296;          // row slices of the two-stage column reductions (2 per SM)
int colsum_gx(long long M) { long long g = (M + 63) / 64; return (int)(g < 1 ? 1 : (g < kColsumBlocks ? g : kColsumBlocks)); }
}  // namespace

extern "C" {

long long lc_nn_bn_scratch_floats(int C) { return (long long)kColsumBlocks * 2 * C + 3 * (long long)C; }

int lc_nn_im2col(const void* src, int src_kind, int N, int H, int W, int C, int ks, int stride, int pad, int korder, void* col_bf16, long long ld_col,
                 void* colT_bf16, long long ld_colT, int Kp, lc_stream_t stream) {
    LC_CHECK_ARG(src && (col_bf16 || colT_bf16) && N >= 1 && H >= 1 && W >= 1 && C >= 1 && ks >= 1 && stride >= 1 && pad >= 0 && src_kind >= 0 && src_kind <= 2 &&
                 (korder == 0 || korder == 1));
    Im2colArgs a{};
    a.src = src; a.src_kind = src_kind; a.N = N; a.H = H; a.W = W; a.C = C; a.ks = ks; a.stride = stride; a.pad = pad; a.korder = korder;
    a.Ho = (H + 2 * pad - ks) / stride + 1; a.Wo = (W + 2 * pad - ks) / stride + 1;
    LC_CHECK_ARG(a.Ho >= 1 && a.Wo >= 1);
    a.K = ks * ks * C; a.Kp = Kp;
    a.M = (long long)N * a.Ho * a.Wo;
    LC_CHECK_ARG(Kp >= a.K && (col_bf16 == nullptr || ld_col >= Kp) && (colT_bf16 == nullptr || ld_colT >= a.M));
    a.col = reinterpret_cast<__nv_bfloat16*>(col_bf16); a.ld_col = ld_col; a.colT = reinterpret_cast<__nv_bfloat16*>(colT_bf16); a.ld_colT = ld_colT;
    const long long mspan = colT_bf16 != nullptr ? ld_colT : a.M;      // the zero tail of the transposed rows is written too
    if (col_bf16 == nullptr && src_kind == SRC_NHWC_BF16 && korder == KORDER_TAP_C && C % 64 == 0 && Kp == a.K && ld_colT % 8 == 0 &&
        ((uintptr_t)colT_bf16 % 16) == 0 && ((uintptr_t)src % 16) == 0) {
        dim3 grid((unsigned)((mspan + 63) / 64), (unsigned)(ks * ks * (C / 64)));
        im2colT_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
        return lc_launch_status();
    }
    dim3 grid((unsigned)((mspan + 63) / 64), (unsigned)((Kp + 63) / 64));
    im2col_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_nn_im2col_f32(const void* src, int src_kind, int N, int H, int W, int C, int ks, int stride, int pad, int korder, float* col, long long ld_col, float* colT,
                     long long ld_colT, lc_stream_t stream) {
    LC_CHECK_ARG(src && (col || colT) && N >= 1 && H >= 1 && W >= 1 && C >= 1 && ks >= 1 && stride >= 1 && pad >= 0 && src_kind >= 0 && src_kind <= 2 &&
                 (korder == 0 || korder == 1));
    Im2colArgs a{};
    a.src = src; a.src_kind = src_kind; a.N = N; a.H = H; a.W = W; a.C = C; a.ks = ks; a.stride = stride; a.pad = pad; a.korder = korder;
    a.Ho = (H + 2 * pad - ks) / stride + 1; a.Wo = (W + 2 * pad - ks) / stride + 1;
    LC_CHECK_ARG(a.Ho >= 1 && a.Wo >= 1);
    a.K = ks * ks * C; a.Kp = a.K;
    a.M = (long long)N * a.Ho * a.Wo;
    LC_CHECK_ARG((col == nullptr || ld_col >= a.K) && (colT == nullptr || ld_colT >= a.M));
    a.col = reinterpret_cast<__nv_bfloat16*>(col); a.ld_col = ld_col; a.colT = reinterpret_cast<__nv_bfloat16*>(colT); a.ld_colT = ld_colT;
    const long long mspan = colT != nullptr ? ld_colT : a.M;
    dim3 grid((unsigned)((mspan + 63) / 64), (unsigned)((a.K + 63) / 64));
    im2col_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_nn_col2im(const void* dcol_bf16, long long ld, const float* addend, float* dx, int N, int H, int W, int C, int ks, int stride, int pad, int korder,
                 lc_stream_t stream) {
    LC_CHECK_ARG(dcol_bf16 && dx && N >= 1 && C >= 1 && ks >= 1 && stride >= 1 && pad >= 0 && ld >= (long long)ks * ks * C);
    Col2imArgs a{};
    a.dcol = reinterpret_cast<const __nv_bfloat16*>(dcol_bf16); a.ld = ld; a.addend = addend; a.dx = dx; a.N = N; a.H = H; a.W = W; a.C = C; a.ks = ks;
    a.stride = stride; a.pad = pad; a.korder = korder;
    a.Ho = (H + 2 * pad - ks) / stride + 1; a.Wo = (W + 2 * pad - ks) / stride + 1;
    col2im_kernel<<<nn_grid((long long)N * H * W * C), 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_nn_bn_stats(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum, float* running, float* aff,
                   float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(y && aff && scratch && M >= 1 && C >= 64 && C % 64 == 0);
    const int gx = colsum_gx(M);
    bn_colsum_kernel<<<dim3(gx, C / 64), 256, 0, (cudaStream_t)stream>>>(y, M, C, scratch);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    bn_finalize_kernel<<<(C + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(scratch, gx, M, C, gamma, beta, eps, momentum, running, aff);
    return lc_launch_status();
}

int lc_nn_bn_eval_affine(const float* running, int C, const float* gamma, const float* beta, float eps, float* aff, lc_stream_t stream) {
    LC_CHECK_ARG(running && aff && C >= 1);
    bn_eval_affine_kernel2<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(running, C, gamma, beta, eps, aff);
    return lc_launch_status();
}

int lc_nn_bn_act(const float* y, const float* aff, const float* res, const float* res_aff, long long M, int C, int relu, float drop_p,
                 const unsigned long long* rng, int rng_stream, void* out_bf16, float* out_f32, lc_stream_t stream) {
    LC_CHECK_ARG(y && aff && (out_bf16 || out_f32) && M >= 1 && C >= 4 && C % 4 == 0 && drop_p >= 0.f && drop_p < 1.f);
    BnActArgs2 a{};
    a.y = y; a.aff = aff; a.res = res; a.res_aff = res_aff; a.M = M; a.C = C; a.relu = relu; a.drop_p = drop_p; a.rng = rng; a.rng_stream = (unsigned long long)rng_stream;
    a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.out_f32 = out_f32;
    bn_act_kernel<<<nn_grid(M * (C / 4)), 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_nn_dropout_mask(const unsigned long long* rng, int rng_stream, float drop_p, long long n, unsigned char* keep, lc_stream_t stream) {
    LC_CHECK_ARG(rng && keep && n >= 1);
    dropout_mask_kernel<<<nn_grid(n), 256, 0, (cudaStream_t)stream>>>(rng, (unsigned long long)rng_stream, drop_p, n, keep);
    return lc_launch_status();
}
int lc_nn_rng_advance(unsigned long long* rng, lc_stream_t stream) {
    LC_CHECK_ARG(rng != nullptr);
    rng_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rng);
    return lc_launch_status();
}

int lc_nn_bn_backward(const float* g, const float* act_f32, const void* act_bf16, float gscale, const float* y, const float* aff, long long M, int C,
                      float* dgamma, float* dbeta, void* dy_bf16, float* dy_f32, float* dz_out, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(g && y && aff && scratch && (dy_bf16 || dy_f32) && M >= 1 && C >= 64 && C % 64 == 0 && !(act_f32 && act_bf16));
    BnBwdArgs2 a{};
    a.g = g; a.act_f32 = act_f32; a.act_bf16 = reinterpret_cast<const __nv_bfloat16*>(act_bf16); a.relu_mask = (act_f32 || act_bf16) ? 1 : 0; a.gscale = gscale;
    a.y = y; a.aff = aff; a.M = M; a.C = C; a.dgamma = dgamma; a.dbeta = dbeta; a.dy_bf16 = reinterpret_cast<__nv_bfloat16*>(dy_bf16); a.dy_f32 = dy_f32;
    a.dz_out = dz_out;
    const int gx = colsum_gx(M);
    a.partial = scratch; a.coef = scratch + (size_t)kColsumBlocks * 2 * C;
    cudaStream_t st = (cudaStream_t)stream;
    bn_bwd_colsum_kernel<<<dim3(gx, C / 64), 256, 0, st>>>(a);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, 8), 0, st>>>(a, gx);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    bn_bwd_apply_kernel2<<<nn_grid(M * (C / 4)), 256, 0, st>>>(a);
    return lc_launch_status();
}

int lc_nn_maxpool_forward(const float* in, int N, int H, int W, int C, int k, int stride, int pad, float* out_f32, void* out_bf16, unsigned char* idx,
                          lc_stream_t stream) {
    LC_CHECK_ARG(in && idx && (out_f32 || out_bf16) && N >= 1 && k >= 1 && k <= 15 && stride >= 1 && pad >= 0 && pad < k);
    PoolArgs a{};
    a.in = in; a.out_f32 = out_f32; a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.idx = idx; a.N = N; a.H = H; a.W = W; a.C = C; a.k = k;
    a.stride = stride; a.pad = pad; a.Ho = (H + 2 * pad - k) / stride + 1; a.Wo = (W + 2 * pad - k) / stride + 1;
    LC_CHECK_ARG(C % 4 == 0);
    maxpool_fwd_kernel<<<nn_grid((long long)N * a.Ho * a.Wo * (C / 4)), 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}
int lc_nn_maxpool_backward(const float* g, const unsigned char* idx, int N, int H, int W, int C, int k, int stride, int pad, float* dx, lc_stream_t stream) {
    LC_CHECK_ARG(g && idx && dx && N >= 1 && k >= 1 && stride >= 1);
    PoolBwdArgs a{};
    a.g = g; a.idx = idx; a.dx = dx; a.N = N; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride; a.pad = pad;
    a.Ho = (H + 2 * pad - k) / stride + 1; a.Wo = (W + 2 * pad - k) / stride + 1;
    LC_CHECK_ARG(C % 4 == 0);
    maxpool_bwd_kernel<<<nn_grid((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_nn_avgpool_forward(const float* in, int N, int HW, int C, float* out, lc_stream_t stream) {
    LC_CHECK_ARG(in && out && N >= 1 && HW >= 1 && C >= 1);
    avgpool_nhwc_fwd_kernel<<<nn_grid((long long)N * C), 256, 0, (cudaStream_t)stream>>>(in, N, HW, C, out);
    return lc_launch_status();
}
int lc_nn_avgpool_backward(const float* dfeat, int N, int HW, int C, float* dx, lc_stream_t stream) {
    LC_CHECK_ARG(dfeat && dx && N >= 1 && HW >= 1 && C >= 1);
    avgpool_nhwc_bwd_kernel<<<nn_grid((long long)N * HW * C), 256, 0, (cudaStream_t)stream>>>(dfeat, N, HW, C, dx);
    return lc_launch_status();
}

int lc_nn_pack_weight(const float* w, int Cout, int Cin, int ks, int korder, int mode, void* out_bf16, long long ld, lc_stream_t stream) {
    LC_CHECK_ARG(w && out_bf16 && Cout >= 1 && Cin >= 1 && ks >= 1 && mode >= 0 && mode <= 2 && (korder == 0 || korder == 1));
    const int K = Cin * ks * ks;
    const long long rows = mode == PACK_FWD ? Cout : (mode == PACK_TRANSPOSED ? K : Cin);
    LC_CHECK_ARG(ld >= (mode == PACK_FWD ? K : (mode == PACK_TRANSPOSED ? Cout : ks * ks * Cout)));
    pack_weight_kernel<<<nn_grid(rows * ld), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, ks, korder, mode, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld, rows);
    return lc_launch_status();
}

int lc_nn_wgrad_reduce(const float* partial, int nsplit, int Cout, int Cin, int ks, int korder, long long ldp, float* dw, lc_stream_t stream) {
    LC_CHECK_ARG(partial && dw && nsplit >= 1 && Cout >= 1 && Cin >= 1 && ks >= 1 && ldp >= (long long)Cin * ks * ks);
    wgrad_reduce_kernel<<<nn_grid((long long)Cout * Cin * ks * ks), 256, 0, (cudaStream_t)stream>>>(partial, nsplit, Cout, Cin, ks, korder, ldp, dw);
    return lc_launch_status();
}

int lc_nn_cast_transpose(const float* src, long long rows, int cols, void* out_bf16, long long ld, void* outT_bf16, long long ldT, lc_stream_t stream) {
    LC_CHECK_ARG(src && (out_bf16 || outT_bf16) && rows >= 1 && cols >= 1 && (out_bf16 == nullptr || ld >= cols) && (outT_bf16 == nullptr || ldT >= rows));
    const long long rspan = outT_bf16 != nullptr ? ldT : rows;
    const long long cspan = out_bf16 != nullptr ? ld : cols;
    dim3 grid((unsigned)((rspan + 63) / 64), (unsigned)((cspan + 63) / 64));
    cast_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld,
                                                                   reinterpret_cast<__nv_bfloat16*>(outT_bf16), ldT);
    return lc_launch_status();
}

}  // extern "C"
