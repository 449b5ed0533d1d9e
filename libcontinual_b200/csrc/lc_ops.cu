// C entry points of the continual-learning specific kernels (cl_ops.cuh).  See include/lc_b200.h.
#include "../../include/lc_b200.h"
#include "cl_ops.cuh"
#include "data_ops.cuh"

using namespace lc;

extern "C" {

int lc_cosine_head_forward(const float* feat, const float* W, const float* sigma, int batch, int ncls, int feat_dim, float* inv_norm,
                           float* scores, float* logits, int ld, lc_stream_t stream) {
    LC_CHECK_ARG(feat && W && inv_norm && scores && logits && batch >= 1 && ncls >= 1 && ld >= ncls && feat_dim == 64);
    cudaStream_t st = (cudaStream_t)stream;
    row_inv_norm_kernel<64><<<(batch + ncls + 3) / 4, 128, 0, st>>>(feat, batch, W, ncls, inv_norm);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    cosine_head_fwd_kernel<64><<<(batch * ncls + 127) / 128, 128, 0, st>>>(feat, W, inv_norm, sigma, batch, ncls, scores, logits, ld);
    return lc_launch_status();
}

int lc_cosine_head_backward(const float* gscores, int ld, const float* feat, const float* W, const float* inv_norm, int batch, int ncls,
                            int feat_dim, float* dfeat, float* dW, lc_stream_t stream) {
    LC_CHECK_ARG(gscores && feat && W && inv_norm && dfeat && dW && batch >= 1 && ncls >= 1 && feat_dim == 64);
    cosine_head_bwd_kernel<64><<<batch + ncls, 64, 0, (cudaStream_t)stream>>>(gscores, ld, feat, W, inv_norm, batch, ncls, dfeat, dW);
    return lc_launch_status();
}

int lc_lucir_loss(const float* logits, const float* scores, int ld, const float* feat, const float* ref_feat, int feat_dim, const int64_t* y,
                  int batch, int ncls, int num_old, int K, float cur_lamda, float margin, float lw_mr, float* dlogits, float* dscores,
                  float* dfeat, int64_t* pred, float* scal, float* dsigma, lc_stream_t stream) {
    LC_CHECK_ARG(logits && scores && feat && ref_feat && y && dlogits && dscores && dfeat && pred && scal);
    LC_CHECK_ARG(batch >= 1 && ncls >= 1 && ld >= ncls && num_old >= 0 && num_old <= ncls && K >= 1 && feat_dim >= 1);
    LucirArgs a{};
    a.logits = logits; a.scores = scores; a.feat = feat; a.ref_feat = ref_feat; a.y = reinterpret_cast<const long long*>(y);
    a.dlogits = dlogits; a.dscores = dscores; a.dfeat = dfeat; a.pred = reinterpret_cast<long long*>(pred); a.scal = scal; a.dsigma = dsigma;
    a.B = batch; a.C = ncls; a.ld = ld; a.D = feat_dim; a.num_old = num_old; a.K = K; a.cur_lamda = cur_lamda; a.margin = margin; a.lw_mr = lw_mr;
    lucir_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_l2p_select(const float* query, const float* key, int batch, int pool, int dim, int top_k, float* sim, int64_t* ids, int* hist,
                  float* reduce_sim, float* dkey, float* scratch, lc_stream_t stream) {
    return lc_l2p_select_phase(query, key, batch, pool, dim, top_k, sim, ids, hist, reduce_sim, dkey, scratch, 0, stream);
}

int lc_l2p_select_phase(const float* query, const float* key, int batch, int pool, int dim, int top_k, float* sim, int64_t* ids, int* hist,
                        float* reduce_sim, float* dkey, float* scratch, int phase, lc_stream_t stream) {
    LC_CHECK_ARG(query && key && sim && ids && hist && reduce_sim && scratch && batch >= 1 && batch <= 4096 && pool >= 1 && pool <= 32 && top_k >= 1 &&
                 top_k <= pool && dim >= 1 && phase >= 0 && phase <= 2);
    L2pArgs a{};
    a.phase = phase;
    a.query = query; a.key = key; a.sim = sim; a.ids = reinterpret_cast<long long*>(ids); a.hist = hist; a.reduce_sim = reduce_sim; a.dkey = dkey;
    a.qsum = scratch; a.B = batch; a.P = pool; a.D = dim; a.top_k = top_k;
    const size_t smem = (32 + batch) * sizeof(float) + 64 * sizeof(int) + 32 * sizeof(float);
    if (phase != 2) l2p_sim_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(a);
    l2p_select_kernel<<<1, kL2pNT, smem, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_l2p_gather(const float* prompt, const int64_t* ids, float* out, int batch, int top_k, int length, int dim, lc_stream_t stream) {
    LC_CHECK_ARG(prompt && ids && out && batch >= 1 && top_k >= 1 && length >= 1 && dim % 4 == 0);
    const long long n4 = (long long)batch * top_k * length * dim / 4;
    const int grid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    l2p_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prompt, reinterpret_cast<const long long*>(ids), out, batch, top_k, length, dim);
    return lc_launch_status();
}

int lc_gpm_project(float* grad, const float* proj, int rows, int dim, lc_stream_t stream) {
    LC_CHECK_ARG(grad && proj && rows >= 1 && dim >= 4 && dim % 4 == 0 && dim <= 4096);
    constexpr int RT = 4;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(gpm_project_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT * 4096 * 4) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    gpm_project_kernel<RT><<<(rows + RT - 1) / RT, 256, (size_t)RT * dim * sizeof(float), (cudaStream_t)stream>>>(grad, proj, rows, dim);
    return lc_launch_status();
}

int lc_lora_merge_qkv(const float* qkv_w, const float* A_k, const float* B_k, const float* A_v, const float* B_v, float* out, int dim, int rank,
                      lc_stream_t stream) {
    LC_CHECK_ARG(qkv_w && A_k && B_k && A_v && B_v && out && dim % 4 == 0 && rank >= 1 && rank <= 64);
    lora_merge_qkv_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(qkv_w, A_k, B_k, A_v, B_v, out, dim, rank);
    return lc_launch_status();
}

int lc_lora_bgrad(const float* dW, const float* A, float* dB, int dim, int rank, lc_stream_t stream) {
    LC_CHECK_ARG(dW && A && dB && dim >= 1 && rank >= 1);
    lora_bgrad_kernel<<<(dim + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dW, A, dB, dim, rank);
    return lc_launch_status();
}

int lc_herding_select(const float* feats, const int* cls_begin, int ncls, int dim, int per_class, float* work, int64_t* out, lc_stream_t stream) {
    LC_CHECK_ARG(feats && cls_begin && work && out && ncls >= 1 && per_class >= 1 && dim == 64);
    herding_select_kernel<64><<<ncls, 256, 0, (cudaStream_t)stream>>>(feats, work, cls_begin, per_class, reinterpret_cast<long long*>(out));
    return lc_launch_status();
}

int lc_ncm_classify(const float* feat, const float* means, int batch, int ncls, int dim, int64_t* pred, lc_stream_t stream) {
    LC_CHECK_ARG(feat && means && pred && batch >= 1 && ncls >= 1 && dim == 64);
    ncm_classify_kernel<64><<<(batch + 3) / 4, 128, 0, (cudaStream_t)stream>>>(feat, means, batch, ncls, reinterpret_cast<long long*>(pred));
    return lc_launch_status();
}

// ---- input pipeline / evaluation meter (data_ops.cuh) ------------------------------------------------------------------------------------------------
static inline int data_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

int lc_augment_cifar_u8(const uint8_t* src, const int64_t* idx, const int* draw, const float* bright, float* out, int batch, int H, int W, int pad,
                        const float* mean3, const float* std3, lc_stream_t stream) {
    LC_CHECK_ARG(src && draw && out && mean3 && std3 && batch >= 1 && H >= 1 && W >= 1 && pad >= 0);
    AugCifarArgs a{};
    a.src = src; a.idx = reinterpret_cast<const long long*>(idx); a.draw = draw; a.bright = bright; a.out = out; a.B = batch; a.H = H; a.W = W; a.pad = pad;
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean3[c]; a.std[c] = std3[c]; }
    augment_cifar_kernel<<<data_grid((long long)batch * H * W), 256, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

long long lc_resize_scratch_bytes(int batch, int H, int out) {
    return (long long)batch * 2 * out * (kResizeKMax + 2) * 4 + (long long)batch * H * out * 3;
}

int lc_resize_crop_u8(const uint8_t* src, const int64_t* idx, const int* draw, const int* flip, float* out, int batch, int H, int W, int out_size,
                      const float* mean3, const float* std3, void* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(src && draw && out && mean3 && std3 && scratch && batch >= 1 && H >= 1 && W >= 1 && out_size >= 1 && ((uintptr_t)scratch % 16 == 0));
    ResizeArgs a{};
    a.src = src; a.idx = reinterpret_cast<const long long*>(idx); a.draw = draw; a.flip = flip; a.out = out; a.B = batch; a.H = H; a.W = W; a.OUT = out_size;
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean3[c]; a.std[c] = std3[c]; }
    a.coef = reinterpret_cast<int*>(scratch);
    a.cmin = a.coef + (size_t)batch * 2 * out_size * kResizeKMax;
    a.tmp = reinterpret_cast<unsigned char*>(a.cmin + (size_t)batch * 2 * out_size * 2);
    cudaStream_t st = (cudaStream_t)stream;
    resize_coeffs_kernel<<<(batch * 2 * out_size + 127) / 128, 128, 0, st>>>(a);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    resize_h_kernel<<<data_grid((long long)batch * H * out_size), 256, 0, st>>>(a);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    resize_v_kernel<<<data_grid((long long)batch * out_size * out_size), 256, 0, st>>>(a);
    return lc_launch_status();
}

int lc_eval_meter(const int64_t* pred, const int64_t* label, int n, const int* bounds, int ntask, int task, unsigned long long* counts, lc_stream_t stream) {
    LC_CHECK_ARG(pred && label && counts && n >= 1 && ((task >= 0) || (bounds && ntask >= 1)));
    eval_meter_kernel<<<data_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(pred), reinterpret_cast<const long long*>(label), n,
                                                                       bounds, ntask, task, counts);
    return lc_launch_status();
}

int lc_eval_fold(unsigned long long* batch_counts, unsigned long long* total, int ntask, int reference_rounding, lc_stream_t stream) {
    LC_CHECK_ARG(batch_counts && total && ntask >= 1 && ntask <= 1024);
    eval_fold_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(batch_counts, total, ntask, reference_rounding);
    return lc_launch_status();
}

}  // extern "C"
