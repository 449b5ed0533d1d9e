/* libcontinual_b200 — C ABI of the B200-native per-step training hot path of RL-VIG/LibContinual.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t (passed as void*), launches
 * asynchronously on that stream, performs no allocation and no host synchronisation (except the *_create / *_destroy
 * calls), and returns 0 on success or a negative errno-style code (-22 invalid argument / unsupported shape, -5 CUDA
 * error).  All tensors are fp32 unless stated; labels are int64.  Activations are NHWC (torch channels_last memory).
 *
 * The reference has no FFI: these calls replace eager ATen/cuDNN/cuBLAS op sequences issued by the Python classes cited on
 * each entry (paths relative to the reference root).  The Python host that mirrors the reference plugin surface and binds
 * this library through ctypes is `libcontinual_b200/`; INTEGRATION.md shows the stub a reference maintainer would add.
 */
#ifndef LC_B200_H
#define LC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_OK 0
#define LC_ERR_INVALID (-22)
#define LC_ERR_CUDA (-5)

typedef void* lc_stream_t; /* cudaStream_t */

const char* lc_version(void);
/* 0 when the current CUDA device is compute capability 10.x (the only target this library is built for). */
int lc_device_check(void);

/* ---------------------------------------------------------------------------------------------------------------------
 * CIFAR ResNet backbone (core/model/backbone/resnet.py:289-412 `CifarResNet`/`ResNetBasicblock`, factory :760-763).
 * Parameter arena: the reference's `named_parameters()` order, conv weights OIHW.  Running-stat arena: per BN layer
 * (mean[C], var[C]) in `named_buffers()` order (num_batches_tracked is kept by the host).  Workspace: opaque scratch of
 * lc_resnet_workspace_floats() floats that the caller zero-fills ONCE after allocation.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct lc_resnet lc_resnet;

int lc_resnet_create(int depth, int in_ch, int img, int max_batch, lc_resnet** out);
void lc_resnet_destroy(lc_resnet* net);
/* Arithmetic mode of the 3x3 stride-1 convolutions (forward + data gradient): 0 = exact fp32 FMA on CUDA cores (default),
 * 1 = tcgen05 tensor cores, fp32 accumulation in TMEM: TF32 operands for forward / data gradient (the arithmetic class cuDNN uses
 * for the reference's convs under PyTorch's default allow_tf32), BF16 operands for the weight gradient.  workspace word 8 (int) is set to 1 if a tensor-core barrier ever timed out. */
int lc_resnet_set_mode(lc_resnet* net, int mode);
int lc_resnet_get_mode(const lc_resnet* net);
/* last_relu = 0: the last residual block omits its final ReLU (LUCIR's `modified_ResNet`, resnet.py:472-502, factory `resnet32_V2`). */
int lc_resnet_set_last_relu(lc_resnet* net, int last_relu);
long long lc_resnet_param_count(const lc_resnet* net);
long long lc_resnet_rstat_count(const lc_resnet* net);
long long lc_resnet_workspace_floats(const lc_resnet* net);
int lc_resnet_num_convs(const lc_resnet* net);
int lc_resnet_num_launches(const lc_resnet* net, int backward);
/* Layout of conv layer `idx` (reference registration order) inside the arenas; any out pointer may be NULL. */
int lc_resnet_conv_info(const lc_resnet* net, int idx, long long* w_off, int* cout, int* cin, int* ksize, int* stride,
                        long long* gamma_off, long long* beta_off, long long* rstat_off);
enum { LC_WS_FEAT = 0, LC_WS_GRAD_LAST = 1, LC_WS_FMAP1 = 2, LC_WS_FMAP2 = 3, LC_WS_FMAP3 = 4, LC_WS_DFEAT = 5 };
/* Float offset inside the workspace of: pooled features [B][64]; gradient w.r.t. the last feature map [B][8][8][64]
 * (written by lc_head_backward, consumed by lc_resnet_backward); the three stage outputs (NHWC). */
long long lc_resnet_ws_offset(const lc_resnet* net, int what);

/* `CifarResNet.forward` up to the last residual block (pooling lives in lc_head_forward).  train != 0: batch statistics,
 * running stats updated when update_running != 0 (nn.BatchNorm2d train mode); train == 0: running statistics. */
int lc_resnet_forward(lc_resnet* net, const float* x_nchw, int batch, const float* params, float* rstat, float* workspace,
                      int train, int update_running, lc_stream_t stream);
/* Autograd of the above for every backbone parameter: consumes LC_WS_GRAD_LAST, writes grads[0 .. param_count). */
int lc_resnet_backward(lc_resnet* net, const float* x_nchw, int batch, const float* params, float* workspace, float* grads,
                       lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Head + losses.
 * lc_head_forward  : nn.AvgPool2d(8)+flatten (resnet.py:389-390) and nn.Linear (ewc.py:52-57, icarl.py:24-38, finetune.py:19).
 * lc_loss_ce_kd    : F.cross_entropy over logits[:, ce_lo:ce_hi) with targets y-ce_lo (ewc.py:90-99, icarl.py:208-209,
 *                    lwf.py:57-62) + kd_w * `_KD_loss`(logits[:, :kd_n], teacher[:, :kd_n], T) (icarl.py:198-206, lwf.py:75-78),
 *                    d(loss)/d(logits), argmax over logits[:, :pred_n), #correct.  scal: [0] loss [1] #correct [2] ce [3] kd.
 * lc_head_backward : Linear + AvgPool autograd: dW[ncls][C], db[ncls], dfeat[B][C], and the broadcast gradient of the last
 *                    feature map (nullable).
 * ------------------------------------------------------------------------------------------------------------------- */
int lc_head_forward(const float* act_nhwc, int batch, int hw, int feat_dim, const float* W, const float* bias, int ncls,
                    float* feat, float* logits, int ldl, lc_stream_t stream);
int lc_loss_ce_kd(const float* logits, int ldl, const float* teacher, int ldt, const int64_t* y, int batch, int ce_lo, int ce_hi,
                  int kd_n, float kd_w, float T, int pred_n, float* dlogits, int64_t* pred, float* scal, lc_stream_t stream);
/* L2P's masked loss (l2p.py:89-106): logits outside [lo, hi) are -inf, so CE and argmax run over the slice; scal[0] = CE + extra_coeff * *extra
 * (the pull-constraint term, extra = reduce_sim of lc_l2p_select, extra_coeff = -pull_constraint_coeff). */
int lc_loss_ce_masked(const float* logits, int ldl, const int64_t* y, int batch, int lo, int hi, const float* extra, float extra_coeff, float* dlogits,
                      int64_t* pred, float* scal, lc_stream_t stream);
int lc_head_backward(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int batch, int feat_dim, float* dW,
                     float* db, float* dfeat, float* gact, int hw, lc_stream_t stream);
/* Stand-alone nn.AvgPool2d(8)+flatten and its backward, for callers that keep their own head on backbone(x)['features']. */
int lc_avgpool_forward(const float* act_nhwc, int batch, int hw, int feat_dim, float* feat, lc_stream_t stream);
int lc_avgpool_backward(const float* dfeat, int batch, int hw, int feat_dim, float* gact_nhwc, lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Flat-arena kernels.  `hp` arrays live in DEVICE memory (CUDA-graph friendly).
 * lc_ewc_penalty_grad : `EWC.compute_ewc` + autograd (ewc.py:207-225,100): grad += lamda*F*(theta-theta*);
 *                       scal[4] = sum F (theta-theta*)^2 / 2 ; scal[0] += lamda * scal[4].   scratch >= 2*296+2 floats.
 * lc_fisher_accumulate: fisher += grad^2 * weight (ewc.py:173).   lc_fisher_merge: fisher/num_samples, alpha-EMA (ewc.py:129-131,202-204).
 * lc_sgd_momentum     : torch.optim.SGD step, hp = {lr, momentum, weight_decay}.
 * lc_adam             : torch.optim.Adam step, hp = {lr, b1, b2, eps, wd, 1-b1^t, 1-b2^t}.
 * lc_adam_tick        : device-side step counter: hp[7] += 1, hp[5] = 1-b1^hp[7], hp[6] = 1-b2^hp[7] with b1, b2 read as DOUBLES from
 *                       hp[8..11] (hp: 12 floats, 8-byte aligned); a captured step replays it in front of lc_adam, so bias corrections
 *                       never depend on host staging.
 * lc_clip_grad_norm   : torch.nn.utils.clip_grad_norm_(.., max_norm) over one arena (l2p.py:104); scratch >= 2*296 floats.
 * ------------------------------------------------------------------------------------------------------------------- */
int lc_ewc_penalty_grad(const float* theta, const float* theta_ref, const float* fisher, float* grad, long long n, const float* hp_lamda,
                        float* scratch, uint32_t* counter, float* scal, lc_stream_t stream);
int lc_fisher_accumulate(float* fisher, const float* grad, long long n, float weight, lc_stream_t stream);
int lc_fisher_merge(float* f_new, const float* f_old, long long n, float num_samples, float alpha, lc_stream_t stream);
int lc_sgd_momentum(float* p, const float* g, float* m, long long n, const float* hp, lc_stream_t stream);
/* lc_sgd_momentum with the elements [freeze_lo, freeze_hi) left untouched (a param group with lr = 0, weight_decay = 0: lucir.py:229-240). */
int lc_sgd_momentum_frozen(float* p, const float* g, float* m, long long n, const float* hp, long long freeze_lo, long long freeze_hi,
                           lc_stream_t stream);
int lc_adam(float* p, const float* g, float* m, float* v, long long n, const float* hp, lc_stream_t stream);
int lc_adam_tick(float* hp, lc_stream_t stream);
int lc_clip_grad_norm(float* g, long long n, float max_norm, float* scratch, float* norm_out, lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Continual-learning specific kernels.
 * lc_cosine_head_forward : `CosineLinear` / `SplitCosineLinear` (resnet.py:418-463): scores = normalize(feat) . normalize(W)^T,
 *                          logits = sigma * scores.  inv_norm: [batch + ncls] floats (kept for the backward).
 * lc_cosine_head_backward: autograd of the above given gscores = d(loss)/d(scores) (caller folds sigma: sigma*dlogits + dscores).
 * lc_lucir_loss          : `LUCIR.observe` task>0 (lucir.py:184-205): cur_lamda * CosineEmbeddingLoss(feat, ref_feat, +1) + CE(logits, y) +
 *                          lw_mr * MarginRankingLoss(gt score, top-K novel scores of old-class samples, margin); gradients w.r.t. logits,
 *                          pre-scale scores and features; argmax / #correct.  scal: [0] loss [1] #correct [2] ce [3] less-forget [5] mr.
 * lc_l2p_select          : `L2P.forward` (prompt.py:369-406): cosine similarity, per-sample top-k, batch-wide majority vote (ties: count
 *                          desc then id asc), pull-constraint `reduce_sim` and its gradient w.r.t. prompt_key.  scratch >= dim floats.
 * lc_l2p_gather          : batched_prompt[b][t*len + l][:] = prompt[ids[t]][l][:].
 * lc_gpm_project         : grad <- grad - (grad.view(rows, dim) @ proj)  (gpm.py:78-81; proj = U U^T, [dim][dim]), in place.
 * lc_lora_merge_qkv      : W' = cat(W_q, W_k + B_k A_k, W_v + B_v A_v)  (transformer.py:246-254).  lc_lora_bgrad: dB = dW' A^T.
 * ------------------------------------------------------------------------------------------------------------------- */
int lc_cosine_head_forward(const float* feat, const float* W, const float* sigma, int batch, int ncls, int feat_dim, float* inv_norm,
                           float* scores, float* logits, int ld, lc_stream_t stream);
int lc_cosine_head_backward(const float* gscores, int ld, const float* feat, const float* W, const float* inv_norm, int batch, int ncls,
                            int feat_dim, float* dfeat, float* dW, lc_stream_t stream);
int lc_lucir_loss(const float* logits, const float* scores, int ld, const float* feat, const float* ref_feat, int feat_dim, const int64_t* y,
                  int batch, int ncls, int num_old, int K, float cur_lamda, float margin, float lw_mr, float* dlogits, float* dscores,
                  float* dfeat, int64_t* pred, float* scal, float* dsigma /* nullable: sum(dlogits*scores) */, lc_stream_t stream);
int lc_l2p_select(const float* query, const float* key, int batch, int pool, int dim, int top_k, float* sim, int64_t* ids, int* hist,
                  float* reduce_sim, float* dkey, float* scratch, lc_stream_t stream);
/* The same selection in two phases for data parallelism: phase 1 = similarities + per-sample top-k -> `hist` (this rank's counts); the caller SUM-all-reduces
 * `hist` over the ranks; phase 2 = majority vote on that histogram, `reduce_sim` and `dkey` of this rank's samples for the voted ids (prompt.py:380-401 on the
 * GLOBAL batch).  phase 0 = lc_l2p_select. */
int lc_l2p_select_phase(const float* query, const float* key, int batch, int pool, int dim, int top_k, float* sim, int64_t* ids, int* hist,
                        float* reduce_sim, float* dkey, float* scratch, int phase, lc_stream_t stream);
int lc_l2p_gather(const float* prompt, const int64_t* ids, float* out, int batch, int top_k, int length, int dim, lc_stream_t stream);
int lc_gpm_project(float* grad, const float* proj, int rows, int dim, lc_stream_t stream);
int lc_lora_merge_qkv(const float* qkv_w, const float* A_k, const float* B_k, const float* A_v, const float* B_v, float* out, int dim, int rank,
                      lc_stream_t stream);
int lc_lora_bgrad(const float* dW, const float* A, float* dB, int dim, int rank, lc_stream_t stream);
/* iCaRL exemplar management.
 * lc_herding_select : greedy herding of `LinearHerdingBuffer.herding_select` (linearherdingbuffer.py:133-163) over L2-normalised
 *                     features sorted by class (class c = rows [cls_begin[c], cls_begin[c+1])): per class `per_class` picks of
 *                     argmin || mean - (f + running_sum)/(i+1) ||, the picked row pushed away by +1e6; out[c][i] = global row
 *                     index (-1 if the class is smaller).  work: scratch of the same size as feats.
 * lc_ncm_classify   : `ICarl.NCM_classify` (icarl.py:122-152): argmin_c ||feat - mean_c||^2 (first minimum). */
int lc_herding_select(const float* feats, const int* cls_begin, int ncls, int dim, int per_class, float* work, int64_t* out, lc_stream_t stream);
int lc_ncm_classify(const float* feat, const float* means, int batch, int ncls, int dim, int64_t* pred, lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * ViT-B/16 building blocks (core/model/backbone/transformer.py:169-197 MultiHeadAttention, :1255-1273 Mlp, :1331-1336 block).
 * lc_gemm_bf16 : C[b][M][N] = alpha * A[b][M][K] . B[b][N][K]^T (+ bias[N]) (+ residual fp32) on tcgen05 (BF16 operands fetched by TMA
 *                into 128B-swizzled smem, fp32 accumulation in TMEM).  A/B/C are bf16 unless out_f32; lda/ldb/ldc/ldr and the batch
 *                strides are in elements (row strides of A and B must be multiples of 8).  out2 (nullable, bf16): GELU(C), exact erf
 *                form (nn.GELU), in which case C keeps the pre-activation.  Replaces nn.Linear / F.linear / torch.matmul.
 * ------------------------------------------------------------------------------------------------------------------- */
int lc_gemm_bf16(const void* A, int lda, long long strideA, const void* B, int ldb, long long strideB, void* C, int ldc, long long strideC, int M, int N,
                 int K, int batch, const float* bias, const float* residual, int ldr, long long strideR, void* out2, int out_f32, float alpha,
                 int* error_flag, lc_stream_t stream);
/* Same GEMM with a two-level batch (z = z_out * batch_in + z_in; every operand has an inner and an outer batch stride; a stride of 0 shares the operand across that batch level): the per-head
 * attention GEMMs (Q K^T, P V) read strided views of the fused QKV buffer [B][T][3][H][64] and write [B][T][H*64] without copies. */
typedef struct lc_gemm_desc {
    const void* A; long long lda, strideA_in, strideA_out;
    const void* B; long long ldb, strideB_in, strideB_out;
    void* C; long long ldc, strideC_in, strideC_out;
    const float* bias; const float* residual; long long ldr, strideR_in, strideR_out;
    void* out2;
    const void* gelu_bwd_aux;   /* nullable bf16 indexed like C: result *= GELU'(aux) */
    int M, N, K, batch_in, batch_out, out_f32;
    float alpha;
    int gelu_mode;              /* bit 0: the GELU side tensor is GELU'(pre-activation): with out2, C stores it instead of the pre-activation; with
                                 * gelu_bwd_aux, aux is a plain multiplier.  bit 1 (with out2): C is not stored at all (no-grad passes).  0 = as above. */
    int ksplit;                 /* > 1: split-K — the K blocks are divided over ksplit fp32 partial outputs, partial s at C + s * strideC_split
                                 * (out_f32 must be 1, no epilogue extras; every split gets at least one K block: ksplit <= ceil(K / 64)) */
    long long strideC_split;
} lc_gemm_desc;
int lc_gemm_bf16_ex(const lc_gemm_desc* desc, int* error_flag, lc_stream_t stream);

/* Implicit-GEMM convolution on the same kernel (F.conv2d of resnet.py:48-64 BasicBlock / :133-150 stem, and its stride-1 data gradient):
 *   Y[(n, ho, wo)][co] = sum_{kh, kw, c} X[n][ho*stride - pad + kh][wo*stride - pad + kw][c] * Wk[co][(kh*ks + kw)*C + c]  (+ bias, + residual)
 * X: BF16 NHWC, C % 64 == 0; Wk: BF16 [Cout][ks*ks*C] (lc_nn_pack_weight, tap-major); Y: fp32 or BF16 [N*Ho*Wo][ldc].  The producer warp reads X
 * through a rank-4 tensor map {C, W, H, N} whose box is one 128-pixel output tile; the filter tap is a coordinate offset of the box and the zero
 * padding is TMA's out-of-bounds fill, so no im2col matrix is ever written.  Wo must divide 128 and Ho*Wo must divide or be a multiple of 128. */
typedef struct lc_conv_desc {
    const void* X; const void* Wk; void* Y;
    const float* bias; const float* residual; long long ldc, ldr;
    int N, H, W, C, Cout, ks, stride, pad, Ho, Wo, out_f32;
} lc_conv_desc;
int lc_conv_gemm_bf16(const lc_conv_desc* desc, int* error_flag, lc_stream_t stream);
/* Row-wise / layout kernels of the ViT forward (transformer.py:2222-2261): im2col of the 16x16 patches (timm PatchEmbed as a GEMM), cls-row
 * assembly, LayerNorm (fp32 in -> bf16 and/or fp32 out, optional (mean, rstd) per row), mean over a row range (L2P: the prompt positions, transformer.py:2256-2259), fp32 linear head,
 * fp32 -> bf16 cast. */
/* Fused multi-head self-attention, head dim 64, T <= 256 tokens (transformer.py:169-197): O = softmax(Q K^T / 8) V on strided views of the
 * fused QKV buffer [B][T][3][H][64] (BF16) -> O [B][T][H*64] (BF16); lse2 [B][H][T] = base-2 log-sum-exp of the scaled score rows, kept for
 * the backward.  The T x T score / probability matrices live in TMEM / shared memory only. */
int lc_attn_forward(const void* qkv_bf16, void* out_bf16, float* lse2, int batch, int T, int heads, int* error_flag, lc_stream_t stream);
/* Its backward: dQ, dK, dV (written into dqkv with the QKV layout) from dO, with the probabilities recomputed on chip from lse2;
 * out_bf16 / rowdot are not read any more (the softmax row term sum_j P_j dP_j is formed on chip from the same probabilities the products use). */
int lc_attn_backward(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2, float* rowdot, void* dqkv_bf16, int batch, int T,
                     int heads, int* error_flag, lc_stream_t stream);
/* The same two kernels with P prefix keys / values per image placed in front of the token keys — `MultiHeadAttention.forward(prompt=(pk, pv))`
 * (transformer.py:175-180; DualPrompt prompt.py:299-316, CodaPrompt prompt.py:199-201): pk / pv BF16 [B][P][H*64], T + P <= 256.  The backward
 * returns the prefix-row gradients dpk / dpv as fp32 [B][P][H*64] (they are prompt-pool parameters) next to dqkv. */
int lc_attn_forward_prefix(const void* qkv_bf16, void* out_bf16, float* lse2, int batch, int T, int heads, const void* pk_bf16, const void* pv_bf16, int P,
                            int* error_flag, lc_stream_t stream);
int lc_attn_backward_prefix(const void* qkv_bf16, const void* dout_bf16, const float* lse2, void* dqkv_bf16, int batch, int T, int heads, const void* pk_bf16,
                             const void* pv_bf16, float* dpk, float* dpv, int P, int* error_flag, lc_stream_t stream);
int lc_vit_patchify(const float* img_nchw, void* out_bf16, int batch, lc_stream_t stream);
int lc_vit_set_rows(float* x, long long batch_stride, int batch, int row0, int nrows, const float* src, const float* add, int dim, lc_stream_t stream);
int lc_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps, long long rows, int dim, void* out_bf16, float* out_f32,
                         float* stat, lc_stream_t stream);
/* Backward of the frozen backbone wrt its input tokens (what carries the loss to the L2P prompt rows; the backbone itself has
 * requires_grad=False: l2p.py:66-71): LayerNorm backward fused with the residual-path gradient (dh either dense fp32 or the pooled-feature
 * gradient broadcast over the first n_active rows of every image) and the batch sum of the shared prompt rows. */
int lc_layernorm_backward(const float* dh, const float* dh_pool, int T, int n_active, const float* x, const float* gamma, float eps, long long rows, int dim,
                          const float* res, float* out_f32, void* out_bf16, lc_stream_t stream);
int lc_sum_batch_rows(const float* x, long long batch_stride, int batch, int nrows, int dim, float* out, lc_stream_t stream);
int lc_vit_pool_rows(const float* y, long long batch_stride, int batch, int r0, int nr, int dim, float* feat, lc_stream_t stream);
int lc_linear_head(const float* feat, const float* W, const float* bias, int batch, int ncls, int dim, float* logits, int ld, lc_stream_t stream);
int lc_cast_bf16(const float* in, void* out_bf16, long long n, lc_stream_t stream);
/* nn.Linear classifier backward on a wide feature (l2p.py:31-40), and the L2P parameter gradients: the prompt-row gradient scattered into the
 * pool layout [pool][length][dim] (zero elsewhere) and dkey_out = coeff * dkey_in. */
int lc_linear_head_backward(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int batch, int dim, float* dW, float* db,
                            float* dfeat, lc_stream_t stream);
int lc_l2p_backward(const float* dprompts, const int64_t* ids, int pool, int top_k, int length, int dim, float* dpool, const float* dkey_in, float coeff,
                    float* dkey_out, lc_stream_t stream);

/* Low-rank adapters on the fused QKV projection (transformer.py:199-274 MultiHeadAttention_LoRA, :276-357 MultiHeadAttention_SDLoRA).
 * lc_lora_merge      : for every layer l and adapted slab s (bit s of slab_mask: 0 = q, 1 = k, 2 = v; arrays hold the adapted slabs in increasing
 *                      order): W'[l][s] = W[l][s] + B[l][s] diag(scale[l][s]) A[l][s]   (A [rank][dim], B [dim][rank], scale nullable = 1), written
 *                      as the BF16 GEMM operands wb [l][3 dim][dim] / wbt [l][dim][3 dim] (either nullable) and/or fp32 w_out (may alias w:
 *                      `merge_weight`, transformer.py:237-242).  Replaces the per-forward `k_weight + lora_B_k.weight @ lora_A_k.weight` (:246-254).
 * lc_lora_bgrad_rows : out[s][c][j] = sum_n X[n][x0 + s*x_slab_stride + c] * Z[n][s*rank + j]  — the adapter gradient in rank form, e.g.
 *                      d lora_B_k.weight = dK^T (h A_k^T) with X = d(qkv) (BF16) and Z = the saved down-projection (fp32); deterministic
 *                      two-stage sum over `nchunk` row chunks (partial >= lc_lora_bgrad_partial_floats floats).  rank <= 16, dim % 768 == 0. */
/* DualPrompt's key-query match over the e-prompt layers (prompt.py:270-292): keys[l] / dkeys[l] are HOST arrays of nlayers DEVICE pointers to the
 * [pool][768] key matrices and their gradients.  task_id >= 0 (training, task-id bootstrap): idx[l][b] = task_id, *loss = sum_l sum_b (1 - cos(q_b,
 * K_l[task_id])), dkeys[l][task_id] = its gradient (the query is detached).  task_id < 0 (inference): idx[l][b] = argmax_k cos(q_b, K_l[k]).
 * lc_gather_rows_bf16: out[b][r][:] = bf16(src[idx[b] * idx_stride + r * dim + :]) (idx nullable = 0): the selected prompt halves as per-image prefix rows. */
int lc_prompt_key_match(const float* query, const float* const* keys, float* const* dkeys, int nlayers, int batch, int pool, int dim, int task_id,
                        int64_t* idx, float* loss, lc_stream_t stream);
/* CodaPrompt's attention-weighted prompt (prompt.py:146-201) for `nlayers` blocks at once; K / A / p / pk / pv (and the gradient arrays) are HOST arrays of
 * DEVICE pointers, one per block: K, A [pool][768], p [pool][length][768] (the first nk components are used), outputs pk / pv BF16 [batch][length/2][768]:
 *   alpha[l][b][k] = cos(q_b * A_k, K_k),  P_[b] = sum_k alpha[l][b][k] p[k],  pk = P_[:, :length/2], pv = P_[:, length/2:]
 * alpha / vnorm ([nlayers][batch][nk], vnorm = |q_b * A_k|) are kept for the backward, which maps the prefix-row gradients dpk / dpv (fp32, from
 * lc_attn_backward_prefix) to dK, dA ([pool][768], rows < nk) and dp ([pool][length][768], components < nk); dalpha is scratch of the same size. */
int lc_coda_prompt_forward(const float* query, const float* const* K, const float* const* A, const float* const* p, void* const* pk_bf16, void* const* pv_bf16,
                           int nlayers, int batch, int nk, int length, int dim, float* alpha, float* vnorm, lc_stream_t stream);
int lc_coda_prompt_backward(const float* query, const float* const* K, const float* const* A, const float* const* p, const float* const* dpk, const float* const* dpv,
                            float* const* dK, float* const* dA, float* const* dp, int nlayers, int batch, int nk, int length, int dim, const float* alpha,
                            const float* vnorm, float* dalpha, lc_stream_t stream);
int lc_gather_rows_bf16(const float* src, const int64_t* idx, long long idx_stride, int rows, int dim, int batch, void* out_bf16, lc_stream_t stream);
/* GPM's gradient projection (gpm.py:78-81) on the tensor cores at fp32-level accuracy: grad <- grad - grad.view(rows, dim) @ proj for a SYMMETRIC proj
 * (= U U^T, gpm.py:124) given as its two-term BF16 split (lc_split_bf16: x = hi + lo): three BF16 tcgen05 GEMMs (hi hi + lo hi + hi lo, fp32 accumulation,
 * applied in place through the residual epilogue) reproduce the fp32 product to ~1e-5 relative.  g_hi / g_lo: BF16 scratch [rows][dim]. */
int lc_split_bf16(const float* x, void* hi_bf16, void* lo_bf16, long long n, lc_stream_t stream);
int lc_gpm_project_tc(float* grad, const void* proj_hi_bf16, const void* proj_lo_bf16, int rows, int dim, void* g_hi_bf16, void* g_lo_bf16, int* error_flag,
                      lc_stream_t stream);
/* out[c][r] = in[r][c] (BF16; columns [rows, ld_out) of out zero-filled): the transposed copy that turns a contraction over token rows
 * (InfLoRA's input matrix sum_n h_n h_n^T, transformer.py:242-244) into the K-major operands of lc_gemm_bf16. */
int lc_transpose_bf16(const void* in_bf16, long long ld_in, long long rows, int cols, void* out_bf16, long long ld_out, lc_stream_t stream);
/* lc_rowouter_bf16   : the general form of lc_lora_bgrad_rows: out[s][c][j] = scale * sum_n X[n][x0 + s*x_slab_stride + c] * Z[n][z0 + s*z_slab_stride + j];
 *                      `transposed` writes out[s][j][c] instead (a lora_A gradient d A = (dY B)^T h with X = h, Z = dY B); scale_dev: nullable DEVICE scalar
 *                      (SD-LoRA's trainable magnitude, sd_lora.py:122-125).
 * lc_coldot_accumulate: dmag[i] += sum_{jj < rank} col_weight[i*rank + jj] * sum_n G[n][i*rank + jj] * Z[n][z0 + i*rank + jj], i < cols / rank — the gradient of
 *                      the per-adapter magnitudes of MultiHeadAttention_SDLoRA (transformer.py:312-332) from G = dY B_cat and Z = h A_cat^T; accumulates
 *                      (zero dmag once per step).  partial >= nchunk * cols floats. */
int lc_rowouter_bf16(const void* x_bf16, long long ldx, int x0, int x_slab_stride, int nslab, int dim, const float* z, int ldz, int z0, int z_slab_stride, int rank,
                     long long rows, float* partial, int nchunk, float* out, int transposed, const float* scale_dev, lc_stream_t stream);
int lc_coldot_accumulate(const float* g, int ldg, const float* z, int ldz, int z0, int cols, int rank, long long rows, const float* col_weight, float* partial, int nchunk,
                         float* dmag, lc_stream_t stream);
int lc_lora_merge(const float* w, const float* A, const float* B, const float* scale, int slab_mask, int layers, int dim, int rank, void* wb_bf16,
                  void* wbt_bf16, float* w_out, lc_stream_t stream);
long long lc_lora_bgrad_partial_floats(int nslab, int dim, int rank, int nchunk);
int lc_lora_bgrad_rows(const void* x_bf16, long long ldx, int x0, int x_slab_stride, int nslab, int dim, const float* z, int ldz, int rank, long long rows, float* partial,
                       int nchunk, float* out, lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Per-kernel entry points (unit-tested individually; the network-level calls above are compositions of these).
 * conv3x3: NHWC fp32, pad 1.  `w_oihw` is the native nn.Conv2d weight; mode 0 = forward, 1 = data gradient (input is dy).
 * Supported (cin, cout, width_out, stride): the CifarResNet layer shapes.  `in_nchw` != 0: input is NCHW (network stem).
 * Optional prologue relu(in*pro_scale+pro_shift) (per input channel), optional addend, optional BN statistics:
 * stat_out = {scale[C], shift[C], mean[C], invstd[C]} and running-stat update (rstat = {mean[C], var[C]}, nullable).
 * scratch: >= lc_conv_scratch_floats(...) floats, first 64 words zero.
 * ------------------------------------------------------------------------------------------------------------------- */
long long lc_conv_scratch_floats(int batch, int cin, int cout, int width_out);
int lc_conv3x3(const float* in, const float* w_oihw, float* out, int batch, int cin, int cout, int width_out, int stride, int mode,
               int in_nchw, const float* pro_scale, const float* pro_shift, const float* addend, const float* gamma,
               const float* beta, float* rstat, float* stat_out, float* scratch, lc_stream_t stream);
/* One launch of the forward conv kernel on pre-packed weights [cin][9][cout] (lc_conv3x3 leaves them at scratch+96). */
int lc_conv3x3_packed(const float* in, const float* wpack, float* out, int batch, int cin, int cout, int width_out, int stride,
                      const float* pro_scale, const float* pro_shift, lc_stream_t stream);
/* tcgen05 (kind::tf32, TMEM accumulators) version for the square stride-1 layers (c, width) in {(16,32),(32,16),(64,8)};
 * same prologue / addend / statistics options; scratch >= lc_conv_tc_scratch_floats, first 64 words zero; scratch word 8 (int)
 * is set to 1 if the MMA completion barrier timed out. */
long long lc_conv_tc_scratch_floats(int batch, int c, int width);
int lc_conv3x3_tc(const float* in, const float* w_oihw, float* out, int batch, int c, int width, int mode, const float* pro_scale,
                  const float* pro_shift, const float* addend, const float* gamma, const float* beta, float* rstat, float* stat_out,
                  float* scratch, lc_stream_t stream);
/* tcgen05 forward of the stride-2 3x3 convs at the stage transitions (cin -> 2*cin; (cin, width_out) in {(16,16),(32,8)}; parity-plane implicit
 * GEMM, csrc/conv_s2_tc.cuh) — replaces the cuDNN fprop of core/model/backbone/resnet.py:341-343 at stride 2.  Optional BatchNorm statistics of the
 * output (gamma / beta / stat_out as lc_conv3x3).  scratch >= lc_conv_scratch_floats(batch, cin, 2*cin, width_out), first 64 words zero. */
int lc_conv3x3s2_tc(const float* in, const float* w_oihw, float* out, int batch, int cin, int width_out, const float* gamma, const float* beta,
                    float* rstat, float* stat_out, float* scratch, lc_stream_t stream);
/* Data gradient of the same stride-2 conv on tcgen05 (one staged dy tile, nine tap groups into four parity-plane accumulators; replaces cuDNN's
 * dgrad, i.e. autograd of resnet.py:341-343 at stride 2): dy [B][wo][wo][2*cin] -> dx [B][2*wo][2*wo][cin]. */
int lc_conv3x3s2_dgrad_tc(const float* dy, const float* w_oihw, float* dx, int batch, int cin, int width_out, float* scratch, lc_stream_t stream);
/* One launch of the tensor-core conv on pre-packed weights (lc_conv3x3_tc leaves the forward packing at scratch+96+2*9*c*c). */
int lc_conv3x3_tc_packed(const float* in, const float* wtc, float* out, int batch, int c, int width, const float* pro_scale,
                         const float* pro_shift, int* error_flag, lc_stream_t stream);
int lc_conv3x3_wgrad(const float* in, const float* dy, float* dw_oihw, int batch, int cin, int cout, int width_out, int stride,
                     int in_nchw, const float* pro_scale, const float* pro_shift, float* scratch, lc_stream_t stream);
/* tcgen05 (kind::tf32, MN-major operands, TMEM accumulators) weight gradient for (c, width) in {(16,32),(32,16),(64,8)}. */
int lc_conv3x3_wgrad_tc(const float* in, const float* dy, float* dw_oihw, int batch, int c, int width, const float* pro_scale,
                        const float* pro_shift, float* scratch, lc_stream_t stream);
/* 1x1 stride-2 shortcut conv: mode 0 forward (+stats as above), 1 data gradient ACCUMULATED into `out` (shape of the conv
 * input), 2 weight gradient into `out` ([cout][cin]). */
int lc_conv1x1s2(const float* a, const float* b, float* out, int batch, int cin, int cout, int width_out, int mode, const float* gamma,
                 const float* beta, float* rstat, float* stat_out, float* scratch, lc_stream_t stream);
/* out = relu(y*scale+shift (+ res | + res*res_scale+res_shift)) over NHWC [npix][C]. */
int lc_bn_act_forward(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale,
                      const float* res_shift, float* out, long long npix, int C, lc_stream_t stream);
/* BatchNorm(+ReLU) backward. mask_mode 0: none, 1: g *= (mask_src > 0), 2: g *= (y*scale+shift > 0).  stat = {scale, shift,
 * mean, invstd}.  Writes dy, dgamma[C], dbeta[C]; g_out (nullable) receives the masked g.  scratch >= 592*2*C + 3*C + 64. */
int lc_bn_backward(const float* g, const float* mask_src, int mask_mode, const float* y, const float* stat, float* dy, float* g_out,
                   float* dgamma, float* dbeta, long long npix, int C, float* scratch, lc_stream_t stream);


/* ---- generic layer kernels for ResNet18 (resnet.py:26-64,110-246; LwF lwf.py:52-70) and AlexNet_TRGP (alexnet.py:94-156; GPM gpm.py:45-204) --------
 * Activations NHWC; a conv output is the row-major matrix [M = N*Ho*Wo][Cout] the GEMM epilogue writes.  korder: 0 = k = (kh*ks + kw)*C + c (tap-major,
 * what lc_conv_gemm_bf16 contracts over), 1 = k = (c*ks + kh)*ks + kw (= weight.view(Cout, -1), the order of GPM's bases, gpm.py:79,163-168).
 * src_kind: 0 NHWC bf16, 1 NCHW fp32 (network input), 2 NHWC fp32. */
long long lc_nn_bn_scratch_floats(int C);
/* Patch matrix of F.conv2d's input: col [M][ld_col] and / or its transpose colT [Kp][ld_colT] (BF16; columns / rows beyond K and M are zero).  Also GPM's
 * representation matrix (the Python triple loop of gpm.py:157-168) when called with korder = 1. */
int lc_nn_im2col(const void* src, int src_kind, int N, int H, int W, int C, int ks, int stride, int pad, int korder, void* col_bf16, long long ld_col,
                 void* colT_bf16, long long ld_colT, int Kp, lc_stream_t stream);
/* The same patch matrix in fp32 (K unpadded): GPM's representation matrices feed an SVD (gpm.py:157-168, 170-204). */
int lc_nn_im2col_f32(const void* src, int src_kind, int N, int H, int W, int C, int ks, int stride, int pad, int korder, float* col, long long ld_col, float* colT,
                     long long ld_colT, lc_stream_t stream);
/* Data gradient of a conv from dcol = dY * W ([M][ld] BF16): dx (fp32 NHWC) = addend + fold(dcol). */
int lc_nn_col2im(const void* dcol_bf16, long long ld, const float* addend, float* dx, int N, int H, int W, int C, int ks, int stride, int pad, int korder,
                 lc_stream_t stream);
/* nn.BatchNorm2d / 1d in train mode over y [M][C] (C % 64 == 0): aff = [scale | shift | mean | invstd]; running = [mean | var] updated when non-null
 * (momentum, unbiased variance); gamma / beta nullable.  scratch >= lc_nn_bn_scratch_floats(C).  Deterministic (fixed-order fp64 finalize). */
int lc_nn_bn_stats(const float* y, long long M, int C, const float* gamma, const float* beta, float eps, float momentum, float* running, float* aff,
                   float* scratch, lc_stream_t stream);
int lc_nn_bn_eval_affine(const float* running, int C, const float* gamma, const float* beta, float eps, float* aff, lc_stream_t stream);
/* out = dropout(relu(y*scale + shift + [res | res*res_scale + res_shift])): BF16 and / or fp32.  Dropout (alexnet.py:130,136,142,149,154): keep mask =
 * counter-based hash of (rng[0] = seed, rng[1] = step, rng_stream = layer, element index), survivors scaled by 1/(1-p); rng == NULL or p == 0: off. */
int lc_nn_bn_act(const float* y, const float* aff, const float* res, const float* res_aff, long long M, int C, int relu, float drop_p,
                 const unsigned long long* rng, int rng_stream, void* out_bf16, float* out_f32, lc_stream_t stream);
int lc_nn_dropout_mask(const unsigned long long* rng, int rng_stream, float drop_p, long long n, unsigned char* keep, lc_stream_t stream);
int lc_nn_rng_advance(unsigned long long* rng, lc_stream_t stream);
/* native_batch_norm_backward + threshold_backward (+ dropout): dz = g * [act > 0] * gscale (act = the stored layer output; NULL, NULL: no mask);
 * dgamma = sum dz*xhat, dbeta = sum dz; dy = scale*(dz - dbeta/M - xhat*dgamma/M) as BF16 and / or fp32; dz_out (nullable) = dz. */
int lc_nn_bn_backward(const float* g, const float* act_f32, const void* act_bf16, float gscale, const float* y, const float* aff, long long M, int C,
                      float* dgamma, float* dbeta, void* dy_bf16, float* dy_f32, float* dz_out, float* scratch, lc_stream_t stream);
/* nn.MaxPool2d(k, stride, pad) on fp32 NHWC; idx keeps the window position of the (first) maximum for the backward. */
int lc_nn_maxpool_forward(const float* in, int N, int H, int W, int C, int k, int stride, int pad, float* out_f32, void* out_bf16, unsigned char* idx,
                          lc_stream_t stream);
int lc_nn_maxpool_backward(const float* g, const unsigned char* idx, int N, int H, int W, int C, int k, int stride, int pad, float* dx, lc_stream_t stream);
/* nn.AdaptiveAvgPool2d((1, 1)) over fp32 NHWC [N][HW][C] and its backward. */
int lc_nn_avgpool_forward(const float* in, int N, int HW, int C, float* out, lc_stream_t stream);
int lc_nn_avgpool_backward(const float* dfeat, int N, int HW, int C, float* dx, lc_stream_t stream);
/* OIHW fp32 weights -> BF16 GEMM operands.  mode 0: [Cout][K] (korder); 1: [K][Cout] (B operand of dcol = dY * W); 2: [Cin][flipped taps][Cout]
 * (stride-1 data gradient as a convolution of dY).  Row length ld, tail zero. */
int lc_nn_pack_weight(const float* w, int Cout, int Cin, int ks, int korder, int mode, void* out_bf16, long long ld, lc_stream_t stream);
/* dW (OIHW fp32) = sum of nsplit split-K partials [nsplit][Cout][ldp] (columns in korder), fixed order. */
int lc_nn_wgrad_reduce(const float* partial, int nsplit, int Cout, int Cin, int ks, int korder, long long ldp, float* dw, lc_stream_t stream);
/* fp32 [rows][cols] -> BF16 [rows][ld] and / or its transpose [cols][ldT] (tails zero). */
int lc_nn_cast_transpose(const float* src, long long rows, int cols, void* out_bf16, long long ld, void* outT_bf16, long long ldT, lc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Input pipeline and evaluation meter on the device.  The uint8 dataset [N][H][W][3] is resident in HBM; one call per batch
 * gathers `idx`, applies the reference's transform chain with PIL's / torchvision's own arithmetic and writes the fp32 NCHW batch.  The random draws
 * come from the host (libcontinual_b200/data.py restates torchvision's get_params).
 * lc_augment_cifar_u8 : RandomCrop(H, padding=pad) -> RandomHorizontalFlip -> ColorJitter(brightness) -> ToTensor -> Normalize (core/data/data.py:11-16);
 *                       draw[b] = {dx, dy, flip, 0} with (dx, dy) the crop origin inside the zero-padded image, bright[b] the PIL brightness factor
 *                       (NULL = 1).  Identity draws {pad, pad, 0} = the test transform (data.py:18).
 * lc_resize_crop_u8   : RandomResizedCrop(out) / Resize (+CenterCrop) -> flip -> ToTensor -> Normalize (data.py:27-36), bit-exact to PIL's two-pass
 *                       fixed-point bilinear resize.  draw[b] = {top, left, h, w, OH, OW, oy, ox}: crop box, resized size, origin of the out x out window.
 *                       scratch: lc_resize_scratch_bytes(batch, H, out) bytes, 16-byte aligned.
 * lc_eval_meter       : counts[t] += {#correct, #seen} per task (Trainer._validate, core/trainer.py:616-720): task >= 0 -> every sample to row `task`;
 *                       task < 0 -> the row whose class range [bounds[t], bounds[t+1]) holds the label.  counts: uint64 [ntask][2].
 * ------------------------------------------------------------------------------------------------------------------- */
int lc_augment_cifar_u8(const uint8_t* src, const int64_t* idx, const int* draw, const float* bright, float* out, int batch, int H, int W, int pad,
                        const float* mean3, const float* std3, lc_stream_t stream);
long long lc_resize_scratch_bytes(int batch, int H, int out);
int lc_resize_crop_u8(const uint8_t* src, const int64_t* idx, const int* draw, const int* flip, float* out, int batch, int H, int W, int out_size,
                      const float* mean3, const float* std3, void* scratch, lc_stream_t stream);
int lc_eval_meter(const int64_t* pred, const int64_t* label, int n, const int* bounds, int ntask, int task, unsigned long long* counts, lc_stream_t stream);
/* total[t] += batch[t] and batch[t] = 0.  reference_rounding = 1 adds int(correct / seen * seen) in double, i.e. `int(acc * batch_size)` (trainer.py:644). */
int lc_eval_fold(unsigned long long* batch_counts, unsigned long long* total, int ntask, int reference_rounding, lc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LC_B200_H */
