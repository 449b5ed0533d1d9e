// Continual-learning specific small kernels (all fp32, HBM / latency bound, warp-shuffle reductions, fixed summation order):
//   cosine_head_fwd / cosine_head_bwd : CosineLinear / SplitCosineLinear (core/model/backbone/resnet.py:418-463) + autograd
//   lucir_loss                        : LUCIR.observe task>0 losses (core/model/lucir.py:184-205): less-forget cosine embedding,
//                                       CE, margin ranking over the top-K novel scores of old-class samples, and their gradients
//   l2p_select / l2p_gather           : L2P prompt pool (core/model/backbone/prompt.py:369-406): key/query cosine similarity,
//                                       per-sample top-k, batch-wide majority vote, pull-constraint loss and its key gradient
//   gpm_project                       : g <- g - (g.view(R,-1) @ M)   (core/model/gpm.py:78-81)
//   lora_merge_qkv / lora_bgrad       : W' = cat(Wq, Wk + Bk Ak, Wv + Bv Av) (core/model/backbone/transformer.py:246-254) and
//                                       dB = dW' A^T (the rank-r gradient autograd derives from it)
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace lc {

// ---------------------------------------------------------------------------------------------------------------------
// cosine head.  One warp per (sample | class) row for the norms, then one thread per logit.
// ---------------------------------------------------------------------------------------------------------------------
// inv_norm[r] = 1 / max(||x_r||, 1e-12)   for the B feature rows followed by the C weight rows
template <int D>
__global__ void __launch_bounds__(128) row_inv_norm_kernel(const float* feat, int B, const float* W, int C, float* inv_norm) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B + C) return;
    const float* x = row < B ? feat + (size_t)row * D : W + (size_t)(row - B) * D;
    float s = 0.f;
#pragma unroll
    for (int j = lane; j < D; j += 32) s = fmaf(x[j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) inv_norm[row] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}

// scores[b][c] = cos(feat_b, W_c) ; logits = sigma * scores
template <int D>
__global__ void __launch_bounds__(128) cosine_head_fwd_kernel(const float* feat, const float* W, const float* inv_norm, const float* sigma, int B, int C,
                                                              float* scores, float* logits, int ld) {
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= B * C) return;
    const int b = idx / C, c = idx % C;
    const float* f = feat + (size_t)b * D;
    const float* w = W + (size_t)c * D;
    float d = 0.f;
#pragma unroll 8
    for (int j = 0; j < D; ++j) d = fmaf(f[j], __ldg(w + j), d);
    const float s = d * inv_norm[b] * inv_norm[B + c];
    scores[(size_t)b * ld + c] = s;
    logits[(size_t)b * ld + c] = (sigma != nullptr ? *sigma : 1.f) * s;
}

// gs[b][c] = d(loss)/d(scores) (sigma already folded in by the caller: gs = sigma*dlogits + dscores).
// blocks [0,B): dfeat rows ; blocks [B, B+C): dW rows.  Normalisation backward: dx = (dxhat - xhat (xhat . dxhat)) / ||x||.
template <int D>
__global__ void __launch_bounds__(D) cosine_head_bwd_kernel(const float* gs, int ld, const float* feat, const float* W, const float* inv_norm, int B, int C,
                                                            float* dfeat, float* dW) {
    __shared__ float s_red[D / 32];
    const int j = threadIdx.x;
    const bool is_feat = (int)blockIdx.x < B;
    const int r = is_feat ? blockIdx.x : blockIdx.x - B;
    float acc = 0.f;     // d(xhat)[j]
    if (is_feat) {
        for (int c = 0; c < C; ++c) acc = fmaf(gs[(size_t)r * ld + c], __ldg(W + (size_t)c * D + j) * inv_norm[B + c], acc);
    } else {
        for (int b = 0; b < B; ++b) acc = fmaf(gs[(size_t)b * ld + r], __ldg(feat + (size_t)b * D + j) * inv_norm[b], acc);
    }
    const float inv = inv_norm[is_feat ? r : B + r];
    const float xhat = (is_feat ? feat[(size_t)r * D + j] : W[(size_t)r * D + j]) * inv;
    float dot = warp_sum(xhat * acc);
    if ((j & 31) == 0) s_red[j >> 5] = dot;
    __syncthreads();
    dot = 0.f;
#pragma unroll
    for (int k = 0; k < D / 32; ++k) dot += s_red[k];
    const float out = (acc - xhat * dot) * inv;
    if (is_feat) dfeat[(size_t)r * D + j] = out; else dW[(size_t)r * D + j] = out;
}

// ---------------------------------------------------------------------------------------------------------------------
// LUCIR losses (one thread per sample, single CTA; B <= 1024)
// ---------------------------------------------------------------------------------------------------------------------
struct LucirArgs {
    const float* logits;     // [B][ld]  sigma * scores
    const float* scores;     // [B][ld]  before scale
    const float* feat;       // [B][D]   student features (classifier input)
    const float* ref_feat;   // [B][D]   frozen reference model features
    const long long* y;
    float* dlogits;          // [B][ld]  out: d(loss)/d(logits) from the CE term
    float* dscores;          // [B][ld]  out: d(loss)/d(scores) from the margin-ranking term
    float* dfeat;            // [B][D]   out: d(loss)/d(feat) from the less-forget term
    long long* pred;
    float* scal;             // [0] loss [1] #correct [2] ce [3] less-forget (x lamda) [5] margin ranking (x lw_mr)
    float* dsigma;           // nullable: d(loss)/d(sigma) = sum(dlogits * scores)   (logits = sigma * scores)
    int B, C, ld, D, num_old, K;
    float cur_lamda, margin, lw_mr;
};

__global__ void __launch_bounds__(256) lucir_loss_kernel(LucirArgs a) {
    __shared__ float s_a[256], s_b[256], s_c[256];
    __shared__ int s_i[256], s_h[256];
    // pass 1: number of "hard" samples (old-class labels) — the margin-ranking mean divides by hard_num * K
    int hard = 0;
    for (int n = threadIdx.x; n < a.B; n += 256) hard += (a.y[n] < a.num_old);
    s_h[threadIdx.x] = hard;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if (threadIdx.x < off) s_h[threadIdx.x] += s_h[threadIdx.x + off]; __syncthreads(); }
    const int hard_num = s_h[0];
    const float invB = 1.f / (float)a.B;
    const float mr_w = hard_num > 0 ? a.lw_mr / (float)(hard_num * a.K) : 0.f;
    float ce_acc = 0.f, lf_acc = 0.f, mr_acc = 0.f, dsig_acc = 0.f;
    int ok = 0;
    for (int n = threadIdx.x; n < a.B; n += 256) {
        const float* lg = a.logits + (size_t)n * a.ld;
        const float* sc = a.scores + (size_t)n * a.ld;
        float* dl = a.dlogits + (size_t)n * a.ld;
        float* ds = a.dscores + (size_t)n * a.ld;
        const int y = (int)a.y[n];
        // CE over all C logits + argmax
        float m = -CUDART_INF_F; int bi = 0;
        for (int k = 0; k < a.C; ++k) { const float v = lg[k]; if (v > m) { m = v; bi = k; } }
        a.pred[n] = bi; ok += (bi == y);
        float se = 0.f;
        for (int k = 0; k < a.C; ++k) se += expf(lg[k] - m);
        ce_acc += m + logf(se) - lg[y];
        const float inv_se = 1.f / se;
        for (int k = 0; k < a.C; ++k) {
            const float d = (expf(lg[k] - m) * inv_se - (k == y ? 1.f : 0.f)) * invB;
            dl[k] = d; ds[k] = 0.f;
            dsig_acc = fmaf(d, sc[k], dsig_acc);
        }
        // less-forget: lamda * mean_b (1 - cos(f, f_ref))   (CosineEmbeddingLoss, target +1; eps 1e-8 on the squared norms as ATen)
        const float* f = a.feat + (size_t)n * a.D;
        const float* r = a.ref_feat + (size_t)n * a.D;
        float ff = 0.f, rr = 0.f, fr = 0.f;
        for (int j = 0; j < a.D; ++j) { ff = fmaf(f[j], f[j], ff); rr = fmaf(r[j], r[j], rr); fr = fmaf(f[j], r[j], fr); }
        const float EPS = 1e-12f;     // at::cosine_embedding_loss: cos = prod / sqrt((mag1 + EPS) * (mag2 + EPS))
        const float denom = sqrtf((ff + EPS) * (rr + EPS));
        const float cosv = fr / denom;
        lf_acc += 1.f - cosv;
        {   // d(-cos)/df = -( r/denom - cos * f/(ff+EPS) )
            float* df = a.dfeat + (size_t)n * a.D;
            const float w = a.cur_lamda * invB;
            for (int j = 0; j < a.D; ++j) df[j] = -w * (r[j] / denom - cosv * f[j] / (ff + EPS));
        }
        // margin ranking: old-class samples only; K largest novel scores (ties: lowest index first, as a stable sort would)
        if (y < a.num_old && hard_num > 0) {
            const float gt = sc[y];
            float last = CUDART_INF_F; int last_idx = -1;
            for (int t = 0; t < a.K; ++t) {
                float best = -CUDART_INF_F; int bidx = -1;
                for (int k = a.num_old; k < a.C; ++k) {
                    const float v = sc[k];
                    const bool after = v < last || (v == last && k > last_idx);
                    if (after && v > best) { best = v; bidx = k; }
                }
                if (bidx < 0) break;
                const float viol = a.margin - (gt - best);      // MarginRankingLoss(x1 = gt, x2 = novel, y = 1): max(0, -(x1-x2)+margin)
                if (viol > 0.f) { mr_acc += viol; ds[y] -= mr_w; ds[bidx] += mr_w; }
                last = best; last_idx = bidx;
            }
        }
    }
    s_a[threadIdx.x] = ce_acc; s_b[threadIdx.x] = lf_acc; s_c[threadIdx.x] = mr_acc; s_i[threadIdx.x] = ok;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_a[threadIdx.x] += s_a[threadIdx.x + off]; s_b[threadIdx.x] += s_b[threadIdx.x + off];
            s_c[threadIdx.x] += s_c[threadIdx.x + off]; s_i[threadIdx.x] += s_i[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float ce = s_a[0] * invB, lf = a.cur_lamda * s_b[0] * invB, mr = hard_num > 0 ? a.lw_mr * s_c[0] / (float)(hard_num * a.K) : 0.f;
        a.scal[0] = lf + ce + mr; a.scal[1] = (float)s_i[0]; a.scal[2] = ce; a.scal[3] = lf; a.scal[5] = mr;
    }
    if (a.dsigma != nullptr) {        // second fixed-order block reduction (uniform branch)
        __syncthreads();
        s_a[threadIdx.x] = dsig_acc;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) { if (threadIdx.x < off) s_a[threadIdx.x] += s_a[threadIdx.x + off]; __syncthreads(); }
        if (threadIdx.x == 0) *a.dsigma = s_a[0];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// L2P prompt selection: the B x P cosine similarities on B CTAs, then one CTA for the vote / pull constraint (B <= 4096 samples, pool <= 32 prompts,
// D = embed dim)
// Tie rules (the reference leaves them to torch.topk): per-sample top-k by (value desc, index asc); majority top-k over the
// histogram by (count desc, id asc) after the reference's "pad with ids[0] / count 0" step.
// ---------------------------------------------------------------------------------------------------------------------
struct L2pArgs {
    const float* query;      // [B][D]  cls features of the frozen pass
    const float* key;        // [P][D]  prompt_key
    float* sim;              // [B][P]  out: cosine similarity
    long long* ids;          // [top_k] out: majority prompt ids (shared by every sample)
    int* hist;               // [P]     out
    float* reduce_sim;       // [1]     out: sum_b sum_{j in ids} khat_j . qhat_b / B
    float* dkey;             // [P][D]  out: d(reduce_sim)/d(key)  (nullable)
    float* qsum;             // [D]     scratch: sum_b qhat_b
    int B, P, D, top_k;
    int phase;               // 0: vote on this batch's own histogram; 1: per-sample top-k -> hist only; 2: vote on the histogram found in `hist`
};                           // (1 + 2 with a SUM all-reduce of `hist` in between = the batch-wide majority vote of the GLOBAL batch under data parallelism)

// (1) similarities: one CTA per sample, warp w owns prompts w, w + 8, ...  (inverse norms recomputed per warp: 2 x D floats out of L1 / L2)
__global__ void __launch_bounds__(256) l2p_sim_kernel(L2pArgs a) {
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* q = a.query + (size_t)b * a.D;
    float s = 0.f;
    for (int j = lane; j < a.D; j += 32) s = fmaf(q[j], q[j], s);
    s = warp_sum(s);
    const float qn = 1.f / fmaxf(sqrtf(s), 1e-12f);
    for (int p = warp; p < a.P; p += 8) {
        const float* k = a.key + (size_t)p * a.D;
        float n2 = 0.f;
        for (int j = lane; j < a.D; j += 32) n2 = fmaf(k[j], k[j], n2);
        n2 = warp_sum(n2);
        const float kn = 1.f / fmaxf(sqrtf(n2), 1e-12f);
        float d = 0.f;
        for (int j = lane; j < a.D; j += 32) d = fmaf(q[j] * qn, k[j] * kn, d);
        d = warp_sum(d);
        if (lane == 0) a.sim[(size_t)b * a.P + p] = d;
    }
}

// (2) vote + pull constraint: single CTA of kL2pNT threads over the [B][P] similarities
constexpr int kL2pNT = 768, kL2pNW = kL2pNT / 32;
__global__ void __launch_bounds__(kL2pNT) l2p_select_kernel(L2pArgs a) {
    extern __shared__ float sm[];
    float* s_knorm = sm;                  // [P] inverse norms
    float* s_qn = sm + 32;                // [B] inverse norms
    int* s_hist = reinterpret_cast<int*>(sm + 32 + a.B);     // [32]
    int* s_ids = s_hist + 32;             // [32]
    float* s_part = reinterpret_cast<float*>(s_ids + 32);    // [32] per-warp partial sums
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) s_hist[tid] = (a.phase == 2 && tid < a.P) ? a.hist[tid] : 0;
    // inverse norms: one warp per row
    if (a.phase != 1)
    for (int r = warp; r < a.P + a.B; r += kL2pNW) {
        const float* x = r < a.P ? a.key + (size_t)r * a.D : a.query + (size_t)(r - a.P) * a.D;
        float s = 0.f;
        for (int j = lane; j < a.D; j += 32) s = fmaf(x[j], x[j], s);
        s = warp_sum(s);
        if (lane == 0) { const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f); if (r < a.P) s_knorm[r] = inv; else s_qn[r - a.P] = inv; }
    }
    __syncthreads();
    // per-sample top-k -> histogram (integer atomics: order-independent)
    if (a.phase != 2)
    for (int b = tid; b < a.B; b += kL2pNT) {
        const float* s = a.sim + (size_t)b * a.P;
        float last = CUDART_INF_F; int last_idx = -1;
        for (int t = 0; t < a.top_k; ++t) {
            float best = -CUDART_INF_F; int bidx = -1;
            for (int p = 0; p < a.P; ++p) {
                const float v = s[p];
                const bool after = v < last || (v == last && p > last_idx);
                if (after && (v > best)) { best = v; bidx = p; }
            }
            if (bidx < 0) break;
            atomicAdd(&s_hist[bidx], 1);
            last = best; last_idx = bidx;
        }
    }
    __syncthreads();
    if (a.phase == 1) {                                  // uniform: the vote happens after the histograms of all ranks have been summed
        if (tid < a.P) a.hist[tid] = s_hist[tid];
        return;
    }
    if (tid == 0) {
        // reference: ids = unique(sorted) padded with ids[0]; counts padded with 0; topk(counts) -> ids
        int present[32], cnt[32], np = 0;
        for (int p = 0; p < a.P; ++p) if (s_hist[p] > 0) { present[np] = p; cnt[np] = s_hist[p]; ++np; }
        for (int i = np; i < a.P; ++i) { present[i] = present[0]; cnt[i] = 0; }
        bool used[32];
        for (int i = 0; i < a.P; ++i) used[i] = false;
        for (int t = 0; t < a.top_k; ++t) {
            int bi = -1;
            for (int i = 0; i < a.P; ++i) if (!used[i] && (bi < 0 || cnt[i] > cnt[bi])) bi = i;
            used[bi] = true;
            s_ids[t] = present[bi];
            a.ids[t] = present[bi];
        }
        for (int p = 0; p < a.P; ++p) a.hist[p] = s_hist[p];
    }
    // qsum[j] = sum_b qhat_b[j]  (fixed order; the loads of a column are independent of each other) ; reduce_sim = sum_{t} khat_{id_t} . qsum / B
    for (int j = tid; j < a.D; j += kL2pNT) {
        float s = 0.f;
#pragma unroll 8
        for (int b = 0; b < a.B; ++b) s = fmaf(a.query[(size_t)b * a.D + j], s_qn[b], s);
        a.qsum[j] = s;
    }
    __syncthreads();
    float part = 0.f;
    for (int t = 0; t < a.top_k; ++t) {
        const int p = s_ids[t];
        for (int j = tid; j < a.D; j += kL2pNT) part = fmaf(a.key[(size_t)p * a.D + j] * s_knorm[p], a.qsum[j], part);
    }
    part = warp_sum(part);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kL2pNW; ++w) t += s_part[w];
        *a.reduce_sim = t / (float)a.B;
    }
    // d(reduce_sim)/d(key_p) = mult_p * (v - khat (khat . v)) / ||k||,  v = qsum / B,  mult_p = #times p appears in ids
    if (a.dkey != nullptr) {
        for (int p = warp; p < a.P; p += kL2pNW) {
            int mult = 0;
            for (int t = 0; t < a.top_k; ++t) mult += (s_ids[t] == p);
            const float* k = a.key + (size_t)p * a.D;
            const float inv = s_knorm[p];
            float dot = 0.f;
            for (int j = lane; j < a.D; j += 32) dot = fmaf(k[j] * inv, a.qsum[j], dot);
            dot = warp_sum(dot);
            for (int j = lane; j < a.D; j += 32)
                a.dkey[(size_t)p * a.D + j] = mult == 0 ? 0.f : (float)mult * (a.qsum[j] - k[j] * inv * dot) * inv / (float)a.B;
        }
    }
}

// batched_prompt[b][t*L + l][:] = prompt[ids[t]][l][:]   (prompt [P][L][D]); one float4 per thread
__global__ void __launch_bounds__(256) l2p_gather_kernel(const float* prompt, const long long* ids, float* out, int B, int top_k, int L, int D) {
    const long long n4 = (long long)B * top_k * L * D / 4;
    const int row4 = D / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const int d4 = (int)(i % row4);
        const long long r = i / row4;
        const int l = (int)(r % L), t = (int)((r / L) % top_k);
        const long long src = (((long long)ids[t] * L + l) * D) / 4 + d4;
        reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(prompt) + src);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// GPM projection: G <- G - G @ M  with G [R][D], M [D][D].  One CTA owns RT full rows (so the update can be in place), stages them
// in shared memory and streams M; fp32 FMA, 4 output columns per thread.
// ---------------------------------------------------------------------------------------------------------------------
template <int RT>
__global__ void __launch_bounds__(256) gpm_project_kernel(float* G, const float* M, int R, int D) {
    extern __shared__ float s_g[];                 // [RT][D]
    const int r0 = blockIdx.x * RT;
    const int rows = R - r0 < RT ? R - r0 : RT;
    for (int e = threadIdx.x; e < RT * D; e += 256) {
        const int r = e / D, j = e % D;
        s_g[e] = r < rows ? G[(size_t)(r0 + r) * D + j] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < D; c += 256 * 4) {       // D % 4 == 0
        float acc[RT][4];
#pragma unroll
        for (int r = 0; r < RT; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        for (int k = 0; k < D; ++k) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(M + (size_t)k * D + c));
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const float g = s_g[r * D + k];
                acc[r][0] = fmaf(g, m.x, acc[r][0]); acc[r][1] = fmaf(g, m.y, acc[r][1]);
                acc[r][2] = fmaf(g, m.z, acc[r][2]); acc[r][3] = fmaf(g, m.w, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (r < rows) {
                float4 o;
                o.x = s_g[r * D + c] - acc[r][0]; o.y = s_g[r * D + c + 1] - acc[r][1];
                o.z = s_g[r * D + c + 2] - acc[r][2]; o.w = s_g[r * D + c + 3] - acc[r][3];
                *reinterpret_cast<float4*>(G + (size_t)(r0 + r) * D + c) = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// LoRA (InfLoRA_OPT): out[3D][D] = cat(Wq, Wk + Bk Ak, Wv + Bv Av); A [r][D], B [D][r], r <= 16
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lora_merge_qkv_kernel(const float* qkv, const float* Ak, const float* Bk, const float* Av, const float* Bv,
                                                             float* out, int D, int r) {
    const long long n4 = (long long)3 * D * D / 4;
    const int row4 = D / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const int row = (int)(i / row4), c = (int)(i % row4) * 4;
        float4 v = __ldg(reinterpret_cast<const float4*>(qkv) + i);
        if (row >= D) {
            const bool is_k = row < 2 * D;
            const float* A = is_k ? Ak : Av;
            const float* Bm = (is_k ? Bk : Bv) + (size_t)(row - (is_k ? D : 2 * D)) * r;
            for (int t = 0; t < r; ++t) {
                const float b = __ldg(Bm + t);
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(A + (size_t)t * D + c));
                v.x = fmaf(b, a4.x, v.x); v.y = fmaf(b, a4.y, v.y); v.z = fmaf(b, a4.z, v.z); v.w = fmaf(b, a4.w, v.w);
            }
        }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// dB[i][t] = sum_j dW[i][j] * A[t][j]   (dW: the [D][D] slice of d(W') for k or v); one warp per output row i
__global__ void __launch_bounds__(128) lora_bgrad_kernel(const float* dW, const float* A, float* dB, int D, int r) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= D) return;
    const float* g = dW + (size_t)i * D;
    for (int t = 0; t < r; ++t) {
        float s = 0.f;
        for (int j = lane; j < D; j += 32) s = fmaf(__ldg(g + j), __ldg(A + (size_t)t * D + j), s);
        s = warp_sum(s);
        if (lane == 0) dB[(size_t)i * r + t] = s;
    }
}

}  // namespace lc

namespace lc {

// ---------------------------------------------------------------------------------------------------------------------
// iCaRL exemplar management (core/model/buffer/linearherdingbuffer.py:133-163, core/model/icarl.py:122-152)
// ---------------------------------------------------------------------------------------------------------------------
// Herding: one CTA per class over its (already L2-normalised) feature rows [begin, end).  Mirrors the reference's fp32 op order:
//   cost_j = || mean - (f_j + running_sum) / (i + 1) ||_2 ,  idx = first argmin,  running_sum += f_idx,  f_idx += 1e6
// `work` is a scratch copy of the features (the algorithm mutates them).  out[c][i] = global row index, -1 when the class has
// fewer than `per_class` rows.
template <int D>
__global__ void __launch_bounds__(256) herding_select_kernel(const float* feats, float* work, const int* cls_begin, int per_class, long long* out) {
    __shared__ float s_mean[D], s_rs[D];
    __shared__ float s_cost[256];
    __shared__ int s_arg[256];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int begin = cls_begin[c], end = cls_begin[c + 1], n = end - begin;
    for (int e = tid; e < n * D; e += 256) work[(size_t)begin * D + e] = feats[(size_t)begin * D + e];
    if (tid < D) {      // torch's mean over dim 0: sum of the column, divided by n
        float s = 0.f;
        for (int j = 0; j < n; ++j) s += feats[(size_t)(begin + j) * D + tid];
        s_mean[tid] = s / (float)n;
        s_rs[tid] = 0.f;
    }
    __syncthreads();
    for (int i = 0; i < per_class; ++i) {
        if (i >= n) { if (tid == 0) out[(size_t)c * per_class + i] = -1; continue; }
        const float inv = (float)(i + 1);
        float best = CUDART_INF_F; int barg = 0x7fffffff;
        for (int j = tid; j < n; j += 256) {
            const float* f = work + (size_t)(begin + j) * D;
            float acc = 0.f;
#pragma unroll 8
            for (int d = 0; d < D; ++d) {
                const float t = s_mean[d] - (f[d] + s_rs[d]) / inv;
                acc = fmaf(t, t, acc);
            }
            const float cost = sqrtf(acc);
            if (cost < best) { best = cost; barg = j; }
        }
        s_cost[tid] = best; s_arg[tid] = barg;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (tid < off) {
                const float oc = s_cost[tid + off]; const int oa = s_arg[tid + off];
                if (oc < s_cost[tid] || (oc == s_cost[tid] && oa < s_arg[tid])) { s_cost[tid] = oc; s_arg[tid] = oa; }
            }
            __syncthreads();
        }
        const int pick = s_arg[0];
        if (tid == 0) out[(size_t)c * per_class + i] = begin + pick;
        if (tid < D) {
            float* f = work + (size_t)(begin + pick) * D;
            s_rs[tid] += f[tid];
            f[tid] = f[tid] + 1e6f;
        }
        __syncthreads();
    }
}

// nearest-class-mean: pred[b] = argmin_c sum_d (feat[b][d] - mean[c][d])^2   (first minimum), one warp per sample
template <int D>
__global__ void __launch_bounds__(128) ncm_classify_kernel(const float* feat, const float* means, int B, int C, long long* pred) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* f = feat + (size_t)b * D;
    float best = CUDART_INF_F; int barg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float* m = means + (size_t)c * D;
        float acc = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) { const float t = f[d] - __ldg(m + d); acc = fmaf(t, t, acc); }
        if (acc < best) { best = acc; barg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oa = __shfl_xor_sync(0xffffffffu, barg, o);
        if (ob < best || (ob == best && oa < barg)) { best = ob; barg = oa; }
    }
    if (lane == 0) pred[b] = barg;
}

}  // namespace lc
