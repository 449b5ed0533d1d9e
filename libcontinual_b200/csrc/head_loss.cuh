// Classifier head + loss kernels (fp32).
//   avgpool_fc_fwd      : nn.AvgPool2d(8) + flatten + nn.Linear          (resnet.py:389-390, ewc.py:52-57, icarl.py:24-38)
//   ce_kd_loss          : F.cross_entropy over a logit slice + `_KD_loss` (ewc.py:90-100, icarl.py:197-221, lwf.py:52-78)
//                         and their gradient w.r.t. the logits, argmax prediction, #correct
//   head_bwd            : Linear backward (dW, db, dfeat) + AvgPool backward
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace lc {

// one CTA per sample; C = feature width (64), HW = pooled pixels (64)
template <int C>
__global__ void __launch_bounds__(C) avgpool_fc_fwd_kernel(const float* act /*[B][HW][C]*/, int HW, const float* W /*[ncls][C]*/,
                                                            const float* bias /*nullable*/, int ncls, float* feat /*[B][C]*/,
                                                            float* logits /*[B][ldl]*/, int ldl) {
    __shared__ float s_f[C];
    const int n = blockIdx.x, c = threadIdx.x;
    const float* src = act + (size_t)n * HW * C + c;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __ldg(src + (size_t)p * C);
    s = s / (float)HW;
    s_f[c] = s;
    feat[(size_t)n * C + c] = s;
    __syncthreads();
    for (int k = c; k < ncls; k += C) {
        const float* w = W + (size_t)k * C;
        float d = 0.f;
#pragma unroll 8
        for (int j = 0; j < C; ++j) d = fmaf(s_f[j], __ldg(w + j), d);
        logits[(size_t)n * ldl + k] = d + (bias != nullptr ? bias[k] : 0.f);
    }
}

struct LossArgs {
    const float* logits;    // [B][ldl] student
    const float* teacher;   // nullable [B][ldt]
    const long long* y;     // [B] absolute labels (int64)
    float* dlogits;         // [B][ldl], columns [0, ncols) written (others untouched)
    long long* pred;        // [B]
    float* scal;            // [0]=loss (ce + kd_w*kd)  [1]=#correct  [2]=ce  [3]=kd
    int B, ldl, ldt, ncols;
    int ce_lo, ce_hi;       // CE over logits[:, ce_lo:ce_hi) with target y - ce_lo
    int kd_n;               // KD over logits[:, 0:kd_n) vs teacher[:, 0:kd_n)   (0 = off)
    int pred_n;             // argmax over logits[:, pred_lo:pred_n)
    int pred_lo;
    const float* extra;     // nullable device scalar: scal[0] += extra_coeff * *extra (L2P pull constraint, l2p.py:99)
    float extra_coeff;
    float kd_w, T;
};

__global__ void __launch_bounds__(256) ce_kd_loss_kernel(LossArgs a) {
    __shared__ float s_ce[256], s_kd[256];
    __shared__ int s_ok[256];
    float ce_acc = 0.f, kd_acc = 0.f;
    int ok_acc = 0;
    const float invB = 1.f / (float)a.B;
    for (int n = threadIdx.x; n < a.B; n += 256) {
        const float* lg = a.logits + (size_t)n * a.ldl;
        float* dl = a.dlogits + (size_t)n * a.ldl;
        const int y = (int)a.y[n];
        // prediction (first maximal index)
        float best = -CUDART_INF_F; int bi = a.pred_lo;
        for (int k = a.pred_lo; k < a.pred_n; ++k) { const float v = lg[k]; if (v > best) { best = v; bi = k; } }
        a.pred[n] = bi;
        ok_acc += (bi == y);
        for (int k = 0; k < a.ncols; ++k) dl[k] = 0.f;
        // cross entropy on the slice
        float m = -CUDART_INF_F;
        for (int k = a.ce_lo; k < a.ce_hi; ++k) m = fmaxf(m, lg[k]);
        float se = 0.f;
        for (int k = a.ce_lo; k < a.ce_hi; ++k) se += expf(lg[k] - m);
        const float lse = m + logf(se);
        ce_acc += lse - lg[y];
        const float inv_se = 1.f / se;
        for (int k = a.ce_lo; k < a.ce_hi; ++k) dl[k] = (expf(lg[k] - m) * inv_se - (k == y ? 1.f : 0.f)) * invB;
        // distillation
        if (a.kd_n > 0) {
            const float* tg = a.teacher + (size_t)n * a.ldt;
            const float invT = 1.f / a.T;
            float ms = -CUDART_INF_F, mt = -CUDART_INF_F;
            for (int k = 0; k < a.kd_n; ++k) { ms = fmaxf(ms, lg[k] * invT); mt = fmaxf(mt, tg[k] * invT); }
            float ss = 0.f, st = 0.f;
            for (int k = 0; k < a.kd_n; ++k) { ss += expf(lg[k] * invT - ms); st += expf(tg[k] * invT - mt); }
            const float lss = ms + logf(ss), inv_ss = 1.f / ss, inv_st = 1.f / st;
            float kd = 0.f;
            for (int k = 0; k < a.kd_n; ++k) {
                const float q = expf(tg[k] * invT - mt) * inv_st;
                const float ps = expf(lg[k] * invT - ms) * inv_ss;
                kd -= q * (lg[k] * invT - lss);
                dl[k] += a.kd_w * invB * invT * (ps - q);
            }
            kd_acc += kd;
        }
    }
    s_ce[threadIdx.x] = ce_acc; s_kd[threadIdx.x] = kd_acc; s_ok[threadIdx.x] = ok_acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_ce[threadIdx.x] += s_ce[threadIdx.x + off];
            s_kd[threadIdx.x] += s_kd[threadIdx.x + off];
            s_ok[threadIdx.x] += s_ok[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float ce = s_ce[0] * invB, kd = s_kd[0] * invB;
        a.scal[0] = ce + a.kd_w * kd + (a.extra != nullptr ? a.extra_coeff * *a.extra : 0.f);
        a.scal[1] = (float)s_ok[0];
        a.scal[2] = ce;
        a.scal[3] = kd;
    }
}

// blocks [0, ncls): dW[k][:], db[k]; blocks [ncls, ncls+B): dfeat[n][:] and the avg-pool backward broadcast into gact
template <int C>
__global__ void __launch_bounds__(C) head_bwd_kernel(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int B,
                                                      float* dW, float* db /*nullable*/, float* dfeat, float* gact /*nullable [B][HW][C]*/, int HW) {
    const int c = threadIdx.x;
    if ((int)blockIdx.x < ncls) {
        const int k = blockIdx.x;
        float acc = 0.f, bacc = 0.f;
        for (int n = 0; n < B; ++n) {
            const float d = __ldg(dlogits + (size_t)n * ldl + k);
            acc = fmaf(d, __ldg(feat + (size_t)n * C + c), acc);
            bacc += d;
        }
        dW[(size_t)k * C + c] = acc;
        if (db != nullptr && c == 0) db[k] = bacc;
    } else {
        const int n = blockIdx.x - ncls;
        float acc = 0.f;
        for (int k = 0; k < ncls; ++k) acc = fmaf(__ldg(dlogits + (size_t)n * ldl + k), __ldg(W + (size_t)k * C + c), acc);
        dfeat[(size_t)n * C + c] = acc;
        if (gact != nullptr) {
            const float g = acc / (float)HW;
            float* dst = gact + (size_t)n * HW * C + c;
            for (int p = 0; p < HW; ++p) dst[(size_t)p * C] = g;
        }
    }
}

// stand-alone global average pool (nn.AvgPool2d(8) + flatten, resnet.py:389-390) and its backward, for callers that put
// their own head on top of backbone(x)['features']
template <int C>
__global__ void __launch_bounds__(C) avgpool_fwd_kernel(const float* act, int HW, float* feat) {
    const int n = blockIdx.x, c = threadIdx.x;
    const float* src = act + (size_t)n * HW * C + c;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __ldg(src + (size_t)p * C);
    feat[(size_t)n * C + c] = s / (float)HW;
}
template <int C>
__global__ void __launch_bounds__(C) avgpool_bwd_kernel(const float* dfeat, int HW, float* gact) {
    const int n = blockIdx.x, c = threadIdx.x;
    const float g = dfeat[(size_t)n * C + c] / (float)HW;
    float* dst = gact + (size_t)n * HW * C + c;
    for (int p = 0; p < HW; ++p) dst[(size_t)p * C] = g;
}

}  // namespace lc
