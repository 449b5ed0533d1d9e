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
#pragma unroll 16                                   // 16 row loads in flight; the additions keep their order
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

// One block of 1024 threads: warp w owns samples w, w + 32, ...; the lanes stride over the classes (max / sum-exp by warp shuffles), so a 128 x 100 batch
// is one pass of 4 samples per warp instead of 100-term serial loops in 128 threads (40 us -> a few us on the critical path of every step).  Sums over
// the samples: per-warp partials in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) ce_kd_loss_kernel(LossArgs a) {
    __shared__ float s_ce[32], s_kd[32];
    __shared__ int s_ok[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ce_acc = 0.f, kd_acc = 0.f;
    int ok_acc = 0;
    const float invB = 1.f / (float)a.B;
    for (int n = warp; n < a.B; n += 32) {
        const float* lg = a.logits + (size_t)n * a.ldl;
        float* dl = a.dlogits + (size_t)n * a.ldl;
        const int y = (int)a.y[n];
        // prediction (first maximal index)
        float best = -CUDART_INF_F; int bi = 0x7fffffff;
        for (int k = a.pred_lo + lane; k < a.pred_n; k += 32) { const float v = lg[k]; if (v > best) { best = v; bi = k; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (bi == 0x7fffffff) bi = a.pred_lo;
        // cross entropy on the slice
        float m = -CUDART_INF_F;
        for (int k = a.ce_lo + lane; k < a.ce_hi; k += 32) m = fmaxf(m, lg[k]);
        m = warp_max(m);
        float se = 0.f;
        for (int k = a.ce_lo + lane; k < a.ce_hi; k += 32) se += expf(lg[k] - m);
        se = warp_sum(se);
        const float lse = m + logf(se), inv_se = 1.f / se;
        // distillation
        float ms = -CUDART_INF_F, mt = -CUDART_INF_F, ss = 0.f, stt = 0.f, kd = 0.f, lss = 0.f, inv_ss = 0.f, inv_st = 0.f;
        const float invT = a.kd_n > 0 ? 1.f / a.T : 0.f;
        const float* tg = a.kd_n > 0 ? a.teacher + (size_t)n * a.ldt : nullptr;
        if (a.kd_n > 0) {
            for (int k = lane; k < a.kd_n; k += 32) { ms = fmaxf(ms, lg[k] * invT); mt = fmaxf(mt, tg[k] * invT); }
            ms = warp_max(ms); mt = warp_max(mt);
            for (int k = lane; k < a.kd_n; k += 32) { ss += expf(lg[k] * invT - ms); stt += expf(tg[k] * invT - mt); }
            ss = warp_sum(ss); stt = warp_sum(stt);
            lss = ms + logf(ss); inv_ss = 1.f / ss; inv_st = 1.f / stt;
        }
        for (int k = lane; k < a.ncols; k += 32) {
            float d = 0.f;
            if (k >= a.ce_lo && k < a.ce_hi) d = (expf(lg[k] - m) * inv_se - (k == y ? 1.f : 0.f)) * invB;
            if (k < a.kd_n) {
                const float q = expf(tg[k] * invT - mt) * inv_st;
                const float ps = expf(lg[k] * invT - ms) * inv_ss;
                kd -= q * (lg[k] * invT - lss);
                d += a.kd_w * invB * invT * (ps - q);
            }
            dl[k] = d;
        }
        kd = warp_sum(kd);
        if (lane == 0) {
            a.pred[n] = bi;
            ok_acc += (bi == y);
            ce_acc += lse - lg[y];
            kd_acc += kd;
        }
    }
    if (lane == 0) { s_ce[warp] = ce_acc; s_kd[warp] = kd_acc; s_ok[warp] = ok_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ces = 0.f, kds = 0.f; int oks = 0;
        for (int w = 0; w < 32; ++w) { ces += s_ce[w]; kds += s_kd[w]; oks += s_ok[w]; }
        const float ce = ces * invB, kd = kds * invB;
        a.scal[0] = ce + a.kd_w * kd + (a.extra != nullptr ? a.extra_coeff * *a.extra : 0.f);
        a.scal[1] = (float)oks;
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
#pragma unroll 16                                   // these loops are chains of L2 latencies, not arithmetic: keep 16 loads in flight (same FMA order)
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
#pragma unroll 10
        for (int k = 0; k < ncls; ++k) acc = fmaf(__ldg(dlogits + (size_t)n * ldl + k), __ldg(W + (size_t)k * C + c), acc);
        dfeat[(size_t)n * C + c] = acc;
        if (gact != nullptr) {
            const float g = acc / (float)HW;
            float* dst = gact + (size_t)n * HW * C + c;
#pragma unroll 16
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
#pragma unroll 16
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
