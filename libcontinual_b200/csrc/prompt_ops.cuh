// Prompt-pool kernels of the prefix-tuning methods (core/model/backbone/prompt.py:231-337 DualPrompt): key-query cosine match with the
// task-id bootstrap loss and its key gradient (training) or the per-sample top-1 selection (inference), and the gather of the selected
// prompt halves into the per-image BF16 prefix key / value rows consumed by the fused attention kernels.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace lc {

constexpr int kPromptMaxLayers = 8;

struct KeyMatchArgs {
    const float* q;                       // [B][D] query features (no-grad pass; detached in the reference: prompt.py:277)
    const float* K[kPromptMaxLayers];     // per e-layer [pool][D]
    float* dK[kPromptMaxLayers];          // per e-layer [pool][D] gradient rows (training: only row `task_id` is written)
    long long* idx;                       // [nl][B] selected pool row per image
    float* loss;                          // [1] training: sum_l sum_b (1 - cos(q_b, K_l[task_id]))   (prompt.py:283)
    int nl, B, pool, D, task_id;          // task_id >= 0: training (bootstrap), < 0: inference (argmax over the pool)
};

// one block, 256 threads (8 warps; a warp owns image rows warp, warp + 8, ...).  D <= 768 * ... handled 32 lanes x D/32 strided.
template <int D>
__global__ void __launch_bounds__(256) prompt_key_match_kernel(KeyMatchArgs a) {
    constexpr int PER = D / 32;
    __shared__ float s_S[8][D];
    __shared__ float s_c[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float total_loss = 0.f;
    for (int l = 0; l < a.nl; ++l) {
        if (a.task_id >= 0) {
            const float* k = a.K[l] + (size_t)a.task_id * D;
            float kv[PER], kn = 0.f;
#pragma unroll
            for (int i = 0; i < PER; ++i) { kv[i] = k[lane + 32 * i]; kn = fmaf(kv[i], kv[i], kn); }
            kn = fmaxf(sqrtf(warp_sum(kn)), 1e-12f);
            float S[PER], csum = 0.f;
#pragma unroll
            for (int i = 0; i < PER; ++i) S[i] = 0.f;
            for (int b = warp; b < a.B; b += 8) {
                const float* q = a.q + (size_t)b * D;
                float qv[PER], qn = 0.f, dot = 0.f;
#pragma unroll
                for (int i = 0; i < PER; ++i) { qv[i] = q[lane + 32 * i]; qn = fmaf(qv[i], qv[i], qn); dot = fmaf(qv[i], kv[i], dot); }
                qn = fmaxf(sqrtf(warp_sum(qn)), 1e-12f);
                dot = warp_sum(dot);
                csum += dot / (qn * kn);
                const float iq = 1.f / qn;
#pragma unroll
                for (int i = 0; i < PER; ++i) S[i] = fmaf(qv[i], iq, S[i]);
                if (lane == 0) a.idx[(size_t)l * a.B + b] = a.task_id;
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) s_S[warp][lane + 32 * i] = S[i];
            if (lane == 0) s_c[warp] = csum;
            __syncthreads();
            float C = 0.f;
            for (int w = 0; w < 8; ++w) C += s_c[w];
            // d/dK[task] of sum_b (1 - qhat_b . khat) = -(sum_b qhat_b - C khat) / |k|
            for (int d = threadIdx.x; d < D; d += 256) {
                float Sd = 0.f;
                for (int w = 0; w < 8; ++w) Sd += s_S[w][d];
                const float kh = k[d] / kn;
                a.dK[l][(size_t)a.task_id * D + d] = -(Sd - C * kh) / kn;
            }
            total_loss += (float)a.B - C;
            __syncthreads();
        } else {
            for (int b = warp; b < a.B; b += 8) {
                const float* q = a.q + (size_t)b * D;
                float qv[PER];
#pragma unroll
                for (int i = 0; i < PER; ++i) qv[i] = q[lane + 32 * i];
                float best = -3.4e38f;
                int bi = 0;
                for (int p = 0; p < a.pool; ++p) {
                    const float* k = a.K[l] + (size_t)p * D;
                    float kn = 0.f, dot = 0.f;
#pragma unroll
                    for (int i = 0; i < PER; ++i) { const float kx = k[lane + 32 * i]; kn = fmaf(kx, kx, kn); dot = fmaf(qv[i], kx, dot); }
                    const float cs = warp_sum(dot) / fmaxf(sqrtf(warp_sum(kn)), 1e-12f);      // the query norm is common to all keys of a row
                    if (cs > best) { best = cs; bi = p; }
                }
                if (lane == 0) a.idx[(size_t)l * a.B + b] = bi;
            }
        }
    }
    if (threadIdx.x == 0 && a.loss != nullptr) a.loss[0] = total_loss;
}

// out[b][r][:] = bf16(src[idx[b] * idx_stride + r * D + :])   (idx nullable: row 0 for every image);  grid (B, rows), D/4 threads
__global__ void __launch_bounds__(192) gather_rows_bf16_kernel(const float* src, const long long* idx, long long idx_stride, int rows, int D, __nv_bfloat16* out) {
    const int b = blockIdx.x, r = blockIdx.y;
    const float* s = src + (idx != nullptr ? (size_t)idx[b] * idx_stride : 0) + (size_t)r * D;
    __nv_bfloat16* o = out + ((size_t)b * rows + r) * D;
    for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(s + d);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(o + d) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

}  // namespace lc

namespace lc {

// ---------------------------------------------------------------------------------------------------------------------
// CodaPrompt (core/model/backbone/prompt.py:37-220): per block l in 0..4 and image b
//     v_k = q_b * A_k (elementwise) ; alpha[b][k] = cos(v_k, K_k) ; P_[b] = sum_k alpha[b][k] p[k]   ([Lp][D]; first half = prefix keys, second = values)
// over the first `nk` pool components (the reference never advances `task_count`, so s = 0, f = pool / n_tasks in every task: prompt.py:166-168).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kCodaMaxK = 16;

struct CodaArgs {
    const float* q;                        // [B][D]
    const float* K[kPromptMaxLayers];      // per layer [pool][D]
    const float* A[kPromptMaxLayers];
    const float* p[kPromptMaxLayers];      // per layer [pool][Lp][D]
    __nv_bfloat16* pk[kPromptMaxLayers];   // out per layer [B][Lp/2][D]
    __nv_bfloat16* pv[kPromptMaxLayers];
    float* alpha;                          // [nl][B][nk]   cosines
    float* vnorm;                          // [nl][B][nk]   |q_b * A_k|
    int B, nk, Lp, D;
};

// grid (B, nl), 192 threads (4 columns each, D = 768)
__global__ void __launch_bounds__(192) coda_prompt_fwd_kernel(CodaArgs a) {
    __shared__ float s_red[3][6];
    __shared__ float s_alpha[kCodaMaxK];
    const int b = blockIdx.x, l = blockIdx.y, d = threadIdx.x * 4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D;
    const float4 q = *reinterpret_cast<const float4*>(a.q + (size_t)b * D + d);
    for (int k = 0; k < a.nk; ++k) {
        const float4 A = ldg4(a.A[l] + (size_t)k * D + d), K = ldg4(a.K[l] + (size_t)k * D + d);
        const float4 v = make_float4(q.x * A.x, q.y * A.y, q.z * A.z, q.w * A.w);
        float vv = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, kk = K.x * K.x + K.y * K.y + K.z * K.z + K.w * K.w,
              vk = v.x * K.x + v.y * K.y + v.z * K.z + v.w * K.w;
        vv = warp_sum(vv); kk = warp_sum(kk); vk = warp_sum(vk);
        __syncthreads();
        if (lane == 0) { s_red[0][warp] = vv; s_red[1][warp] = kk; s_red[2][warp] = vk; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float sv = 0.f, sk = 0.f, svk = 0.f;
            for (int w = 0; w < 6; ++w) { sv += s_red[0][w]; sk += s_red[1][w]; svk += s_red[2][w]; }
            const float nv = fmaxf(sqrtf(sv), 1e-12f), nK = fmaxf(sqrtf(sk), 1e-12f);
            const float al = svk / (nv * nK);
            s_alpha[k] = al;
            const size_t o = ((size_t)l * a.B + b) * a.nk + k;
            a.alpha[o] = al; a.vnorm[o] = nv;
        }
    }
    __syncthreads();
    const int half = a.Lp / 2;
    for (int r = 0; r < a.Lp; ++r) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < a.nk; ++k) {
            const float4 pv = ldg4(a.p[l] + ((size_t)k * a.Lp + r) * D + d);
            const float al = s_alpha[k];
            acc.x = fmaf(al, pv.x, acc.x); acc.y = fmaf(al, pv.y, acc.y); acc.z = fmaf(al, pv.z, acc.z); acc.w = fmaf(al, pv.w, acc.w);
        }
        __nv_bfloat16* o = (r < half ? a.pk[l] + ((size_t)b * half + r) * D : a.pv[l] + ((size_t)b * half + (r - half)) * D) + d;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
        *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

struct CodaBwdArgs {
    const float* q;
    const float* K[kPromptMaxLayers];
    const float* A[kPromptMaxLayers];
    const float* p[kPromptMaxLayers];
    const float* dpk[kPromptMaxLayers];    // per layer [B][Lp/2][D] fp32 (attention backward)
    const float* dpv[kPromptMaxLayers];
    float* dK[kPromptMaxLayers];           // per layer [pool][D]: rows [0, nk) written
    float* dA[kPromptMaxLayers];
    float* dp[kPromptMaxLayers];           // per layer [pool][Lp][D]: components [0, nk) written
    const float* alpha;
    const float* vnorm;
    float* dalpha;                         // [nl][B][nk] scratch
    int B, nk, Lp, D;
};

// dalpha[l][b][k] = <dP_[b], p[k]> ; grid (B, nl), 192 threads
__global__ void __launch_bounds__(192) coda_prompt_bwd_alpha_kernel(CodaBwdArgs a) {
    __shared__ float s_red[kCodaMaxK][6];
    const int b = blockIdx.x, l = blockIdx.y, d = threadIdx.x * 4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D, half = a.Lp / 2;
    float acc[kCodaMaxK];
#pragma unroll
    for (int k = 0; k < kCodaMaxK; ++k) acc[k] = 0.f;
    for (int r = 0; r < a.Lp; ++r) {
        const float* g = (r < half ? a.dpk[l] + ((size_t)b * half + r) * D : a.dpv[l] + ((size_t)b * half + (r - half)) * D) + d;
        const float4 gv = *reinterpret_cast<const float4*>(g);
#pragma unroll
        for (int k = 0; k < kCodaMaxK; ++k) {
            if (k < a.nk) {
                const float4 pv = ldg4(a.p[l] + ((size_t)k * a.Lp + r) * D + d);
                acc[k] += gv.x * pv.x + gv.y * pv.y + gv.z * pv.z + gv.w * pv.w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kCodaMaxK; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) s_red[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < a.nk) {
        float s = 0.f;
        for (int w = 0; w < 6; ++w) s += s_red[threadIdx.x][w];
        a.dalpha[((size_t)l * a.B + b) * a.nk + threadIdx.x] = s;
    }
}

// grid (Lp + nk, nl), 192 threads.  Blocks [0, Lp): dp[k][r][:] = sum_b alpha[b][k] dP_[b][r][:].  Blocks [Lp, Lp + nk): the key / attention-vector
// gradients of component k from dalpha through the cosine (saved alpha, |v|):
//   dK_k = sum_b dalpha (v / |v| - alpha Khat) / |K| ,  dA_k = sum_b dalpha ((Khat - alpha v / |v|) / |v|) * q_b ,  v = q_b * A_k
__global__ void __launch_bounds__(192) coda_prompt_bwd_param_kernel(CodaBwdArgs a) {
    __shared__ float s_red[6];
    __shared__ float s_nK;
    const int l = blockIdx.y, d = threadIdx.x * 4, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D, half = a.Lp / 2, B = a.B, nk = a.nk;
    const float* alpha = a.alpha + (size_t)l * B * nk;
    if ((int)blockIdx.x < a.Lp) {
        const int r = blockIdx.x;
        float4 acc[kCodaMaxK];
#pragma unroll
        for (int k = 0; k < kCodaMaxK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < B; ++b) {
            const float* g = (r < half ? a.dpk[l] + ((size_t)b * half + r) * D : a.dpv[l] + ((size_t)b * half + (r - half)) * D) + d;
            const float4 gv = *reinterpret_cast<const float4*>(g);
#pragma unroll
            for (int k = 0; k < kCodaMaxK; ++k) {
                if (k < nk) {
                    const float al = __ldg(alpha + (size_t)b * nk + k);
                    acc[k].x = fmaf(al, gv.x, acc[k].x); acc[k].y = fmaf(al, gv.y, acc[k].y); acc[k].z = fmaf(al, gv.z, acc[k].z); acc[k].w = fmaf(al, gv.w, acc[k].w);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kCodaMaxK; ++k)
            if (k < nk) *reinterpret_cast<float4*>(a.dp[l] + ((size_t)k * a.Lp + r) * D + d) = acc[k];
        return;
    }
    const int k = blockIdx.x - a.Lp;
    const float4 K = ldg4(a.K[l] + (size_t)k * D + d), A = ldg4(a.A[l] + (size_t)k * D + d);
    float kk = warp_sum(K.x * K.x + K.y * K.y + K.z * K.z + K.w * K.w);
    if (lane == 0) s_red[warp] = kk;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 6; ++w) s += s_red[w]; s_nK = fmaxf(sqrtf(s), 1e-12f); }
    __syncthreads();
    const float inK = 1.f / s_nK;
    const float4 Kh = make_float4(K.x * inK, K.y * inK, K.z * inK, K.w * inK);
    const float* dal = a.dalpha + (size_t)l * B * nk;
    const float* vn = a.vnorm + (size_t)l * B * nk;
    float4 gK = make_float4(0.f, 0.f, 0.f, 0.f), gA = gK;
    for (int b = 0; b < B; ++b) {
        const float4 q = *reinterpret_cast<const float4*>(a.q + (size_t)b * D + d);
        const float da = __ldg(dal + (size_t)b * nk + k), al = __ldg(alpha + (size_t)b * nk + k), inv = 1.f / __ldg(vn + (size_t)b * nk + k);
        const float4 vh = make_float4(q.x * A.x * inv, q.y * A.y * inv, q.z * A.z * inv, q.w * A.w * inv);
        gK.x = fmaf(da, (vh.x - al * Kh.x) * inK, gK.x); gK.y = fmaf(da, (vh.y - al * Kh.y) * inK, gK.y);
        gK.z = fmaf(da, (vh.z - al * Kh.z) * inK, gK.z); gK.w = fmaf(da, (vh.w - al * Kh.w) * inK, gK.w);
        gA.x = fmaf(da, (Kh.x - al * vh.x) * inv * q.x, gA.x); gA.y = fmaf(da, (Kh.y - al * vh.y) * inv * q.y, gA.y);
        gA.z = fmaf(da, (Kh.z - al * vh.z) * inv * q.z, gA.z); gA.w = fmaf(da, (Kh.w - al * vh.w) * inv * q.w, gA.w);
    }
    *reinterpret_cast<float4*>(a.dK[l] + (size_t)k * D + d) = gK;
    *reinterpret_cast<float4*>(a.dA[l] + (size_t)k * D + d) = gA;
}

}  // namespace lc
