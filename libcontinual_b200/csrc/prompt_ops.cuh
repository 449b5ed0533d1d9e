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
