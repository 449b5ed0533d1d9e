// Row-wise / layout kernels around the ViT-B/16 GEMMs (core/model/backbone/transformer.py:2222-2261 `VisionTransformer.forward`,
// :1331-1336 block, :169-197 attention).  fp32 residual stream, BF16 GEMM operands.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <math_constants.h>

namespace lc {

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// timm PatchEmbed input side: image NCHW fp32 [B][3][224][224] -> im2col rows bf16 [B*196][768], column = c*256 + py*16 + px
// (the flattening order of the conv weight [768][3][16][16]), so that patch-embed is a plain GEMM.
__global__ void __launch_bounds__(256) patchify_kernel(const float* img, __nv_bfloat16* out, int B) {
    const long long n8 = (long long)B * 196 * 768 / 8;          // 8 consecutive px of one (patch, c, py) per thread
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
        const int col8 = (int)(i % 96), row = (int)(i / 96);     // 768/8 = 96 groups per row
        const int c = col8 / 32, py = (col8 % 32) / 2, px0 = (col8 % 2) * 8;
        const int b = row / 196, p = row % 196, gy = p / 14, gx = p % 14;
        const float* src = img + (((size_t)b * 3 + c) * 224 + gy * 16 + py) * 224 + gx * 16 + px0;
        const float4 a = ldg4(src), d = ldg4(src + 4);
        *reinterpret_cast<uint4*>(out + (size_t)row * 768 + col8 * 8) = make_uint4(pack2_bf16(a.x, a.y), pack2_bf16(a.z, a.w), pack2_bf16(d.x, d.y), pack2_bf16(d.z, d.w));
    }
}

// x[b][row0 + r][:] = src[r][:] (+ add[r][:])   for every batch element: the cls token row (cls_token + pos_embed[0]) and the shared
// L2P prompt rows in front of it
__global__ void __launch_bounds__(192) set_rows_kernel(float* x, long long batch_stride, int row0, const float* src, const float* add, int D) {
    const int b = blockIdx.x, r = blockIdx.y;
    for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
        float4 v = ldg4(src + (size_t)r * D + j);
        if (add != nullptr) { const float4 a = ldg4(add + (size_t)r * D + j); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        *reinterpret_cast<float4*>(x + (size_t)b * batch_stride + (size_t)(row0 + r) * D + j) = v;
    }
}

// LayerNorm over D = 768: one warp per row, fp32 in; bf16 and/or fp32 out.  Two-pass (mean, then centred variance) in registers.
template <int D>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const float* x, const float* gamma, const float* beta, float eps, long long rows,
                                                            __nv_bfloat16* out_bf16, float* out_f32, float* stat /*nullable [rows][2] mean, rstd*/) {
    constexpr int PER = D / 32;       // 24 values per lane, as 6 float4
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * D;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PER / 4; ++k) {
        const float4 t = *reinterpret_cast<const float4*>(xr + (k * 32 + lane) * 4);
        v[k * 4] = t.x; v[k * 4 + 1] = t.y; v[k * 4 + 2] = t.z; v[k * 4 + 3] = t.w;
        s += t.x + t.y + t.z + t.w;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (stat != nullptr && lane == 0) { stat[row * 2] = mean; stat[row * 2 + 1] = rstd; }
#pragma unroll
    for (int k = 0; k < PER / 4; ++k) {
        const int j = (k * 32 + lane) * 4;
        const float4 g = ldg4(gamma + j), b = ldg4(beta + j);
        const float o0 = (v[k * 4] - mean) * rstd * g.x + b.x, o1 = (v[k * 4 + 1] - mean) * rstd * g.y + b.y;
        const float o2 = (v[k * 4 + 2] - mean) * rstd * g.z + b.z, o3 = (v[k * 4 + 3] - mean) * rstd * g.w + b.w;
        if (out_bf16 != nullptr) *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D + j) = make_uint2(pack2_bf16(o0, o1), pack2_bf16(o2, o3));
        if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + (size_t)row * D + j) = make_float4(o0, o1, o2, o3);
    }
}

// softmax over the first T columns of every row of S (fp32, row stride ld), written as bf16 P with the padding columns [T, ld) zeroed
// (they are the K-tail of the P.V GEMM).  One warp per row.
__global__ void __launch_bounds__(128) softmax_rows_kernel(const float* S, __nv_bfloat16* P, long long rows, int T, int ld) {
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = S + (size_t)row * ld;
    float m = -CUDART_INF_F;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) sum += expf(s[j] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __nv_bfloat16* p = P + (size_t)row * ld;
    for (int j = lane; j < ld; j += 32) p[j] = __float2bfloat16_rn(j < T ? expf(s[j] - m) * inv : 0.f);
}

// V of the fused QKV buffer [B][T][3][H][64] (bf16) -> V^T [B*H][64][ld] (keys contiguous, padding keys zeroed): the K-major B operand of P.V
__global__ void __launch_bounds__(256) transpose_v_kernel(const __nv_bfloat16* qkv, __nv_bfloat16* vt, int B, int T, int H, int ld) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int bh = blockIdx.y, b = bh / H, h = bh % H;
    const int t0 = blockIdx.x * 64;
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int tt = e / 64, d = e % 64;
        const int t = t0 + tt;
        tile[tt][d] = t < T ? qkv[((size_t)(b * T + t) * 3 + 2) * (H * 64) + h * 64 + d] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        const int d = e / 64, tt = e % 64;
        if (t0 + tt < ld) vt[((size_t)bh * 64 + d) * ld + t0 + tt] = tile[tt][d];
    }
}

// features[b][:] = mean over rows [r0, r0+nr) of y[b][:][:]  (L2P: the 25 prompt positions; otherwise the cls row) ; fp32
__global__ void __launch_bounds__(192) pool_rows_kernel(const float* y, long long batch_stride, int r0, int nr, int D, float* feat) {
    const int b = blockIdx.x;
    for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < r0 + nr; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(y + (size_t)b * batch_stride + (size_t)r * D + j);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        const float inv = 1.f / (float)nr;
        *reinterpret_cast<float4*>(feat + (size_t)b * D + j) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
}

// small fp32 linear head: logits[b][k] = feat[b] . W[k] + bias[k]   (one warp per output)
__global__ void __launch_bounds__(128) linear_head_kernel(const float* feat, const float* W, const float* bias, int B, int C, int D, float* logits, int ld) {
    const int idx = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= B * C) return;
    const int b = idx / C, k = idx % C;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s = fmaf(feat[(size_t)b * D + j], __ldg(W + (size_t)k * D + j), s);
    s = warp_sum(s);
    if (lane == 0) logits[(size_t)b * ld + k] = s + (bias != nullptr ? bias[k] : 0.f);
}

// fp32 -> bf16 cast (weights, once per load)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* in, __nv_bfloat16* out, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = __float2bfloat16_rn(in[i]);
}

}  // namespace lc
