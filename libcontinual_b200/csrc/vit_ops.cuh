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

// ---- backward (input gradients only: the backbone is frozen, l2p.py:66-71) ------------------------------------------------------------

// LayerNorm backward wrt its input, fused with the residual-path gradient:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dh * gamma ;   out = dx + res
// dh is fp32 [rows][D], or — for the final LayerNorm in front of the L2P prompt-position mean (transformer.py:2253-2258) — the pooled
// gradient broadcast: dh(row (b, t)) = t < n_active ? dfeat[b] / n_active : 0  (dh_pool = dfeat [B][D], T = tokens per image).
// (mean, rstd) are recomputed from x (x is read anyway).  Writes fp32 and/or a BF16 copy (the A operand of the next backward GEMM).
template <int D>
__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const float* dh, const float* dh_pool, int T, int n_active, const float* x, const float* gamma,
                                                            float eps, long long rows, const float* res, float* out_f32, __nv_bfloat16* out_bf16) {
    constexpr int PER = D / 32;
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * D;
    float v[PER], g[PER];
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
    for (int i = 0; i < PER; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    const float* dsrc = nullptr;
    float dscale = 1.f;
    if (dh != nullptr) dsrc = dh + (size_t)row * D;
    else if ((int)(row % T) < n_active) { dsrc = dh_pool + (size_t)(row / T) * D; dscale = 1.f / (float)n_active; }
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int k = 0; k < PER / 4; ++k) {
        const int j = (k * 32 + lane) * 4;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dsrc != nullptr) d = *reinterpret_cast<const float4*>(dsrc + j);
        const float4 gm = ldg4(gamma + j);
        g[k * 4] = d.x * dscale * gm.x; g[k * 4 + 1] = d.y * dscale * gm.y; g[k * 4 + 2] = d.z * dscale * gm.z; g[k * 4 + 3] = d.w * dscale * gm.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[k * 4 + i] *= rstd; sg += g[k * 4 + i]; sgx = fmaf(g[k * 4 + i], v[k * 4 + i], sgx); }
    }
    const float mg = warp_sum(sg) / (float)D, mgx = warp_sum(sgx) / (float)D;
#pragma unroll
    for (int k = 0; k < PER / 4; ++k) {
        const int j = (k * 32 + lane) * 4;
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = rstd * (g[k * 4 + i] - mg - v[k * 4 + i] * mgx);
        if (res != nullptr) {
            const float4 r = *reinterpret_cast<const float4*>(res + (size_t)row * D + j);
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + (size_t)row * D + j) = make_float4(o[0], o[1], o[2], o[3]);
        if (out_bf16 != nullptr) *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D + j) = make_uint2(pack2_bf16(o[0], o[1]), pack2_bf16(o[2], o[3]));
    }
}

// out[r][:] = sum over the batch of x[b][r][:]   (gradient of the shared prompt rows; fixed summation order)
__global__ void __launch_bounds__(192) sum_batch_rows_kernel(const float* x, long long batch_stride, int B, int D, float* out) {
    const int r = blockIdx.x;
    for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < B; ++b) {
            const float4 t = *reinterpret_cast<const float4*>(x + (size_t)b * batch_stride + (size_t)r * D + j);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)r * D + j) = acc;
    }
}

// nn.Linear backward for a wide feature (classifier on the 768-d pooled feature, l2p.py:31-40):
//   blocks [0, ncls): dW[k][:] = sum_n dlogits[n][k] feat[n][:], db[k] ; blocks [ncls, ncls + B): dfeat[n][:] = sum_k dlogits[n][k] W[k][:]
__global__ void __launch_bounds__(192) linear_head_bwd_kernel(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int B, int D,
                                                              float* dW, float* db, float* dfeat) {
    if ((int)blockIdx.x < ncls) {
        const int k = blockIdx.x;
        for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8                                                 // eight independent row loads in flight (the chain of FMAs keeps its order)
            for (int n = 0; n < B; ++n) {
                const float d = __ldg(dlogits + (size_t)n * ldl + k);
                const float4 f = ldg4(feat + (size_t)n * D + j);
                acc.x = fmaf(d, f.x, acc.x); acc.y = fmaf(d, f.y, acc.y); acc.z = fmaf(d, f.z, acc.z); acc.w = fmaf(d, f.w, acc.w);
            }
            *reinterpret_cast<float4*>(dW + (size_t)k * D + j) = acc;
        }
        if (db != nullptr && threadIdx.x == 0) {
            float b = 0.f;
            for (int n = 0; n < B; ++n) b += __ldg(dlogits + (size_t)n * ldl + k);
            db[k] = b;
        }
    } else {
        const int n = blockIdx.x - ncls;
        for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            // classes outside the task mask carry exactly zero gradient: fma(0, w, acc) == acc, so no branch (a branch would serialise the row loads)
#pragma unroll 8
            for (int k = 0; k < ncls; ++k) {
                const float d = __ldg(dlogits + (size_t)n * ldl + k);
                const float4 w = ldg4(W + (size_t)k * D + j);
                acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
            }
            *reinterpret_cast<float4*>(dfeat + (size_t)n * D + j) = acc;
        }
    }
}

// L2P parameter gradients: dpool[P][L][D] = 0 except dpool[ids[t]][l][:] = dprompts[t*L + l][:] (ids are distinct);
// dkey_out = coeff * dkey_in  (coeff = -pull_constraint_coeff: l2p.py:99).  One block per pool row (p, l) + P blocks for the keys.
__global__ void __launch_bounds__(192) l2p_backward_kernel(const float* dprompts, const long long* ids, int pool, int top_k, int L, int D, float* dpool,
                                                           const float* dkey_in, float coeff, float* dkey_out) {
    const int r = blockIdx.x;
    if (r < pool * L) {
        const int p = r / L, l = r % L;
        int t = -1;
        for (int i = 0; i < top_k; ++i) if ((int)ids[i] == p) t = i;
        for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0) v = *reinterpret_cast<const float4*>(dprompts + (size_t)(t * L + l) * D + j);
            *reinterpret_cast<float4*>(dpool + (size_t)r * D + j) = v;
        }
    } else {
        const int p = r - pool * L;
        for (int j = threadIdx.x * 4; j < D; j += 192 * 4) {
            const float4 v = *reinterpret_cast<const float4*>(dkey_in + (size_t)p * D + j);
            *reinterpret_cast<float4*>(dkey_out + (size_t)p * D + j) = make_float4(coeff * v.x, coeff * v.y, coeff * v.z, coeff * v.w);
        }
    }
}

// out[c][r] = in[r][c] for a bf16 matrix [rows][cols] (row stride ld_in) -> [cols][ld_out]; columns r in [rows, ld_out) of `out` are zero-filled
// (the K padding of the token-contraction GEMM h^T h of InfLoRA's input matrix, transformer.py:242-244).  grid (ceil(ld_out/64), ceil(cols/64)), block (64, 4)... 32x32 tiles of 2-byte elements.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* in, long long ld_in, long long rows, int cols, __nv_bfloat16* out, long long ld_out) {
    __shared__ __nv_bfloat16 tile[64][66];
    const long long r0 = (long long)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
        const long long r = r0 + i;
        tile[i][tx] = (r < rows && c0 + tx < cols) ? in[(size_t)r * ld_in + c0 + tx] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i;
        const long long r = r0 + tx;
        if (c < cols && r < ld_out) out[(size_t)c * ld_out + r] = tile[tx][i];
    }
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the two-term BF16 split that lets three BF16 tensor-core products (hi hi + lo hi + hi lo) reproduce an
// fp32 product to ~2^-16 relative (the dropped lo lo term and the residual of the split are both below that)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

}  // namespace lc
