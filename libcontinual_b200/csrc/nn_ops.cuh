// Generic layer kernels around the tcgen05 GEMM for the backbones whose shapes are not the CIFAR-ResNet's: torchvision-style ResNet18
// (core/model/backbone/resnet.py:26-64,110-246; LwF at 64x64) and AlexNet_TRGP (core/model/backbone/alexnet.py:94-156; GPM).
// Activations are NHWC; a conv output is the row-major matrix [M = N*Ho*Wo][Cout] that the GEMM epilogue writes.
//   im2col            : explicit patch matrix (BF16), row-major [M][K] and / or transposed [K][M] (the weight-gradient GEMM contracts over M,
//                       so both of its operands must be M-contiguous); also GPM's representation matrix (gpm.py:157-168) on the device
//   col2im            : data gradient of a strided / unpadded conv from the patch-gradient matrix dcol = dY * W   (gather form, no atomics)
//   bn_colsum/finalize: BatchNorm batch statistics of a [M][C] matrix (two launches, fixed-order fp64 finalize: deterministic)
//   bn_act            : y*scale+shift (+ residual (optionally with its own affine)) -> ReLU -> dropout -> BF16 and / or fp32
//   bn_bwd_colsum/finalize/apply : native_batch_norm_backward + threshold_backward (+ dropout scale) in two passes
//   maxpool fwd/bwd   : nn.MaxPool2d(k, s, p) on NHWC with the argmax kept as a uint8 window index
//   pack_weight       : OIHW fp32 -> BF16 GEMM operands (forward [Cout][K], transposed [K][Cout], flipped data-gradient [Cin][taps][Cout])
//   wgrad_reduce      : split-K partial sums [S][Cout][K] -> fp32 OIHW gradient (fixed order)
// All HBM-bound: coalesced along the contiguous dimension of whatever they write.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace lc {
namespace nn {

__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// counter-based uniform in [0, 1): splitmix64 of (seed, stream offset, element index).  Used for dropout: the mask of a step is a pure function
// of (seed, offset, index), so the test-suite can ask for the very mask a step used (lc_nn_dropout_mask).
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long offset, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (offset + 1) + idx * 0xD1342543DE82EF95ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// ---- im2col -----------------------------------------------------------------------------------------------------------------------------
enum { SRC_NHWC_BF16 = 0, SRC_NCHW_F32 = 1, SRC_NHWC_F32 = 2 };
enum { KORDER_TAP_C = 0, KORDER_C_TAP = 1 };      // k = (kh*ks + kw)*C + c   |   k = (c*ks + kh)*ks + kw  (= weight.view(Cout, -1), gpm.py:79)

struct Im2colArgs {
    const void* src;
    int src_kind;
    int N, H, W, C, ks, stride, pad, Ho, Wo;
    int korder;
    int K, Kp;                     // K = ks*ks*C; rows / columns [K, Kp) are written as zeros
    __nv_bfloat16* col;            // nullable [M][ld_col]
    long long ld_col;
    __nv_bfloat16* colT;           // nullable [Kp][ld_colT]
    long long ld_colT;
    long long M;
};

__device__ __forceinline__ float im2col_fetch(const Im2colArgs& a, long long m, int k) {
    if (m >= a.M || k >= a.K) return 0.f;
    const int wo = (int)(m % a.Wo);
    const long long t = m / a.Wo;
    const int ho = (int)(t % a.Ho);
    const int n = (int)(t / a.Ho);
    int c, kh, kw;
    if (a.korder == KORDER_TAP_C) { c = k % a.C; const int tap = k / a.C; kh = tap / a.ks; kw = tap % a.ks; }
    else { kw = k % a.ks; const int q = k / a.ks; kh = q % a.ks; c = q / a.ks; }
    const int hi = ho * a.stride - a.pad + kh, wi = wo * a.stride - a.pad + kw;
    if (hi < 0 || hi >= a.H || wi < 0 || wi >= a.W) return 0.f;
    if (a.src_kind == SRC_NHWC_BF16) return bf16_to_f(reinterpret_cast<const __nv_bfloat16*>(a.src)[(((size_t)n * a.H + hi) * a.W + wi) * a.C + c]);
    if (a.src_kind == SRC_NCHW_F32) return reinterpret_cast<const float*>(a.src)[(((size_t)n * a.C + c) * a.H + hi) * a.W + wi];
    return reinterpret_cast<const float*>(a.src)[(((size_t)n * a.H + hi) * a.W + wi) * a.C + c];
}

// 64 (m) x 64 (k) tile per block: the load phase runs k-fastest (coalesced reads of NHWC channels, coalesced writes of `col`), the transposed
// copy leaves through shared memory m-fastest
template <typename OutT> __device__ __forceinline__ OutT im2col_cvt(float v);
template <> __device__ __forceinline__ __nv_bfloat16 im2col_cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float im2col_cvt<float>(float v) { return v; }

// OutT = __nv_bfloat16 (GEMM operands) or float (GPM's representation matrices, which feed an SVD: gpm.py:157-168)
template <typename OutT>
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs a) {
    __shared__ OutT tile[64][65];
    OutT* col = reinterpret_cast<OutT*>(a.col);
    OutT* colT = reinterpret_cast<OutT*>(a.colT);
    const long long m0 = (long long)blockIdx.x * 64;
    const int k0 = blockIdx.y * 64;
    const int lo = threadIdx.x & 63, hi4 = threadIdx.x >> 6;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int ml = hi4 + 4 * i;
        const long long m = m0 + ml;
        const int k = k0 + lo;
        const OutT v = im2col_cvt<OutT>(im2col_fetch(a, m, k));
        tile[ml][lo] = v;
        if (col != nullptr && m < a.M && k < a.Kp) col[(size_t)m * a.ld_col + k] = v;
    }
    if (colT == nullptr) return;
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int kl = hi4 + 4 * i;
        const int k = k0 + kl;
        const long long m = m0 + lo;
        if (k < a.Kp && m < a.ld_colT) colT[(size_t)k * a.ld_colT + m] = tile[lo][kl];     // columns [M, ld_colT) get the zeros of im2col_fetch
    }
}

// Fast path of the transposed patch matrix (the B operand of every weight-gradient GEMM): NHWC BF16 source, C % 64 == 0, tap-major K.  A block moves
// 64 pixels x 64 channels of ONE tap: 16-byte loads of whole channel runs (two index decodes per thread instead of one per element), a padded
// shared-memory transpose, 16-byte stores of 8 consecutive pixels per channel row.  grid (ceil(ld_colT / 64), taps * C / 64).
__global__ void __launch_bounds__(256) im2colT_nhwc_kernel(Im2colArgs a) {
    __shared__ __align__(16) __nv_bfloat16 tile[64][72];
    const long long m0 = (long long)blockIdx.x * 64;
    const int cch = a.C >> 6;
    const int tap = blockIdx.y / cch, c0 = (blockIdx.y - tap * cch) << 6;
    const int kh = tap / a.ks, kw = tap - kh * a.ks;
    const int v = threadIdx.x & 7, r = threadIdx.x >> 3;
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int row = r + 32 * half;
        const long long m = m0 + row;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (m < a.M) {
            const int wo = (int)(m % a.Wo);
            const long long t = m / a.Wo;
            const int ho = (int)(t % a.Ho);
            const int n = (int)(t / a.Ho);
            const int hi = ho * a.stride - a.pad + kh, wi = wo * a.stride - a.pad + kw;
            if (hi >= 0 && hi < a.H && wi >= 0 && wi < a.W)
                val = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)n * a.H + hi) * a.W + wi) * a.C + c0 + v * 8));
        }
        *reinterpret_cast<uint4*>(&tile[row][v * 8]) = val;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cr = r + 32 * half;                       // channel row of this tap
        const long long m = m0 + v * 8;
        if (m >= a.ld_colT) continue;
        __align__(16) __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = tile[v * 8 + j][cr];
        __nv_bfloat16* dst = a.colT + (size_t)(tap * a.C + c0 + cr) * a.ld_colT + m;
        if (m + 8 <= a.ld_colT) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
        else for (int j = 0; j < 8 && m + j < a.ld_colT; ++j) dst[j] = o[j];
    }
}

// ---- col2im: dX[n][hi][wi][c] = (addend) + sum over (kh, kw) with (hi + pad - kh) % stride == 0 of dcol[(n, ho, wo)][k(kh, kw, c)] -----------
struct Col2imArgs {
    const __nv_bfloat16* dcol;     // [M][ld]
    long long ld;
    const float* addend;           // nullable, same shape as dx
    float* dx;                     // fp32 NHWC [N][H][W][C]
    int N, H, W, C, ks, stride, pad, Ho, Wo, korder;
};
__global__ void __launch_bounds__(256) col2im_kernel(Col2imArgs a) {
    const long long total = (long long)a.N * a.H * a.W * a.C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % a.C);
        long long t = e / a.C;
        const int wi = (int)(t % a.W); t /= a.W;
        const int hi = (int)(t % a.H);
        const int n = (int)(t / a.H);
        float acc = a.addend != nullptr ? a.addend[e] : 0.f;
        // only taps with (hi + pad - kh) % stride == 0 contribute: start at the right residue and step by the stride
        for (int kh = (hi + a.pad) % a.stride; kh < a.ks; kh += a.stride) {
            const int hn = hi + a.pad - kh;
            if (hn < 0) break;
            const int ho = hn / a.stride;
            if (ho >= a.Ho) continue;
            for (int kw = (wi + a.pad) % a.stride; kw < a.ks; kw += a.stride) {
                const int wn = wi + a.pad - kw;
                if (wn < 0) break;
                const int wo = wn / a.stride;
                if (wo >= a.Wo) continue;
                const int k = a.korder == KORDER_TAP_C ? (kh * a.ks + kw) * a.C + c : (c * a.ks + kh) * a.ks + kw;
                acc += bf16_to_f(a.dcol[(((size_t)n * a.Ho + ho) * a.Wo + wo) * a.ld + k]);
            }
        }
        a.dx[e] = acc;
    }
}

// ---- BatchNorm forward statistics of y[M][C] -------------------------------------------------------------------------------------------------
// grid (GX, C/64), 256 threads = 16 column quads x 16 row lanes; partial[(bx*2 + {0,1})*C + c] = sum y, sum y^2 over the block's rows
__global__ void __launch_bounds__(256) bn_colsum_kernel(const float* __restrict__ y, long long M, int C, float* __restrict__ partial) {
    __shared__ float4 s1[256], s2[256];
    const int cq = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const int c = blockIdx.y * 64 + cq * 4;
    const long long per = (M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < M ? r0 + per : M;
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    for (long long r = r0 + rl; r < r1; r += 16) {
        const float4 v = ldg4(y + (size_t)r * C + c);
        a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
        a2.x = fmaf(v.x, v.x, a2.x); a2.y = fmaf(v.y, v.y, a2.y); a2.z = fmaf(v.z, v.z, a2.z); a2.w = fmaf(v.w, v.w, a2.w);
    }
    s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
    __syncthreads();
    for (int off = 8; off > 0; off >>= 1) {
        if (rl < off) {
            float4 p = s1[threadIdx.x], q = s1[threadIdx.x + off * 16];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w; s1[threadIdx.x] = p;
            p = s2[threadIdx.x]; q = s2[threadIdx.x + off * 16];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w; s2[threadIdx.x] = p;
        }
        __syncthreads();
    }
    if (rl == 0) {
        *reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 2 + 0) * C + c) = s1[threadIdx.x];
        *reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 2 + 1) * C + c) = s2[threadIdx.x];
    }
}
// aff = [scale | shift | mean | invstd] (4*C); running = [mean | var] (nullable); unbiased variance for the running update (nn.BatchNorm)
// fixed-order sum of the per-block partials of channel c: 8 slices (threadIdx.y) of the partial rows, then a fixed-order combine in shared memory
// block (32, 8): 32 channels per block
__device__ __forceinline__ void colsum_combine(const float* __restrict__ partial, int nparts, int C, int c, double& S1, double& S2) {
    __shared__ double sh[2][8][32];
    double a1 = 0.0, a2 = 0.0;
    if (c < C)
        for (int p = threadIdx.y; p < nparts; p += 8) { a1 += (double)partial[((size_t)p * 2 + 0) * C + c]; a2 += (double)partial[((size_t)p * 2 + 1) * C + c]; }
    sh[0][threadIdx.y][threadIdx.x] = a1; sh[1][threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    S1 = 0.0; S2 = 0.0;
    for (int q = 0; q < 8; ++q) { S1 += sh[0][q][threadIdx.x]; S2 += sh[1][q][threadIdx.x]; }
}
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nparts, long long M, int C, const float* gamma, const float* beta, float eps,
                                   float momentum, float* running, float* aff) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double S1, S2;
    colsum_combine(partial, nparts, C, c, S1, S2);
    if (c >= C || threadIdx.y != 0) return;
    const double mean = S1 / (double)M;
    double var = S2 / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const double istd = 1.0 / sqrt(var + (double)eps);
    const float sc = (float)((double)(gamma ? gamma[c] : 1.f) * istd);
    aff[c] = sc;
    aff[C + c] = (float)((double)(beta ? beta[c] : 0.f) - mean * (double)sc);
    aff[2 * C + c] = (float)mean;
    aff[3 * C + c] = (float)istd;
    if (running != nullptr) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running[c] = (float)((1.0 - momentum) * (double)running[c] + momentum * mean);
        running[C + c] = (float)((1.0 - momentum) * (double)running[C + c] + momentum * unb);
    }
}
// eval mode: the same affine from running statistics
__global__ void bn_eval_affine_kernel2(const float* running, int C, const float* gamma, const float* beta, float eps, float* aff) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double istd = 1.0 / sqrt((double)running[C + c] + (double)eps);
    const float sc = (float)((double)(gamma ? gamma[c] : 1.f) * istd);
    aff[c] = sc;
    aff[C + c] = (beta ? beta[c] : 0.f) - running[c] * sc;
    aff[2 * C + c] = running[c];
    aff[3 * C + c] = (float)istd;
}

// ---- BN apply (+ residual) + ReLU + dropout ---------------------------------------------------------------------------------------------------
struct BnActArgs2 {
    const float* y;                // [M][C]
    const float* aff;              // scale = aff, shift = aff + C
    const float* res;              // nullable fp32 [M][C]
    const float* res_aff;          // nullable: the residual is itself a raw conv output with its own BN affine (downsample path)
    __nv_bfloat16* out_bf16;       // nullable
    float* out_f32;                // nullable
    const unsigned long long* rng; // nullable: {seed, offset} on the device -> dropout with keep probability 1 - drop_p, survivors / (1 - drop_p)
    unsigned long long rng_stream; // distinguishes the dropout layers of one step
    float drop_p;
    long long M;
    int C, relu;
};
__global__ void __launch_bounds__(256) bn_act_kernel(BnActArgs2 a) {
    const int c4n = a.C >> 2;
    const long long n4 = a.M * c4n;
    unsigned long long seed = 0, off = 0;
    const bool drop = a.rng != nullptr && a.drop_p > 0.f;
    if (drop) { seed = a.rng[0]; off = a.rng[1] * 64ull + a.rng_stream; }
    const float keep_scale = drop ? 1.f / (1.f - a.drop_p) : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        float4 v = ldg4(a.y + i * 4);
        const float4 sc = ldg4(a.aff + c), sh = ldg4(a.aff + a.C + c);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (a.res != nullptr) {
            float4 r = ldg4(a.res + i * 4);
            if (a.res_aff != nullptr) {
                const float4 rs = ldg4(a.res_aff + c), rh = ldg4(a.res_aff + a.C + c);
                r.x = fmaf(r.x, rs.x, rh.x); r.y = fmaf(r.y, rs.y, rh.y); r.z = fmaf(r.z, rs.z, rh.z); r.w = fmaf(r.w, rs.w, rh.w);
            }
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (drop) {
            const unsigned long long e = (unsigned long long)i * 4;
            v.x = uniform01(seed, off, e) >= a.drop_p ? v.x * keep_scale : 0.f;
            v.y = uniform01(seed, off, e + 1) >= a.drop_p ? v.y * keep_scale : 0.f;
            v.z = uniform01(seed, off, e + 2) >= a.drop_p ? v.z * keep_scale : 0.f;
            v.w = uniform01(seed, off, e + 3) >= a.drop_p ? v.w * keep_scale : 0.f;
        }
        if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + i * 4) = v;
        if (a.out_bf16 != nullptr) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned int*>(&lo); pk.y = *reinterpret_cast<const unsigned int*>(&hi);
            *reinterpret_cast<uint2*>(a.out_bf16 + i * 4) = pk;
        }
    }
}
__global__ void dropout_mask_kernel(const unsigned long long* rng, unsigned long long rng_stream, float drop_p, long long n, unsigned char* keep) {
    const unsigned long long seed = rng[0], off = rng[1] * 64ull + rng_stream;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        keep[i] = uniform01(seed, off, (unsigned long long)i) >= drop_p ? 1 : 0;
}
__global__ void rng_advance_kernel(unsigned long long* rng) { if (threadIdx.x == 0 && blockIdx.x == 0) rng[1] += 1ull; }

// ---- BN backward ------------------------------------------------------------------------------------------------------------------------------
// dz = g * [act > 0] * gscale (ReLU + dropout in one mask: a dropped unit is stored as 0 and a kept one carries 1/(1-p));
// pass 1: S1[c] = sum dz, S2[c] = sum dz * xhat ; pass 2: dy = scale * (dz - S1/M - xhat * S2/M)
struct BnBwdArgs2 {
    const float* g;                // [M][C] gradient w.r.t. the layer output (after ReLU / dropout)
    const float* act_f32;          // nullable mask source (the stored output); exactly one of act_f32 / act_bf16 when relu_mask
    const __nv_bfloat16* act_bf16;
    const float* y;                // raw conv output
    const float* aff;              // scale | shift | mean | invstd
    float* partial;                // [GX][2][C]
    float* coef;                   // [3][C]  dy = c0*dz + c1*y + c2
    float* dgamma; float* dbeta;   // nullable [C]
    __nv_bfloat16* dy_bf16;        // nullable
    float* dy_f32;                 // nullable
    float* dz_out;                 // nullable: the masked gradient (what flows into an identity shortcut)
    long long M;
    int C, relu_mask;
    float gscale;
};
__device__ __forceinline__ float4 bn2_masked(const BnBwdArgs2& a, long long e) {
    float4 g = ldg4(a.g + e);
    if (a.relu_mask) {
        float4 o;
        if (a.act_f32 != nullptr) o = ldg4(a.act_f32 + e);
        else {
            const uint2 pk = __ldg(reinterpret_cast<const uint2*>(a.act_bf16 + e));
            o.x = __uint_as_float(pk.x << 16); o.y = __uint_as_float(pk.x & 0xffff0000u); o.z = __uint_as_float(pk.y << 16); o.w = __uint_as_float(pk.y & 0xffff0000u);
        }
        g.x = o.x > 0.f ? g.x * a.gscale : 0.f; g.y = o.y > 0.f ? g.y * a.gscale : 0.f;
        g.z = o.z > 0.f ? g.z * a.gscale : 0.f; g.w = o.w > 0.f ? g.w * a.gscale : 0.f;
    }
    return g;
}
__global__ void __launch_bounds__(256) bn_bwd_colsum_kernel(BnBwdArgs2 a) {
    __shared__ float4 s1[256], s2[256];
    const int cq = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const int c = blockIdx.y * 64 + cq * 4;
    const long long per = (a.M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < a.M ? r0 + per : a.M;
    const float4 mu = ldg4(a.aff + 2 * a.C + c), is = ldg4(a.aff + 3 * a.C + c);
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    for (long long r = r0 + rl; r < r1; r += 16) {
        const long long e = r * a.C + c;
        const float4 yv = ldg4(a.y + e);
        const float4 g = bn2_masked(a, e);
        a1.x += g.x; a1.y += g.y; a1.z += g.z; a1.w += g.w;
        a2.x = fmaf(g.x, (yv.x - mu.x) * is.x, a2.x); a2.y = fmaf(g.y, (yv.y - mu.y) * is.y, a2.y);
        a2.z = fmaf(g.z, (yv.z - mu.z) * is.z, a2.z); a2.w = fmaf(g.w, (yv.w - mu.w) * is.w, a2.w);
    }
    s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
    __syncthreads();
    for (int off = 8; off > 0; off >>= 1) {
        if (rl < off) {
            float4 p = s1[threadIdx.x], q = s1[threadIdx.x + off * 16];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w; s1[threadIdx.x] = p;
            p = s2[threadIdx.x]; q = s2[threadIdx.x + off * 16];
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w; s2[threadIdx.x] = p;
        }
        __syncthreads();
    }
    if (rl == 0) {
        *reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.x * 2 + 0) * a.C + c) = s1[threadIdx.x];
        *reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.x * 2 + 1) * a.C + c) = s2[threadIdx.x];
    }
}
__global__ void bn_bwd_finalize_kernel(BnBwdArgs2 a, int nparts) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double S1, S2;
    colsum_combine(a.partial, nparts, a.C, c, S1, S2);
    if (c >= a.C || threadIdx.y != 0) return;
    const double N = (double)a.M, sc = (double)a.aff[c], m = (double)a.aff[2 * a.C + c], istd = (double)a.aff[3 * a.C + c];
    const double c1 = -sc * S2 / N * istd;
    a.coef[c] = (float)sc;
    a.coef[a.C + c] = (float)c1;
    a.coef[2 * a.C + c] = (float)(-sc * S1 / N - c1 * m);
    if (a.dgamma != nullptr) a.dgamma[c] = (float)S2;
    if (a.dbeta != nullptr) a.dbeta[c] = (float)S1;
}
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel2(BnBwdArgs2 a) {
    const int C = a.C, c4n = C >> 2;
    const long long n4 = a.M * c4n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const long long e = i * 4;
        const float4 yv = ldg4(a.y + e);
        const float4 g = bn2_masked(a, e);
        const float4 c0 = ldg4(a.coef + c), c1 = ldg4(a.coef + C + c), c2 = ldg4(a.coef + 2 * C + c);
        float4 d;
        d.x = fmaf(c0.x, g.x, fmaf(c1.x, yv.x, c2.x)); d.y = fmaf(c0.y, g.y, fmaf(c1.y, yv.y, c2.y));
        d.z = fmaf(c0.z, g.z, fmaf(c1.z, yv.z, c2.z)); d.w = fmaf(c0.w, g.w, fmaf(c1.w, yv.w, c2.w));
        if (a.dy_f32 != nullptr) *reinterpret_cast<float4*>(a.dy_f32 + e) = d;
        if (a.dy_bf16 != nullptr) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned int*>(&lo); pk.y = *reinterpret_cast<const unsigned int*>(&hi);
            *reinterpret_cast<uint2*>(a.dy_bf16 + e) = pk;
        }
        if (a.dz_out != nullptr) *reinterpret_cast<float4*>(a.dz_out + e) = g;
    }
}

// ---- max pooling on NHWC (nn.MaxPool2d(k, s, p), floor mode) -----------------------------------------------------------------------------------
struct PoolArgs {
    const float* in;               // fp32 [N][H][W][C]
    float* out_f32;                // nullable [N][Ho][Wo][C]
    __nv_bfloat16* out_bf16;       // nullable
    unsigned char* idx;            // [N][Ho][Wo][C] window index kh*k + kw of the maximum (first maximum in scan order, like ATen)
    int N, H, W, C, k, stride, pad, Ho, Wo;
};
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(PoolArgs a) {      // one thread = 4 channels of one output pixel (C % 4 == 0)
    const int c4n = a.C >> 2;
    const long long total = (long long)a.N * a.Ho * a.Wo * c4n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % c4n) * 4;
        long long t = e / c4n;
        const int wo = (int)(t % a.Wo); t /= a.Wo;
        const int ho = (int)(t % a.Ho);
        const int n = (int)(t / a.Ho);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        bool any = false;
        for (int kh = 0; kh < a.k; ++kh) {
            const int hi = ho * a.stride - a.pad + kh;
            if (hi < 0 || hi >= a.H) continue;
            for (int kw = 0; kw < a.k; ++kw) {
                const int wi = wo * a.stride - a.pad + kw;
                if (wi < 0 || wi >= a.W) continue;
                const float4 v = ldg4(a.in + (((size_t)n * a.H + hi) * a.W + wi) * a.C + c);
                const int id = kh * a.k + kw;
                if (!any || v.x > best.x) { best.x = v.x; b0 = id; }
                if (!any || v.y > best.y) { best.y = v.y; b1 = id; }
                if (!any || v.z > best.z) { best.z = v.z; b2 = id; }
                if (!any || v.w > best.w) { best.w = v.w; b3 = id; }
                any = true;
            }
        }
        const size_t o = (size_t)e * 4;
        if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + o) = best;
        if (a.out_bf16 != nullptr) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(best.x, best.y), hi2 = __floats2bfloat162_rn(best.z, best.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned int*>(&lo); pk.y = *reinterpret_cast<const unsigned int*>(&hi2);
            *reinterpret_cast<uint2*>(a.out_bf16 + o) = pk;
        }
        *reinterpret_cast<uchar4*>(a.idx + o) = make_uchar4((unsigned char)b0, (unsigned char)b1, (unsigned char)b2, (unsigned char)b3);
    }
}
// gather form: every input element sums the gradients of the windows whose argmax it is
struct PoolBwdArgs {
    const float* g;                // [N][Ho][Wo][C]
    const unsigned char* idx;
    float* dx;                     // [N][H][W][C]
    int N, H, W, C, k, stride, pad, Ho, Wo;
};
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(PoolBwdArgs a) {      // one thread = 4 channels of one input pixel
    const int c4n = a.C >> 2;
    const long long total = (long long)a.N * a.H * a.W * c4n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % c4n) * 4;
        long long t = e / c4n;
        const int wi = (int)(t % a.W); t /= a.W;
        const int hi = (int)(t % a.H);
        const int n = (int)(t / a.H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kh = (hi + a.pad) % a.stride; kh < a.k; kh += a.stride) {        // windows whose row kh lands on hi
            const int hn = hi + a.pad - kh;
            if (hn < 0) break;
            const int ho = hn / a.stride;
            if (ho >= a.Ho) continue;
            for (int kw = (wi + a.pad) % a.stride; kw < a.k; kw += a.stride) {
                const int wn = wi + a.pad - kw;
                if (wn < 0) break;
                const int wo = wn / a.stride;
                if (wo >= a.Wo) continue;
                const size_t o = (((size_t)n * a.Ho + ho) * a.Wo + wo) * a.C + c;
                const uchar4 id = *reinterpret_cast<const uchar4*>(a.idx + o);
                const float4 g = ldg4(a.g + o);
                const int me = kh * a.k + kw;
                if (id.x == me) acc.x += g.x;
                if (id.y == me) acc.y += g.y;
                if (id.z == me) acc.z += g.z;
                if (id.w == me) acc.w += g.w;
            }
        }
        *reinterpret_cast<float4*>(a.dx + (size_t)e * 4) = acc;
    }
}

// global average pool of NHWC fp32 [N][HW][C] -> [N][C], and its backward (broadcast / HW) added into dx or written
__global__ void __launch_bounds__(256) avgpool_nhwc_fwd_kernel(const float* in, int N, int HW, int C, float* out) {
    const long long total = (long long)N * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int n = (int)(e / C);
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += in[((size_t)n * HW + p) * C + c];
        out[e] = s / (float)HW;
    }
}
__global__ void __launch_bounds__(256) avgpool_nhwc_bwd_kernel(const float* dfeat, int N, int HW, int C, float* dx) {
    const long long total = (long long)N * HW * C;
    const float inv = 1.f / (float)HW;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int n = (int)(e / ((long long)HW * C));
        dx[e] = dfeat[(size_t)n * C + c] * inv;
    }
}

// ---- weights ------------------------------------------------------------------------------------------------------------------------------------
enum { PACK_FWD = 0, PACK_TRANSPOSED = 1, PACK_DGRAD_FLIPPED = 2 };
// w: fp32 [Cout][Cin][ks][ks] (a Linear is ks = 1).
//   PACK_FWD           out[co][k]                       k in `korder`, row length ld (>= K; [K, ld) zero)
//   PACK_TRANSPOSED    out[k][co]                       row length ld (>= Cout)                      (dcol = dY * W: B operand rows = k)
//   PACK_DGRAD_FLIPPED out[ci][((ks-1-kh)*ks + (ks-1-kw))*Cout + co]   row length ld                   (dX = conv(dY, flipped W), stride 1)
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int ks, int korder, int mode, __nv_bfloat16* out,
                                                           long long ld, long long rows) {
    const long long total = rows * ld;
    const int K = Cin * ks * ks;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / ld;
        const int col = (int)(e % ld);
        float v = 0.f;
        if (mode == PACK_FWD) {
            if (col < K && r < Cout) {
                int c, kh, kw;
                if (korder == KORDER_TAP_C) { c = col % Cin; const int tap = col / Cin; kh = tap / ks; kw = tap % ks; }
                else { kw = col % ks; const int q = col / ks; kh = q % ks; c = q / ks; }
                v = w[(((size_t)r * Cin + c) * ks + kh) * ks + kw];
            }
        } else if (mode == PACK_TRANSPOSED) {
            if (col < Cout && r < K) {
                const int k = (int)r;
                int c, kh, kw;
                if (korder == KORDER_TAP_C) { c = k % Cin; const int tap = k / Cin; kh = tap / ks; kw = tap % ks; }
                else { kw = k % ks; const int q = k / ks; kh = q % ks; c = q / ks; }
                v = w[(((size_t)col * Cin + c) * ks + kh) * ks + kw];
            }
        } else {
            if (col < ks * ks * Cout && r < Cin) {
                const int co = col % Cout, tapf = col / Cout;
                const int kh = ks - 1 - tapf / ks, kw = ks - 1 - tapf % ks;
                v = w[(((size_t)co * Cin + (int)r) * ks + kh) * ks + kw];
            }
        }
        out[e] = __float2bfloat16_rn(v);
    }
}

// dW (OIHW fp32) = sum over S split-K partials [S][Cout][ldp] whose columns are in `korder`
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int S, int Cout, int Cin, int ks, int korder, long long ldp,
                                                            float* __restrict__ dw) {
    const int K = Cin * ks * ks;
    const long long total = (long long)Cout * K;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(e / K);
        const int j = (int)(e % K);             // OIHW flat index within the filter: (c*ks + kh)*ks + kw
        int k = j;
        if (korder == KORDER_TAP_C) { const int kw = j % ks; const int q = j / ks; const int kh = q % ks; const int c = q / ks; k = (kh * ks + kw) * Cin + c; }
        float s = 0.f;
        for (int p = 0; p < S; ++p) s += partial[((size_t)p * Cout + co) * ldp + k];
        dw[e] = s;
    }
}

// fp32 [rows][cols] -> bf16 [rows][ld] (columns [cols, ld) zero) and / or its transpose bf16 [cols][ldT] (columns [rows, ldT) zero)
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ src, long long rows, int cols, __nv_bfloat16* out, long long ld,
                                                              __nv_bfloat16* outT, long long ldT) {
    __shared__ __nv_bfloat16 tile[64][66];
    const long long r0 = (long long)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const int lo = threadIdx.x & 63, hi4 = threadIdx.x >> 6;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int rl = hi4 + 4 * i;
        const long long r = r0 + rl;
        const int c = c0 + lo;
        const __nv_bfloat16 v = __float2bfloat16_rn((r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f);
        tile[rl][lo] = v;
        if (out != nullptr && r < rows && c < ld) out[(size_t)r * ld + c] = v;
    }
    if (outT == nullptr) return;
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int cl = hi4 + 4 * i;
        const int c = c0 + cl;
        const long long r = r0 + lo;
        if (c < cols && r < ldT) outT[(size_t)c * ldT + r] = tile[lo][cl];
    }
}

static inline int nn_grid(long long n, int per_block = 256) {
    long long b = (n + per_block - 1) / per_block;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace nn
}  // namespace lc
