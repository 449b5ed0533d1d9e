// fp32 NHWC direct-convolution kernels (CUDA-core path).  This is the "exact" arithmetic class: plain fp32 FMA,
// deterministic reduction order, used as the parity anchor against the CPU oracle and by the exact-mode step.
//
// Replaces, for the CIFAR ResNet of core/model/backbone/resnet.py:289-412, the cuDNN fprop / dgrad / wgrad calls made
// by nn.Conv2d(k=3, pad=1, bias=False) (resnet.py:295,298,334) and the 1x1 stride-2 shortcut conv (resnet.py:365-366),
// with the BatchNorm batch-statistics reduction (resnet.py:296,299,335) fused into the conv epilogue and the
// BN-apply + ReLU of the producer layer fused into the conv prologue.
#pragma once
#include "common.cuh"

namespace lc {

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm statistics: per-CTA partial (sum, sumsq) -> last CTA reduces in double, writes the affine form and
// updates running statistics (nn.BatchNorm2d train-mode semantics: biased var for normalisation, unbiased for the
// running estimate, momentum 0.1).
// ---------------------------------------------------------------------------------------------------------------
struct BnStatArgs {
    float* partial;            // [nparts][2][C]  (null -> no statistics)
    unsigned int* counter;     // self-resetting election counter
    const float* gamma;        // [C]
    const float* beta;         // [C]
    float* running_mean;       // [C]  (updated when update_running != 0)
    float* running_var;        // [C]
    float* scale;              // out [C] : gamma * invstd
    float* shift;              // out [C] : beta - mean * gamma * invstd
    float* mean;               // out [C]
    float* invstd;             // out [C]
    float momentum;
    float eps;
    int update_running;
    int defer;                 // 1: only write this CTA's partial row; the consumers reduce the rows themselves (BnLazy) — no election, no serial tail
};

// Consumer-side BatchNorm finalisation.  A convolution that produced `partial` rows ([nparts][2][C]: per-CTA sum and sum of squares of its output)
// with stat.defer = 1 leaves the reduction to whoever needs scale / shift next: every consumer CTA sums the rows itself, in the same fixed order and
// in fp64, so all CTAs (and the end-of-forward bn_finalize_layers_kernel that serves the backward pass and the running statistics) obtain bit-identical
// coefficients — and the producer has no threadfence + atomic + last-CTA tail (measured: ~7 us of a 17 us stage-1 convolution).
struct BnLazy {
    const float* partial;      // null -> not lazy (use the finalised scale / shift arrays)
    const float* gamma;
    const float* beta;
    int nparts;
    float count;               // elements per channel (B*H*W)
    float eps;
};

// Fixed-order fp64 column sums of partial[nparts][2][C] by a 256-thread CTA.  red: 1024 doubles of shared memory; on return red[0..2C) holds the sums
// (visible to all threads).  C in {16, 32, 64}.
// Phase 1 (no barrier): this thread's slice of the rows -> red.  Phase 2 (three barriers): combine.  Split so that a kernel can put independent
// work (its tile copies) between the two and so that the 12 row registers of phase 1 are dead before that work's registers go live.
__device__ __forceinline__ void bn_partial_sums_load(const float* partial, int nparts, int C, double* red) {
    if (threadIdx.x >= 256) return;                    // CTAs may carry extra (non-worker) warps: the reduction is laid out for 256 threads
    const int CQ = C >> 1, NSL = 256 / CQ;             // float4 column groups; row slices
    const int cq = threadIdx.x % CQ, sl = threadIdx.x / CQ;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const float4* base = reinterpret_cast<const float4*>(partial) + cq;
    // batches of 12 rows: all loads of a batch are issued before the first add, so a batch costs one L2 round trip (B = 128: one batch per thread)
    for (int p0 = sl; p0 < nparts; p0 += 12 * NSL) {
        float4 v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int p = p0 + k * NSL;
            v[k] = p < nparts ? __ldcg(base + (size_t)p * CQ) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) { a0 += (double)v[k].x; a1 += (double)v[k].y; a2 += (double)v[k].z; a3 += (double)v[k].w; }
    }
    red[threadIdx.x * 4 + 0] = a0; red[threadIdx.x * 4 + 1] = a1; red[threadIdx.x * 4 + 2] = a2; red[threadIdx.x * 4 + 3] = a3;
}
__device__ __forceinline__ void bn_partial_sums_finish(int C, double* red) {
    const int CQ = C >> 1, NSL = 256 / CQ;
    __syncthreads();
    double t = 0.0;
    if ((int)threadIdx.x < 2 * C) {
        const int cqi = threadIdx.x >> 2, k = threadIdx.x & 3;
        for (int q = 0; q < NSL; ++q) t += red[(q * CQ + cqi) * 4 + k];
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * C) red[threadIdx.x] = t;
    __syncthreads();
}
__device__ __forceinline__ void bn_partial_sums(const float* partial, int nparts, int C, double* red) {
    bn_partial_sums_load(partial, nparts, C, red);
    bn_partial_sums_finish(C, red);
}

// scale / shift / mean / invstd of one channel from its sums (the one place this arithmetic lives for the lazy path)
__device__ __forceinline__ void bn_affine_from_sums(double s1, double s2, double count, float gamma, float beta, float eps, float* sc, float* sh, float* mean,
                                                    float* invstd, double* var_out) {
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    const double istd = 1.0 / sqrt(var + (double)eps);
    *sc = (float)((double)gamma * istd);
    *sh = (float)((double)beta - m * (double)gamma * istd);
    *mean = (float)m;
    *invstd = (float)istd;
    *var_out = var;
}

// lazy scale / shift into shared memory: s_aff[0..C) = scale, s_aff[C..2C) = shift (all 256 threads must call; ends with a barrier)
__device__ __forceinline__ void bn_lazy_affine_finish(const BnLazy& z, int C, double* red, float* s_aff) {
    bn_partial_sums_finish(C, red);
    if ((int)threadIdx.x < C) {
        float sc, sh, m, is; double var;
        bn_affine_from_sums(red[threadIdx.x], red[C + threadIdx.x], (double)z.count, z.gamma[threadIdx.x], z.beta[threadIdx.x], z.eps, &sc, &sh, &m, &is, &var);
        s_aff[threadIdx.x] = sc; s_aff[C + threadIdx.x] = sh;
    }
    __syncthreads();
}
__device__ __forceinline__ void bn_lazy_affine(const BnLazy& z, int C, double* red, float* s_aff) {
    bn_partial_sums_load(z.partial, z.nparts, C, red);
    bn_lazy_affine_finish(z, C, red, s_aff);
}

// The same for the BatchNorm BACKWARD sums: partial rows [nparts][2][C] = per-CTA (sum g, sum g*xhat) left by the fused epilogue of a tensor-core
// data-gradient conv.  Consumers (the next data-gradient conv, the weight-gradient kernel, bn_bwd_apply_kernel) derive dy = c0*g + c1*y + c2 themselves;
// CTA 0 of a consumer launched with write_grads also stores dgamma / dbeta.
struct BnBwdLazy {
    const float* partial;      // null -> not lazy (use the finalised coefficient array)
    const float* scale;        // [C] forward affine of the BatchNorm
    const float* mean;
    const float* invstd;
    float* dgamma;             // [C]
    float* dbeta;
    int nparts;
    float count;
    int write_grads;
};
// s_coef[0..3C) = c0, c1, c2 (all 256 threads must call; ends with a barrier)
__device__ __forceinline__ void bn_bwd_lazy_coef_finish(const BnBwdLazy& z, int C, double* red, float* s_coef) {
    bn_partial_sums_finish(C, red);
    if ((int)threadIdx.x < C) {
        const int ch = threadIdx.x;
        const double S1 = red[ch], S2 = red[C + ch], N = (double)z.count;
        const double sc = (double)z.scale[ch], istd = (double)z.invstd[ch], m = (double)z.mean[ch];
        const double c1 = -sc * S2 / N * istd;
        s_coef[ch] = (float)sc;
        s_coef[C + ch] = (float)c1;
        s_coef[2 * C + ch] = (float)(-sc * S1 / N - c1 * m);
        if (z.write_grads && blockIdx.x == 0) { z.dgamma[ch] = (float)S2; z.dbeta[ch] = (float)S1; }
    }
    __syncthreads();
}
__device__ __forceinline__ void bn_bwd_lazy_coef(const BnBwdLazy& z, int C, double* red, float* s_coef) {
    bn_partial_sums_load(z.partial, z.nparts, C, red);
    bn_bwd_lazy_coef_finish(z, C, red, s_coef);
}

// End-of-forward finalisation of every deferred layer (one CTA per layer): scale / shift / mean / invstd for the backward pass + running statistics.
struct BnFinEntry { long long part_off, gamma_off, beta_off, rstat_off, aff_off; int C, pp, mrows, hw, tmax, nsm; };   // nparts = ceil(B*pp / mrows), or (tmax > 0)
                                                                                                                        // the persistent conv's grid; count = B*hw
static __global__ void __launch_bounds__(256) bn_finalize_layers_kernel(const BnFinEntry* tab, const float* params, float* rstat, float* ws, int batch,
                                                                        float momentum, float eps, int update_running) {
    __shared__ double red[1024];
    const BnFinEntry e = tab[blockIdx.x];
    int nparts = (int)(((long long)batch * e.pp + e.mrows - 1) / e.mrows);
    if (e.tmax > 0) {      // conv_tcp_grid(): one row per CTA of the persistent kernel
        const long long ntiles = ((long long)batch * e.pp + 127) / 128;
        long long g = (ntiles + e.tmax - 1) / e.tmax;
        g = g < e.nsm ? e.nsm : g;
        nparts = (int)(g > ntiles ? ntiles : g);
    }
    const float cnt = (float)((long long)batch * e.hw);
    bn_partial_sums(ws + e.part_off, nparts, e.C, red);
    if ((int)threadIdx.x < e.C) {
        const int c = threadIdx.x;
        float sc, sh, m, is; double var;
        bn_affine_from_sums(red[c], red[e.C + c], (double)cnt, params[e.gamma_off + c], params[e.beta_off + c], eps, &sc, &sh, &m, &is, &var);
        float* aff = ws + e.aff_off;
        aff[c] = sc; aff[e.C + c] = sh; aff[2 * e.C + c] = m; aff[3 * e.C + c] = is;
        if (update_running && rstat != nullptr) {
            const double count = (double)cnt;
            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
            float* rm = rstat + e.rstat_off; float* rv = rm + e.C;
            rm[c] = (float)((1.0 - (double)momentum) * (double)rm[c] + (double)momentum * (double)(red[c] / count));
            rv[c] = (float)((1.0 - (double)momentum) * (double)rv[c] + (double)momentum * unb);
        }
    }
}

template <int C, int NT>
__device__ __forceinline__ void bn_finalize_last_block(const BnStatArgs& s, int nparts, double count, float* s_red /* >= max(2*NT, 4*C) floats, 8B aligned */) {
    // thread -> one float4 column group (COLS/4 of them) and one of NSL slices of the partial rows; independent float4 loads
    // (4 accumulators, unrolled) keep enough requests in flight that the last block's tail is a few microseconds, not tens
    constexpr int COLS = 2 * C;
    constexpr int CQ = COLS / 4;
    double* red = reinterpret_cast<double*>(s_red);
    if (NT >= CQ) {
        constexpr int NSL = NT >= CQ ? NT / CQ : 1;
        static_assert(NT % CQ == 0 || NT < CQ, "thread count vs channel count");
        const int cq = threadIdx.x % CQ, sl = threadIdx.x / CQ;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const float4* base = reinterpret_cast<const float4*>(s.partial) + cq;
#pragma unroll 4
        for (int p = sl; p < nparts; p += NSL) {
            const float4 v = __ldcg(base + (size_t)p * CQ);
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
        // fixed-order combine over the slices, 4 columns at a time (smem holds NT doubles)
        for (int k = 0; k < 4; ++k) {
            __syncthreads();
            red[threadIdx.x] = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
            __syncthreads();
            if (threadIdx.x < CQ) {
                double t = 0.0;
                for (int q = 0; q < NSL; ++q) t += red[q * CQ + threadIdx.x];
                a0 = k == 0 ? t : a0; a1 = k == 1 ? t : a1; a2 = k == 2 ? t : a2; a3 = k == 3 ? t : a3;
            }
        }
        __syncthreads();
        if (threadIdx.x < CQ) {
            red[threadIdx.x * 4 + 0] = a0; red[threadIdx.x * 4 + 1] = a1; red[threadIdx.x * 4 + 2] = a2; red[threadIdx.x * 4 + 3] = a3;
        }
        __syncthreads();
    } else {
        for (int j = threadIdx.x; j < COLS; j += NT) {
            double acc = 0.0;
            for (int p = 0; p < nparts; ++p) acc += (double)__ldcg(s.partial + (size_t)p * COLS + j);
            red[j] = acc;
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < C; c += NT) {
        const double m = red[c] / count;
        double var = red[C + c] / count - m * m;
        if (var < 0.0) var = 0.0;
        const double istd = 1.0 / sqrt(var + (double)s.eps);
        const float g = s.gamma[c], b = s.beta[c];
        const float sc = (float)((double)g * istd);
        s.scale[c] = sc;
        s.shift[c] = (float)((double)b - m * (double)g * istd);
        s.mean[c] = (float)m;
        s.invstd[c] = (float)istd;
        if (s.update_running) {
            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
            s.running_mean[c] = (float)((1.0 - (double)s.momentum) * (double)s.running_mean[c] + (double)s.momentum * m);
            s.running_var[c] = (float)((1.0 - (double)s.momentum) * (double)s.running_var[c] + (double)s.momentum * unb);
        }
    }
}

// Last CTA of a BatchNorm-backward reduction (256 threads): fixed-order fp64 sum of the per-CTA partial rows [nparts][2][C] -> the apply
// coefficients dy = c0*g + c1*y + c2, dgamma, dbeta.  Shared by bn_bwd_reduce_kernel and the fused epilogue of the tensor-core data-gradient conv.
template <int C>
__device__ __forceinline__ void bn_bwd_finalize_last_block(const float* partial, int nparts, double N, const float* scale, const float* mean,
                                                           const float* invstd, float* coef, float* dgamma, float* dbeta, float* s_red /* 512 floats */) {
    // float4 column groups x slices of the partial rows: few, wide, independent loads (see bn_finalize_last_block)
    double* red = reinterpret_cast<double*>(s_red);
    constexpr int COLS = 2 * C;
    constexpr int CQ = COLS / 4;
    constexpr int NSL = 256 / CQ;
    const int cq2 = threadIdx.x % CQ, sl = threadIdx.x / CQ;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const float4* base = reinterpret_cast<const float4*>(partial) + cq2;
#pragma unroll 4
    for (int p = sl; p < nparts; p += NSL) {
        const float4 v = __ldcg(base + (size_t)p * CQ);
        a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
    }
    for (int k = 0; k < 4; ++k) {
        __syncthreads();
        red[threadIdx.x] = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
        __syncthreads();
        if (threadIdx.x < CQ) {
            double t = 0.0;
            for (int q = 0; q < NSL; ++q) t += red[q * CQ + threadIdx.x];
            a0 = k == 0 ? t : a0; a1 = k == 1 ? t : a1; a2 = k == 2 ? t : a2; a3 = k == 3 ? t : a3;
        }
    }
    __syncthreads();
    if (threadIdx.x < CQ) {
        red[threadIdx.x * 4 + 0] = a0; red[threadIdx.x * 4 + 1] = a1; red[threadIdx.x * 4 + 2] = a2; red[threadIdx.x * 4 + 3] = a3;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        const int ch = threadIdx.x;
        const double S1 = red[ch], S2 = red[C + ch];
        const double sc = (double)scale[ch], istd = (double)invstd[ch], m = (double)mean[ch];
        const double c1 = -sc * S2 / N * istd;
        coef[ch] = (float)sc;
        coef[C + ch] = (float)c1;
        coef[2 * C + ch] = (float)(-sc * S1 / N - c1 * m);
        dgamma[ch] = (float)S2;
        dbeta[ch] = (float)S1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 convolution, pad 1, stride 1 or 2; forward and (with packed-transposed weights) data-gradient.
// ---------------------------------------------------------------------------------------------------------------
struct Conv3x3Args {
    const float* in;          // NHWC [B][HI][WI][CIN] (HI = HO*STRIDE; DILATE: real tensor is [B][HO/2][WO/2][CIN]) or NCHW (IN_NCHW)
    const float* wpack;       // [CIN][9][COUT]
    float* out;               // NHWC [B][HO][WO][COUT]
    const float* pro_scale;   // nullable: input transform relu(x*scale[c]+shift[c]) applied on load (zero padding stays zero)
    const float* pro_shift;
    const float* addend;      // nullable: out += addend   (may alias out)
    BnStatArgs stat;          // stat.partial nullable
    int B;
};

constexpr int conv_row_pitch(int iw_t, int lr, int pt, int stride) {
    // smallest pitch >= iw_t that keeps the LR row-groups of a warp on disjoint banks
    int want = lr == 1 ? -1 : (lr == 2 ? 16 : 8);
    for (int rp = iw_t; rp < iw_t + 33; ++rp) {
        if (lr == 1) return rp;
        int m = (pt * stride * rp) % 32;
        if (m == want || (lr == 4 && m == 24)) return rp;
    }
    return iw_t;
}
constexpr int conv_plane_stride(int n, int cin) {
    if (cin == 16) { while (n % 8 != 2) ++n; return n; }
    return n | 1;
}

template <int CIN, int COUT, int WO, int PT, int CT, int RS, int COUT_CTA, int CIN_CHUNK, int STRIDE, bool DILATE, bool IN_NCHW>
struct Conv3x3Cfg {
    static constexpr int HO = WO;
    static constexpr int LR = 32 / WO;
    static constexpr int TILE_H = RS * LR * PT;
    static constexpr int CG = COUT_CTA / CT;
    static constexpr int NWARP = RS * CG;
    static constexpr int NT = NWARP * 32;
    static constexpr int IH_T = (TILE_H - 1) * STRIDE + 3;
    static constexpr int IW_T = (WO - 1) * STRIDE + 3;
    static constexpr int RP = conv_row_pitch(IW_T, LR, PT, STRIDE);
    static constexpr int PS = conv_plane_stride(IH_T * RP, CIN);
    static constexpr int TILES_PER_IMG = HO / TILE_H;
    static constexpr int COUT_SPLIT = COUT / COUT_CTA;
    static constexpr int HI = HO * STRIDE;   // virtual input extent
    static constexpr int WI = WO * STRIDE;
    static constexpr int SMEM_IN = ((CIN * PS + 3) / 4) * 4;   // keeps s_w 16-byte aligned
    static constexpr int SMEM_W = CIN_CHUNK * 9 * COUT_CTA;
    static constexpr int SMEM_STAT = NWARP * CT * 2;
    static constexpr int SMEM_RED = 2 * NT > 4 * COUT ? 2 * NT : 4 * COUT;   // doubles for the finalize
    static constexpr size_t SMEM_BYTES = sizeof(float) * (SMEM_IN + SMEM_W + (SMEM_STAT > SMEM_RED ? SMEM_STAT : SMEM_RED) + 4);
    static_assert(32 % WO == 0 && HO % TILE_H == 0 && COUT % COUT_CTA == 0 && COUT_CTA % CT == 0 && CT % 4 == 0, "tiling");
    static_assert(CIN % CIN_CHUNK == 0, "cin chunk");
    static_assert(!(DILATE && STRIDE != 1), "dilated input implies stride 1");
    static_assert(!DILATE || (PT % 2 == 0), "the dilated kernel skips zero rows by the parity of i + r: tile rows must start on even rows");
};

template <int CIN, int COUT, int WO, int PT, int CT, int RS, int COUT_CTA, int CIN_CHUNK, int STRIDE, bool DILATE, bool IN_NCHW>
__global__ void __launch_bounds__(Conv3x3Cfg<CIN, COUT, WO, PT, CT, RS, COUT_CTA, CIN_CHUNK, STRIDE, DILATE, IN_NCHW>::NT)
conv3x3_kernel(Conv3x3Args a) {
    using K = Conv3x3Cfg<CIN, COUT, WO, PT, CT, RS, COUT_CTA, CIN_CHUNK, STRIDE, DILATE, IN_NCHW>;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;
    float* s_w = smem + K::SMEM_IN;
    float* s_x = s_w + K::SMEM_W;   // stats / finalize scratch (8-byte aligned: SMEM_IN + SMEM_W is even, see below)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y;
    const int tile = blockIdx.x % K::TILES_PER_IMG;
    const int cosplit = blockIdx.x / K::TILES_PER_IMG;
    const int co_base = cosplit * COUT_CTA;
    const int oh0 = tile * K::TILE_H;
    const int ih0 = oh0 * STRIDE - 1;   // first (virtual) input row of the tile

    // ---- stage the input tile, channel-planar, applying the producer's BN+ReLU on the fly ----------------------
    if constexpr (IN_NCHW) {
        constexpr int TOT = CIN * K::IH_T * K::IW_T;
        for (int e = tid; e < TOT; e += K::NT) {
            const int c = e % K::IW_T, r = (e / K::IW_T) % K::IH_T, ci = e / (K::IW_T * K::IH_T);
            const int ih = ih0 + r, iw = c - 1;
            float v = 0.f;
            if (ih >= 0 && ih < K::HI && iw >= 0 && iw < K::WI) v = __ldg(a.in + (((size_t)n * CIN + ci) * K::HI + ih) * K::WI + iw);
            s_in[ci * K::PS + r * K::RP + c] = v;
        }
    } else {
        constexpr int G4 = CIN / 4;
        constexpr int TOT = K::IH_T * K::IW_T * G4;
        const bool pro = a.pro_scale != nullptr;
        for (int e = tid; e < TOT; e += K::NT) {
            const int g = e % G4, c = (e / G4) % K::IW_T, r = e / (G4 * K::IW_T);
            const int ih = ih0 + r, iw = c - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            bool ok = ih >= 0 && ih < K::HI && iw >= 0 && iw < K::WI;
            size_t off;
            if constexpr (DILATE) {
                ok = ok && ((ih & 1) == 0) && ((iw & 1) == 0);
                off = (((size_t)n * (K::HI / 2) + (ih >> 1)) * (K::WI / 2) + (iw >> 1)) * CIN + g * 4;
            } else {
                off = (((size_t)n * K::HI + ih) * K::WI + iw) * CIN + g * 4;
            }
            if (ok) {
                v = ldg4(a.in + off);
                if (pro) {
                    const float4 sc = ldg4(a.pro_scale + g * 4), sh = ldg4(a.pro_shift + g * 4);
                    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
                    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
                    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
                    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
                }
            }
            float* d = s_in + (g * 4) * K::PS + r * K::RP + c;
            d[0] = v.x; d[K::PS] = v.y; d[2 * K::PS] = v.z; d[3 * K::PS] = v.w;
        }
    }

    // ---- thread tile --------------------------------------------------------------------------------------------
    const int strip = warp / K::CG, cg = warp % K::CG;
    const int c = lane % WO, rsub = lane / WO;
    const int orow0 = (strip * K::LR + rsub) * PT;       // tile-local first output row of this thread
    constexpr int NV = (PT - 1) * STRIDE + 3;
    float acc[PT][CT];
#pragma unroll
    for (int i = 0; i < PT; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k) acc[i][k] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += CIN_CHUNK) {
        __syncthreads();   // previous chunk's weights consumed (and, first time, nothing)
        {
            constexpr int W4 = COUT_CTA / 4;
            constexpr int TOTW = CIN_CHUNK * 9 * W4;
            for (int e = tid; e < TOTW; e += K::NT) {
                const int q = e % W4, row = e / W4;                 // row = ci_local*9 + tap
                const float4 w = ldg4(a.wpack + ((size_t)(c0 * 9 + row)) * COUT + co_base + q * 4);
                *reinterpret_cast<float4*>(s_w + row * COUT_CTA + q * 4) = w;
            }
        }
        __syncthreads();   // weights (and, first time, the input tile) visible
#pragma unroll 1
        for (int ci = 0; ci < CIN_CHUNK; ++ci) {
            const float* plane = s_in + (c0 + ci) * K::PS + (orow0 * STRIDE) * K::RP + c * STRIDE;
            const float* wrow = s_w + ci * 9 * COUT_CTA + cg * CT;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                float v[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) v[j] = plane[j * K::RP + s];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float w[CT];
#pragma unroll
                    for (int k4 = 0; k4 < CT / 4; ++k4) {
                        const float4 t = *reinterpret_cast<const float4*>(wrow + (r * 3 + s) * COUT_CTA + k4 * 4);
                        w[k4 * 4 + 0] = t.x; w[k4 * 4 + 1] = t.y; w[k4 * 4 + 2] = t.z; w[k4 * 4 + 3] = t.w;
                    }
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        // dilated input (data gradient of a stride-2 conv): input row oh0 + orow0 + i + r - 1 holds data only when it is even; oh0 and
                        // orow0 are multiples of PT (even), so the parity of i + r decides at compile time — half of the FMAs of the zero-inserted
                        // tensor are never issued (the column parity varies per lane and is left alone)
                        if (DILATE && ((i + r) & 1) == 0) continue;
#pragma unroll
                        for (int k = 0; k < CT; ++k) acc[i][k] = fmaf(v[i * STRIDE + r], w[k], acc[i][k]);
                    }
                }
            }
        }
    }

    // ---- epilogue: optional addend, store, optional BN statistics ----------------------------------------------------
    const int co0 = co_base + cg * CT;
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        const int oh = oh0 + orow0 + i;
        const size_t o = (((size_t)n * K::HO + oh) * WO + c) * COUT + co0;
        if (a.addend != nullptr) {
#pragma unroll
            for (int k4 = 0; k4 < CT / 4; ++k4) {
                const float4 t = *reinterpret_cast<const float4*>(a.addend + o + k4 * 4);
                acc[i][k4 * 4 + 0] += t.x; acc[i][k4 * 4 + 1] += t.y; acc[i][k4 * 4 + 2] += t.z; acc[i][k4 * 4 + 3] += t.w;
            }
        }
#pragma unroll
        for (int k4 = 0; k4 < CT / 4; ++k4)
            *reinterpret_cast<float4*>(a.out + o + k4 * 4) = make_float4(acc[i][k4 * 4], acc[i][k4 * 4 + 1], acc[i][k4 * 4 + 2], acc[i][k4 * 4 + 3]);
    }

    if (a.stat.partial != nullptr) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CT; ++k) {
            float sm = 0.f, sq = 0.f;
#pragma unroll
            for (int i = 0; i < PT; ++i) { sm += acc[i][k]; sq = fmaf(acc[i][k], acc[i][k], sq); }
            sm = warp_sum(sm); sq = warp_sum(sq);
            if (lane == 0) { s_x[(warp * CT + k) * 2] = sm; s_x[(warp * CT + k) * 2 + 1] = sq; }
        }
        __syncthreads();
        const int part = n * K::TILES_PER_IMG + tile;
        const int nparts = a.B * K::TILES_PER_IMG;
        if (tid < COUT_CTA * 2) {
            const int stat = tid / COUT_CTA, ch = tid % COUT_CTA;          // ch within this CTA's cout slice
            const int cgi = ch / CT, k = ch % CT;
            float t = 0.f;
#pragma unroll
            for (int st = 0; st < RS; ++st) t += s_x[((st * K::CG + cgi) * CT + k) * 2 + stat];
            a.stat.partial[((size_t)part * 2 + stat) * COUT + co_base + ch] = t;
        }
        const unsigned int nblocks = gridDim.x * gridDim.y;
        if (last_block_done(a.stat.counter, nblocks)) {
            bn_finalize_last_block<COUT, K::NT>(a.stat, nparts, (double)a.B * K::HO * WO, s_x);
        }
    }
}

template <typename KCfg, typename Kern>
static inline int conv_launch(Kern kern, const Conv3x3Args& a, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        if (KCfg::SMEM_BYTES > 48 * 1024) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KCfg::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        }
        attr_done = true;
    }
    dim3 grid(KCfg::TILES_PER_IMG * KCfg::COUT_SPLIT, a.B);
    kern<<<grid, KCfg::NT, KCfg::SMEM_BYTES, st>>>(a);
    return lc_launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 weight gradient:  dW[o][i][r][s] = sum_{n,oh,ow} dy[n,oh,ow,o] * a[n, oh*S+r-1, ow*S+s-1, i]
// Each CTA accumulates a K-slice (a strided set of (image,row-tile)s) in registers and writes one partial
// [COUT][CIN][9]; all layers' partials are summed in fixed order by wgrad_reduce_all_kernel (deterministic).
// ---------------------------------------------------------------------------------------------------------------
struct WgradArgs {
    const float* in;          // NHWC [B][HI][WI][CIN] or NCHW (IN_NCHW)
    const float* dy;          // NHWC [B][HO][WO][COUT]
    float* partial;           // [nsplit][COUT][CIN][9]
    const float* pro_scale;   // nullable prologue on `in`
    const float* pro_shift;
    int B;
    int nsplit;
};

template <int CIN, int COUT, int WO, int STRIDE, int CI_CTA, int KS, int TILE_H, bool IN_NCHW>
struct WgradCfg {
    static constexpr int HO = WO;
    static constexpr int OG = COUT / 4;
    static constexpr int NT = CI_CTA * OG * KS;
    static constexpr int IH_T = (TILE_H - 1) * STRIDE + 3;
    static constexpr int IW_T = (WO - 1) * STRIDE + 3;
    static constexpr int RP = IW_T;
    static constexpr int PS = (IH_T * RP) | 1;
    static constexpr int HI = HO * STRIDE, WI = WO * STRIDE;
    static constexpr int TILES_PER_IMG = HO / TILE_H;
    static constexpr int SMEM_A = ((CI_CTA * PS + 3) / 4) * 4;
    static constexpr int SMEM_DY = TILE_H * WO * COUT;
    static constexpr int SMEM_RED = CI_CTA * OG * 36;
    static constexpr int SMEM_MAIN = SMEM_A + SMEM_DY;
    static constexpr size_t SMEM_BYTES = sizeof(float) * (SMEM_MAIN > SMEM_RED ? SMEM_MAIN : SMEM_RED);
    static constexpr int CIN_SPLIT = (CIN + CI_CTA - 1) / CI_CTA;
    static_assert(HO % TILE_H == 0 && COUT % 4 == 0 && NT <= 1024 && NT % 32 == 0, "wgrad tiling");
};

template <int CIN, int COUT, int WO, int STRIDE, int CI_CTA, int KS, int TILE_H, bool IN_NCHW>
__global__ void __launch_bounds__(WgradCfg<CIN, COUT, WO, STRIDE, CI_CTA, KS, TILE_H, IN_NCHW>::NT)
wgrad3x3_kernel(WgradArgs a) {
    using K = WgradCfg<CIN, COUT, WO, STRIDE, CI_CTA, KS, TILE_H, IN_NCHW>;
    extern __shared__ __align__(16) float smem[];
    float* s_a = smem;
    float* s_dy = smem + K::SMEM_A;   // 16-byte aligned

    const int tid = threadIdx.x;
    const int ci_l = tid % CI_CTA;
    const int og = (tid / CI_CTA) % K::OG;
    const int ks = tid / (CI_CTA * K::OG);
    const int ci0 = blockIdx.y * CI_CTA;
    const bool ci_ok = (ci0 + ci_l) < CIN;

    float acc[4][9];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;

    const int ntiles = a.B * K::TILES_PER_IMG;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n = t / K::TILES_PER_IMG, tile = t % K::TILES_PER_IMG;
        const int oh0 = tile * TILE_H, ih0 = oh0 * STRIDE - 1;
        __syncthreads();
        // input tile, planar [ci][row][col]
        if constexpr (IN_NCHW) {
            constexpr int TOT = CI_CTA * K::IH_T * K::IW_T;
            for (int e = tid; e < TOT; e += K::NT) {
                const int c = e % K::IW_T, r = (e / K::IW_T) % K::IH_T, ci = e / (K::IW_T * K::IH_T);
                const int ih = ih0 + r, iw = c - 1;
                float v = 0.f;
                if ((ci0 + ci) < CIN && ih >= 0 && ih < K::HI && iw >= 0 && iw < K::WI)
                    v = __ldg(a.in + (((size_t)n * CIN + ci0 + ci) * K::HI + ih) * K::WI + iw);
                s_a[ci * K::PS + r * K::RP + c] = v;
            }
        } else {
            constexpr int G4 = CI_CTA / 4;
            constexpr int TOT = K::IH_T * K::IW_T * G4;
            const bool pro = a.pro_scale != nullptr;
            for (int e = tid; e < TOT; e += K::NT) {
                const int g = e % G4, c = (e / G4) % K::IW_T, r = e / (G4 * K::IW_T);
                const int ih = ih0 + r, iw = c - 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ih >= 0 && ih < K::HI && iw >= 0 && iw < K::WI) {
                    v = ldg4(a.in + (((size_t)n * K::HI + ih) * K::WI + iw) * CIN + ci0 + g * 4);
                    if (pro) {
                        const float4 sc = ldg4(a.pro_scale + ci0 + g * 4), sh = ldg4(a.pro_shift + ci0 + g * 4);
                        v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
                        v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
                        v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
                        v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
                    }
                }
                float* d = s_a + (g * 4) * K::PS + r * K::RP + c;
                d[0] = v.x; d[K::PS] = v.y; d[2 * K::PS] = v.z; d[3 * K::PS] = v.w;
            }
        }
        // dy tile [pixel][COUT] (contiguous in NHWC)
        {
            constexpr int TOT4 = TILE_H * WO * COUT / 4;
            const float* src = a.dy + (((size_t)n * K::HO + oh0) * WO) * COUT;
            for (int e = tid; e < TOT4; e += K::NT) *reinterpret_cast<float4*>(s_dy + e * 4) = ldg4(src + e * 4);
        }
        __syncthreads();

        const float* plane = s_a + ci_l * K::PS;
        for (int ohl = ks; ohl < TILE_H; ohl += KS) {
            const float* prow = plane + (ohl * STRIDE) * K::RP;
            if constexpr (STRIDE == 1) {
                float w0[3], w1[3], w2[3];   // window columns: w0 = iw-1, w1 = iw, w2 = iw+1 (each 3 rows)
#pragma unroll
                for (int r = 0; r < 3; ++r) { w1[r] = prow[r * K::RP + 0]; w2[r] = prow[r * K::RP + 1]; }
#pragma unroll 4
                for (int ow = 0; ow < WO; ++ow) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) { w0[r] = w1[r]; w1[r] = w2[r]; w2[r] = prow[r * K::RP + ow + 2]; }
                    const float4 d = *reinterpret_cast<const float4*>(s_dy + (ohl * WO + ow) * COUT + og * 4);
                    const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            acc[k][r * 3 + 0] = fmaf(dd[k], w0[r], acc[k][r * 3 + 0]);
                            acc[k][r * 3 + 1] = fmaf(dd[k], w1[r], acc[k][r * 3 + 1]);
                            acc[k][r * 3 + 2] = fmaf(dd[k], w2[r], acc[k][r * 3 + 2]);
                        }
                }
            } else {
#pragma unroll 2
                for (int ow = 0; ow < WO; ++ow) {
                    float w[9];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int s = 0; s < 3; ++s) w[r * 3 + s] = prow[r * K::RP + ow * STRIDE + s];
                    const float4 d = *reinterpret_cast<const float4*>(s_dy + (ohl * WO + ow) * COUT + og * 4);
                    const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int q = 0; q < 9; ++q) acc[k][q] = fmaf(dd[k], w[q], acc[k][q]);
                }
            }
        }
    }

    // ---- reduce the KS in-CTA slices in fixed order, then write this CTA's partial ---------------------------------
    if constexpr (KS > 1) {
        float* s_red = smem;
        const int slot = (og * CI_CTA + ci_l) * 36;
        for (int q = 1; q < KS; ++q) {
            __syncthreads();
            if (ks == q) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int t = 0; t < 9; ++t) s_red[slot + k * 9 + t] = acc[k][t];
            }
            __syncthreads();
            if (ks == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int t = 0; t < 9; ++t) acc[k][t] += s_red[slot + k * 9 + t];
            }
        }
    }
    if (ks == 0 && ci_ok) {
        float* dst = a.partial + (size_t)blockIdx.x * COUT * CIN * 9;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float* p = dst + ((size_t)(og * 4 + k) * CIN + ci0 + ci_l) * 9;
#pragma unroll
            for (int t = 0; t < 9; ++t) p[t] = acc[k][t];
        }
    }
}

template <typename KCfg, typename Kern>
static inline int wgrad_launch(Kern kern, const WgradArgs& a, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        if (KCfg::SMEM_BYTES > 48 * 1024) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KCfg::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        }
        attr_done = true;
    }
    dim3 grid(a.nsplit, KCfg::CIN_SPLIT);
    kern<<<grid, KCfg::NT, KCfg::SMEM_BYTES, st>>>(a);
    return lc_launch_status();
}

}  // namespace lc
