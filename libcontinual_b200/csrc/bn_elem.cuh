// Elementwise / per-channel kernels on fp32 NHWC activations: BN-apply + residual + ReLU forward, and the two-pass
// BatchNorm backward (per-channel reduction, then affine apply).  All HBM-bound: float4 accesses, channel index = idx % C.
//
// Reference op sequence replaced: F.relu(bn(x)), F.relu(residual + bn_b(...)) in resnet.py:306-316,382 and their autograd
// (native_batch_norm_backward + threshold_backward + add).
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"

namespace lc {

struct BnActArgs {
    const float* y;           // raw conv output
    const float* scale;       // [C]
    const float* shift;       // [C]
    const float* res;         // nullable residual (same shape)
    const float* res_scale;   // nullable: residual is itself a raw conv output to be BN'ed (downsample path)
    const float* res_shift;
    float* out;
    long long n4;             // number of float4 elements
    int C;
    int no_relu;              // 1: skip the final ReLU (last block of LUCIR's modified_ResNet, resnet.py:501-502)
    BnLazy lazy;              // lazy.partial != null: scale / shift of `y` are reduced here from the producing conv's partial rows (conv_simt.cuh)
};

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(BnActArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent tensor-core conv may start its data-independent prologue now
    __shared__ double s_red[1024];
    __shared__ __align__(16) float s_aff[128];
    const int c4n = a.C >> 2;
    const float* scale = a.scale;
    const float* shift = a.shift;
    if (a.lazy.partial != nullptr) {
        bn_lazy_affine(a.lazy, a.C, s_red, s_aff);
        scale = s_aff; shift = s_aff + a.C;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        float4 v = ldg4(a.y + i * 4);
        const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (a.res != nullptr) {
            float4 r = ldg4(a.res + i * 4);
            if (a.res_scale != nullptr) {
                const float4 rs = ldg4(a.res_scale + c), rh = ldg4(a.res_shift + c);
                r.x = fmaf(r.x, rs.x, rh.x); r.y = fmaf(r.y, rs.y, rh.y); r.z = fmaf(r.z, rs.z, rh.z); r.w = fmaf(r.w, rs.w, rh.w);
            }
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (!a.no_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(a.out + i * 4) = v;
    }
}

// mask modes for the ReLU that follows the BN being differentiated
enum { LC_MASK_NONE = 0, LC_MASK_FROM_OUT = 1, LC_MASK_FROM_BN = 2 };

struct BnBwdArgs {
    const float* g;           // incoming gradient (w.r.t. the ReLU output, or w.r.t. the BN output when mask_mode == NONE)
    const float* mask_src;    // LC_MASK_FROM_OUT: materialised ReLU output
    const float* y;           // raw conv output (BN input)
    const float* scale;       // [C] forward affine of this BN (gamma*invstd), used by LC_MASK_FROM_BN and for c0
    const float* shift;       // [C]
    const float* mean;        // [C]
    const float* invstd;      // [C]
    float* partial;           // [nblk][2][C]
    unsigned int* counter;
    float* coef;              // out [3][C]: dy = c0*g + c1*y + c2
    float* dgamma;            // out [C]
    float* dbeta;             // out [C]
    float* dy;                // apply: output
    float* g_out;             // apply: nullable, masked gradient written back (may alias g)
    long long npix;
    int C;
    int mask_mode;
    int reduce_writes_g;      // 1: the reduction pass stores the masked gradient to g_out (fused flow: no apply pass follows)
    BnBwdLazy blazy;          // apply pass: .partial != null -> coefficients reduced here from a fused epilogue's partial rows (conv_simt.cuh)
};

__device__ __forceinline__ float4 bn_masked_grad(const BnBwdArgs& a, long long e, int c, float4 yv) {
    float4 g = ldg4(a.g + e);
    if (a.mask_mode == LC_MASK_FROM_OUT) {
        const float4 o = ldg4(a.mask_src + e);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    } else if (a.mask_mode == LC_MASK_FROM_BN) {
        const float4 sc = ldg4(a.scale + c), sh = ldg4(a.shift + c);
        g.x = fmaf(yv.x, sc.x, sh.x) > 0.f ? g.x : 0.f;
        g.y = fmaf(yv.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(yv.z, sc.z, sh.z) > 0.f ? g.z : 0.f;
        g.w = fmaf(yv.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
    }
    return g;
}

// pass 1: s1[c] = sum g, s2[c] = sum g * xhat ; last block turns them into the apply coefficients
template <int C>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(BnBwdArgs a) {
    constexpr int C4 = C / 4;
    constexpr int PL = 256 / C4;
    __shared__ __align__(16) float s_red[2 * 256 * 4];
    const int cq = threadIdx.x % C4, pl = threadIdx.x / C4, c = cq * 4;
    const long long per = (a.npix + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per;
    const long long p1 = p0 + per < a.npix ? p0 + per : a.npix;
    const float4 mu = ldg4(a.mean + c), is = ldg4(a.invstd + c);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (long long p = p0 + pl; p < p1; p += PL) {
        const long long e = p * C + c;
        const float4 yv = ldg4(a.y + e);
        const float4 g = bn_masked_grad(a, e, c, yv);
        if (a.reduce_writes_g) *reinterpret_cast<float4*>(a.g_out + e) = g;       // masked gradient (may alias g: one reader per element)
        s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
        s2.x = fmaf(g.x, (yv.x - mu.x) * is.x, s2.x);
        s2.y = fmaf(g.y, (yv.y - mu.y) * is.y, s2.y);
        s2.z = fmaf(g.z, (yv.z - mu.z) * is.z, s2.z);
        s2.w = fmaf(g.w, (yv.w - mu.w) * is.w, s2.w);
    }
    float4* r1 = reinterpret_cast<float4*>(s_red);
    float4* r2 = r1 + 256;
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int off = PL / 2; off > 0; off >>= 1) {      // fixed-order tree over the pixel lanes
        if (pl < off) {
            float4 x = r1[threadIdx.x], y2 = r1[threadIdx.x + off * C4];
            x.x += y2.x; x.y += y2.y; x.z += y2.z; x.w += y2.w; r1[threadIdx.x] = x;
            x = r2[threadIdx.x]; y2 = r2[threadIdx.x + off * C4];
            x.x += y2.x; x.y += y2.y; x.z += y2.z; x.w += y2.w; r2[threadIdx.x] = x;
        }
        __syncthreads();
    }
    if (pl == 0) {
        *reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.x * 2 + 0) * C + c) = r1[threadIdx.x];
        *reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.x * 2 + 1) * C + c) = r2[threadIdx.x];
    }
    if (last_block_done(a.counter, gridDim.x))
        bn_bwd_finalize_last_block<C>(a.partial, (int)gridDim.x, (double)a.npix, a.scale, a.mean, a.invstd, a.coef, a.dgamma, a.dbeta, s_red);
}

// pass 2: dy = c0*g + c1*y + c2  (g masked as in pass 1); optionally writes the masked g back
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BnBwdArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent tensor-core conv may start its data-independent prologue now
    __shared__ double s_red[1024];
    __shared__ __align__(16) float s_coef[192];
    const int C = a.C, c4n = C >> 2;
    const long long n4 = a.npix * c4n;
    const float* coef = a.coef;
    if (a.blazy.partial != nullptr) {
        bn_bwd_lazy_coef(a.blazy, C, s_red, s_coef);
        coef = s_coef;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const long long e = i * 4;
        const float4 yv = ldg4(a.y + e);
        const float4 g = bn_masked_grad(a, e, c, yv);
        const float4 c0 = *reinterpret_cast<const float4*>(coef + c), c1 = *reinterpret_cast<const float4*>(coef + C + c),
                     c2 = *reinterpret_cast<const float4*>(coef + 2 * C + c);
        float4 d;
        d.x = fmaf(c0.x, g.x, fmaf(c1.x, yv.x, c2.x));
        d.y = fmaf(c0.y, g.y, fmaf(c1.y, yv.y, c2.y));
        d.z = fmaf(c0.z, g.z, fmaf(c1.z, yv.z, c2.z));
        d.w = fmaf(c0.w, g.w, fmaf(c1.w, yv.w, c2.w));
        *reinterpret_cast<float4*>(a.dy + e) = d;
        if (a.g_out != nullptr) *reinterpret_cast<float4*>(a.g_out + e) = g;
    }
}

// eval-mode affine of every BN layer from its running statistics (one launch for the whole network)
struct BnEvalEntry { long long gamma_off, beta_off, rstat_off, aff_off; int C; int pad; };
__global__ void bn_eval_affine_kernel(const BnEvalEntry* tab, int nlayers, const float* params, const float* rstat, float* aff, float eps) {
    const int l = blockIdx.x;
    if (l >= nlayers) return;
    const BnEvalEntry e = tab[l];
    for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
        const float m = rstat[e.rstat_off + c], v = rstat[e.rstat_off + e.C + c];
        const float istd = (float)(1.0 / sqrt((double)v + (double)eps));
        const float sc = params[e.gamma_off + c] * istd;
        aff[e.aff_off + c] = sc;
        aff[e.aff_off + e.C + c] = params[e.beta_off + c] - m * sc;
        aff[e.aff_off + 2 * e.C + c] = m;
        aff[e.aff_off + 3 * e.C + c] = istd;
    }
}

static inline int elem_grid(long long n4) {
    long long b = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace lc
