// 1x1 stride-2 shortcut convolution (resnet.py:365-366) forward / data-grad / weight-grad, plus the two table-driven
// multi-layer kernels: weight packing (OIHW -> [ci][tap][co] and the flipped/transposed dgrad form) and the final,
// fixed-order reduction of all weight-gradient partials into the OIHW gradient arena.
#pragma once
#include "conv_simt.cuh"

namespace lc {

// ---------------------------------------------------------------------------------------------------------------
// forward: out[n,oh,ow,:] = W[COUT][CIN] * in[n,2oh,2ow,:]   (+ BN statistics)
// ---------------------------------------------------------------------------------------------------------------
struct Conv1x1Args {
    const float* in;      // NHWC [B][2*WO][2*WO][CIN]
    const float* w;       // fwd: packed [CIN][COUT]; dgrad/wgrad: see kernels
    float* out;
    BnStatArgs stat;
    int B;
};

template <int CIN, int COUT, int WO>
__global__ void __launch_bounds__(128) conv1x1s2_fwd_kernel(Conv1x1Args a) {
    constexpr int NT = 128;
    __shared__ __align__(16) float s_w[CIN * COUT];
    __shared__ __align__(16) float s_x[(4 * 2 * COUT > 4 * COUT ? 4 * 2 * COUT : 4 * COUT) + 2 * NT];
    for (int e = threadIdx.x; e < CIN * COUT / 4; e += NT) *reinterpret_cast<float4*>(s_w + e * 4) = ldg4(a.w + e * 4);
    __syncthreads();
    const long long npix = (long long)a.B * WO * WO;
    const long long p = (long long)blockIdx.x * NT + threadIdx.x;
    const bool ok = p < npix;
    float acc[COUT];
#pragma unroll
    for (int k = 0; k < COUT; ++k) acc[k] = 0.f;
    if (ok) {
        const int ow = (int)(p % WO), oh = (int)((p / WO) % WO);
        const long long n = p / (WO * WO);
        const float* src = a.in + ((n * (2 * WO) + 2 * oh) * (2 * WO) + 2 * ow) * CIN;
#pragma unroll 1
        for (int c4 = 0; c4 < CIN / 4; ++c4) {
            const float4 xv = ldg4(src + c4 * 4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* wr = s_w + (c4 * 4 + j) * COUT;
#pragma unroll
                for (int k4 = 0; k4 < COUT / 4; ++k4) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + k4 * 4);
                    acc[k4 * 4 + 0] = fmaf(xs[j], w.x, acc[k4 * 4 + 0]);
                    acc[k4 * 4 + 1] = fmaf(xs[j], w.y, acc[k4 * 4 + 1]);
                    acc[k4 * 4 + 2] = fmaf(xs[j], w.z, acc[k4 * 4 + 2]);
                    acc[k4 * 4 + 3] = fmaf(xs[j], w.w, acc[k4 * 4 + 3]);
                }
            }
        }
        float* dst = a.out + p * COUT;
#pragma unroll
        for (int k4 = 0; k4 < COUT / 4; ++k4)
            *reinterpret_cast<float4*>(dst + k4 * 4) = make_float4(acc[k4 * 4], acc[k4 * 4 + 1], acc[k4 * 4 + 2], acc[k4 * 4 + 3]);
    }
    if (a.stat.partial != nullptr) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < COUT; ++k) {
            float sm = warp_sum(acc[k]);            // out-of-range pixels hold zeros
            float sq = warp_sum(acc[k] * acc[k]);
            if (lane == 0) { s_x[(warp * 2 + 0) * COUT + k] = sm; s_x[(warp * 2 + 1) * COUT + k] = sq; }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * COUT; j += NT) {
            const int stat = j / COUT, ch = j % COUT;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) t += s_x[(w * 2 + stat) * COUT + ch];
            a.stat.partial[(size_t)blockIdx.x * 2 * COUT + j] = t;
        }
        if (last_block_done(a.stat.counter, gridDim.x)) {
            bn_finalize_last_block<COUT, NT>(a.stat, (int)gridDim.x, (double)npix, s_x);
        }
    }
}

// data gradient: gin[n,2oh,2ow,i] += sum_o dy[n,oh,ow,o] * W[o][i]     (W in its native OIHW = [COUT][CIN] form)
template <int CIN, int COUT, int WO>
__global__ void __launch_bounds__(128) conv1x1s2_dgrad_accum_kernel(const float* dy, const float* w, float* gin, int B) {
    constexpr int NT = 128;
    __shared__ __align__(16) float s_w[COUT * CIN];
    for (int e = threadIdx.x; e < CIN * COUT / 4; e += NT) *reinterpret_cast<float4*>(s_w + e * 4) = ldg4(w + e * 4);
    __syncthreads();
    const long long npix = (long long)B * WO * WO;
    const long long p = (long long)blockIdx.x * NT + threadIdx.x;
    if (p >= npix) return;
    const int ow = (int)(p % WO), oh = (int)((p / WO) % WO);
    const long long n = p / (WO * WO);
    float acc[CIN];
#pragma unroll
    for (int k = 0; k < CIN; ++k) acc[k] = 0.f;
    const float* src = dy + p * COUT;
#pragma unroll 1
    for (int o4 = 0; o4 < COUT / 4; ++o4) {
        const float4 dv = ldg4(src + o4 * 4);
        const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* wr = s_w + (o4 * 4 + j) * CIN;
#pragma unroll
            for (int k4 = 0; k4 < CIN / 4; ++k4) {
                const float4 wv = *reinterpret_cast<const float4*>(wr + k4 * 4);
                acc[k4 * 4 + 0] = fmaf(ds[j], wv.x, acc[k4 * 4 + 0]);
                acc[k4 * 4 + 1] = fmaf(ds[j], wv.y, acc[k4 * 4 + 1]);
                acc[k4 * 4 + 2] = fmaf(ds[j], wv.z, acc[k4 * 4 + 2]);
                acc[k4 * 4 + 3] = fmaf(ds[j], wv.w, acc[k4 * 4 + 3]);
            }
        }
    }
    float* dst = gin + ((n * (2 * WO) + 2 * oh) * (2 * WO) + 2 * ow) * CIN;
#pragma unroll
    for (int k4 = 0; k4 < CIN / 4; ++k4) {
        float4 t = *reinterpret_cast<const float4*>(dst + k4 * 4);
        t.x += acc[k4 * 4]; t.y += acc[k4 * 4 + 1]; t.z += acc[k4 * 4 + 2]; t.w += acc[k4 * 4 + 3];
        *reinterpret_cast<float4*>(dst + k4 * 4) = t;
    }
}

// weight gradient partials: partial[split][o][i] = sum_{pix in split} dy[pix][o] * in[n,2oh,2ow,i]
template <int CIN, int COUT, int WO>
__global__ void __launch_bounds__(256) conv1x1s2_wgrad_kernel(const float* in, const float* dy, float* partial, int B) {
    constexpr int NT = 256, P = 64;
    constexpr int OGN = NT / CIN;          // thread groups along COUT
    constexpr int OPT = COUT / OGN;        // couts per thread
    static_assert(NT % CIN == 0 && COUT % OGN == 0, "1x1 wgrad tiling");
    __shared__ __align__(16) float s_x[P * CIN];
    __shared__ __align__(16) float s_d[P * COUT];
    const int i = threadIdx.x % CIN, og = threadIdx.x / CIN;
    float acc[OPT];
#pragma unroll
    for (int k = 0; k < OPT; ++k) acc[k] = 0.f;
    const long long npix = (long long)B * WO * WO;
    for (long long base = (long long)blockIdx.x * P; base < npix; base += (long long)gridDim.x * P) {
        __syncthreads();
        for (int e = threadIdx.x; e < P * CIN / 4; e += NT) {
            const int pl = e / (CIN / 4), c4 = e % (CIN / 4);
            const long long p = base + pl;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < npix) {
                const int ow = (int)(p % WO), oh = (int)((p / WO) % WO);
                const long long n = p / (WO * WO);
                v = ldg4(in + ((n * (2 * WO) + 2 * oh) * (2 * WO) + 2 * ow) * CIN + c4 * 4);
            }
            *reinterpret_cast<float4*>(s_x + pl * CIN + c4 * 4) = v;
        }
        for (int e = threadIdx.x; e < P * COUT / 4; e += NT) {
            const int pl = e / (COUT / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base + pl < npix) v = ldg4(dy + base * COUT + (long long)e * 4);
            *reinterpret_cast<float4*>(s_d + e * 4) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int pl = 0; pl < P; ++pl) {
            const float xv = s_x[pl * CIN + i];
#pragma unroll
            for (int k = 0; k < OPT; ++k) acc[k] = fmaf(s_d[pl * COUT + og * OPT + k], xv, acc[k]);
        }
    }
    float* dst = partial + (size_t)blockIdx.x * COUT * CIN;
#pragma unroll
    for (int k = 0; k < OPT; ++k) dst[(size_t)(og * OPT + k) * CIN + i] = acc[k];
}

// ---------------------------------------------------------------------------------------------------------------
// table-driven multi-layer kernels
// ---------------------------------------------------------------------------------------------------------------
struct ConvTabEntry {
    long long w_off;      // OIHW weights / gradient offset in the parameter (gradient) arena
    long long wf_off;     // packed forward weights  [ci][tap][co]
    long long wd_off;     // packed dgrad weights    [co][ntap-1-tap][ci]   (-1: not needed)
    long long part_off;   // weight-gradient partials [nsplit][n]
    long long wtf_off;    // tensor-core forward packing  [tap][ci/4][co][ci%4]   (-1: layer not on the tensor-core path)
    long long wtd_off;    // tensor-core dgrad packing    [8-tap][co/4][ci][co%4]
    long long wts_off1;   // 1 + offset of the tensor-core forward packing of a STRIDE-2 layer (conv_s2_tc.cuh; same [tap][ci/4][co][ci%4] order); 0: none
    long long wtsd_off1;  // 1 + offset of the stride-2 layer's tensor-core DATA-GRADIENT packing [tap][co/4][ci][co%4] (tap not flipped); 0: none
    int cout, cin, ntap, nsplit;
    int blk_begin;        // first block of this layer in the table kernels
    int nsplit_tc;        // split count used by the tensor-core weight-gradient kernel (<= nsplit)
};

__device__ __forceinline__ int tab_find_layer(const ConvTabEntry* tab, int nlayers, int blk) {
    int l = 0;
    while (l + 1 < nlayers && tab[l + 1].blk_begin <= blk) ++l;
    return l;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const ConvTabEntry* tab, int nlayers, const float* params, float* packed) {
    const int l = tab_find_layer(tab, nlayers, blockIdx.x);
    const ConvTabEntry t = tab[l];
    const int n = t.cout * t.cin * t.ntap;
    const int e = (blockIdx.x - t.blk_begin) * 256 + threadIdx.x;
    if (e >= n) return;
    const int tap = e % t.ntap, ci = (e / t.ntap) % t.cin, co = e / (t.ntap * t.cin);
    const float w = params[t.w_off + e];
    packed[t.wf_off + ((size_t)ci * t.ntap + tap) * t.cout + co] = w;
    if (t.wd_off >= 0) packed[t.wd_off + ((size_t)co * t.ntap + (t.ntap - 1 - tap)) * t.cin + ci] = w;
    if (t.wts_off1 > 0) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w));
        packed[t.wts_off1 - 1 + (((size_t)tap * (t.cin >> 2) + (ci >> 2)) * t.cout + co) * 4 + (ci & 3)] = __uint_as_float(r);
        if (t.wtsd_off1 > 0) packed[t.wtsd_off1 - 1 + (((size_t)tap * (t.cout >> 2) + (co >> 2)) * t.cin + ci) * 4 + (co & 3)] = __uint_as_float(r);
    }
    if (t.wtf_off >= 0) {      // tensor-core operands are pre-rounded to TF32 (round-to-nearest-away) once per step
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w));
        const float wr = __uint_as_float(r);
        packed[t.wtf_off + (((size_t)tap * (t.cin >> 2) + (ci >> 2)) * t.cout + co) * 4 + (ci & 3)] = wr;
        packed[t.wtd_off + (((size_t)(t.ntap - 1 - tap) * (t.cout >> 2) + (co >> 2)) * t.cin + ci) * 4 + (co & 3)] = wr;
    }
}

// grad[w_off + e] = sum_{s < nsplit} partial[part_off + s*n + e]   (fixed order => deterministic).
// tc_mode != 0: layers on the tensor-core path wrote their partials as [tap][ci][co]; read them in that (coalesced) order and
// scatter the sums to OIHW.
__global__ void __launch_bounds__(256) wgrad_reduce_all_kernel(const ConvTabEntry* tab, int nlayers, const float* partial, float* grad, int tc_mode) {
    const int l = tab_find_layer(tab, nlayers, blockIdx.x);
    const ConvTabEntry t = tab[l];
    const int n = t.cout * t.cin * t.ntap;
    const int e = (blockIdx.x - t.blk_begin) * 256 + threadIdx.x;
    if (e >= n) return;
    const float* p = partial + t.part_off + e;
    const bool tc_layer = tc_mode != 0 && t.wtf_off >= 0;
    const int ns = tc_layer ? t.nsplit_tc : t.nsplit;
    // 8 independent partial sums (8 loads in flight), combined in a fixed order
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int s = 0;
    for (; s + 8 <= ns; s += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += __ldg(p + (size_t)(s + k) * n);
    }
    for (int k = 0; s < ns; ++s, ++k) a[k] += __ldg(p + (size_t)s * n);
    const float acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    int out = e;
    if (tc_layer) {
        const int co = e % t.cout, ci = (e / t.cout) % t.cin, tap = e / (t.cout * t.cin);
        out = (co * t.cin + ci) * t.ntap + tap;
    }
    grad[t.w_off + out] = acc;
}

}  // namespace lc
