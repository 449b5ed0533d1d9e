// 3x3 / pad 1 / stride 1 weight gradient on the tcgen05 tensor cores (kind::f16 with BF16 operands, fp32 accumulation in TMEM),
// NHWC fp32 activations / gradients in HBM, converted to BF16 while staging.
//
//   dW[co][ci][dr][dc] = sum_Q  dY[Q][co] * Xs[Q + dr*(W+2) + dc][ci]          (Q: flattened padded rows, see conv_tc.cuh)
//
// The contraction runs over pixels, so both operands are consumed MN-major (channels contiguous, the reduction index strided).
// The channel-chunk-planar staging of the forward kernel (plane = one 16-byte chunk of channels, row r at plane + 16*r) IS the
// canonical MN-major / no-swizzle UMMA layout: core matrix = 8 consecutive rows x 16 B, SBO = plane size (next 8 channels),
// LBO = 128 B (next 8 rows).  [Measured on B200 with tools/tc_probe.cu: MN-major + SWIZZLE_NONE is exact for kind::f16 (bf16),
// M = 64 and 128, but yields all-zero accumulators for kind::tf32 — hence BF16 operands here while forward / dgrad stay TF32.]
// One MMA covers K = 16 pixels; the dr tap is a descriptor row offset.
//
// The output tile is tiny (COUT x 9*CIN), so the M side is made as tall as the instruction allows by staging the input tile
// THREE times, copy c displaced by (c-1) rows: block b = c*CH8 + j of the M operand is then "tap dc = c-1, channels 8j..8j+7",
// and one MMA covers all three dc taps of a dr.  D rows = (dc, ci), D columns = co; one accumulator per dr.
//   C = 16: M64 (48 rows used), N16, 3 accumulators x 16 TMEM columns
//   C = 32: M128 (96 rows used), N32, 3 x 32 columns
//   C = 64: M128 (dc = -1, 0) + M64 (dc = +1), N64, 6 x 64 columns
// Unused M blocks read whatever follows the copies in this CTA's shared memory; they only produce D rows nobody reads.
//
// Each CTA owns a strided set of 128-pixel tiles, accumulates them all in TMEM (no per-tile epilogue), and finally writes ONE
// partial [9][CIN][COUT]; wgrad_reduce_all_kernel sums the partials in fixed order (deterministic split-K) and restores OIHW.
#pragma once
#include "conv_tc.cuh"
#include <cuda_bf16.h>

#ifndef LC_WGRAD16_MAXNREG
#define LC_WGRAD16_MAXNREG 88      // stage-1 weight-gradient CTA (256 x 88) + two data-gradient CTAs (288 x 72) = one SM register file
#endif

namespace lc {
namespace tc {

// instruction descriptor: D=F32, A=B=BF16, both MN-major (bits 15, 16), M, N
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);      // .x = lo (low 16 bits), round-to-nearest-even
    return *reinterpret_cast<const uint32_t*>(&v);
}

struct WgradTcArgs {
    const float* in;          // NHWC [B][W][W][C]  (conv input; optional BN+ReLU prologue)
    const float* dy;          // NHWC [B][W][W][C]
    float* partial;           // [gridDim.x][9][C][C]  (tap, ci, co)
    const float* pro_scale;   // nullable
    const float* pro_shift;
    const float* dy_y;        // nullable: dY = coef0[c]*dy + coef1[c]*dy_y + coef2[c] evaluated while staging (fused BatchNorm backward apply,
    const float* dy_coef;     //           see BnBwdFuse in conv_tc.cuh); dy_coef = [3][C]
    BnBwdLazy dy_blazy;       // .partial != null: the coefficients are reduced here from the producing epilogue's partial rows (CTA 0 also writes dgamma / dbeta)
    int* error_flag;
    int B;
};

template <int C, int W>
struct WgradTcCfg {
    static constexpr int NT = 256;
    static constexpr int WP = W + 2;
    static constexpr int PP = WP * WP;
    static constexpr int HALO = WP + 1;
    static constexpr int CH8 = C / 8;                          // 16-byte BF16 chunks (8 channels) per row
    static constexpr int ROWS_X = 128 + 2 * HALO;
    static constexpr int PLANE_X = ROWS_X * 16;
    static constexpr int PLANE_Y = 128 * 16;
    static constexpr int X_BYTES = 3 * CH8 * PLANE_X;          // three displaced copies
    static constexpr int Y_BYTES = CH8 * PLANE_Y;
    static constexpr int MBLK = C == 16 ? 8 : 16;              // 8-row blocks spanned by the (first) MMA of a dr
    static constexpr int SPAN = MBLK + (C == 64 ? 8 : 0);
    static constexpr int PAD_BYTES = (SPAN > 3 * CH8 ? SPAN - 3 * CH8 : 0) * PLANE_X;
    static constexpr int OFF_Y = X_BYTES;
    static constexpr int OFF_ROWTAB = OFF_Y + (Y_BYTES > PAD_BYTES ? Y_BYTES : PAD_BYTES);
    static constexpr int ROWTAB_BYTES = ((ROWS_X * 4 + 15) / 16) * 16;
    static constexpr int OFF_BAR = OFF_ROWTAB + ROWTAB_BYTES;
    static constexpr int OFF_RED = OFF_BAR + 64;                 // 1024 doubles (bn_partial_sums) + c0, c1, c2
    static constexpr size_t SMEM_BYTES = OFF_RED + 8192 + 5 * C * 4;
    static constexpr int NACC = C == 64 ? 6 : 3;
    static constexpr uint32_t TMEM_COLS = C == 16 ? 64 : (C == 32 ? 128 : 512);
    static constexpr int RSTEP = NT / CH8;
    static constexpr int NEX = (ROWS_X + RSTEP - 1) / RSTEP;   // staging iterations for the input tile
    static constexpr int NEY = (128 + RSTEP - 1) / RSTEP;
    static_assert(C == 16 || C == 32 || C == 64, "C");
};

template <int C, int W>
__global__ void __launch_bounds__(256) __maxnreg__(C == 16 ? LC_WGRAD16_MAXNREG : 168) wgrad3x3_tc_kernel(WgradTcArgs a) {
    using K = WgradTcCfg<C, W>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sX = smem_raw;
    unsigned char* sY = smem_raw + K::OFF_Y;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    float* s_coef = reinterpret_cast<float*>(smem_raw + K::OFF_RED + 8192);      // c0, c1, c2 of the fused BatchNorm-backward apply
    float* s_pro = s_coef + 3 * C;                                              // scale, shift of the producer BatchNorm (input prologue)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = a.B * K::PP;
    const int ntiles = (total + 127) / 128;

    // lazy BatchNorm-backward coefficients, phase 1 (independent loads first; combined below)
    const bool dyaff = a.dy_y != nullptr;
    const bool dylazy = dyaff && a.dy_blazy.partial != nullptr;
    if (dylazy) bn_partial_sums_load(a.dy_blazy.partial, a.dy_blazy.nparts, C, reinterpret_cast<double*>(smem_raw + K::OFF_RED));
    if (tid == 32) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);

    const int j = tid % K::CH8, r0 = tid / K::CH8;               // this thread's 8-channel chunk column and first row
    const bool pro = a.pro_scale != nullptr;
    // flattened padded position -> source pixel index (-1: border / outside the batch); PP and WP are compile-time, so the divisions are multiplies
    auto src_of = [&](int Q) -> int {
        if (Q < 0 || Q >= total) return -1;
        const int n = Q / K::PP, rem = Q - n * K::PP;
        const int hp = rem / K::WP, wp = rem - hp * K::WP;
        return (hp >= 1 && hp <= W && wp >= 1 && wp <= W) ? (n * W + (hp - 1)) * W + (wp - 1) : -1;
    };
    // ---- software pipeline: the global loads of tile k+1 (raw fp32, registers) are issued right after tile k's MMAs and land while the tensor core
    //      works; affine transforms + BF16 conversion happen when the registers are stored to shared memory -------------------------------------
    float4 xa[K::NEX], xb[K::NEX], ya[K::NEY], yb[K::NEY], ua[K::NEY], ub[K::NEY];
    uint32_t xvalid = 0, yvalid = 0;
    auto load_tile = [&](int tile) {
        const int q0 = tile * 128;
        xvalid = 0; yvalid = 0;
#pragma unroll
        for (int i = 0; i < K::NEX; ++i) {
            const int r = r0 + i * K::RSTEP;
            const int src = r < K::ROWS_X ? src_of(q0 - K::HALO + r) : -1;
            if (src >= 0) {
                const float* g = a.in + (size_t)src * C + j * 8;
                xa[i] = ldg4(g); xb[i] = ldg4(g + 4);
                xvalid |= 1u << i;
            }
        }
#pragma unroll
        for (int i = 0; i < K::NEY; ++i) {
            const int m = r0 + i * K::RSTEP;
            const int src = m < 128 ? src_of(q0 + m) : -1;
            if (src >= 0) {
                const float* g = a.dy + (size_t)src * C + j * 8;
                ya[i] = ldg4(g); yb[i] = ldg4(g + 4);
                if (dyaff) { const float* yy = a.dy_y + (size_t)src * C + j * 8; ua[i] = ldg4(yy); ub[i] = ldg4(yy + 4); }
                yvalid |= 1u << i;
            }
        }
    };
    auto store_tile = [&]() {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < K::NEX; ++i) {
            const int r = r0 + i * K::RSTEP;
            if (r < K::ROWS_X) {
                float4 p = z4, q = z4;
                if (xvalid & (1u << i)) {
                    p = xa[i]; q = xb[i];
                    if (pro) {
                        const float4 sc0 = *reinterpret_cast<const float4*>(s_pro + j * 8), sc1 = *reinterpret_cast<const float4*>(s_pro + j * 8 + 4);
                        const float4 sh0 = *reinterpret_cast<const float4*>(s_pro + C + j * 8), sh1 = *reinterpret_cast<const float4*>(s_pro + C + j * 8 + 4);
                        p.x = fmaxf(fmaf(p.x, sc0.x, sh0.x), 0.f); p.y = fmaxf(fmaf(p.y, sc0.y, sh0.y), 0.f);
                        p.z = fmaxf(fmaf(p.z, sc0.z, sh0.z), 0.f); p.w = fmaxf(fmaf(p.w, sc0.w, sh0.w), 0.f);
                        q.x = fmaxf(fmaf(q.x, sc1.x, sh1.x), 0.f); q.y = fmaxf(fmaf(q.y, sc1.y, sh1.y), 0.f);
                        q.z = fmaxf(fmaf(q.z, sc1.z, sh1.z), 0.f); q.w = fmaxf(fmaf(q.w, sc1.w, sh1.w), 0.f);
                    }
                }
                const uint4 v = make_uint4(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w), pack_bf16(q.x, q.y), pack_bf16(q.z, q.w));
#pragma unroll
                for (int c = 0; c < 3; ++c) {      // copy c holds tile row r at copy row r-(c-1)
                    const int rr = r - (c - 1);
                    if (rr >= 0 && rr < K::ROWS_X) *reinterpret_cast<uint4*>(sX + (size_t)((c * K::CH8 + j) * K::PLANE_X + rr * 16)) = v;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < K::NEY; ++i) {
            const int m = r0 + i * K::RSTEP;
            if (m < 128) {
                float4 p = z4, q = z4;
                if (yvalid & (1u << i)) {
                    p = ya[i]; q = yb[i];
                    if (dyaff) {      // dY = c0*g + c1*y + c2 (same fma order as bn_bwd_apply_kernel)
                        const float4 k0a = *reinterpret_cast<const float4*>(s_coef + j * 8), k0b = *reinterpret_cast<const float4*>(s_coef + j * 8 + 4);
                        const float4 k1a = *reinterpret_cast<const float4*>(s_coef + C + j * 8), k1b = *reinterpret_cast<const float4*>(s_coef + C + j * 8 + 4);
                        const float4 k2a = *reinterpret_cast<const float4*>(s_coef + 2 * C + j * 8), k2b = *reinterpret_cast<const float4*>(s_coef + 2 * C + j * 8 + 4);
                        p.x = fmaf(k0a.x, p.x, fmaf(k1a.x, ua[i].x, k2a.x)); p.y = fmaf(k0a.y, p.y, fmaf(k1a.y, ua[i].y, k2a.y));
                        p.z = fmaf(k0a.z, p.z, fmaf(k1a.z, ua[i].z, k2a.z)); p.w = fmaf(k0a.w, p.w, fmaf(k1a.w, ua[i].w, k2a.w));
                        q.x = fmaf(k0b.x, q.x, fmaf(k1b.x, ub[i].x, k2b.x)); q.y = fmaf(k0b.y, q.y, fmaf(k1b.y, ub[i].y, k2b.y));
                        q.z = fmaf(k0b.z, q.z, fmaf(k1b.z, ub[i].z, k2b.z)); q.w = fmaf(k0b.w, q.w, fmaf(k1b.w, ub[i].w, k2b.w));
                    }
                }
                *reinterpret_cast<uint4*>(sY + (size_t)(j * K::PLANE_Y + m * 16)) =
                    make_uint4(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w), pack_bf16(q.x, q.y), pack_bf16(q.z, q.w));
            }
        }
    };

    int tile = blockIdx.x;
    if (tile < ntiles) load_tile(tile);          // in flight across the coefficient reduction below

    if (pro && tid < 2 * C) s_pro[tid] = tid < C ? a.pro_scale[tid] : a.pro_shift[tid - C];
    if (dylazy) bn_bwd_lazy_coef_finish(a.dy_blazy, C, reinterpret_cast<double*>(smem_raw + K::OFF_RED), s_coef);
    else if (dyaff && tid < 3 * C) s_coef[tid] = a.dy_coef[tid];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sX_u = smem_u32(sX), sY_u = smem_u32(sY);

    uint32_t phase = 0, first = 1;
    for (; tile < ntiles; tile += gridDim.x) {
        store_tile();
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

        // ---- MMAs: per dr, 8 K-steps of 16 pixels.  M operand = displaced input copies, N operand = dY ---------------------------
        if (tid == 0) {
            const uint64_t x_hi = make_desc(0, 128, K::PLANE_X);      // LBO = 128 B (next 8 rows), SBO = plane (next 8 channels)
            const uint64_t y_hi = make_desc(0, 128, K::PLANE_Y);
            const uint32_t xBase = sX_u >> 4, yBase = sY_u >> 4;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const uint32_t xrow = xBase + (uint32_t)(K::HALO + (d - 1) * K::WP);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                    const uint64_t yd = y_hi | (uint64_t)((yBase + (uint32_t)(ks * 16)) & 0x3FFF);
                    const uint64_t xd = x_hi | (uint64_t)((xrow + (uint32_t)(ks * 16)) & 0x3FFF);
                    if constexpr (C == 16) {
                        mma_f16(tmem_base + (uint32_t)(d * 16), xd, yd, make_idesc_bf16_mn(64, 16), acc);
                    } else if constexpr (C == 32) {
                        mma_f16(tmem_base + (uint32_t)(d * 32), xd, yd, make_idesc_bf16_mn(128, 32), acc);
                    } else {
                        mma_f16(tmem_base + (uint32_t)(d * 128), xd, yd, make_idesc_bf16_mn(128, 64), acc);
                        const uint64_t xd2 = x_hi | (uint64_t)((xrow + (uint32_t)(ks * 16) + (uint32_t)(2 * K::CH8) * (K::PLANE_X >> 4)) & 0x3FFF);
                        mma_f16(tmem_base + (uint32_t)(d * 128 + 64), xd2, yd, make_idesc_bf16_mn(64, 64), acc);
                    }
                }
            }
            mma_commit(bar);
        }
        first = 0;
        if (tile + (int)gridDim.x < ntiles) load_tile(tile + (int)gridDim.x);      // next tile's loads fly while the tensor core reads this one
        const bool done = mbar_wait(bar, phase);      // MMAs complete: shared memory may be overwritten
        phase ^= 1;
        fence_after_sync();
        if (!done && tid == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);
    }

    // ---- epilogue: D rows = (dc, ci), columns = co  ->  partial[co][ci][tap] ------------------------------------------------------
    float* dst = a.partial + (size_t)blockIdx.x * C * C * 9;
    const bool has_tiles = (int)blockIdx.x < ntiles;
    const int quarter = warp & 3, grp = warp >> 2;
    for (int acc_i = grp; acc_i < K::NACC; acc_i += 2) {
        int d, col0, mrows, copy0;
        if constexpr (C == 64) { d = acc_i >> 1; col0 = d * 128 + (acc_i & 1) * 64; mrows = (acc_i & 1) ? 64 : 128; copy0 = (acc_i & 1) ? 2 : 0; }
        else { d = acc_i; col0 = d * C; mrows = C == 16 ? 64 : 128; copy0 = 0; }
        // M=128: row = TMEM lane.  M=64: row 16*q + l lives in lane 32*q + l (l < 16)   [verified by tools/tc_probe.cu]
        const int m = mrows == 128 ? quarter * 32 + lane : quarter * 16 + lane;
        const int useful = C == 16 ? 48 : (C == 32 ? 96 : mrows);      // rows backed by real (dc, ci) blocks
        const bool row_ok = (mrows == 128 || lane < 16) && m < useful;
        const int blk = m >> 3, ci = (blk % K::CH8) * 8 + (m & 7), dc = copy0 + blk / K::CH8;     // dc index 0..2
        const int tap = d * 3 + dc;
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col0 + c0), v);
            if (row_ok) {      // partial layout [tap][ci][co]: 64 contiguous bytes per thread (the reduce kernel restores OIHW)
                float4* o4 = reinterpret_cast<float4*>(dst + ((size_t)tap * C + ci) * C + c0);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    o4[k4] = has_tiles ? make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

template <int C, int W>
static inline int wgrad_tc_launch(const WgradTcArgs& a, int nsplit, cudaStream_t st) {
    using K = WgradTcCfg<C, W>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(wgrad3x3_tc_kernel<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    wgrad3x3_tc_kernel<C, W><<<nsplit, K::NT, K::SMEM_BYTES, st>>>(a);
    return lc_launch_status();
}

}  // namespace tc
}  // namespace lc
