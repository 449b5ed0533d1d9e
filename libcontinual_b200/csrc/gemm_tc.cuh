// Dense GEMM on the tcgen05 tensor cores for the ViT-B/16 blocks (core/model/backbone/transformer.py:169-197 MultiHeadAttention,
// :1255-1273 Mlp, :1331-1336 ResidualAttentionBlock): D[M][N] = A[M][K] * B[N][K]^T, BF16 operands, fp32 accumulation in TMEM,
// fused epilogues (bias, exact GELU, fp32 residual add, BF16 / fp32 output, optional transposed per-head store).
//
// Structure (one 128 x BN output tile per CTA, warp specialised, mbarrier pipelines):
//   warp 0 : TMA producer — cp.async.bulk.tensor (rank-4 maps: K, rows, inner batch, outer batch) loads of A (128 x 64) and B (BN x 64) K-blocks into a STAGES-deep ring of
//            128-byte-swizzled shared-memory tiles (CU_TENSOR_MAP_SWIZZLE_128B), arriving on full[stage] with expect_tx bytes
//   warp 1 : MMA issuer — one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M128 x BN x K16) per K-block with
//            SWIZZLE_128B K-major shared-memory descriptors (SBO = 1024 B, start advanced by 32 B per K16 step), commits the stage back
//            to the producer (empty[stage]) and, after the last K-block, the accumulator to the epilogue (tmem_full)
//   warps 2-5 : epilogue — tcgen05.ld 32x32b (warp w owns TMEM lanes 32*(w%4)..), bias / GELU / residual, vectorised stores
// Operands are addressed through tensor maps (row stride and batch strides arbitrary), so the same kernel runs the per-head
// batched attention GEMMs (Q K^T, P V) on strided views of the fused QKV buffer.
#pragma once
#include "wgrad_tc.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace lc {
namespace tc {

enum { GEMM_OUT_BF16 = 0, GEMM_OUT_F32 = 1 };

struct GemmArgs {
    void* out;                 // [batch][M][ldc] (bf16 or fp32)
    const float* bias;         // nullable [N]
    const float* residual;     // nullable fp32 [batch][M][ldr], added before the store
    void* out2;                // nullable bf16: GELU(out) (out then holds the pre-activation, needed by the backward)
    const void* gelu_aux;      // nullable bf16, indexed like out: the result is multiplied by GELU'(gelu_aux) (fc2 backward-data -> d pre-activation)
    int M, N, K;
    int ldc, ldr;
    long long c_stride_in, c_stride_out, r_stride_in, r_stride_out;   // elements; batch index z = z_out * batch_in + z_in
    int batch_in;
    int a_bcast, b_bcast;      // bit 0 / bit 1: the operand is shared across the inner / outer batch (its tensor map has extent 1 there)
    int out_dtype;
    float alpha;               // scale applied to the accumulator before bias (attention: 1/sqrt(d))
    int* error_flag;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(tmap),
                 "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// K-major, SWIZZLE_128B (layout type 2), SBO = 1024 B (8 rows x 128 B), LBO unused (1)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int NT = 192;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr uint32_t TMEM_COLS = BN;   // 128 or 256 (power of two >= 32)
    static_assert(BN == 128 || BN == 256 || BN == 64, "BN");
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs a) {
    using K = GemmCfg<BN>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base_u = (smem_u32(smem_dyn) + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* base = smem_dyn + (base_u - smem_u32(smem_dyn));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + K::STAGES * K::STAGE_BYTES);
    uint64_t* empty = full + K::STAGES;
    uint64_t* tmem_full = empty + K::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * K::BM;
    const int z_in = (int)blockIdx.z % a.batch_in, z_out = (int)blockIdx.z / a.batch_in;
    const int nkb = (a.K + K::BK - 1) / K::BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
    }
    if (warp == 1) tmem_alloc(tmem_slot, K::TMEM_COLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % K::STAGES;
                const uint32_t ph = (uint32_t)(kb / K::STAGES) & 1u;
                mbar_wait(empty + s, ph ^ 1u);                                     // slot free (first lap passes immediately)
                mbar_expect_tx(full + s, (uint32_t)K::STAGE_BYTES);
                const uint32_t sa = base_u + (uint32_t)s * K::STAGE_BYTES;
                tma_load_4d(sa, &tmA, full + s, kb * K::BK, m0, (a.a_bcast & 1) ? 0 : z_in, (a.a_bcast & 2) ? 0 : z_out);
                tma_load_4d(sa + K::A_BYTES, &tmB, full + s, kb * K::BK, n0, (a.b_bcast & 1) ? 0 : z_in, (a.b_bcast & 2) ? 0 : z_out);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % K::STAGES;
                const uint32_t ph = (uint32_t)(kb / K::STAGES) & 1u;
                mbar_wait(full + s, ph);
                fence_after_sync();
                const uint32_t sa = base_u + (uint32_t)s * K::STAGE_BYTES;
                const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + K::A_BYTES);
#pragma unroll
                for (int k = 0; k < K::BK / 16; ++k)                               // +32 B (2 x 16-byte units) per K16 step inside the swizzle atom
                    mma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                mma_commit(empty + s);                                             // frees the stage when these MMAs have read it
            }
            mma_commit(tmem_full);
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4 --------------------------------------------------------------------
        const bool done = mbar_wait(tmem_full, 0);
        fence_after_sync();
        if (!done && lane == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 2);
        const int quarter = warp & 3;
        const int m = m0 + quarter * 32 + lane;
        const bool row_ok = m < a.M;
        const size_t crow = (size_t)z_out * a.c_stride_out + (size_t)z_in * a.c_stride_in + (size_t)m * a.ldc;
        const size_t rrow = (size_t)z_out * a.r_stride_out + (size_t)z_in * a.r_stride_in + (size_t)m * a.ldr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            const int n = n0 + c0;
            if (row_ok && n < a.N && n + 16 > a.N) {      // ragged last chunk (attention: N = #keys; head: N = #classes): scalar path
                for (int i = 0; i < 16 && n + i < a.N; ++i) {
                    float x = v[i] * a.alpha + (a.bias != nullptr ? a.bias[n + i] : 0.f);
                    if (a.residual != nullptr) x += a.residual[rrow + n + i];
                    if (a.gelu_aux != nullptr) x *= dgelu_erf(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.gelu_aux)[crow + n + i]));
                    if (a.out_dtype == GEMM_OUT_F32) reinterpret_cast<float*>(a.out)[crow + n + i] = x;
                    else reinterpret_cast<__nv_bfloat16*>(a.out)[crow + n + i] = __float2bfloat16_rn(x);
                    if (a.out2 != nullptr) reinterpret_cast<__nv_bfloat16*>(a.out2)[crow + n + i] = __float2bfloat16_rn(gelu_erf(x));
                }
            } else if (row_ok && n < a.N) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= a.alpha;
                if (a.bias != nullptr) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 b4 = ldg4(a.bias + n + k4 * 4);
                        v[k4 * 4] += b4.x; v[k4 * 4 + 1] += b4.y; v[k4 * 4 + 2] += b4.z; v[k4 * 4 + 3] += b4.w;
                    }
                }
                if (a.residual != nullptr) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const float4 r4 = *reinterpret_cast<const float4*>(a.residual + rrow + n + k4 * 4);
                        v[k4 * 4] += r4.x; v[k4 * 4 + 1] += r4.y; v[k4 * 4 + 2] += r4.z; v[k4 * 4 + 3] += r4.w;
                    }
                }
                if (a.gelu_aux != nullptr) {
                    const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.gelu_aux) + crow + n);
                    const uint4 a0 = ap[0], a1 = ap[1];
                    const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[2 * i] *= dgelu_erf(__uint_as_float(aw[i] << 16));
                        v[2 * i + 1] *= dgelu_erf(__uint_as_float(aw[i] & 0xffff0000u));
                    }
                }
                if (a.out_dtype == GEMM_OUT_F32) {
                    float* o = reinterpret_cast<float*>(a.out) + crow + n;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) *reinterpret_cast<float4*>(o + k4 * 4) = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + crow + n;
                    uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
                    *reinterpret_cast<uint4*>(o) = p0; *reinterpret_cast<uint4*>(o + 8) = p1;
                }
                if (a.out2 != nullptr) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out2) + crow + n;
                    float g[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) g[i] = gelu_erf(v[i]);
                    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
                    *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_bf16(g[8], g[9]), pack_bf16(g[10], g[11]), pack_bf16(g[12], g[13]), pack_bf16(g[14], g[15]));
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

}  // namespace tc
}  // namespace lc
