// Dense GEMM on the tcgen05 tensor cores for the ViT-B/16 blocks (core/model/backbone/transformer.py:169-197 MultiHeadAttention,
// :1255-1273 Mlp, :1331-1336 ResidualAttentionBlock): D[M][N] = A[M][K] * B[N][K]^T, BF16 operands, fp32 accumulation in TMEM,
// fused epilogues (bias, exact GELU, fp32 residual add, BF16 / fp32 output, optional transposed per-head store).
//
// Structure (persistent CTAs, one per SM, looping over 128 x BN output tiles; warp specialised, mbarrier pipelines; two TMEM accumulators so
// that the epilogue of one tile overlaps the MMAs of the next):
//   warp 0 : TMA producer — cp.async.bulk.tensor (rank-4 maps: K, rows, inner batch, outer batch) loads of A (128 x 64) and B (BN x 64) K-blocks into a STAGES-deep ring of
//            128-byte-swizzled shared-memory tiles (CU_TENSOR_MAP_SWIZZLE_128B), arriving on full[stage] with expect_tx bytes
//   warp 1 : MMA issuer — one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M128 x BN x K16) per K-block with
//            SWIZZLE_128B K-major shared-memory descriptors (SBO = 1024 B, start advanced by 32 B per K16 step), commits the stage back
//            to the producer (empty[stage]) and, after the last K-block, the accumulator to the epilogue (tmem_full)
//   warps 2-9 (2-17 for the GELU variants) : epilogue (two / four warps per TMEM lane quarter, interleaved 32-column slabs) — tcgen05.ld 32x32b (warp w owns TMEM lanes 32*(w%4)..), staged through shared memory so that bias / GELU /
//            residual and the global stores run row-wise (128-byte row segments per 8 lanes), then the accumulator is handed back (tmem_empty)
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
    int batch_in, batch_total;
    int a_bcast, b_bcast;      // bit 0 / bit 1: the operand is shared across the inner / outer batch (its tensor map has extent 1 there)
    int out_dtype;
    float alpha;               // scale applied to the accumulator before bias (attention: 1/sqrt(d))
    int* error_flag;
    // split-K: the K blocks are divided over `ksplit` partial outputs (fp32, partial s at out + s * c_stride_split); 0 / 1 = off
    int ksplit;
    long long c_stride_split;
    // implicit convolution (conv_ks > 0): A is an NHWC BF16 activation addressed through a rank-4 {C, W, H, N} tensor map whose box is one 128-pixel
    // output tile (Nt images x Ht rows x Wo columns, element strides = conv stride); K block kb = tap * conv_cchunks + channel chunk, and the tap is a
    // coordinate offset of the box (zero fill outside the image = the padding).  No im2col matrix exists anywhere.
    int conv_ks, conv_cchunks, conv_stride, conv_pad, conv_hw, conv_wo;
    int gelu_mode;             // bit 0: the GELU side tensor holds GELU'(pre-activation) instead of the pre-activation itself — `out` of the out2 variant
                               //        stores it, `gelu_aux` is then a plain multiplier (the backward epilogue drops from 25 to ~8 instructions per
                               //        element; the frozen-backbone backward needs the pre-activation for nothing else).  bit 1: do not store `out`
                               //        at all (no-grad passes keep only GELU(x) in out2)
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
// Exact (erf) GELU and its derivative with erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the BF16 rounding of every
// consumer) on the MUFU approximations: one rcp, one ex2 and ~12 FMA-pipe instructions per element — the GELU epilogues were
// instruction-issue-bound with libdevice erff (ncu: 43k warp instructions per 128x256 tile).  exp2(-x^2 log2(e)/2) serves both the erf
// tail and the Gaussian density of the derivative.
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& gauss) {
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    gauss = ex2_approx(-0.72134752044448170368f * x * x);            // exp(-x^2 / 2)
    const float erf_abs = fmaf(-p * t, gauss, 1.f);                    // erf(|x| / sqrt(2))
    cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);                     // Phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) { float c, g; gelu_parts(x, c, g); return x * c; }
__device__ __forceinline__ float dgelu_erf(float x) { float c, g; gelu_parts(x, c, g); return fmaf(x * 0.3989422804014327f, g, c); }

enum { EPI_F32 = 1, EPI_RES = 2, EPI_GELU2 = 4, EPI_DGELU = 8 };     // compile-time epilogue variants (keeps each instance's code small)

// CG = 2: a pair of CTAs (a 2-CTA cluster on one TPC) computes a 256 x BN block with tcgen05.mma.cta_group::2 — each CTA stages its own 128 rows of A
// and HALF of the B panel (BN / 2 rows), the tensor cores of both SMs read both halves, each CTA keeps its 128 accumulator rows in its own TMEM.
// Per K block a CTA pulls 32 KB through its L2 port instead of 48 KB for the same math: ncu showed the single-CTA kernel at 51 % tensor pipe with
// no DRAM / L2 limiter except the per-SM fill rate (profiles/r2k_ncu_gemm_fc1.txt).
template <int BN, int EPI = 0, int CG = 1>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / CG) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // GELU / GELU' epilogues are instruction-issue bound (ncu: 71 M warp instructions vs 9 M for the plain store, 25 per element): with two
    // epilogue warps per scheduler their dependent MUFU / FMA chains leave the issue slots idle, so those variants run four per scheduler
    static constexpr int EPI_WARPS = ((EPI & (EPI_GELU2 | EPI_DGELU)) != 0 && (EPI & EPI_RES) == 0) ? 16 : 8;
    static constexpr int STAGES = CG == 2 ? (EPI_WARPS == 16 ? 4 : 5) : (BN == 256 ? 3 : (EPI_WARPS == 16 ? 4 : 5));
    static constexpr int NT = 64 + 32 * EPI_WARPS;
    static constexpr int SLAB = 32;                                   // epilogue column slab
    static constexpr int STG_LD = SLAB + 4;                           // staging row stride (floats): conflict-free float4 rows
    static constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;     // per epilogue warp: 32 rows
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + STG_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr uint32_t TMEM_COLS = 2 * BN;                     // two accumulators: the epilogue of tile i overlaps the MMAs of tile i+1
    static_assert(BN == 128 || BN == 256, "BN");
    static_assert(CG == 1 || (CG == 2 && BN == 256), "the CTA-pair variant exists for 256-wide panels");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

#ifdef LC_GEMM_TIMING
// debug build only (tools/gemm_timing.py): per CTA and tile, globaltimer stamps of the three roles
__device__ unsigned long long g_gemm_tstamp[148 * 32 * 8];
__device__ __forceinline__ unsigned long long gemm_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LC_GSTAMP(tile_local, slot) do { if ((tile_local) < 32 && blockIdx.x < 148) g_gemm_tstamp[((size_t)blockIdx.x * 32 + (tile_local)) * 8 + (slot)] = gemm_gtimer(); } while (0)
#else
#define LC_GSTAMP(tile_local, slot) do { } while (0)
#endif

// mbar_arrive: conv_tc.cuh
// ---- CTA-pair (cta_group::2) primitives ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address with the CTA-pair peer bit cleared = the same offset in the leader CTA
// executed by both CTAs of a pair: the tile lands in THIS CTA's shared memory, the transaction bytes are reported to the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(tmap), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of every MMA issued so far by this thread -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Persistent: grid = min(#tiles, #SMs); CTA c processes tiles c, c + grid, ...  Tile order: n fastest inside an m row-block, so the CTAs that run
// concurrently share A row-blocks through L2.  A (the token matrix, up to 155 MB) is the operand that does not fit the 126 MB L2, the weights (<= 4.7 MB)
// always do: with m fastest every n panel streamed A from DRAM again (ncu: 548 MB read for 237 MB algorithmic on fc2, N = 768 -> 3 panels).
template <int BN, int EPI, int CG = 1>
__global__ void __launch_bounds__(GemmCfg<BN, EPI, CG>::NT, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs a) {
    using K = GemmCfg<BN, EPI, CG>;
    // CG == 2: blockIdx.x = 2 * pair + rank; a pair walks the 256 x BN blocks, rank r owns rows [128 r, 128 r + 128) of a block and rows
    // [128 r, ..) of its B panel; only the leader (rank 0) issues MMAs, both ranks load, both drain their own TMEM half
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
    const int worker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nworkers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int MB = K::BM * CG;                  // rows of an output block
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base_u = (smem_u32(smem_dyn) + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* base = smem_dyn + (base_u - smem_u32(smem_dyn));
    float* staging = reinterpret_cast<float*>(base + K::STAGES * K::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + K::STAGES * K::STAGE_BYTES + K::STG_BYTES);
    uint64_t* empty = full + K::STAGES;
    uint64_t* tmem_full = empty + K::STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (a.K + K::BK - 1) / K::BK;
    const int m_tiles = (a.M + MB - 1) / MB, n_tiles = (a.N + BN - 1) / BN;
    const int per_z = m_tiles * n_tiles;
    const int ksplit = a.ksplit > 1 ? a.ksplit : 1;
    const int kper = (nkb + ksplit - 1) / ksplit;
    const int total = per_z * a.batch_total * ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, K::EPI_WARPS * CG); }    // CG == 2: both CTAs' epilogue warps
    }
    if (warp == 1) { if (CG == 2) tmem_alloc_2sm(tmem_slot, K::TMEM_COLS); else tmem_alloc(tmem_slot, K::TMEM_COLS); }
    fence_before_sync();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // (pair) barriers initialised before any remote arrival / multicast commit
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    bool ok = true;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            int tl = 0; (void)tl;
            for (int tile = worker; tile < total; tile += nworkers) {
                const int sp = tile % ksplit, tq = tile / ksplit;
                const int z = tq / per_z, r = tq % per_z;
                const int n0 = (r % n_tiles) * BN, m0 = (r / n_tiles) * MB + (int)crank * K::BM;
                const int z_in = z % a.batch_in, z_out = z / a.batch_in;
                const int kb0 = sp * kper, kb1 = kb0 + kper < nkb ? kb0 + kper : nkb;
                const int cv_n = a.conv_ks > 0 ? m0 / a.conv_hw : 0, cv_h = a.conv_ks > 0 ? (m0 % a.conv_hw) / a.conv_wo * a.conv_stride - a.conv_pad : 0;
                LC_GSTAMP(tl, 0);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % K::STAGES;
                    const uint32_t ph = (it / K::STAGES) & 1u;
                    ok = mbar_wait(empty + s, ph ^ 1u) && ok;
                    const uint32_t sa = base_u + (uint32_t)s * K::STAGE_BYTES;
                    if constexpr (CG == 2) {
                        // both CTAs' bytes are reported to the leader's full[s]; only the leader arrives on it
                        if (crank == 0) mbar_expect_tx(full + s, (uint32_t)(2 * K::STAGE_BYTES));
                        tma_load_4d_2sm(sa, &tmA, full + s, kb * K::BK, m0, (a.a_bcast & 1) ? 0 : z_in, (a.a_bcast & 2) ? 0 : z_out);
                        tma_load_4d_2sm(sa + K::A_BYTES, &tmB, full + s, kb * K::BK, n0 + (int)crank * (BN / 2), (a.b_bcast & 1) ? 0 : z_in,
                                        (a.b_bcast & 2) ? 0 : z_out);
                        continue;
                    }
                    mbar_expect_tx(full + s, (uint32_t)K::STAGE_BYTES);
                    if (a.conv_ks > 0) {
                        const int tap = kb / a.conv_cchunks, chunk = kb - tap * a.conv_cchunks;
                        const int kh = tap / a.conv_ks, kw = tap - kh * a.conv_ks;
                        tma_load_4d(sa, &tmA, full + s, chunk * K::BK, kw - a.conv_pad, cv_h + kh, cv_n);
                    } else {
                        tma_load_4d(sa, &tmA, full + s, kb * K::BK, m0, (a.a_bcast & 1) ? 0 : z_in, (a.a_bcast & 2) ? 0 : z_out);
                    }
                    tma_load_4d(sa + K::A_BYTES, &tmB, full + s, kb * K::BK, n0, (a.b_bcast & 1) ? 0 : z_in, (a.b_bcast & 2) ? 0 : z_out);
                }
                LC_GSTAMP(tl, 1);
                ++tl;
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN);
            uint32_t it = 0, t = 0;
            for (int tile = worker; tile < total; tile += nworkers, ++t) {
                const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
                LC_GSTAMP(t, 2);
                ok = mbar_wait(tmem_empty + acc, aph ^ 1u) && ok;                   // the epilogue has drained this accumulator
                fence_after_sync();
                LC_GSTAMP(t, 3);
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int sp = tile % ksplit;
                const int kb0 = sp * kper, kb1 = kb0 + kper < nkb ? kb0 + kper : nkb;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % K::STAGES;
                    const uint32_t ph = (it / K::STAGES) & 1u;
                    ok = mbar_wait(full + s, ph) && ok;
                    fence_after_sync();
                    const uint32_t sa = base_u + (uint32_t)s * K::STAGE_BYTES;
                    const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + K::A_BYTES);
#pragma unroll
                    for (int k = 0; k < K::BK / 16; ++k) {                         // +32 B (2 x 16-byte units) per K16 step inside the swizzle atom
                        if constexpr (CG == 2) mma_f16_2sm(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                        else mma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (CG == 2) mma_commit_2sm(empty + s); else mma_commit(empty + s);     // frees the stage (in both CTAs) when these MMAs have read it
                }
                if constexpr (CG == 2) mma_commit_2sm(tmem_full + acc); else mma_commit(tmem_full + acc);
                LC_GSTAMP(t, 4);
            }
        }
    } else {
        // ---- epilogue warps 2..9: TMEM lane quarter = warp % 4, slab parity = (warp - 2) / 4.  Per 32-column slab: TMEM -> registers (thread = row) -> shared staging ->
        //      row-wise pass (8 lanes x float4 = one 128-byte row segment, 4 rows per instruction): bias / residual / GELU, coalesced stores
        const int quarter = warp & 3, spar = (warp - 2) >> 2;
        float* stg = staging + (size_t)(warp - 2) * 32 * K::STG_LD;
        const int rr = lane >> 3, c4 = (lane & 7) * 4;
        uint32_t t = 0;
        for (int tile = worker; tile < total; tile += nworkers, ++t) {
            const int sp = tile % ksplit, tq = tile / ksplit;
            const int z = tq / per_z, r = tq % per_z;
            const int n0 = (r % n_tiles) * BN, m0 = (r / n_tiles) * MB + (int)crank * K::BM;
            const int z_in = z % a.batch_in, z_out = z / a.batch_in;
            const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
            if (warp == 2 && lane == 0) LC_GSTAMP(t, 5);
            ok = mbar_wait(tmem_full + acc, aph) && ok;
            fence_after_sync();
            if (warp == 2 && lane == 0) LC_GSTAMP(t, 6);
            const uint32_t trow = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            const size_t cz = (size_t)z_out * a.c_stride_out + (size_t)z_in * a.c_stride_in + (size_t)sp * (size_t)a.c_stride_split;
            const size_t rz = (size_t)z_out * a.r_stride_out + (size_t)z_in * a.r_stride_in;
            const int mrow0 = m0 + quarter * 32;
#pragma unroll 1
            for (int c0 = spar * K::SLAB; c0 < BN && n0 + c0 < a.N; c0 += (K::EPI_WARPS / 4) * K::SLAB) {
                const int n = n0 + c0 + c4;
                const bool vec = n + 3 < a.N;
                float4 res[8];
                if constexpr ((EPI & EPI_RES) != 0) {                                // all of the slab's residual loads in flight before the TMEM read
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int m = mrow0 + j * 4 + rr;
                        res[j] = (vec && m < a.M) ? *reinterpret_cast<const float4*>(a.residual + rz + (size_t)m * a.ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                uint2 gaux[8];
                if constexpr ((EPI & EPI_DGELU) != 0) {                              // likewise the pre-activations GELU' is evaluated at: issued here, consumed
#pragma unroll                                                                        // after the TMEM read (inside the store loop they would serialise behind the
                    for (int j = 0; j < 8; ++j) {                                    // stores: `out` and `gelu_aux` may alias as far as the compiler knows)
                        const int m = mrow0 + j * 4 + rr;
                        gaux[j] = (vec && m < a.M) ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.gelu_aux) + cz + (size_t)m * a.ldc + n))
                                                   : make_uint2(0u, 0u);
                    }
                }
                float v[32];
                tmem_ld16(trow + (uint32_t)c0, v);
                tmem_ld16(trow + (uint32_t)(c0 + 16), v + 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stg + lane * K::STG_LD + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.bias != nullptr) {
                    if (vec) b4 = ldg4(a.bias + n);
                    else { if (n < a.N) b4.x = a.bias[n]; if (n + 1 < a.N) b4.y = a.bias[n + 1]; if (n + 2 < a.N) b4.z = a.bias[n + 2]; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r0 = j * 4;
                    const int m = mrow0 + r0 + rr;
                    if (m >= a.M || n >= a.N) continue;
                    const float4 acc4 = *reinterpret_cast<const float4*>(stg + (r0 + rr) * K::STG_LD + c4);
                    float x0 = fmaf(acc4.x, a.alpha, b4.x), x1 = fmaf(acc4.y, a.alpha, b4.y), x2 = fmaf(acc4.z, a.alpha, b4.z), x3 = fmaf(acc4.w, a.alpha, b4.w);
                    const size_t ci = cz + (size_t)m * a.ldc + n;
                    if (vec) {
                        if constexpr ((EPI & EPI_RES) != 0) { x0 += res[j].x; x1 += res[j].y; x2 += res[j].z; x3 += res[j].w; }
                        if constexpr ((EPI & EPI_DGELU) != 0) {
                            const uint2 g = gaux[j];
                            const float a0 = __uint_as_float(g.x << 16), a1 = __uint_as_float(g.x & 0xffff0000u), a2 = __uint_as_float(g.y << 16),
                                        a3 = __uint_as_float(g.y & 0xffff0000u);
                            if (a.gelu_mode & 1) { x0 *= a0; x1 *= a1; x2 *= a2; x3 *= a3; }
                            else { x0 *= dgelu_erf(a0); x1 *= dgelu_erf(a1); x2 *= dgelu_erf(a2); x3 *= dgelu_erf(a3); }
                        }
                        if constexpr ((EPI & EPI_GELU2) != 0) {
                            float c0_, c1_, c2_, c3_, g0_, g1_, g2_, g3_;
                            gelu_parts(x0, c0_, g0_); gelu_parts(x1, c1_, g1_); gelu_parts(x2, c2_, g2_); gelu_parts(x3, c3_, g3_);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out2) + ci) = make_uint2(pack_bf16(x0 * c0_, x1 * c1_), pack_bf16(x2 * c2_, x3 * c3_));
                            if (a.gelu_mode & 1) {                                   // GELU'(x) = Phi(x) + x phi(x)
                                x0 = fmaf(x0 * 0.3989422804014327f, g0_, c0_); x1 = fmaf(x1 * 0.3989422804014327f, g1_, c1_);
                                x2 = fmaf(x2 * 0.3989422804014327f, g2_, c2_); x3 = fmaf(x3 * 0.3989422804014327f, g3_, c3_);
                            }
                            if (a.gelu_mode & 2) continue;
                        }
                        if constexpr ((EPI & EPI_F32) != 0) *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + ci) = make_float4(x0, x1, x2, x3);
                        else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out) + ci) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
                    } else {
                        const float xs[4] = {x0, x1, x2, x3};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {                               // ragged last columns: scalar
                            if (n + i >= a.N) break;
                            float xi = xs[i];
                            if constexpr ((EPI & EPI_RES) != 0) xi += a.residual[rz + (size_t)m * a.ldr + n + i];
                            if constexpr ((EPI & EPI_DGELU) != 0) {
                                const float av = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.gelu_aux)[ci + i]);
                                xi *= (a.gelu_mode & 1) ? av : dgelu_erf(av);
                            }
                            if constexpr ((EPI & EPI_GELU2) != 0) {
                                reinterpret_cast<__nv_bfloat16*>(a.out2)[ci + i] = __float2bfloat16_rn(gelu_erf(xi));
                                if (a.gelu_mode & 1) xi = dgelu_erf(xi);
                                if (a.gelu_mode & 2) continue;
                            }
                            if constexpr ((EPI & EPI_F32) != 0) reinterpret_cast<float*>(a.out)[ci + i] = xi;
                            else reinterpret_cast<__nv_bfloat16*>(a.out)[ci + i] = __float2bfloat16_rn(xi);
                        }
                    }
                }
                __syncwarp();
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tmem_empty + acc, 0); else mbar_arrive(tmem_empty + acc); }     // the leader's MMA thread waits on it
            if (warp == 2 && lane == 0) LC_GSTAMP(t, 7);
        }
    }
    if (!ok && lane == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 2);
    fence_before_sync();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // (pair) the peer's shared memory and TMEM stay alive until the last MMA / remote arrival
    if (warp == 1) { if (CG == 2) tmem_dealloc_2sm(tmem_base, K::TMEM_COLS); else tmem_dealloc(tmem_base, K::TMEM_COLS); }
}

}  // namespace tc
}  // namespace lc
