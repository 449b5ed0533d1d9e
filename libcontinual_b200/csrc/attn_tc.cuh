// Fused multi-head self-attention for short sequences (<= 256 keys: ViT-B/16 with or without L2P prompts: 197 / 222 tokens, plus the <= 10 prefix
// keys / values of DualPrompt / CodaPrompt) on tcgen05:
//   forward : O = softmax(Q K^T / sqrt(64)) V           (core/model/backbone/transformer.py:169-197)
//   backward: dQ, dK, dV from dO with P recomputed on chip from the saved row log-sum-exp (no T x T matrix ever touches HBM)
// Q, K, V are strided views of the fused QKV buffer [B][T][3][H][64] (BF16); O is [B][T][H*64] (BF16).
//
// Shared-memory operand tiles use ONE layout for every role: "planar" no-swizzle core matrices,
//     element (r, c) of a tile with R rows  ->  (c / 8) * PLANE + r * 16 + (c % 8) * 2 bytes,   PLANE = R * 16 + 16
// (the +16 B pad spreads the 8 chunks of a row over different banks for the staging stores).  The same bytes are
//   * a K-major  operand with MN = r, K = c : descriptor LBO = PLANE (next 8 K elements), SBO = 128 (next 8 rows),   K16 step = +2 planes
//   * an MN-major operand with MN = c, K = r : descriptor LBO = 128 (next 8 K rows),      SBO = PLANE (next 8 MN),    K16 step = +256 B
// so V [keys][d] is the MN-major B operand of P V, K [keys][d] the MN-major B operand of dS K, and the probability tile
// P [queries][keys] is the K-major A operand of P V and the MN-major A operand of P^T dO without any transposed copy.
#pragma once
#include "gemm_tc.cuh"
#include <math_constants.h>

namespace lc {
namespace tc {

constexpr int kAttnD = 64;                       // head dim

#ifdef LC_ATTN_TIMING
// debug build only (tools/attn_timing.py): globaltimer stamps of thread 0 of the first 2048 CTAs, 16 slots each
__device__ unsigned long long g_attn_tstamp[2048 * 16];
__device__ __forceinline__ unsigned long long attn_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LC_ASTAMP(slot) do { const unsigned cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
    if (threadIdx.x == 0 && cta_ < 2048 && (slot) < 16) g_attn_tstamp[cta_ * 16 + (slot)] = attn_gtimer(); } while (0)
#else
#define LC_ASTAMP(slot) do { } while (0)
#endif
constexpr float kLog2e = 1.4426950408889634f;

__host__ __device__ constexpr int attn_plane(int rows) { return rows * 16 + 16; }
__host__ __device__ constexpr int attn_np(int T) { return (T + 15) / 16 * 16; }          // keys padded to the MMA N / K granularity

// instruction descriptor: D = F32, A = B = BF16, M, N, operand majorness (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t attn_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int plane, int kstep) { return make_desc(base + (uint32_t)(kstep * 2 * plane), plane, 128); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int plane, int kstep, int mn0) {
    return make_desc(base + (uint32_t)((mn0 >> 3) * plane + kstep * 256), 128, plane);
}

// stage rows [r0, r0 + nrows) x 64 columns of a token-major bf16 matrix (row stride ld elements) into a planar tile; rows >= rmax are zero
__device__ __forceinline__ void stage_tile(uint32_t dst, int plane, const __nv_bfloat16* src, long long ld, int r0, int nrows, int rmax, int tid, int nt) {
    for (int idx = tid; idx < nrows * 8; idx += nt) {
        const int r = idx >> 3, ch = idx & 7;
        const bool ok = r0 + r < rmax;
        cp_async16(dst + (uint32_t)(ch * plane + r * 16), src + (size_t)(ok ? r0 + r : 0) * ld + ch * 8, ok ? 16u : 0u);
    }
}

// stage the key-side rows of one (batch, head): rows [0, P) from the prefix matrix (row stride pld), rows [P, P + T) from the token-major matrix (row stride
// ld); rows >= P + T up to nrows are zero.  (Prefix keys / values of DualPrompt / CodaPrompt: transformer.py:175-180.)
__device__ __forceinline__ void stage_keys(uint32_t dst, int plane, const __nv_bfloat16* pre, long long pld, int P, const __nv_bfloat16* src, long long ld, int T,
                                           int nrows, int tid, int nt) {
    for (int idx = tid; idx < nrows * 8; idx += nt) {
        const int r = idx >> 3, ch = idx & 7;
        const bool in_pre = r < P, ok = r < P + T;
        const __nv_bfloat16* g = in_pre ? pre + (size_t)r * pld + ch * 8 : src + (size_t)(ok ? r - P : 0) * ld + ch * 8;
        cp_async16(dst + (uint32_t)(ch * plane + r * 16), g, ok ? 16u : 0u);
    }
}
__device__ __forceinline__ float ex2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }


template <int W> __device__ __forceinline__ void tmem_ldw(uint32_t taddr, float* v) {
    if constexpr (W == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}
// One W-column group of a score row: running maximum over the valid keys
template <int W> __device__ __forceinline__ float attn_max_group(uint32_t trow, int c0, int NK, float m) {
    float v[W];
    tmem_ldw<W>(trow + (uint32_t)c0, v);
    if (c0 + W <= NK) {
#pragma unroll
        for (int i = 0; i < W; ++i) m = fmaxf(m, v[i]);
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) if (c0 + i < NK) m = fmaxf(m, v[i]);
    }
    return m;
}
// One W-column group of P = exp2(S c - off) written as BF16 into the planar tile (row `row`); returns the sum of the ROUNDED probabilities
template <int W> __device__ __forceinline__ float attn_p_group(uint32_t trow, int c0, int NK, float c, float off, unsigned char* tile, int PQ, int row) {
    float v[W];
    tmem_ldw<W>(trow + (uint32_t)c0, v);
    const bool full = c0 + W <= NK;
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < W / 8; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = g * 8 + 2 * i;
            float p0 = ex2_fast(fmaf(v[k], c, -off)), p1 = ex2_fast(fmaf(v[k + 1], c, -off));
            if (!full) { p0 = c0 + k < NK ? p0 : 0.f; p1 = c0 + k + 1 < NK ? p1 : 0.f; }
            const uint32_t w = pack_bf16(p0, p1);
            sum += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
            pk[i] = w;
        }
        *reinterpret_cast<uint4*>(tile + (size_t)(((c0 >> 3) + g) * PQ + row * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    return sum;
}
// sum_j P_j dP_j over one W-column group (P from the BF16 tile, dP from TMEM)
template <int W> __device__ __forceinline__ float attn_d_group(uint32_t trow, int c0, const unsigned char* tile, int PQ, int row, float acc) {
    float v[W];
    tmem_ldw<W>(trow + (uint32_t)c0, v);
#pragma unroll
    for (int g = 0; g < W / 8; ++g) {
        const uint4 pa = *reinterpret_cast<const uint4*>(tile + (size_t)(((c0 >> 3) + g) * PQ + row * 16));
        const uint32_t pw[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc = fmaf(__uint_as_float(pw[i] << 16), v[g * 8 + 2 * i], acc);
            acc = fmaf(__uint_as_float(pw[i] & 0xffff0000u), v[g * 8 + 2 * i + 1], acc);
        }
    }
    return acc;
}
// dS = P (dP - D) for one W-column group, written as BF16 into the dS tile
template <int W> __device__ __forceinline__ void attn_ds_group(uint32_t trow, int c0, const unsigned char* ptile, unsigned char* dstile, int PQ, int row, float Dr) {
    float v[W];
    tmem_ldw<W>(trow + (uint32_t)c0, v);
#pragma unroll
    for (int g = 0; g < W / 8; ++g) {
        const size_t o = (size_t)(((c0 >> 3) + g) * PQ + row * 16);
        const uint4 pa = *reinterpret_cast<const uint4*>(ptile + o);
        const uint32_t pw[4] = {pa.x, pa.y, pa.z, pa.w};
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)                                                    // P is exactly 0 on padding rows / columns
            pk[i] = pack_bf16(__uint_as_float(pw[i] << 16) * (v[g * 8 + 2 * i] - Dr), __uint_as_float(pw[i] & 0xffff0000u) * (v[g * 8 + 2 * i + 1] - Dr));
        *reinterpret_cast<uint4*>(dstile + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

struct AttnFwdArgs {
    const __nv_bfloat16* qkv;   // [B][T][3][H][64]
    __nv_bfloat16* out;         // [B][T][H*64]
    float* lse2;                // [B][H][T]  row log-sum-exp of the scaled scores, base 2:  max*c + log2(sum exp2(s*c - max*c)), c = log2(e)/8
    int T, H;
    int* error_flag;
    const __nv_bfloat16* pk;    // nullable [B][P][H*64]: prefix keys / values placed in front of the token keys (transformer.py:175-180)
    const __nv_bfloat16* pv;
    int P;
};

// grid (ceil(T/128), H, B), 256 threads, 2 CTAs / SM (256 TMEM columns, ~85 KB shared memory each).  Two threads share a query row (= TMEM lane): thread
// (row, chalf) owns one half of the key columns for the softmax and 32 of the 64 output columns.
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int T = a.T, P = a.pk != nullptr ? a.P : 0, NK = T + P, NP = attn_np(NK);
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, chalf = tid >> 7;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    constexpr int PQ = attn_plane(128);                 // Q tile and P tile planes (128 rows)
    const int PK = attn_plane(NP);                       // K / V planes (NP rows)
    const uint32_t s_base = smem_u32(smem);
    const int regionA = max(8 * PQ + 8 * PK, (NP / 8) * PQ);
    const uint32_t sQ = s_base, sK = s_base + 8 * PQ, sP = s_base, sV = s_base + (uint32_t)regionA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + regionA + 8 * PK);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_red = reinterpret_cast<float*>(smem + regionA + 8 * PK + 64);          // [256] row maxima, then [256] row sums
    const long long ld = 3LL * a.H * kAttnD;
    const int HD = a.H * kAttnD;
    const __nv_bfloat16* base = a.qkv + (size_t)b * T * ld + h * kAttnD;
    const __nv_bfloat16* pkb = P ? a.pk + (size_t)b * P * HD + h * kAttnD : base;
    const __nv_bfloat16* pvb = P ? a.pv + (size_t)b * P * HD + h * kAttnD : base;
    LC_ASTAMP(0);
    stage_tile(sQ, PQ, base, ld, q0, 128, T, tid, 256);
    stage_keys(sK, PK, pkb, HD, P, base + HD, ld, T, NP, tid, 256);
    cp_async_commit();
    stage_keys(sV, PK, pvb, HD, P, base + 2 * HD, ld, T, NP, tid, 256);      // needed only by P V: lands while the softmax runs
    cp_async_commit();
    if (tid == 32) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
    if (warp == 0) tmem_alloc(slot, 256);
    cp_async_wait_group<1>();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *slot;
    LC_ASTAMP(1);
    if (tid == 0) {
        const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sQ, PQ, k), desc_kmajor(sK, PK, k), idesc, k != 0);
        mma_commit(bars);
    }
    bool ok = mbar_wait(bars, 0);
    fence_after_sync();
    LC_ASTAMP(2);
    // ---- softmax over this thread's half of the row held by its TMEM lane -----------------------------------------------------------
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int NPh = ((NP >> 1) + 15) & ~15;
    const int cbeg = chalf ? NPh : 0, cend = chalf ? NP : NPh;
    const float c = kLog2e * 0.125f;
    float m = -CUDART_INF_F;
    {
        int c0 = cbeg;
        for (; c0 + 32 <= cend; c0 += 32) m = attn_max_group<32>(trow, c0, NK, m);
        if (c0 < cend) m = attn_max_group<16>(trow, c0, NK, m);
    }
    s_red[tid] = m;
    __syncthreads();
    LC_ASTAMP(3);
    m = fmaxf(s_red[row], s_red[row + 128]);             // NK >= 1 key in the first half: finite
    const float mc = m * c;
    float sum = 0.f;
    // P overlays the Q / K tiles, which MMA 1 (completed: bars[0]) has finished reading; rows are normalised by what the tensor core will actually multiply
    {
        int c0 = cbeg;
        for (; c0 + 32 <= cend; c0 += 32) sum += attn_p_group<32>(trow, c0, NK, c, mc, smem, PQ, row);
        if (c0 < cend) sum += attn_p_group<16>(trow, c0, NK, c, mc, smem, PQ, row);
    }
    s_red[256 + tid] = sum;
    cp_async_wait_all();                                 // this thread's V chunks
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    LC_ASTAMP(4);
    if (tid == 0) {
        const uint32_t idesc = attn_idesc(128, kAttnD, 0, 1);
        for (int k = 0; k < NP / 16; ++k) mma_f16(tmem, desc_kmajor(sP, PQ, k), desc_mnmajor(sV, PK, k, 0), idesc, k != 0);
        mma_commit(bars + 1);
    }
    sum = s_red[256 + row] + s_red[256 + row + 128];
    ok = mbar_wait(bars + 1, 0) && ok;
    fence_after_sync();
    LC_ASTAMP(5);
    if (!ok && a.error_flag != nullptr && (tid & 31) == 0) atomicExch(a.error_flag, 3);
    const int t = q0 + row;
    const float inv = 1.f / sum;
    {
        // O rows go out through a per-warp stage in the (now dead) P tile region: 4 lanes write one 64-byte row segment
        unsigned char* stg = smem + (size_t)warp * 32 * 80;
        float v[32];
        tmem_ld32(trow + (uint32_t)(chalf * 32), v);        // warp-collective (.sync.aligned): rows past the end take part, only the store is predicated
#pragma unroll
        for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(stg + (tid & 31) * 80 + i * 2) = make_uint4(pack_bf16(v[i] * inv, v[i + 1] * inv), pack_bf16(v[i + 2] * inv, v[i + 3] * inv),
                                                                                 pack_bf16(v[i + 4] * inv, v[i + 5] * inv), pack_bf16(v[i + 6] * inv, v[i + 7] * inv));
        __syncwarp();
        const int lane = tid & 31, rr = lane >> 2, ch = lane & 3;
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
            const int r = p4 * 8 + rr;
            const int tq = q0 + (warp & 3) * 32 + r;
            if (tq < T) *reinterpret_cast<uint4*>(a.out + ((size_t)b * T + tq) * HD + h * kAttnD + chalf * 32 + ch * 8) = *reinterpret_cast<const uint4*>(stg + r * 80 + ch * 16);
        }
    }
    if (t < T && chalf == 0) a.lse2[((size_t)b * a.H + h) * T + t] = mc + log2f(sum);
    LC_ASTAMP(6);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

inline size_t attn_fwd_smem(int NK) {
    const int NP = attn_np(NK), PQ = attn_plane(128), PK = attn_plane(NP);
    const int regionA = (8 * PQ + 8 * PK) > (NP / 8) * PQ ? (8 * PQ + 8 * PK) : (NP / 8) * PQ;
    return (size_t)regionA + 8 * PK + 64 + 512 * sizeof(float);
}

struct AttnBwdArgs {
    const __nv_bfloat16* qkv;   // [B][T][3][H][64]
    const __nv_bfloat16* dout;  // [B][T][H*64]
    const float* lse2;          // [B][H][T]
    __nv_bfloat16* dqkv;        // [B][T][3][H][64]
    int T, H;
    int* error_flag;
    const __nv_bfloat16* pk;    // nullable [B][P][H*64] prefix keys / values (as in the forward)
    const __nv_bfloat16* pv;
    float* dpk;                 // [B][P][H*64] fp32: gradients of the prefix rows (they are parameters of the prompt pools)
    float* dpv;
    int P;
};

// grid (H, B), 256 threads, 1 CTA / SM (all 512 TMEM columns, ~202 KB shared memory).  Per 128-query tile:
//   S = Q K^T -> P = exp2(S c - lse2) -> dP = dO V^T -> D = sum_j P_j dP_j -> dS = P (dP - D) -> dQ = dS K / 8 ; dK += dS^T Q / 8 ; dV += P^T dO
// (dK, dV stay in TMEM across the query tiles).  TMEM columns: [0, NP) S then dP then (first 64) dQ ; [256, 384) dK (two 128-key halves x 64) ; [384, 512) dV.
// Two threads share a query row (= TMEM lane): thread (row, chalf) owns one half of the key columns.
__global__ void __launch_bounds__(256) attn_bwd_kernel(AttnBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int T = a.T, P = a.pk != nullptr ? a.P : 0, NK = T + P, NP = attn_np(NK);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, chalf = tid >> 7;
    const int h = blockIdx.x, b = blockIdx.y;
    constexpr int PQ = attn_plane(128);
    const int PK = attn_plane(NP);
    const int nplanes = NP / 8;
    // sP then sdS are adjacent on purpose: the second key half (M = 128 keys from 128) of the MN-major A operands reads up to 4 planes past the
    // end of its tile (keys >= NP): those reads must stay inside the allocation; the rows they feed are never stored.
    const uint32_t s_base = smem_u32(smem);
    const uint32_t sP = s_base, sdS = sP + nplanes * PQ, sK = sdS + nplanes * PQ, sV = sK + 8 * PK, sQ = sV + 8 * PK, sdO = sQ + 8 * PQ;
    unsigned char* g_sP = smem;
    unsigned char* g_sdS = smem + nplanes * PQ;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * nplanes * PQ + 16 * PK + 16 * PQ);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    float* s_dpart = reinterpret_cast<float*>(smem + 2 * nplanes * PQ + 16 * PK + 16 * PQ + 64);
    const long long ld = 3LL * a.H * kAttnD;
    const int HD = a.H * kAttnD;
    const __nv_bfloat16* base = a.qkv + (size_t)b * T * ld + h * kAttnD;
    const __nv_bfloat16* pkb = P ? a.pk + (size_t)b * P * HD + h * kAttnD : base;
    const __nv_bfloat16* pvb = P ? a.pv + (size_t)b * P * HD + h * kAttnD : base;
    LC_ASTAMP(0);
    stage_keys(sK, PK, pkb, HD, P, base + HD, ld, T, NP, tid, 256);
    if (tid == 32) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(slot, 512);
    uint32_t phase = 0;
    bool ok = true;
    const float c = kLog2e * 0.125f;
    const int NPh = ((NP >> 1) + 15) & ~15;
    const int cbeg = chalf ? NPh : 0, cend = chalf ? NP : NPh;
    const int ntile = (T + 127) / 128;
    uint32_t tmem = 0;
    for (int qt = 0; qt < ntile; ++qt) {
        const int q0 = qt * 128;
        // two copy groups: {K (first tile only), Q} feed S; {V (first tile only), dO} feed dP and land while S / P are being computed
        stage_tile(sQ, PQ, base, ld, q0, 128, T, tid, 256);
        cp_async_commit();
        if (qt == 0) stage_keys(sV, PK, pvb, HD, P, base + 2 * HD, ld, T, NP, tid, 256);
        stage_tile(sdO, PQ, a.dout + (size_t)b * T * HD + h * kAttnD, HD, q0, 128, T, tid, 256);
        cp_async_commit();
        const int t = q0 + row;
        const bool rv = t < T;
        const float l2 = rv ? a.lse2[((size_t)b * a.H + h) * T + t] : CUDART_INF_F;   // in flight with the tile copies; +inf: P = exp2(-inf) = 0 on padding rows
        cp_async_wait_group<1>();
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        tmem = *slot;
        LC_ASTAMP(1 + 7 * qt);
        if (tid == 0) {
            const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
            for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sQ, PQ, k), desc_kmajor(sK, PK, k), idesc, k != 0);
            mma_commit(bar);
        }
        ok = mbar_wait(bar, phase & 1) && ok; ++phase;
        fence_after_sync();
        LC_ASTAMP(2 + 7 * qt);
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // ---- P = exp2(S c - lse2) for this thread's half of the key columns ------------------------------------------------------------
        {
            int c0 = cbeg;
            for (; c0 + 32 <= cend; c0 += 32) attn_p_group<32>(trow, c0, NK, c, l2, g_sP, PQ, row);
            if (c0 < cend) attn_p_group<16>(trow, c0, NK, c, l2, g_sP, PQ, row);
        }
        cp_async_wait_all();                               // this thread's V / dO chunks
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();                                   // everyone is done reading S: its columns may now receive dP
        fence_after_sync();
        LC_ASTAMP(3 + 7 * qt);
        if (tid == 0) {
            const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
            for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sdO, PQ, k), desc_kmajor(sV, PK, k), idesc, k != 0);
            mma_commit(bar);
        }
        ok = mbar_wait(bar, phase & 1) && ok; ++phase;
        fence_after_sync();
        LC_ASTAMP(4 + 7 * qt);
        // ---- D = sum_j P_j dP_j in fp32 from the very probabilities the dS / dV products use.  (The usual shortcut D = dO . O inherits the BF16
        //      rounding of the stored O as a common-mode error of the whole row, which dS = P (dP - D) does not average out when the values of
        //      a head are nearly alike; the row sum over the keys has no such term.)  Partial sums of the two threads of a row meet in shared memory.
        float dpart = 0.f;
        {
            int c0 = cbeg;
            for (; c0 + 32 <= cend; c0 += 32) dpart = attn_d_group<32>(trow, c0, g_sP, PQ, row, dpart);
            if (c0 < cend) dpart = attn_d_group<16>(trow, c0, g_sP, PQ, row, dpart);
        }
        s_dpart[tid] = dpart;
        __syncthreads();
        const float Dr = s_dpart[row] + s_dpart[row + 128];
        // ---- dS = P (dP - D) ---------------------------------------------------------------------------------------------------------
        {
            int c0 = cbeg;
            for (; c0 + 32 <= cend; c0 += 32) attn_ds_group<32>(trow, c0, g_sP, g_sdS, PQ, row, Dr);
            if (c0 < cend) attn_ds_group<16>(trow, c0, g_sP, g_sdS, PQ, row, Dr);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();                                   // P and dS tiles complete; dP consumed
        fence_after_sync();
        LC_ASTAMP(5 + 7 * qt);
        if (tid == 0) {
            const uint32_t id_q = attn_idesc(128, kAttnD, 0, 1), id_kv = attn_idesc(128, kAttnD, 1, 1);
            for (int k = 0; k < NP / 16; ++k) mma_f16(tmem, desc_kmajor(sdS, PQ, k), desc_mnmajor(sK, PK, k, 0), id_q, k != 0);              // dQ
            for (int half = 0; half < 2; ++half) {
                if (half * 128 >= NP) break;
                for (int k = 0; k < 128 / 16; ++k) {
                    const uint32_t acc = (qt != 0 || k != 0) ? 1u : 0u;
                    mma_f16(tmem + 256 + half * 64, desc_mnmajor(sdS, PQ, k, half * 128), desc_mnmajor(sQ, PQ, k, 0), id_kv, acc);          // dK += dS^T Q
                    mma_f16(tmem + 384 + half * 64, desc_mnmajor(sP, PQ, k, half * 128), desc_mnmajor(sdO, PQ, k, 0), id_kv, acc);          // dV += P^T dO
                }
            }
            mma_commit(bar);
        }
        ok = mbar_wait(bar, phase & 1) && ok; ++phase;
        fence_after_sync();
        LC_ASTAMP(6 + 7 * qt);
        // ---- dQ tile out: thread (row, chalf) stores 32 of the 64 columns ----------------------------------------------------------------
        {
            // staged through the (now dead) P tile region: 4 lanes write one 64-byte row segment
            unsigned char* stg = smem + (size_t)warp * 32 * 80;
            float v[32];
            tmem_ld32(trow + (uint32_t)(chalf * 32), v);
#pragma unroll
            for (int i = 0; i < 32; i += 8)
                *reinterpret_cast<uint4*>(stg + (tid & 31) * 80 + i * 2) = make_uint4(pack_bf16(v[i] * 0.125f, v[i + 1] * 0.125f), pack_bf16(v[i + 2] * 0.125f, v[i + 3] * 0.125f),
                                                                                     pack_bf16(v[i + 4] * 0.125f, v[i + 5] * 0.125f), pack_bf16(v[i + 6] * 0.125f, v[i + 7] * 0.125f));
            __syncwarp();
            const int lane = tid & 31, rr = lane >> 2, ch = lane & 3;
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) {
                const int r = p4 * 8 + rr;
                const int tq = q0 + (warp & 3) * 32 + r;
                if (tq < T)
                    *reinterpret_cast<uint4*>(a.dqkv + ((size_t)b * T + tq) * ld + h * kAttnD + chalf * 32 + ch * 8) = *reinterpret_cast<const uint4*>(stg + r * 80 + ch * 16);
            }
        }
        fence_before_sync();
        __syncthreads();                                   // dQ columns and the Q / dO tiles are free for the next query tile
        fence_after_sync();
        LC_ASTAMP(7 + 7 * qt);
    }
    // ---- dK, dV out: TMEM lane = key (half * 128 + row); thread (row, chalf): chalf 0 -> dK, chalf 1 -> dV.  Keys [0, P) are the prefix rows (fp32,
    //      their own matrices), keys [P, P + T) the tokens ----------------------------------------------------------------------------------------
    {
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float sc = chalf == 0 ? 0.125f : 1.f;
        for (int half = 0; half < 2; ++half) {
            if (half * 128 >= NP) break;
            const int key = half * 128 + row;
            const bool pre = key < P;
            float* op = pre ? (chalf == 0 ? a.dpk : a.dpv) + ((size_t)b * P + key) * HD + h * kAttnD : nullptr;
            // token rows go out through a per-warp shared-memory stage so that 8 lanes write one 128-byte row segment (thread-per-row 16-byte stores
            // touch every 32-byte sector twice); the few prefix rows are written directly in fp32
            unsigned char* stg = smem + (size_t)warp * 32 * 144;
#pragma unroll
            for (int c0 = 0; c0 < kAttnD; c0 += 32) {
                float v[32];
                tmem_ld32(trow + (uint32_t)(256 + chalf * 128 + half * 64 + c0), v);
                if (pre) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + c0 + i) = make_float4(v[i] * sc, v[i + 1] * sc, v[i + 2] * sc, v[i + 3] * sc);
                }
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    *reinterpret_cast<uint4*>(stg + (tid & 31) * 144 + (c0 + i) * 2) = make_uint4(pack_bf16(v[i] * sc, v[i + 1] * sc), pack_bf16(v[i + 2] * sc, v[i + 3] * sc),
                                                                                                 pack_bf16(v[i + 4] * sc, v[i + 5] * sc), pack_bf16(v[i + 6] * sc, v[i + 7] * sc));
            }
            __syncwarp();
            {
                const int lane = tid & 31, rr = lane >> 3, ch = lane & 7;
#pragma unroll
                for (int p8 = 0; p8 < 8; ++p8) {
                    const int r = p8 * 4 + rr;
                    const int key_r = half * 128 + (warp & 3) * 32 + r;
                    if (key_r >= P && key_r < NK)
                        *reinterpret_cast<uint4*>(a.dqkv + ((size_t)b * T + (key_r - P)) * ld + (1 + chalf) * HD + h * kAttnD + ch * 8) =
                            *reinterpret_cast<const uint4*>(stg + r * 144 + ch * 16);
                }
            }
            __syncwarp();
        }
    }
    LC_ASTAMP(15);
    if (!ok && a.error_flag != nullptr && (tid & 31) == 0) atomicExch(a.error_flag, 4);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

inline size_t attn_bwd_smem(int NK) {
    const int NP = attn_np(NK), PQ = attn_plane(128), PK = attn_plane(NP);
    return (size_t)2 * (NP / 8) * PQ + 16 * PK + 16 * PQ + 64 + 256 * sizeof(float);
}

}  // namespace tc
}  // namespace lc
