// Fused multi-head self-attention for short sequences (T <= 256 tokens: ViT-B/16 with or without L2P prompts: 197 / 222) on tcgen05:
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

struct AttnFwdArgs {
    const __nv_bfloat16* qkv;   // [B][T][3][H][64]
    __nv_bfloat16* out;         // [B][T][H*64]
    float* lse2;                // [B][H][T]  row log-sum-exp of the scaled scores, base 2:  max*c + log2(sum exp2(s*c - max*c)), c = log2(e)/8
    int T, H;
    int* error_flag;
};

// grid (ceil(T/128), H, B), 128 threads, 2 CTAs / SM (256 TMEM columns, ~85 KB shared memory each)
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int T = a.T, NP = attn_np(T);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    constexpr int PQ = attn_plane(128);                 // Q tile and P tile planes (128 rows)
    const int PK = attn_plane(NP);                       // K / V planes (NP rows)
    const uint32_t s_base = smem_u32(smem);
    const int regionA = max(8 * PQ + 8 * PK, (NP / 8) * PQ);
    const uint32_t sQ = s_base, sK = s_base + 8 * PQ, sP = s_base, sV = s_base + (uint32_t)regionA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + regionA + 8 * PK);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    const long long ld = 3LL * a.H * kAttnD;
    const __nv_bfloat16* base = a.qkv + (size_t)b * T * ld + h * kAttnD;
    stage_tile(sQ, PQ, base, ld, q0, 128, T, tid, 128);
    stage_tile(sK, PK, base + a.H * kAttnD, ld, 0, NP, T, tid, 128);
    stage_tile(sV, PK, base + 2 * a.H * kAttnD, ld, 0, NP, T, tid, 128);
    if (tid == 32) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
    if (warp == 0) tmem_alloc(slot, 256);
    cp_async_wait_all();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sQ, PQ, k), desc_kmajor(sK, PK, k), idesc, k != 0);
        mma_commit(bars);
    }
    bool ok = mbar_wait(bars, 0);
    fence_after_sync();
    // ---- softmax over the row held by this thread's TMEM lane ------------------------------------------------------------------
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const float c = kLog2e * 0.125f;
    float m = -CUDART_INF_F;
    for (int c0 = 0; c0 < NP; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) if (c0 + i < T) m = fmaxf(m, v[i]);
    }
    const float mc = m * c;
    float sum = 0.f;
    // P overlays the Q / K tiles, which MMA 1 (completed: bars[0]) has finished reading
    for (int c0 = 0; c0 < NP; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p0 = c0 + 2 * i < T ? exp2f(fmaf(v[2 * i], c, -mc)) : 0.f;
            const float p1 = c0 + 2 * i + 1 < T ? exp2f(fmaf(v[2 * i + 1], c, -mc)) : 0.f;
            const __nv_bfloat162 pr = __floats2bfloat162_rn(p0, p1);
            sum += __low2float(pr) + __high2float(pr);           // normalise by what the tensor core will actually multiply
            pk[i] = *reinterpret_cast<const uint32_t*>(&pr);
        }
        unsigned char* dst = smem + (size_t)((c0 >> 3) * PQ + tid * 16);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + PQ) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = attn_idesc(128, kAttnD, 0, 1);
        for (int k = 0; k < NP / 16; ++k) mma_f16(tmem, desc_kmajor(sP, PQ, k), desc_mnmajor(sV, PK, k, 0), idesc, k != 0);
        mma_commit(bars + 1);
    }
    ok = mbar_wait(bars + 1, 0) && ok;
    fence_after_sync();
    if (!ok && a.error_flag != nullptr && (tid & 31) == 0) atomicExch(a.error_flag, 3);
    const int t = q0 + tid;
    const float inv = 1.f / sum;
    __nv_bfloat16* o = a.out + ((size_t)b * T + (t < T ? t : 0)) * (a.H * kAttnD) + h * kAttnD;
#pragma unroll
    for (int c0 = 0; c0 < kAttnD; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);                  // warp-collective (.sync.aligned): rows past the end take part, only the store is predicated
        if (t < T) {
            *reinterpret_cast<uint4*>(o + c0) = make_uint4(pack_bf16(v[0] * inv, v[1] * inv), pack_bf16(v[2] * inv, v[3] * inv),
                                                           pack_bf16(v[4] * inv, v[5] * inv), pack_bf16(v[6] * inv, v[7] * inv));
            *reinterpret_cast<uint4*>(o + c0 + 8) = make_uint4(pack_bf16(v[8] * inv, v[9] * inv), pack_bf16(v[10] * inv, v[11] * inv),
                                                               pack_bf16(v[12] * inv, v[13] * inv), pack_bf16(v[14] * inv, v[15] * inv));
        }
    }
    if (t < T) a.lse2[((size_t)b * a.H + h) * T + t] = mc + log2f(sum);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

inline size_t attn_fwd_smem(int T) {
    const int NP = attn_np(T), PQ = attn_plane(128), PK = attn_plane(NP);
    const int regionA = (8 * PQ + 8 * PK) > (NP / 8) * PQ ? (8 * PQ + 8 * PK) : (NP / 8) * PQ;
    return (size_t)regionA + 8 * PK + 64;
}

// D[b][h][t] = sum_d dO[b][t][h*64+d] * O[b][t][h*64+d]   (= sum_j P_j dP_j, the softmax-backward row term); one warp per token row
__global__ void __launch_bounds__(128) attn_rowdot_kernel(const __nv_bfloat16* dO, const __nv_bfloat16* O, float* D, long long rows, int T, int H) {
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int HD = H * kAttnD;
    const long long b = row / T;
    const int t = (int)(row % T);
    for (int h0 = 0; h0 < H; h0 += 4) {                 // a warp covers 4 heads per pass: lane -> (head h0 + lane/8, 8 elements)
        const int h = h0 + (lane >> 3);
        float s = 0.f;
        if (h < H) {
            const uint4 x = *reinterpret_cast<const uint4*>(dO + (size_t)row * HD + h * kAttnD + (lane & 7) * 8);
            const uint4 y = *reinterpret_cast<const uint4*>(O + (size_t)row * HD + h * kAttnD + (lane & 7) * 8);
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s = fmaf(__uint_as_float(xs[i] << 16), __uint_as_float(ys[i] << 16), s);
                s = fmaf(__uint_as_float(xs[i] & 0xffff0000u), __uint_as_float(ys[i] & 0xffff0000u), s);
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (h < H && (lane & 7) == 0) D[((size_t)b * H + h) * T + t] = s;
    }
}

struct AttnBwdArgs {
    const __nv_bfloat16* qkv;   // [B][T][3][H][64]
    const __nv_bfloat16* dout;  // [B][T][H*64]
    const float* lse2;          // [B][H][T]
    const float* D;             // unused (kept for ABI stability): the row term sum_j P_j dP_j is formed on chip
    __nv_bfloat16* dqkv;        // [B][T][3][H][64]
    int T, H;
    int* error_flag;
};

// grid (H, B), 256 threads, 1 CTA / SM (all 512 TMEM columns, ~202 KB shared memory).  Per 128-query tile:
//   S = Q K^T -> P = exp2(S c - lse2) -> dP = dO V^T -> dS = P (dP - D) -> dQ = dS K / 8 ; dK += dS^T Q / 8 ; dV += P^T dO   (dK, dV stay in TMEM)
// TMEM columns: [0, NP) S then dP then (first 64) dQ ; [256, 384) dK (two 128-key halves x 64) ; [384, 512) dV.
__global__ void __launch_bounds__(256) attn_bwd_kernel(AttnBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int T = a.T, NP = attn_np(T);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, chalf = tid >> 7;          // TMEM lane / query row of the tile, and which half of the key columns this thread handles
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
    stage_tile(sK, PK, base + HD, ld, 0, NP, T, tid, 256);
    stage_tile(sV, PK, base + 2 * HD, ld, 0, NP, T, tid, 256);
    if (tid == 32) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(slot, 512);
    uint32_t phase = 0;
    bool ok = true;
    const float c = kLog2e * 0.125f;
    const int cbeg = chalf * (NP / 2), cend = cbeg + NP / 2;      // NP / 2 is a multiple of 8
    const int ntile = (T + 127) / 128;
    uint32_t tmem = 0;
    for (int qt = 0; qt < ntile; ++qt) {
        const int q0 = qt * 128;
        stage_tile(sQ, PQ, base, ld, q0, 128, T, tid, 256);
        stage_tile(sdO, PQ, a.dout + (size_t)b * T * HD + h * kAttnD, HD, q0, 128, T, tid, 256);
        cp_async_wait_all();
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        tmem = *slot;
        if (tid == 0) {
            const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
            for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sQ, PQ, k), desc_kmajor(sK, PK, k), idesc, k != 0);
            mma_commit(bar);
        }
        ok = mbar_wait(bar, phase & 1) && ok; ++phase;
        fence_after_sync();
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int t = q0 + row;
        const bool rv = t < T;
        const float l2 = rv ? a.lse2[((size_t)b * a.H + h) * T + t] : 0.f;
        // ---- P = exp2(S c - lse2) for this thread's half of the key columns ------------------------------------------------------------
        for (int g0 = cbeg & ~15; g0 < cend; g0 += 16) {     // tcgen05.ld x16 granularity: aligned 16-column groups, 8-column sub-groups inside [cbeg, cend)
            float v[16];
            tmem_ld16(trow + (uint32_t)g0, v);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                const int c0 = g0 + sub * 8;
                if (c0 < cbeg || c0 >= cend) continue;
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float p0 = (rv && c0 + 2 * i < T) ? exp2f(fmaf(v[sub * 8 + 2 * i], c, -l2)) : 0.f;
                    const float p1 = (rv && c0 + 2 * i + 1 < T) ? exp2f(fmaf(v[sub * 8 + 2 * i + 1], c, -l2)) : 0.f;
                    pk[i] = pack_bf16(p0, p1);
                }
                *reinterpret_cast<uint4*>(g_sP + (size_t)((c0 >> 3) * PQ + row * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        fence_before_sync();
        __syncthreads();                                   // everyone is done reading S: its columns may now receive dP
        fence_after_sync();
        if (tid == 0) {
            const uint32_t idesc = attn_idesc(128, NP, 0, 0);
#pragma unroll
            for (int k = 0; k < kAttnD / 16; ++k) mma_f16(tmem, desc_kmajor(sdO, PQ, k), desc_kmajor(sV, PK, k), idesc, k != 0);
            mma_commit(bar);
        }
        ok = mbar_wait(bar, phase & 1) && ok; ++phase;
        fence_after_sync();
        // ---- D = sum_j P_j dP_j in fp32 from the very probabilities the dS / dV products use.  (The usual shortcut D = dO . O inherits the BF16
        //      rounding of the stored O as a common-mode error of the whole row, which dS = P (dP - D) does not average out when the values of
        //      a head are nearly alike; the row sum over the keys has no such term.)  Two threads share a row: partial sums meet in shared memory.
        float Dr;
        {
            float dpart = 0.f;
            for (int g0 = cbeg & ~15; g0 < cend; g0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)g0, v);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int c0 = g0 + sub * 8;
                    if (c0 < cbeg || c0 >= cend) continue;
                    const uint4 pv = *reinterpret_cast<const uint4*>(g_sP + (size_t)((c0 >> 3) * PQ + row * 16));
                    const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dpart = fmaf(__uint_as_float(pw[i] << 16), v[sub * 8 + 2 * i], dpart);
                        dpart = fmaf(__uint_as_float(pw[i] & 0xffff0000u), v[sub * 8 + 2 * i + 1], dpart);
                    }
                }
            }
            s_dpart[tid] = dpart;
            __syncthreads();
            Dr = s_dpart[row] + s_dpart[row + 128];
        }
        // ---- dS = P (dP - D) ---------------------------------------------------------------------------------------------------------
        for (int g0 = cbeg & ~15; g0 < cend; g0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)g0, v);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                const int c0 = g0 + sub * 8;
                if (c0 < cbeg || c0 >= cend) continue;
                const uint4 pv = *reinterpret_cast<const uint4*>(g_sP + (size_t)((c0 >> 3) * PQ + row * 16));
                const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float p0 = __uint_as_float(pw[i] << 16), p1 = __uint_as_float(pw[i] & 0xffff0000u);
                    pk[i] = pack_bf16(p0 * (v[sub * 8 + 2 * i] - Dr), p1 * (v[sub * 8 + 2 * i + 1] - Dr));      // P is exactly 0 on padding rows / columns
                }
                *reinterpret_cast<uint4*>(g_sdS + (size_t)((c0 >> 3) * PQ + row * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();                                   // P and dS tiles complete; dP consumed
        fence_after_sync();
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
        // ---- dQ tile out: thread (row, chalf) stores 32 of the 64 columns ----------------------------------------------------------------
        {
            __nv_bfloat16* o = a.dqkv + ((size_t)b * T + (rv ? t : 0)) * ld + h * kAttnD + chalf * 32;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)(chalf * 32 + c0), v);
                if (rv) {
                    *reinterpret_cast<uint4*>(o + c0) = make_uint4(pack_bf16(v[0] * 0.125f, v[1] * 0.125f), pack_bf16(v[2] * 0.125f, v[3] * 0.125f),
                                                                   pack_bf16(v[4] * 0.125f, v[5] * 0.125f), pack_bf16(v[6] * 0.125f, v[7] * 0.125f));
                    *reinterpret_cast<uint4*>(o + c0 + 8) = make_uint4(pack_bf16(v[8] * 0.125f, v[9] * 0.125f), pack_bf16(v[10] * 0.125f, v[11] * 0.125f),
                                                                       pack_bf16(v[12] * 0.125f, v[13] * 0.125f), pack_bf16(v[14] * 0.125f, v[15] * 0.125f));
                }
            }
        }
        fence_before_sync();
        __syncthreads();                                   // dQ columns and the Q / dO tiles are free for the next query tile
        fence_after_sync();
    }
    // ---- dK, dV out: TMEM lane = key (half * 128 + row); thread (row, chalf): chalf 0 -> dK, chalf 1 -> dV ---------------------------------
    {
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float sc = chalf == 0 ? 0.125f : 1.f;
        for (int half = 0; half < 2; ++half) {
            if (half * 128 >= NP) break;
            const int key = half * 128 + row;
            const bool kv = key < T;
            __nv_bfloat16* o = a.dqkv + ((size_t)b * T + (kv ? key : 0)) * ld + (1 + chalf) * HD + h * kAttnD;
#pragma unroll
            for (int c0 = 0; c0 < kAttnD; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)(256 + chalf * 128 + half * 64 + c0), v);
                if (kv) {
                    *reinterpret_cast<uint4*>(o + c0) = make_uint4(pack_bf16(v[0] * sc, v[1] * sc), pack_bf16(v[2] * sc, v[3] * sc), pack_bf16(v[4] * sc, v[5] * sc),
                                                                   pack_bf16(v[6] * sc, v[7] * sc));
                    *reinterpret_cast<uint4*>(o + c0 + 8) = make_uint4(pack_bf16(v[8] * sc, v[9] * sc), pack_bf16(v[10] * sc, v[11] * sc),
                                                                       pack_bf16(v[12] * sc, v[13] * sc), pack_bf16(v[14] * sc, v[15] * sc));
                }
            }
        }
    }
    if (!ok && a.error_flag != nullptr && (tid & 31) == 0) atomicExch(a.error_flag, 4);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

inline size_t attn_bwd_smem(int T) {
    const int NP = attn_np(T), PQ = attn_plane(128), PK = attn_plane(NP);
    return (size_t)2 * (NP / 8) * PQ + 16 * PK + 16 * PQ + 64 + 256 * sizeof(float);
}

}  // namespace tc
}  // namespace lc
