// 3x3 / pad 1 / stride 1 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators in TMEM), NHWC fp32 activations in HBM.  Forward and (with the flipped/transposed packing) data-gradient of
// the CifarResNet 16->16 @32x32, 32->32 @16x16 and 64->64 @8x8 layers (core/model/backbone/resnet.py:295,298).
//
// Formulation ("flattened padded rows"): give every image a one-pixel zero border and flatten (n, hp, wp) into one row index
// Q = (n*(H+2) + hp)*(W+2) + wp.  For a filter tap (dr, dc) the im2col row of output position Q is simply input row
// Q + dr*(W+2) + dc: a constant row SHIFT.  A CTA owns 128 consecutive Q (the MMA M dimension), stages input rows
// [Q0-(W+3), Q0+128+(W+3)) once in shared memory in the canonical K-major / no-swizzle UMMA layout (16-byte channel chunks in
// separate planes, row r at plane + 16*r) and issues, per tap and per 8 input channels, one
//     tcgen05.mma.cta_group::1.kind::tf32  D[128 x COUT] += A[128 x 8] * B[COUT x 8]^T
// whose A descriptor start address is the plane base advanced by the tap's row shift (no im2col copy, no re-load).  Rows that
// land on border positions compute garbage that the epilogue discards ((W+2)^2/W^2 - 1 = 13% / 27% / 56% extra MMA rows at
// W = 32 / 16 / 8; the kernel is HBM/latency-bound, not MMA-bound).
//
// Prologue (optional): relu(x*scale[c]+shift[c]) of the producer BatchNorm applied while staging (borders stay zero).
// Epilogue: TMEM -> registers (tcgen05.ld 32x32b), optional addend, NHWC store of the valid rows, optional BatchNorm
// statistics (per-CTA partial sums, last CTA finalises — same deterministic scheme as the CUDA-core kernels).
#pragma once
#include "conv_simt.cuh"

namespace lc {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: returns false if the phase never completed (a mis-programmed MMA must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// round-to-nearest fp32 -> tf32 (the tensor core itself truncates the low 13 mantissa bits; rounding first removes the bias)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4   [16,30) LBO>>4 (16-byte chunk to the next chunk along K)   [32,46) SBO>>4 (8-row group to the next)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp receives row (lane base + i), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct ConvTcArgs {
    const float* in;         // NHWC [B][W][W][C]
    const float* wtc;        // packed [9][C/4][COUT][4]  (tap, 16-byte K chunk, output channel, 4 input channels)
    float* out;              // NHWC [B][W][W][COUT]
    const float* pro_scale;  // nullable
    const float* pro_shift;
    const float* addend;     // nullable (may alias out)
    BnStatArgs stat;         // stat.partial nullable: [ntiles][2][COUT]
    int* error_flag;         // set to 1 if the MMA completion barrier timed out
    int B;
};

template <int C, int W>
struct ConvTcCfg {
    static constexpr int N = C;                    // square layers: COUT == CIN == C
    static constexpr int WP = W + 2;
    static constexpr int PP = (W + 2) * WP;        // padded positions per image
    static constexpr int HALO = WP + 1;
    static constexpr int ROWS = 128 + 2 * HALO;
    static constexpr int CH = C / 4;               // 16-byte chunks per row
    static constexpr int PLANE = ROWS * 16;        // bytes
    static constexpr int A_BYTES = CH * PLANE;
    static constexpr int BTAP = CH * N * 16;       // bytes per tap
    static constexpr int B_BYTES = 9 * BTAP;
    static constexpr int T_FLOATS = 128 * (N + 1) + 256;               // epilogue transpose + column partials (aliases A/B)
    static constexpr int MAIN_BYTES = A_BYTES + B_BYTES > T_FLOATS * 4 ? A_BYTES + B_BYTES : T_FLOATS * 4;
    static constexpr int RED_FLOATS = 2 * 128 > 4 * N ? 2 * 128 : 4 * N;
    static constexpr size_t SMEM_BYTES = MAIN_BYTES + RED_FLOATS * 4 + 64;
    static_assert((ROWS * CH + 127) / 128 <= 32, "valid-mask width");
    static constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : 64;
    static_assert(C % 8 == 0 && (C == 16 || C == 32 || C == 64), "tensor-core conv: C in {16,32,64}");
};

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills the destination (border rows)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// one bulk (TMA, 1-D) copy global->shared, completion reported on an mbarrier as transaction bytes
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

template <int C, int W>
__global__ void __launch_bounds__(128) conv3x3_tc_kernel(ConvTcArgs a) {
    using K = ConvTcCfg<C, W>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + K::A_BYTES;
    float* s_red = reinterpret_cast<float*>(smem_raw + K::MAIN_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::MAIN_BYTES + K::RED_FLOATS * 4);     // [0] MMA done, [1] weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int total = a.B * K::PP;
    const int q0 = (int)blockIdx.x * 128;

    if (tid == 32) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        // weights are already in UMMA order in global memory: one bulk copy, tracked by bar[1]
        bulk_load(smem_u32(sB), a.wtc, (uint32_t)K::B_BYTES, bar + 1);
    }
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);

    // ---- stage A: rows Q in [q0-HALO, q0+128+HALO), 16-byte chunks, channel-chunk-planar.  All copies are issued before any
    //      is waited for (cp.async), then each thread transforms the chunks it copied itself (BN+ReLU prologue, TF32 rounding).
    constexpr int NE = (K::ROWS * K::CH + 127) / 128;
    uint32_t validmask = 0;
    const uint32_t sA_u = smem_u32(sA);
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        const int e = tid + i * 128;
        if (e < K::ROWS * K::CH) {
            const int j = e % K::CH, r = e / K::CH;
            const int Q = q0 - K::HALO + r;
            const float* src = a.in;
            uint32_t nbytes = 0;
            if (Q >= 0 && Q < total) {
                const int n = Q / K::PP, rem = Q % K::PP;
                const int hp = rem / K::WP, wp = rem % K::WP;
                if (hp >= 1 && hp <= W && wp >= 1 && wp <= W) {
                    src = a.in + (((size_t)n * W + (hp - 1)) * W + (wp - 1)) * C + j * 4;
                    nbytes = 16;
                    validmask |= 1u << i;
                }
            }
            cp_async16(sA_u + (uint32_t)j * K::PLANE + (uint32_t)r * 16, src, nbytes);
        }
    }
    cp_async_wait_all();
    {
        const bool pro = a.pro_scale != nullptr;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            if (validmask & (1u << i)) {
                const int e = tid + i * 128;
                const int j = e % K::CH, r = e / K::CH;
                float4* p4 = reinterpret_cast<float4*>(sA + (size_t)j * K::PLANE + (size_t)r * 16);
                float4 v = *p4;
                if (pro) {
                    const float4 sc = ldg4(a.pro_scale + j * 4), sh = ldg4(a.pro_shift + j * 4);
                    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
                    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
                    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
                    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
                }
                v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                *p4 = v;
            }
        }
    }
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // ---- MMA issue: one thread, 9 taps x C/8 K-steps ---------------------------------------------------------------------
    if (tid == 0) {
        mbar_wait(bar + 1, 0);     // weights landed (async proxy write, ordered by the mbarrier)
        constexpr uint32_t idesc = make_idesc_tf32(K::N);
        const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB);
        uint32_t acc = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dr = tap / 3 - 1, dc = tap % 3 - 1;
            const uint32_t shift = (uint32_t)(K::HALO + dr * K::WP + dc);
#pragma unroll
            for (int kc = 0; kc < C / 8; ++kc) {
                const uint64_t ad = make_desc(aBase + (uint32_t)(2 * kc) * K::PLANE + shift * 16, K::PLANE, 128);
                const uint64_t bd = make_desc(bBase + (uint32_t)tap * K::BTAP + (uint32_t)(2 * kc) * (K::N * 16), K::N * 16, 128);
                mma_tf32(tmem_base, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        mma_commit(bar);           // implies tcgen05.fence::before_thread_sync
    }
    const bool done = mbar_wait(bar, 0);
    fence_after_sync();
    if (!done && tid == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);

    // ---- epilogue ------------------------------------------------------------------------------------------------------------
    const int m = tid;                                   // accumulator row == TMEM lane
    const int Q = q0 + m;
    bool valid = false;
    size_t obase = 0;
    if (Q < total) {
        const int n = Q / K::PP, rem = Q % K::PP;
        const int hp = rem / K::WP, wp = rem % K::WP;
        valid = hp >= 1 && hp <= W && wp >= 1 && wp <= W;
        obase = (((size_t)n * W + (hp - 1)) * W + (wp - 1)) * K::N;
    }
    float* sT = reinterpret_cast<float*>(smem_raw);     // [128][N+1] transpose buffer (A/B tiles are dead now)
    const bool stats = a.stat.partial != nullptr;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < K::N; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        if (valid) {
            if (a.addend != nullptr) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 t = *reinterpret_cast<const float4*>(a.addend + obase + c0 + k4 * 4);
                    v[k4 * 4] += t.x; v[k4 * 4 + 1] += t.y; v[k4 * 4 + 2] += t.z; v[k4 * 4 + 3] += t.w;
                }
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<float4*>(a.out + obase + c0 + k4 * 4) = make_float4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
        }
        if (stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sT[m * (K::N + 1) + c0 + i] = valid ? v[i] : 0.f;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);

    if (stats) {
        // column sums in fixed order: thread t -> channel t % N, row slice t / N
        constexpr int SL = 128 / K::N;                  // slices per channel (8 / 4 / 2)
        constexpr int RPS = 128 / SL;                   // rows per slice
        float* sP = sT + 128 * (K::N + 1);              // [2][SL][N]
        const int ch = tid % K::N, sl = tid / K::N;
        float sm = 0.f, sq = 0.f;
#pragma unroll 4
        for (int r = sl * RPS; r < (sl + 1) * RPS; ++r) {
            const float x = sT[r * (K::N + 1) + ch];
            sm += x; sq = fmaf(x, x, sq);
        }
        sP[sl * K::N + ch] = sm;
        sP[(SL + sl) * K::N + ch] = sq;
        __syncthreads();
        if (tid < 2 * K::N) {
            const int stat = tid / K::N, c = tid % K::N;
            float t = 0.f;
#pragma unroll
            for (int s = 0; s < SL; ++s) t += sP[(stat * SL + s) * K::N + c];
            a.stat.partial[((size_t)blockIdx.x * 2 + stat) * K::N + c] = t;
        }
        if (last_block_done(a.stat.counter, gridDim.x)) {
            bn_finalize_last_block<K::N, 128>(a.stat, (int)gridDim.x, (double)a.B * W * W, s_red);
        }
    }
}

template <int C, int W>
static inline int conv_tc_launch(const ConvTcArgs& a, cudaStream_t st) {
    using K = ConvTcCfg<C, W>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(conv3x3_tc_kernel<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    const long long total = (long long)a.B * K::PP;
    const int grid = (int)((total + 127) / 128);
    conv3x3_tc_kernel<C, W><<<grid, 128, K::SMEM_BYTES, st>>>(a);
    return lc_launch_status();
}

static inline long long conv_tc_tiles(int B, int W) { return ((long long)B * (W + 2) * (W + 2) + 127) / 128; }

}  // namespace tc
}  // namespace lc
