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

#ifndef LC_BWD_MAXNREG
#define LC_BWD_MAXNREG 72      // two data-gradient CTAs + one weight-gradient CTA per SM fit the register file
#endif

namespace lc {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// bounded wait: returns false if the phase never completed (a mis-programmed MMA must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
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

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills the destination (border rows)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// one bulk (TMA, 1-D) copy global->shared, completion reported on an mbarrier as transaction bytes
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

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

// 32 consecutive columns in one instruction (one wait instead of two)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

#ifdef LC_TC_TIMING
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LC_TSTAMP(slot) do { if (a.timing != nullptr && threadIdx.x == 0) a.timing[(size_t)blockIdx.x * 8 + (slot)] = gtimer(); } while (0)
#else
#define LC_TSTAMP(slot) do { } while (0)
#endif

// Fused BatchNorm backward (kernel variant BWD = 1).  The data-gradient conv of layer L produces g = d(loss)/d(ReLU output of the
// BatchNorm in front of L); its epilogue applies that ReLU's mask, stores the masked g and accumulates the two per-channel sums
// of the BatchNorm backward (sum g, sum g*xhat) -> per-CTA partials -> last CTA: the coefficients dy = c0*g + c1*y + c2, dgamma, dbeta.
// The consumer kernels (the data-gradient conv and the weight-gradient kernel of the layer below) evaluate dy while staging, so
// neither native_batch_norm_backward's reduction nor its elementwise pass exists as a launch (resnet.py:306-316 + autograd).
struct BnBwdFuse {
    const float* y;          // raw conv output feeding the BatchNorm being differentiated (same shape as `out`)
    const float* mask_out;   // nullable: materialised ReLU output (mask = mask_out > 0); else mask = scale*y+shift > 0 when `scale` is set
    const float* scale;      // [C] forward affine of that BatchNorm
    const float* shift;
    const float* mean;
    const float* invstd;
    float* partial;          // [grid][2][C]  (null -> no reduction)
    int defer;               // 1: only write the partial row (consumers use BnBwdLazy): no election, no last-CTA tail
    unsigned int* counter;
    float* coef;             // out [3][C]
    float* dgamma;           // out [C]
    float* dbeta;            // out [C]
};

struct ConvTcArgs {
    const float* in;         // NHWC [B][W][W][C]
    const float* wtc;        // packed [9][C/4][COUT][4]  (tap, 16-byte K chunk, output channel, 4 input channels)
    float* out;              // NHWC [B][W][W][COUT]
    const float* pro_scale;  // nullable
    const float* pro_shift;
    BnLazy pro_lazy;         // forward variant: pro_lazy.partial != null -> the prologue's scale / shift are reduced here from the producer's partial rows
    const float* addend;     // nullable (may alias out)
    const float* pro_res;    // MODE 2: residual operand of the prologue (the previous block's input), same shape as `in`
    float* pro_out;          // MODE 2: the previous block's output, written for the rows each CTA owns
    const float* pro_y;      // BWD variant, nullable: staged value = coef0[c]*in + coef1[c]*pro_y + coef2[c] (BatchNorm backward apply)
    const float* pro_coef;   // [3][C]
    BnBwdLazy pro_blazy;     // BWD variant: .partial != null -> the coefficients are reduced here from the producing epilogue's partial rows
    BnBwdFuse bw;            // BWD variant
    BnStatArgs stat;         // stat.partial nullable: [ntiles][2][COUT]
    int* error_flag;         // set to 1 if the MMA completion barrier timed out
    int B;
    unsigned long long* timing;   // debug builds (-DLC_TC_TIMING): [grid][8] globaltimer stamps per phase
};

template <int C, int W>
struct ConvTcCfg {
    static constexpr int N = C;                    // square layers: COUT == CIN == C
    static constexpr int T = C == 16 ? 4 : (C == 32 ? 2 : 1);   // 128-row MMA tiles per CTA (halo + weight loads amortised)
    static constexpr int NT = 256;
    static constexpr int WP = W + 2;
    static constexpr int PP = (W + 2) * WP;        // padded positions per image
    static constexpr int HALO = WP + 1;
    static constexpr int MROWS = T * 128;          // output rows per CTA
    static constexpr int ROWS = MROWS + 2 * HALO;
    static constexpr int CH = C / 4;               // 16-byte chunks per row
    static constexpr int RSTEP = NT / CH;          // rows advanced per staging iteration
    static constexpr int NE = (ROWS + RSTEP - 1) / RSTEP;
    static constexpr int PLANE = ROWS * 16;        // bytes
    static constexpr int A_BYTES = CH * PLANE;
    static constexpr int BTAP = CH * N * 16;       // bytes per tap
    static constexpr int B_BYTES = 9 * BTAP;
    static constexpr int ROWTAB_BYTES = ((ROWS * 4 + 15) / 16) * 16;
    static constexpr int PART_FLOATS = 8 * 16 * 2 * 2;                 // [warp][16 cols][sum,sq] x 2 items
    static constexpr int RED_FLOATS = 2048;                            // 1024 doubles: bn_partial_sums / the last-CTA finalisers
    static constexpr int OFF_B = A_BYTES;
    static constexpr int OFF_ROWTAB = OFF_B + B_BYTES;
    static constexpr int OFF_PART = OFF_ROWTAB + ROWTAB_BYTES;
    static constexpr int OFF_RED = (OFF_PART + PART_FLOATS * 4 + 15) / 16 * 16;
    static constexpr int OFF_AFF = OFF_RED + RED_FLOATS * 4;          // BWD variant: scale, shift, mean, invstd of the differentiated BatchNorm
    static constexpr int OFF_COEF = OFF_AFF + 4 * N * 4;               // BWD variant: c0, c1, c2 of the apply evaluated in the prologue (lazy path)
    static constexpr int OFF_BAR = OFF_COEF + 3 * N * 4;
    static constexpr size_t SMEM_BYTES = OFF_BAR + 64;
    static constexpr uint32_t TMEM_COLS = 64;
    static constexpr int CB = C == 16 ? 16 : 32;   // columns per epilogue work item (64 B / 128 B of an output row)
    static constexpr int ITEMS = T * (N / CB);     // (tile, CB-column block) epilogue work items: 4 (C = 16) or 2
    static constexpr int NIT = ITEMS / 2;          // items per warp: the two warp groups (warps 0-3 / 4-7) alternate
    static_assert(C % 8 == 0 && (C == 16 || C == 32 || C == 64), "tensor-core conv: C in {16,32,64}");
    static_assert(NT % CH == 0 && NE <= 32 && T * N == 64 && NIT * 2 == ITEMS && NIT * CB == 32 && PLANE >= 2048, "tiling");
};

// round-to-nearest (ties away, = cvt.rna.tf32.f32 on finite values) with two integer operations: the cvt instruction runs at a fraction of the ALU rate
// and the staging transform converts every element of the tile
__device__ __forceinline__ float to_tf32_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// 288 threads: warps 0-7 stage / transform / run the epilogue, lane 0 of warp 8 issues every MMA.  The CTA's T tiles are processed as two halves
// (T >= 2): the copies of both halves are issued up front as two cp.async groups; while the tensor core works on the first half the workers transform
// the second, and the first half's epilogue overlaps the second half's MMAs.
// MODE 0: forward (optional BN + ReLU prologue, optional statistics);  1: data gradient with the fused BatchNorm backward (BnBwdFuse);
// MODE 2: forward whose prologue also finishes the PREVIOUS residual block — staged value = relu(scale*in + shift + pro_res) with `in` the raw conv_b
//         output of that block and pro_res its input — and writes that block output to pro_out for the rows this CTA owns: `F.relu(residual + bn_b(..))`
//         (resnet.py:382) never runs as a launch of its own between two stride-1 blocks.
template <int C, int W, int MODE>
__global__ void __launch_bounds__(288) __maxnreg__(MODE == 1 ? LC_BWD_MAXNREG : (MODE == 2 ? 80 : 64)) conv3x3_tc_kernel(ConvTcArgs a) {
    constexpr int BWD = MODE == 1 ? 1 : 0;
    constexpr bool RES = MODE == 2;
    using K = ConvTcCfg<C, W>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + K::OFF_B;
    int* s_rowsrc = reinterpret_cast<int*>(smem_raw + K::OFF_ROWTAB);     // per staged row: source pixel index, or -1 (border)
    float* s_part = reinterpret_cast<float*>(smem_raw + K::OFF_PART);
    float* s_red = reinterpret_cast<float*>(smem_raw + K::OFF_RED);
    float* s_aff = reinterpret_cast<float*>(smem_raw + K::OFF_AFF);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);   // [0] first half's MMAs done, [1] weights landed, [2] second half's MMAs done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = tid < K::NT;
    const int total = a.B * K::PP;
    const int q0 = (int)blockIdx.x * K::MROWS;
    constexpr int HALF = K::T >= 2 ? K::T / 2 : 1;                         // tiles in the first half
    constexpr int ROWS0 = HALF * 128 + 2 * K::HALO;                       // staged rows the first half reads
    constexpr int I0 = K::T >= 2 ? ((ROWS0 + K::RSTEP - 1) / K::RSTEP < K::NE ? (ROWS0 + K::RSTEP - 1) / K::RSTEP : K::NE) : K::NE;
    LC_TSTAMP(0);

    if (tid == 32) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
    }
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);

    // ---- row table: the only place that divides.  Row r of the staged tile <-> padded position Q = q0 - HALO + r ----------
    for (int r = tid; r < K::ROWS; r += 288) {
        const int Q = q0 - K::HALO + r;
        int src = -1;
        if (Q >= 0 && Q < total) {
            const int n = Q / K::PP, rem = Q - n * K::PP;
            const int hp = rem / K::WP, wp = rem - hp * K::WP;
            if (hp >= 1 && hp <= W && wp >= 1 && wp <= W) src = (n * W + (hp - 1)) * W + (wp - 1);
        }
        s_rowsrc[r] = src;
    }
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, the row table) touches nothing a predecessor kernel writes, so
    // it overlaps the predecessor's tail; weights, activations, BN coefficients and the addend are only read after griddepcontrol.wait
    // (= predecessor complete and flushed).  Our own dependent may start its prologue right away.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // weights are already in UMMA order in global memory: one bulk copy, tracked by bar[1]
    if (tid == 32) bulk_load(smem_u32(sB), a.wtc, (uint32_t)K::B_BYTES, bar + 1);
    if (BWD && a.bw.y != nullptr && tid < K::N) {
        const bool bn = a.bw.scale != nullptr;
        s_aff[tid] = bn ? a.bw.scale[tid] : 1.f; s_aff[K::N + tid] = bn ? a.bw.shift[tid] : 0.f;
        s_aff[2 * K::N + tid] = a.bw.mean[tid]; s_aff[3 * K::N + tid] = a.bw.invstd[tid];
    }
    __syncthreads();
    LC_TSTAMP(7);

    // lazy BatchNorm coefficients, phase 1: this thread's rows of the producer's partial sums (their registers die before the tile copies start)
    const bool lazy = !BWD && a.pro_lazy.partial != nullptr;
    const bool blazy = BWD && a.pro_y != nullptr && a.pro_blazy.partial != nullptr;
    if (lazy) bn_partial_sums_load(a.pro_lazy.partial, a.pro_lazy.nparts, C, reinterpret_cast<double*>(s_red));
    if (blazy) bn_partial_sums_load(a.pro_blazy.partial, a.pro_blazy.nparts, C, reinterpret_cast<double*>(s_red));

    // ---- stage A: worker thread -> fixed 16-byte channel chunk j, rows r0, r0+RSTEP, ...  Source rows first (registers), then every copy of both
    //      halves is issued before any wait: iterations [0, I0) form cp.async group 0 (all rows the first half's MMAs read), the rest group 1 ---------
    const int j = tid % K::CH, r0 = tid / K::CH;
    const uint32_t sA_col = smem_u32(sA) + (uint32_t)j * K::PLANE;
    const float* in_col = a.in + j * 4;
    uint32_t validmask = 0;
    const bool proy = BWD && a.pro_y != nullptr;
    const bool pres = RES && a.pro_res != nullptr;
    float4 yreg[(BWD || RES) ? K::NE : 1];      // BWD / RES: the second operand of the prologue (BatchNorm input / residual) stays in registers
    if (worker) {
        int srcs[K::NE];
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            const int r = r0 + i * K::RSTEP;
            srcs[i] = r < K::ROWS ? s_rowsrc[r] : -2;
        }
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            const int r = r0 + i * K::RSTEP;
            const int src = srcs[i];
            if (src != -2) {
                const bool ok = src >= 0;
                cp_async16(sA_col + (uint32_t)r * 16, ok ? in_col + (size_t)src * C : a.in, ok ? 16u : 0u);
                validmask |= (ok ? 1u : 0u) << i;
            }
            if (i == I0 - 1) cp_async_commit();
        }
        cp_async_commit();
        if (BWD || RES) {
            if (proy || pres) {
                const float* second = BWD ? a.pro_y : a.pro_res;
#pragma unroll
                for (int i = 0; i < K::NE; ++i)
                    if (validmask & (1u << i)) yreg[i] = ldg4(second + (size_t)srcs[i] * C + j * 4);
            }
        }
    }
    LC_TSTAMP(1);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), c2 = sh;
    const bool pro = !BWD && (a.pro_scale != nullptr || lazy);
    if (lazy) {      // while the tile copies are in flight: this CTA's own reduction of the producer's statistics
        bn_lazy_affine_finish(a.pro_lazy, C, reinterpret_cast<double*>(s_red), s_aff);
        sc = *reinterpret_cast<const float4*>(s_aff + j * 4); sh = *reinterpret_cast<const float4*>(s_aff + C + j * 4);
    } else if (pro) { sc = ldg4(a.pro_scale + j * 4); sh = ldg4(a.pro_shift + j * 4); }
    if (proy) {
        if (blazy) {
            float* s_coef = reinterpret_cast<float*>(smem_raw + K::OFF_COEF);
            bn_bwd_lazy_coef_finish(a.pro_blazy, C, reinterpret_cast<double*>(s_red), s_coef);
            sc = *reinterpret_cast<const float4*>(s_coef + j * 4); sh = *reinterpret_cast<const float4*>(s_coef + C + j * 4);
            c2 = *reinterpret_cast<const float4*>(s_coef + 2 * C + j * 4);
        } else { sc = ldg4(a.pro_coef + j * 4); sh = ldg4(a.pro_coef + C + j * 4); c2 = ldg4(a.pro_coef + 2 * C + j * 4); }
    }
    // each worker transforms the chunks it copied itself: producer BN + ReLU (forward) or the BatchNorm-backward apply (BWD), then TF32 rounding
    auto transform = [&](int ibeg, int iend) {
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            if (i >= ibeg && i < iend && (validmask & (1u << i))) {
                float4* p4 = reinterpret_cast<float4*>(sA + (size_t)j * K::PLANE + (size_t)(r0 + i * K::RSTEP) * 16);
                float4 v = *p4;
                if (BWD) {
                    if (proy) {       // dy = c0*g + c1*y + c2 (same fma order as bn_bwd_apply_kernel)
                        v.x = fmaf(sc.x, v.x, fmaf(sh.x, yreg[i].x, c2.x));
                        v.y = fmaf(sc.y, v.y, fmaf(sh.y, yreg[i].y, c2.y));
                        v.z = fmaf(sc.z, v.z, fmaf(sh.z, yreg[i].z, c2.z));
                        v.w = fmaf(sc.w, v.w, fmaf(sh.w, yreg[i].w, c2.w));
                    }
                } else if (RES) {
                    if (pres) {       // the previous block's output: relu(bn_b(y) + residual), same operation order as bn_act_fwd_kernel
                        v.x = fmaxf(fmaf(v.x, sc.x, sh.x) + yreg[i].x, 0.f);
                        v.y = fmaxf(fmaf(v.y, sc.y, sh.y) + yreg[i].y, 0.f);
                        v.z = fmaxf(fmaf(v.z, sc.z, sh.z) + yreg[i].z, 0.f);
                        v.w = fmaxf(fmaf(v.w, sc.w, sh.w) + yreg[i].w, 0.f);
                        const int r = r0 + i * K::RSTEP;
                        if (r >= K::HALO && r < K::HALO + K::MROWS)       // rows this CTA owns (the halo rows belong to its neighbours)
                            *reinterpret_cast<float4*>(a.pro_out + (size_t)s_rowsrc[r] * C + j * 4) = v;
                    }
                } else if (pro) {
                    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
                    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
                    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
                    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
                }
                v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                *p4 = v;
            }
        }
    };
    // MMAs of tiles [t0, t1): fully unrolled so that every descriptor is base + immediate; one commit per half
    auto issue = [&](int t0, int t1, uint64_t* done, uint32_t tmem_base) {
        constexpr uint32_t idesc = make_idesc_tf32(K::N);
        const uint64_t a_hi = make_desc(0, K::PLANE, 128), b_hi = make_desc(0, K::N * 16, 128);
        const uint32_t bBase = smem_u32(sB) >> 4;
#pragma unroll
        for (int t = 0; t < K::T; ++t) {
            if (t < t0 || t >= t1) continue;
            const uint32_t aBase = (smem_u32(sA) >> 4) + (uint32_t)(t * 128 + K::HALO);     // 16-byte units: one row each
            const uint32_t dcol = tmem_base + (uint32_t)(t * K::N);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int kc = 0; kc < C / 8; ++kc) {
                    const int shift = (tap / 3 - 1) * K::WP + (tap % 3 - 1);
                    const uint64_t ad = a_hi | (uint64_t)((aBase + (uint32_t)(shift + 2 * kc * (K::PLANE >> 4))) & 0x3FFF);
                    const uint64_t bd = b_hi | (uint64_t)((bBase + (uint32_t)(tap * (K::BTAP >> 4) + 2 * kc * K::N)) & 0x3FFF);
                    mma_tf32(dcol, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
                }
            }
        }
        mma_commit(done);          // implies tcgen05.fence::before_thread_sync
    };

    if (worker) {
        if (K::T >= 2) cp_async_wait_group<1>(); else cp_async_wait_all();
    }
    LC_TSTAMP(2);
    if (worker) transform(0, I0);
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    LC_TSTAMP(3);
    if (tid == K::NT) {             // the issuer: first half (every tile when T == 1)
        mbar_wait(bar + 1, 0);      // weights landed (async-proxy write, ordered by the mbarrier)
        issue(0, HALF, bar, tmem_base);
    }
    if (K::T >= 2) {
        if (worker) { cp_async_wait_all(); transform(I0, K::NE); }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == K::NT) issue(HALF, K::T, bar + 2, tmem_base);
    }
    LC_TSTAMP(4);

    // ---- epilogue: work items = (tile, CB-column block), first-half items first; warp w reads TMEM lanes 32*(w%4).., warp group w/4 takes item it*2+grp.
    // An accumulator row (one output pixel) arrives in ONE thread, but a warp storing 32 x 16 B at a pixel stride touches 16-32 cache lines per
    // instruction — and the data gradient reads up to three more tensors the same way.  So the rows go through a warp-private staging buffer (dead
    // rows of the A tile, XOR-swizzled 64-byte rows) and every global access of the epilogue runs lane-linear instead: the valid rows of a warp are
    // CONSECUTIVE pixels in memory (borders hold no pixel), i.e. one contiguous block, read / written 512 B per instruction.  In that layout a lane
    // keeps one 4-channel group (c = lane % CPR) of rows p = k*(32/CPR) + lane/CPR, so the BatchNorm sums are 4 running sums + log2(32/CPR) shuffles.
    const bool stats = BWD ? (a.bw.partial != nullptr) : (a.stat.partial != nullptr);
    const int quarter = warp & 3, grp = warp >> 2;
    constexpr int CB = K::CB, CPR = CB / 4, NIT = K::NIT;
    auto stage = [&](int p, int c) -> float4* {      // 2 KB pieces (plane, 128-row block) the MMAs of the other half never read
        const int plane = C == 16 ? quarter : (C == 32 ? 2 * quarter + (c >> 2) : 2 * warp + (c >> 2));
        const int beta = C == 64 ? 0 : grp;
        return reinterpret_cast<float4*>(sA + (size_t)plane * K::PLANE + beta * 2048 + p * 64 + (((c & 3) ^ ((p >> 1) & 3)) << 4));
    };
    bool done = true;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        if (!worker) break;           // the issuer warp owns no TMEM lanes; it only keeps the barriers below company
        const int item = it * 2 + grp;
        const int t = item / (K::N / CB), c0 = (item % (K::N / CB)) * CB;
        const int src = s_rowsrc[K::HALO + t * 128 + quarter * 32 + lane];    // accumulator row == TMEM lane
        const unsigned vmask = __ballot_sync(0xffffffffu, src >= 0);
        const int nvalid = __popc(vmask), rank = __popc(vmask & ((1u << lane) - 1u));
        const int src_first = __shfl_sync(0xffffffffu, src, vmask ? __ffs(vmask) - 1 : 0);
        done = mbar_wait((K::T >= 2 && t >= HALF) ? bar + 2 : bar, 0) && done;
        fence_after_sync();
        if (it == 0) LC_TSTAMP(5);
        {
            float v[CB];
            if (CB == 16) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * K::N + c0), v);
            else tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * K::N + c0), v);
            if (src >= 0) {
#pragma unroll
                for (int c = 0; c < CPR; ++c) *stage(rank, c) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
            }
        }
        __syncwarp();
        const int c = lane % CPR, cc = c0 + c * 4;
        float4 asc = make_float4(0.f, 0.f, 0.f, 0.f), ash = asc, amean = asc, aistd = asc;
        const bool by = BWD && a.bw.y != nullptr;
        if (by) {
            asc = *reinterpret_cast<const float4*>(s_aff + cc); ash = *reinterpret_cast<const float4*>(s_aff + K::N + cc);
            amean = *reinterpret_cast<const float4*>(s_aff + 2 * K::N + cc); aistd = *reinterpret_cast<const float4*>(s_aff + 3 * K::N + cc);
        }
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
        for (int k = 0; k < CPR; ++k) {
            const int p = k * (32 / CPR) + lane / CPR;
            if (p < nvalid) {
                float4 x = *stage(p, c);
                const size_t g = (size_t)(src_first + p) * K::N + cc;
                float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.addend != nullptr) {
                    const float4 a4 = *reinterpret_cast<const float4*>(a.addend + g);
                    x.x += a4.x; x.y += a4.y; x.z += a4.z; x.w += a4.w;
                }
                if (by) {
                    y4 = ldg4(a.bw.y + g);
                    if (a.bw.mask_out != nullptr) {
                        const float4 o4 = ldg4(a.bw.mask_out + g);
                        x.x = o4.x > 0.f ? x.x : 0.f; x.y = o4.y > 0.f ? x.y : 0.f; x.z = o4.z > 0.f ? x.z : 0.f; x.w = o4.w > 0.f ? x.w : 0.f;
                    } else if (a.bw.scale != nullptr) {
                        x.x = fmaf(y4.x, asc.x, ash.x) > 0.f ? x.x : 0.f; x.y = fmaf(y4.y, asc.y, ash.y) > 0.f ? x.y : 0.f;
                        x.z = fmaf(y4.z, asc.z, ash.z) > 0.f ? x.z : 0.f; x.w = fmaf(y4.w, asc.w, ash.w) > 0.f ? x.w : 0.f;
                    }
                }
                *reinterpret_cast<float4*>(a.out + g) = x;
                if (stats) {      // forward: sum, sum of squares;  BWD: sum g, sum g * xhat
                    float4 m2 = x;
                    if (BWD) m2 = make_float4((y4.x - amean.x) * aistd.x, (y4.y - amean.y) * aistd.y, (y4.z - amean.z) * aistd.z, (y4.w - amean.w) * aistd.w);
                    s1.x += x.x; s1.y += x.y; s1.z += x.z; s1.w += x.w;
                    s2.x = fmaf(x.x, m2.x, s2.x); s2.y = fmaf(x.y, m2.y, s2.y); s2.z = fmaf(x.z, m2.z, s2.z); s2.w = fmaf(x.w, m2.w, s2.w);
                }
            }
        }
        if (stats) {      // lanes sharing a channel group differ in the bits above log2(CPR): fixed-order butterfly
#pragma unroll
            for (int off = CPR; off < 32; off <<= 1) {
                s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
                s1.z += __shfl_xor_sync(0xffffffffu, s1.z, off); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, off);
                s2.x += __shfl_xor_sync(0xffffffffu, s2.x, off); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, off);
                s2.z += __shfl_xor_sync(0xffffffffu, s2.z, off); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, off);
            }
            if (lane < CPR) {
                float* sp = s_part + (size_t)((warp * NIT + it) * 2) * CB + c * 4;
                *reinterpret_cast<float4*>(sp) = s1; *reinterpret_cast<float4*>(sp + CB) = s2;
            }
        }
        __syncwarp();      // the next item's rows overwrite this warp's staging piece
    }
    if (!done && tid == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);
    fence_before_sync();
    __syncthreads();
    LC_TSTAMP(6);
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);

    if (stats) {
        // channel c of tile t lives in item (t, c/CB) = it * 2 + grp: sum the 4 warp quarters of every tile in fixed order
        if (tid < 2 * K::N) {
            const int stat = tid / K::N, c = tid % K::N;
            float tsum = 0.f;
#pragma unroll
            for (int t = 0; t < K::T; ++t) {
                const int item = t * (K::N / CB) + c / CB, g = item & 1, it = item >> 1;
#pragma unroll
                for (int q = 0; q < 4; ++q) tsum += s_part[(size_t)((((g * 4 + q) * NIT + it) * 2) + stat) * CB + (c % CB)];
            }
            (BWD ? a.bw.partial : a.stat.partial)[((size_t)blockIdx.x * 2 + stat) * K::N + c] = tsum;
        }
        if (BWD ? a.bw.defer : a.stat.defer) return;       // consumers reduce the partial rows themselves (BnLazy / BnBwdLazy)
        if (last_block_done(BWD ? a.bw.counter : a.stat.counter, gridDim.x)) {      // (the 32 issuer-warp threads only add barrier arrivals below)
            if (BWD) bn_bwd_finalize_last_block<K::N>(a.bw.partial, (int)gridDim.x, (double)a.B * W * W, a.bw.scale, a.bw.mean, a.bw.invstd, a.bw.coef,
                                                      a.bw.dgamma, a.bw.dbeta, s_red);
            else bn_finalize_last_block<K::N, K::NT>(a.stat, (int)gridDim.x, (double)a.B * W * W, s_red);
        }
    }
}

template <int C, int W, int BWD = 0>      // BWD = the kernel's MODE (0 forward, 1 fused data gradient, 2 forward + previous block's residual output)
static inline int conv_tc_launch(const ConvTcArgs& a, cudaStream_t st) {
    using K = ConvTcCfg<C, W>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(conv3x3_tc_kernel<C, W, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    const long long total = (long long)a.B * K::PP;
    const int grid = (int)((total + K::MROWS - 1) / K::MROWS);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(K::NT + 32); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<C, W, BWD>, a) != cudaSuccess) return LC_ERR_CUDA;
    return lc_launch_status();
}

// upper bound on the CTA count (statistics partial rows) for any supported shape
static inline long long conv_tc_tiles(int B, int W) { return ((long long)B * (W + 2) * (W + 2) + 127) / 128; }

}  // namespace tc
}  // namespace lc
