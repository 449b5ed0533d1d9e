// 3x3 / pad 1 / STRIDE 2 convolution (CIN -> 2*CIN, the first conv of stages 2 and 3: core/model/backbone/resnet.py:341-343 with stride 2) as an
// implicit GEMM on tcgen05 (kind::tf32, TMEM accumulators) — forward only; NHWC fp32 in HBM.
//
// A stride-2 tap is not a constant row shift of the input, but it IS one of each PARITY PLANE of the input:  P[pr][pc][n][a][b] = X[n][2a+pr][2b+pc].
// Output (i, j) reads, for tap (dr, dc) in {-1,0,1}^2, plane (dr & 1, dc & 1) at (a, b) = (i + (dr < 0 ? -1 : 0), j + (dc < 0 ? -1 : 0)).  Give every plane
// one leading pad row / column (a = -1, b = -1 -> zeros) and flatten (n, a+1, b+1) over a (WO+1) x (WO+1) grid: Q = (n*(WO+1) + a+1)*(WO+1) + b+1.  Then
// the operand of tap (dr, dc) is plane (dr&1, dc&1) shifted by di*(WO+1) + dj rows, di, dj in {-1, 0}: the formulation of conv_tc.cuh with four staged
// tiles instead of one (and a halo on the low side only).  A CTA owns 128 consecutive Q; rows with a+1 == 0 or b+1 == 0 compute garbage that the
// epilogue drops ((WO+1)^2 / WO^2 - 1 = 13 % / 27 % extra MMA rows).  The four parity tiles are gathered with 16-byte cp.async (zero fill for pad rows).
// Epilogue as in conv_tc.cuh: accumulator rows -> swizzled staging in the dead A tile -> lane-linear stores; BatchNorm statistics of the output as
// per-CTA partial rows (deferred, consumer-side reduction) or finalised by the last CTA.
#pragma once
#include "conv_tc.cuh"

namespace lc {
namespace tc {

struct ConvS2Args {
    const float* in;         // NHWC [B][2*WO][2*WO][CIN]  (an activated block output: no prologue)
    const float* wtc;        // packed [9][CIN/4][COUT][4], TF32-rounded
    float* out;              // NHWC [B][WO][WO][COUT]
    BnStatArgs stat;         // stat.partial nullable: [grid][2][COUT]
    // optional fused shortcut: the block's 1x1 / stride-2 downsample conv (resnet.py:365) reads X[n][2i][2j] = parity plane (0,0) at shift 0, which is already
    // staged: CIN/8 more MMAs into a second accumulator, its own output and BatchNorm statistics (always finalised here: its consumer wants scale / shift)
    const float* w1;         // nullable: packed [CIN/4][COUT][4], TF32-rounded
    float* out1;             // NHWC [B][WO][WO][COUT]
    BnStatArgs stat1;
    int* error_flag;
    int B;
};

template <int CIN, int WO>
struct ConvS2Cfg {
    static constexpr int N = 2 * CIN;                // COUT
    static constexpr int WIN = 2 * WO;
    static constexpr int WP = WO + 1;
    static constexpr int PP = WP * WP;
    static constexpr int HALO = WP + 1;              // low side only
    static constexpr int ROWS = 128 + HALO;
    static constexpr int CH = CIN / 4;
    static constexpr int NT = 256;
    static constexpr int RSTEP = NT / CH;
    static constexpr int NE = (ROWS + RSTEP - 1) / RSTEP;
    static constexpr int PLANE = ROWS * 16;
    static constexpr int A_BYTES = 4 * CH * PLANE;   // four parity tiles
    static constexpr int BTAP = CH * N * 16;
    static constexpr int B_BYTES = 9 * BTAP;
    static constexpr int CB = N / 2;                 // output columns per epilogue work item (one per warp: quarter x column half)
    static constexpr int CPR = CB / 4;
    static constexpr int OFF_B = (A_BYTES + 127) / 128 * 128;
    static constexpr int B1_BYTES = CH * N * 16;     // the 1x1 shortcut weights
    static constexpr int OFF_B1 = OFF_B + B_BYTES;
    static constexpr int OFF_ROWTAB = OFF_B1 + B1_BYTES;                     // [ROWS] pixel index of X[n][2a][2b], or -1
    static constexpr int OFF_DST = OFF_ROWTAB + (ROWS * 4 + 15) / 16 * 16;   // [128] output pixel index, or -1
    static constexpr int OFF_PART = OFF_DST + 512;                           // [2 outputs][8 warps][2][CB]
    static constexpr int OFF_RED = OFF_PART + 2 * 8 * 2 * CB * 4;           // 1024 doubles (last-CTA finaliser)
    static constexpr int OFF_BAR = OFF_RED + 8192;
    static constexpr size_t SMEM_BYTES = OFF_BAR + 64;
    static constexpr uint32_t TMEM_COLS = 2 * N;     // conv accumulator + shortcut accumulator: 64 / 128
    static_assert((CIN == 16 && WO == 16) || (CIN == 32 && WO == 8), "stride-2 tensor-core conv: stage transitions of the CIFAR ResNet");
    static_assert(8 * 32 * CB * 4 <= A_BYTES && PLANE >= 2048 && NE <= 8, "epilogue staging lives in the dead A tile");
};

template <int CIN, int WO>
__global__ void __launch_bounds__(288) conv3x3s2_tc_kernel(ConvS2Args a) {
    using K = ConvS2Cfg<CIN, WO>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + K::OFF_B;
    int* s_rowsrc = reinterpret_cast<int*>(smem_raw + K::OFF_ROWTAB);
    int* s_dst = reinterpret_cast<int*>(smem_raw + K::OFF_DST);
    float* s_part = reinterpret_cast<float*>(smem_raw + K::OFF_PART);
    float* s_red = reinterpret_cast<float*>(smem_raw + K::OFF_RED);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);      // [0] MMAs done, [1] weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = tid < K::NT;
    const int total = a.B * K::PP;
    const int q0 = (int)blockIdx.x * 128;

    if (tid == 32) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);
    for (int r = tid; r < K::ROWS; r += 288) {
        const int Q = q0 - K::HALO + r;
        int src = -1, dst = -1;
        if (Q >= 0 && Q < total) {
            const int n = Q / K::PP, rem = Q - n * K::PP;
            const int ap = rem / K::WP, bp = rem - ap * K::WP;
            if (ap >= 1 && bp >= 1) {
                src = (n * K::WIN + 2 * (ap - 1)) * K::WIN + 2 * (bp - 1);
                dst = (n * WO + (ap - 1)) * WO + (bp - 1);
            }
        }
        s_rowsrc[r] = src;
        if (r >= K::HALO) s_dst[r - K::HALO] = dst;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool sc1 = a.w1 != nullptr;
    if (tid == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"((uint32_t)(K::B_BYTES + (sc1 ? K::B1_BYTES : 0))) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(a.wtc), "r"((uint32_t)K::B_BYTES),
                     "r"(smem_u32(bar + 1)) : "memory");
        if (sc1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_raw + K::OFF_B1)), "l"(a.w1),
                         "r"((uint32_t)K::B1_BYTES), "r"(smem_u32(bar + 1)) : "memory");
    }
    __syncthreads();

    // ---- stage the four parity tiles: worker thread -> fixed 16-byte channel chunk j, rows r0, r0 + RSTEP, ... of every plane ----------------------
    const int j = tid % K::CH, r0 = tid / K::CH;
    uint32_t validmask = 0;
    if (worker) {
        int srcs[K::NE];
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            const int r = r0 + i * K::RSTEP;
            srcs[i] = r < K::ROWS ? s_rowsrc[r] : -2;
        }
#pragma unroll
        for (int par = 0; par < 4; ++par) {
            const int poff = (par >> 1) * K::WIN + (par & 1);             // pixel offset of parity (pr, pc)
            const uint32_t dst = smem_u32(sA) + (uint32_t)((par * K::CH + j) * K::PLANE);
#pragma unroll
            for (int i = 0; i < K::NE; ++i) {
                const int r = r0 + i * K::RSTEP, src = srcs[i];
                if (src != -2) {
                    const bool ok = src >= 0;
                    cp_async16(dst + (uint32_t)r * 16, ok ? a.in + (size_t)(src + poff) * CIN + j * 4 : a.in, ok ? 16u : 0u);
                    if (par == 0) validmask |= (ok ? 1u : 0u) << i;
                }
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        // TF32 rounding of the chunks this thread copied (the tensor core would truncate)
#pragma unroll
        for (int par = 0; par < 4; ++par) {
#pragma unroll
            for (int i = 0; i < K::NE; ++i) {
                if (validmask & (1u << i)) {
                    float4* p4 = reinterpret_cast<float4*>(sA + (size_t)(par * K::CH + j) * K::PLANE + (size_t)(r0 + i * K::RSTEP) * 16);
                    float4 v = *p4;
                    v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    *p4 = v;
                }
            }
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == K::NT) {          // the issuer
        mbar_wait(bar + 1, 0);
        constexpr uint32_t idesc = make_idesc_tf32(K::N);
        const uint64_t a0 = make_desc(0, K::PLANE, 128) | (uint64_t)(smem_u32(sA) >> 4), b0 = make_desc(0, K::N * 16, 128) | (uint64_t)(smem_u32(sB) >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dr = tap / 3 - 1, dc = tap % 3 - 1;
            const int par = (dr & 1) * 2 + (dc & 1);
            const int shift = (dr < 0 ? -K::WP : 0) + (dc < 0 ? -1 : 0);
#pragma unroll
            for (int kc = 0; kc < CIN / 8; ++kc) {
                const uint64_t ad = a0 + (uint64_t)((par * K::CH + 2 * kc) * (K::PLANE >> 4) + K::HALO + shift);
                const uint64_t bd = b0 + (uint64_t)(tap * (K::BTAP >> 4) + 2 * kc * K::N);
                mma_tf32(tmem_base, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
            }
        }
        if (sc1) {       // shortcut: parity plane (0,0), no shift
            const uint64_t b1 = make_desc(0, K::N * 16, 128) | (uint64_t)(smem_u32(smem_raw + K::OFF_B1) >> 4);
#pragma unroll
            for (int kc = 0; kc < CIN / 8; ++kc)
                mma_tf32(tmem_base + (uint32_t)K::N, a0 + (uint64_t)(2 * kc * (K::PLANE >> 4) + K::HALO), b1 + (uint64_t)(2 * kc * K::N), idesc, kc != 0 ? 1u : 0u);
        }
        mma_commit(bar);
    }

    // ---- epilogue: warp (quarter, grp) takes accumulator rows 32*quarter.. and columns [grp*CB, (grp+1)*CB) ---------------------------------------------
    const bool stats = a.stat.partial != nullptr;
    const int quarter = warp & 3, grp = warp >> 2;
    bool done = true;
    if (worker) {
        const int c0 = grp * K::CB;
        const int dst = s_dst[quarter * 32 + lane];
        const unsigned vmask = __ballot_sync(0xffffffffu, dst >= 0);
        const int nvalid = __popc(vmask), rank = __popc(vmask & ((1u << lane) - 1u));
        const int dst_first = __shfl_sync(0xffffffffu, dst, vmask ? __ffs(vmask) - 1 : 0);
        done = mbar_wait(bar, 0);
        fence_after_sync();
        unsigned char* stg = sA + (size_t)warp * (32 * K::CB * 4);
        auto stage = [&](int p, int cc) -> float4* {
            return reinterpret_cast<float4*>(stg + (size_t)(cc >> 2) * 2048 + p * 64 + (((cc & 3) ^ ((p >> 1) & 3)) << 4));
        };
        const int c = lane % K::CPR;
#pragma unroll
        for (int o = 0; o < 2; ++o) {            // o = 0: the 3x3 conv; o = 1: the fused 1x1 shortcut
            if (o == 1 && !sc1) break;
            float* outp = o == 0 ? a.out : a.out1;
            const bool st_o = o == 0 ? stats : (a.stat1.partial != nullptr);
            {
                float v[K::CB];
                if (K::CB == 16) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(o * K::N + c0), v);
                else tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(o * K::N + c0), v);
                if (dst >= 0) {
#pragma unroll
                    for (int cc = 0; cc < K::CPR; ++cc) *stage(rank, cc) = make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]);
                }
            }
            __syncwarp();
            float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
            for (int kk = 0; kk < K::CPR; ++kk) {
                const int p = kk * (32 / K::CPR) + lane / K::CPR;
                if (p < nvalid) {
                    const float4 x = *stage(p, c);
                    *reinterpret_cast<float4*>(outp + (size_t)(dst_first + p) * K::N + c0 + c * 4) = x;
                    s1.x += x.x; s1.y += x.y; s1.z += x.z; s1.w += x.w;
                    s2.x = fmaf(x.x, x.x, s2.x); s2.y = fmaf(x.y, x.y, s2.y); s2.z = fmaf(x.z, x.z, s2.z); s2.w = fmaf(x.w, x.w, s2.w);
                }
            }
            if (st_o) {
#pragma unroll
                for (int off = K::CPR; off < 32; off <<= 1) {
                    s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
                    s1.z += __shfl_xor_sync(0xffffffffu, s1.z, off); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, off);
                    s2.x += __shfl_xor_sync(0xffffffffu, s2.x, off); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, off);
                    s2.z += __shfl_xor_sync(0xffffffffu, s2.z, off); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, off);
                }
                if (lane < K::CPR) {
                    float* sp = s_part + (size_t)(o * 8 + warp) * 2 * K::CB + c * 4;
                    *reinterpret_cast<float4*>(sp) = s1; *reinterpret_cast<float4*>(sp + K::CB) = s2;
                }
            }
            __syncwarp();      // the second output reuses this warp's staging block
        }
    }
    if (!done && lane == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);

    const bool stats1 = sc1 && a.stat1.partial != nullptr;
    if (stats || stats1) {
        if (tid < 2 * K::N) {
            const int stat = tid / K::N, ch = tid % K::N, g = ch / K::CB;
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (o == 0 ? !stats : !stats1) continue;
                float tsum = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) tsum += s_part[(size_t)((o * 8 + g * 4 + q) * 2 + stat) * K::CB + (ch % K::CB)];
                (o == 0 ? a.stat.partial : a.stat1.partial)[((size_t)blockIdx.x * 2 + stat) * K::N + ch] = tsum;
            }
        }
        const bool fin0 = stats && !a.stat.defer;
        if (!fin0 && !stats1) return;       // deferred: consumers reduce the partial rows themselves (BnLazy)
        if (last_block_done(stats1 ? a.stat1.counter : a.stat.counter, gridDim.x)) {
            if (fin0) bn_finalize_last_block<K::N, 256>(a.stat, (int)gridDim.x, (double)a.B * WO * WO, s_red);
            if (fin0 && stats1) __syncthreads();
            if (stats1) bn_finalize_last_block<K::N, 256>(a.stat1, (int)gridDim.x, (double)a.B * WO * WO, s_red);
        }
    }
}

static inline int conv_s2_tc_grid(long long batch, int wo) { return (int)((batch * (wo + 1) * (wo + 1) + 127) / 128); }

template <int CIN, int WO>
static inline int conv_s2_tc_launch(const ConvS2Args& a, cudaStream_t st) {
    using K = ConvS2Cfg<CIN, WO>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(conv3x3s2_tc_kernel<CIN, WO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)conv_s2_tc_grid(a.B, WO)); cfg.blockDim = dim3(K::NT + 32); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, conv3x3s2_tc_kernel<CIN, WO>, a) != cudaSuccess) return LC_ERR_CUDA;
    return lc_launch_status();
}

// ---------------------------------------------------------------------------------------------------------------------------------------------------------
// Data gradient of the same stride-2 conv: dX[n][y][x][ci] = sum over taps / co of dY[n][i][j][co] * W[co][ci][dr+1][dc+1] with y = 2i + dr, x = 2j + dc.
// Per PARITY PLANE of dX (py, px) = (y & 1, x & 1) only the taps with dr = py, dc = px (mod 2) contribute, each a constant row shift of the flattened dY
// grid (one TRAILING pad row / column: Q = (n*(WO+1) + i)*(WO+1) + j; tap dr = -1 reads i = a + 1, dr = 0 / +1 read i = a):
//     plane (0,0): tap (0,0)            plane (0,1): taps (0,-1), (0,+1)            plane (1,0): taps (-1,0), (+1,0)            plane (1,1): the four corners
// so the nine taps are nine MMA groups (K = COUT) over ONE staged dY tile into four TMEM accumulators (128 positions x CIN each) — no multiplication by
// the zeros of a dilated gradient, which is what the CUDA-core kernel it replaces spends 3/4 of its time on.  The epilogue writes pixel PAIRS: planes
// (py, 0) and (py, 1) of a position are neighbouring pixels of dX, 2*CIN contiguous floats.
struct DgradS2Args {
    const float* dy;         // NHWC [B][WO][WO][COUT]
    const float* wtc;        // packed [9][COUT/4][CIN][4] (tap NOT flipped), TF32-rounded
    float* out;              // NHWC [B][2*WO][2*WO][CIN]  (every element written)
    // optional fused data gradient of the block's 1x1 / stride-2 shortcut conv: dX[n][2i][2j] += W1^T dY1[n][i][j], i.e. KC/8 more MMAs on a second staged
    // tile into the accumulator of parity plane (0,0)
    const float* dy1;        // nullable: NHWC [B][WO][WO][COUT]
    const float* w1;         // packed [COUT/4][CIN][4], TF32-rounded
    int* error_flag;
    int B;
};

template <int CIN, int WO>
struct DgradS2Cfg {
    static constexpr int KC = 2 * CIN;               // contraction = COUT
    static constexpr int WIN = 2 * WO;
    static constexpr int WP = WO + 1;
    static constexpr int PP = WP * WP;
    static constexpr int HALO = WP + 1;              // high side only
    static constexpr int ROWS = 128 + HALO;
    static constexpr int CH = KC / 4;
    static constexpr int NT = 256;
    static constexpr int RSTEP = NT / CH;
    static constexpr int NE = (ROWS + RSTEP - 1) / RSTEP;
    static constexpr int PLANE = ROWS * 16;
    static constexpr int A_BYTES = CH * PLANE;
    static constexpr int BTAP = CH * CIN * 16;
    static constexpr int B_BYTES = 9 * BTAP;
    static constexpr int PAIRB = 2 * CIN * 4;        // bytes of one output pixel pair
    static constexpr int CPR = PAIRB / 16;           // 16-byte chunks per pair: 8 / 16
    static constexpr int STG_WARP = 32 * PAIRB;      // 4 KB / 8 KB
    static constexpr int OFF_B = (A_BYTES + 127) / 128 * 128;
    static constexpr int A1_PLANE = 128 * 16;        // shortcut gradient tile: 128 rows, no halo
    static constexpr int A1_BYTES = CH * A1_PLANE;
    static constexpr int B1_BYTES = CH * CIN * 16;
    static constexpr int OFF_A1 = (OFF_B + B_BYTES + 127) / 128 * 128;
    static constexpr int OFF_B1 = OFF_A1 + A1_BYTES;
    static constexpr int OFF_STG = (OFF_B1 + B1_BYTES + 127) / 128 * 128;
    static constexpr int OFF_ROWTAB = OFF_STG + 8 * STG_WARP;                // [ROWS] dY pixel index or -1
    static constexpr int OFF_DST = OFF_ROWTAB + (ROWS * 4 + 15) / 16 * 16;   // [128] dX pixel index of (2a, 2b) or -1
    static constexpr int OFF_DSTW = OFF_DST + 512;                           // [8 warps][32] compacted per warp
    static constexpr int OFF_BAR = OFF_DSTW + 8 * 32 * 4;
    static constexpr size_t SMEM_BYTES = OFF_BAR + 64;
    static constexpr uint32_t TMEM_COLS = 4 * CIN;   // 64 / 128
    static_assert((CIN == 16 && WO == 16) || (CIN == 32 && WO == 8), "stage transitions of the CIFAR ResNet");
    static_assert(NE <= 16, "staging");
};

template <int CIN, int WO>
__global__ void __launch_bounds__(288) dgrad3x3s2_tc_kernel(DgradS2Args a) {
    using K = DgradS2Cfg<CIN, WO>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + K::OFF_B;
    int* s_rowsrc = reinterpret_cast<int*>(smem_raw + K::OFF_ROWTAB);
    int* s_dst = reinterpret_cast<int*>(smem_raw + K::OFF_DST);
    int* s_dstw = reinterpret_cast<int*>(smem_raw + K::OFF_DSTW);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = tid < K::NT;
    const int total = a.B * K::PP;
    const int q0 = (int)blockIdx.x * 128;

    if (tid == 32) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);
    for (int r = tid; r < K::ROWS; r += 288) {
        const int Q = q0 + r;
        int src = -1, dst = -1;
        if (Q < total) {
            const int n = Q / K::PP, rem = Q - n * K::PP;
            const int i = rem / K::WP, j = rem - i * K::WP;
            if (i < WO && j < WO) {
                src = (n * WO + i) * WO + j;
                dst = (n * K::WIN + 2 * i) * K::WIN + 2 * j;
            }
        }
        s_rowsrc[r] = src;
        if (r < 128) s_dst[r] = dst;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool sc1 = a.dy1 != nullptr;
    if (tid == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"((uint32_t)(K::B_BYTES + (sc1 ? K::B1_BYTES : 0))) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(a.wtc), "r"((uint32_t)K::B_BYTES),
                     "r"(smem_u32(bar + 1)) : "memory");
        if (sc1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_raw + K::OFF_B1)), "l"(a.w1),
                         "r"((uint32_t)K::B1_BYTES), "r"(smem_u32(bar + 1)) : "memory");
    }
    __syncthreads();

    const int j = tid % K::CH, r0 = tid / K::CH;
    if (worker) {
        uint32_t validmask = 0;
        const uint32_t dstp = smem_u32(sA) + (uint32_t)(j * K::PLANE);
        if (sc1) {       // shortcut gradient tile: rows [0, 128) only (the 1x1 conv has no taps to shift to)
            const uint32_t dst1 = smem_u32(smem_raw + K::OFF_A1) + (uint32_t)(j * K::A1_PLANE);
#pragma unroll
            for (int i = 0; i < (128 + K::RSTEP - 1) / K::RSTEP; ++i) {
                const int r = r0 + i * K::RSTEP;
                if (r < 128) {
                    const int src = s_rowsrc[r];
                    cp_async16(dst1 + (uint32_t)r * 16, src >= 0 ? a.dy1 + (size_t)src * K::KC + j * 4 : a.dy1, src >= 0 ? 16u : 0u);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            const int r = r0 + i * K::RSTEP;
            if (r < K::ROWS) {
                const int src = s_rowsrc[r];
                const bool ok = src >= 0;
                cp_async16(dstp + (uint32_t)r * 16, ok ? a.dy + (size_t)src * K::KC + j * 4 : a.dy, ok ? 16u : 0u);
                validmask |= (ok ? 1u : 0u) << i;
            }
        }
        cp_async_commit();
        cp_async_wait_all();
#pragma unroll
        for (int i = 0; i < K::NE; ++i) {
            if (validmask & (1u << i)) {
                float4* p4 = reinterpret_cast<float4*>(sA + (size_t)j * K::PLANE + (size_t)(r0 + i * K::RSTEP) * 16);
                float4 v = *p4;
                v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                *p4 = v;
            }
        }
        if (sc1) {
#pragma unroll
            for (int i = 0; i < (128 + K::RSTEP - 1) / K::RSTEP; ++i) {
                const int r = r0 + i * K::RSTEP;
                if (r < 128 && s_rowsrc[r] >= 0) {
                    float4* p4 = reinterpret_cast<float4*>(smem_raw + K::OFF_A1 + (size_t)j * K::A1_PLANE + (size_t)r * 16);
                    float4 v = *p4;
                    v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    *p4 = v;
                }
            }
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == K::NT) {
        mbar_wait(bar + 1, 0);
        constexpr uint32_t idesc = make_idesc_tf32(CIN);
        const uint64_t a0 = make_desc(0, K::PLANE, 128) | (uint64_t)(smem_u32(sA) >> 4), b0 = make_desc(0, CIN * 16, 128) | (uint64_t)(smem_u32(sB) >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dr = tap / 3 - 1, dc = tap % 3 - 1;
            const int par = (dr & 1) * 2 + (dc & 1);
            const int shift = (dr < 0 ? K::WP : 0) + (dc < 0 ? 1 : 0);
            const bool first = tap == 0 || tap == 1 || tap == 3 || tap == 4;       // first tap of its parity plane's accumulator
#pragma unroll
            for (int kc = 0; kc < K::KC / 8; ++kc) {
                const uint64_t ad = a0 + (uint64_t)(2 * kc * (K::PLANE >> 4) + shift);
                const uint64_t bd = b0 + (uint64_t)(tap * (K::BTAP >> 4) + 2 * kc * CIN);
                mma_tf32(tmem_base + (uint32_t)(par * CIN), ad, bd, idesc, (first && kc == 0) ? 0u : 1u);
            }
        }
        if (sc1) {       // shortcut: accumulate into parity plane (0,0) (after its own tap above)
            const uint64_t a1 = make_desc(0, K::A1_PLANE, 128) | (uint64_t)(smem_u32(smem_raw + K::OFF_A1) >> 4);
            const uint64_t b1 = make_desc(0, CIN * 16, 128) | (uint64_t)(smem_u32(smem_raw + K::OFF_B1) >> 4);
#pragma unroll
            for (int kc = 0; kc < K::KC / 8; ++kc)
                mma_tf32(tmem_base, a1 + (uint64_t)(2 * kc * (K::A1_PLANE >> 4)), b1 + (uint64_t)(2 * kc * CIN), idesc, 1u);
        }
        mma_commit(bar);
    }

    // ---- epilogue: warp (quarter, py): positions 32*quarter.. , planes (py, 0) and (py, 1) = one pixel pair per position -----------------------------------
    bool done = true;
    if (worker) {
        const int quarter = warp & 3, py = warp >> 2;
        const int dst = s_dst[quarter * 32 + lane];
        const unsigned vmask = __ballot_sync(0xffffffffu, dst >= 0);
        const int nvalid = __popc(vmask), rank = __popc(vmask & ((1u << lane) - 1u));
        if (dst >= 0) s_dstw[warp * 32 + rank] = dst + py * K::WIN;
        done = mbar_wait(bar, 0);
        fence_after_sync();
        unsigned char* stg = smem_raw + K::OFF_STG + (size_t)warp * K::STG_WARP;
        auto stage = [&](int p, int cc) -> float4* {      // 2 KB pieces of 32 x 64-byte sub-rows, XOR-swizzled
            return reinterpret_cast<float4*>(stg + (size_t)(cc >> 2) * 2048 + p * 64 + (((cc & 3) ^ ((p >> 1) & 3)) << 4));
        };
#pragma unroll
        for (int h = 0; h < 2 * CIN / 32; ++h) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(py * 2 * CIN + h * 32), v);
            if (dst >= 0) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) *stage(rank, h * 8 + cc) = make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]);
            }
        }
        __syncwarp();
        const int c = lane % K::CPR;
#pragma unroll
        for (int kk = 0; kk < K::CPR; ++kk) {
            const int p = kk * (32 / K::CPR) + lane / K::CPR;
            if (p < nvalid) *reinterpret_cast<float4*>(a.out + (size_t)s_dstw[warp * 32 + p] * CIN + c * 4) = *stage(p, c);
        }
    }
    if (!done && lane == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

template <int CIN, int WO>
static inline int dgrad_s2_tc_launch(const DgradS2Args& a, cudaStream_t st) {
    using K = DgradS2Cfg<CIN, WO>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(dgrad3x3s2_tc_kernel<CIN, WO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)conv_s2_tc_grid(a.B, WO)); cfg.blockDim = dim3(K::NT + 32); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, dgrad3x3s2_tc_kernel<CIN, WO>, a) != cudaSuccess) return LC_ERR_CUDA;
    return lc_launch_status();
}

}  // namespace tc
}  // namespace lc
