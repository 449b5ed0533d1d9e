// Persistent, warp-specialised forward variant of the 3x3 / pad 1 / stride 1 tensor-core convolution (conv_tc.cuh holds the formulation: flattened padded
// rows, tap = descriptor row shift, planar no-swizzle K-major tiles, kind::tf32, TMEM accumulators).  Layers 16->16 @32x32 and 32->32 @16x16
// (core/model/backbone/resnet.py:295,298), kernel MODE 0 (optional BatchNorm + ReLU prologue) and MODE 2 (prologue also finishes the previous residual
// block, resnet.py:382); the data-gradient variant (MODE 1) and the 64-channel stage stay on conv3x3_tc_kernel.
//
// Why: conv3x3_tc_kernel is ONE wave of CTAs that all load, then all compute, then all store — measured (tools/tc_variants.cu) as a fixed ~6.5 us plus the
// bytes at HALF of the HBM rate, because 16-byte cp.async gathers cap the bytes an SM keeps in flight and no phase overlaps another.  Here one CTA per SM
// owns a CONTIGUOUS range of up to TMAX 128-row tiles (so consecutive tiles share their halo rows through shared memory) and four roles run concurrently:
//   warp 18     loader   : the valid rows of a 128-row chunk are consecutive pixels in HBM (border positions hold no pixel), so a chunk of an operand is
//                          ONE cp.async.bulk (TMA 1-D) of up to 8 / 16 KB into a ring slot; up to 64 KB per SM in flight, completion on mbarriers
//   warps 0-7   transform: ring slot (row-major, compact) -> BatchNorm + ReLU (+ residual, + the previous block's output store) -> TF32 rounding ->
//                          planar UMMA tile (border rows = zeros); releases the slot, signals "chunk k staged"
//   warps 16-17 issuers  : one thread each, alternating tiles (a single thread issues a tcgen05.mma every ~40 ns — measured — which alone would bound
//                          the kernel); tile k needs chunks k and k + 1 (its upper halo): 9 taps x C/8 MMAs into TMEM columns [k*C, (k+1)*C)
//   warps 8-15  epilogue : two warps per TMEM lane quarter, alternating tiles: tcgen05.ld -> warp-private swizzled staging -> lane-linear 512-byte stores;
//                          BatchNorm sums stay in registers across all tiles
// so the stores of tile k, the MMAs of tile k + 1, the transform of chunk k + 2 and the loads of chunks k + 3.. overlap.
#pragma once
#include "conv_tc.cuh"

namespace lc {
namespace tc {

#ifdef LC_TC_TIMING
#define LC_PSTAMP(slot) do { if (a.timing != nullptr) a.timing[(size_t)blockIdx.x * 64 + (slot)] = gtimer(); } while (0)
#else
#define LC_PSTAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ void tcp_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // the 8 transform warps

template <int C, int W>
struct ConvTcpCfg {
    static constexpr int N = C;
    static constexpr int CH = C / 4;                 // 16-byte chunks per row
    static constexpr int ROWB = C * 4;               // bytes per pixel row
    static constexpr int WP = W + 2;
    static constexpr int PP = WP * WP;
    static constexpr int HALO = WP + 1;
    static constexpr int TMAX = C == 16 ? 8 : 3;     // tiles per CTA (B = 128: 1156 / 324 tiles over 148 CTAs)
    static constexpr int RMAX = TMAX * 128 + 2 * HALO;
    static constexpr int PLANE = RMAX * 16;
    static constexpr int A_BYTES = CH * PLANE;
    static constexpr int BTAP = CH * N * 16;
    static constexpr int B_BYTES = 9 * BTAP;
    static constexpr int SLOT = 128 * ROWB;          // one 128-row chunk of one operand
    static constexpr int NS = 65536 / SLOT;          // ring slots: 8 (C = 16) / 4 (C = 32)
    static constexpr int EPI_WARP = 32 * ROWB;       // epilogue staging per warp
    static constexpr int NTHREADS = 608;             // 8 transform + 8 epilogue warps (one warp per scheduler cannot hide its own latencies) + 2 issuers + loader
    static constexpr int NTRANS = 256;
    // Two measured-and-rejected variants stay behind compile-time switches (profiles/r3j_tcp_variants.txt, r3j / r3k_step_ab.txt): neither changes the kernel
    // timed alone on cold inputs (it is bound by the arrival of its chunks and the shared-memory re-reads, not by barrier traffic or the transform's
    // dependency chain); inside the step the elected arrival is neutral as well and the two groups cost 4.5 us per forward pass.
#ifdef LC_TCP_TWO_GROUPS
    static constexpr int NGROUP = 2;                // transform warps 0-3 stage the even chunks, 4-7 the odd ones
#else
    static constexpr int NGROUP = 1;                // all 8 transform warps work on the same chunk
#endif
    static constexpr int GTHREADS = NTRANS / NGROUP;
#ifdef LC_TCP_ELECTED_ARRIVE
    static constexpr int NARRIVE = GTHREADS / 32;   // one elected arrival per transform warp
#else
    static constexpr int NARRIVE = GTHREADS;        // every transform thread arrives on empty[] / staged[] right behind its own proxy fence
#endif
    static constexpr int OFF_B = (A_BYTES + 127) / 128 * 128;
    static constexpr int OFF_RING = (OFF_B + B_BYTES + 127) / 128 * 128;
    static constexpr int OFF_EPI = OFF_RING + NS * SLOT;
    static constexpr int OFF_PART = OFF_EPI + 8 * EPI_WARP;           // [8 epilogue warps][2][N] floats
    static constexpr int OFF_RED = OFF_PART + 8 * 2 * N * 4;         // 1024 doubles: bn_partial_sums / last-CTA finaliser
    static constexpr int OFF_AFF = OFF_RED + 8192;                   // scale, shift of the prologue BatchNorm
    static constexpr int OFF_ROWTAB = OFF_AFF + 2 * N * 4;           // per staged row: pixel index, or -1 (border / outside the batch)
    static constexpr int OFF_LO = OFF_ROWTAB + (RMAX * 4 + 15) / 16 * 16;   // pixels below the first row of chunk k (k <= TMAX + 1)
    static constexpr int OFF_BAR = OFF_LO + 64;
    static constexpr int NBAR = 1 + 2 * NS + (TMAX + 1) + TMAX;      // weights | full[NS] | empty[NS] | staged[TMAX + 1] | mma_done[TMAX]
    static constexpr size_t SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
    static constexpr uint32_t TMEM_COLS = 128;
    static constexpr int CPR = N / 4;                // 16-byte chunks of an output row (the epilogue handles whole rows)
    static_assert(C == 16 || C == 32, "persistent conv: stages 1 and 2");
    static_assert(TMAX * N <= 128 && 65536 % SLOT == 0 && OFF_PART % 16 == 0 && OFF_BAR % 8 == 0, "layout");
};

// number of pixels whose flattened padded position is < Q (Q clamped to [0, total]); for a valid position this IS its pixel index
template <int W>
__device__ __forceinline__ int tcp_pixels_below(int Q, int total) {
    constexpr int WP = W + 2, PP = WP * WP;
    Q = Q < 0 ? 0 : (Q > total ? total : Q);
    const int n = Q / PP, rem = Q - n * PP;
    const int hp = rem / WP, wp = rem - hp * WP;
    const int rows_below = hp - 1 < 0 ? 0 : (hp - 1 > W ? W : hp - 1);
    int cnt = (n * W + rows_below) * W;
    if (hp >= 1 && hp <= W) cnt += wp - 1 < 0 ? 0 : (wp - 1 > W ? W : wp - 1);
    return cnt;
}
template <int W>
__device__ __forceinline__ bool tcp_valid(int Q, int total) {
    constexpr int WP = W + 2, PP = WP * WP;
    if (Q < 0 || Q >= total) return false;
    const int rem = Q % PP;
    const int hp = rem / WP, wp = rem - hp * WP;
    return hp >= 1 && hp <= W && wp >= 1 && wp <= W;
}

// grid size = number of statistics partial rows: every CTA gets floor or ceil of ntiles / grid consecutive tiles, never more than TMAX
static inline int conv_tcp_grid(long long batch, int c, int w, int nsm) {
    const long long ntiles = (batch * (w + 2) * (w + 2) + 127) / 128;
    const long long tmax = c == 16 ? 8 : 3;
    long long g = (ntiles + tmax - 1) / tmax;
    if (g < nsm) g = nsm;
    if (g > ntiles) g = ntiles;
    return (int)g;
}

template <int C, int W, int MODE>
__global__ void __launch_bounds__(608, 1) conv3x3_tcp_kernel(ConvTcArgs a) {
    static_assert(MODE == 0 || MODE == 2, "forward variants");
    constexpr bool RES = MODE == 2;
    using K = ConvTcpCfg<C, W>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sA = smem_raw;
    unsigned char* sB = smem_raw + K::OFF_B;
    unsigned char* sRing = smem_raw + K::OFF_RING;
    float* s_part = reinterpret_cast<float*>(smem_raw + K::OFF_PART);
    float* s_red = reinterpret_cast<float*>(smem_raw + K::OFF_RED);
    float* s_aff = reinterpret_cast<float*>(smem_raw + K::OFF_AFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + K::OFF_BAR);
    uint64_t* wbar = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + K::NS;
    uint64_t* staged = empty + K::NS;
    uint64_t* mdone = staged + (K::TMAX + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + K::NBAR);
    int* s_rowsrc = reinterpret_cast<int*>(smem_raw + K::OFF_ROWTAB);
    int* s_lo = reinterpret_cast<int*>(smem_raw + K::OFF_LO);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) LC_PSTAMP(0);
    const int total = a.B * K::PP;
    const int ntiles = (total + 127) / 128;
    const int t_begin = (int)((long long)blockIdx.x * ntiles / gridDim.x), t_end = (int)((long long)(blockIdx.x + 1) * ntiles / gridDim.x);
    const int nT = t_end - t_begin;                          // 1 .. TMAX
    const int Qbase = t_begin * 128 - K::HALO;               // padded position of staged row 0
    const int R = nT * 128 + 2 * K::HALO;                    // staged rows
    const bool pres = RES && a.pro_res != nullptr;
    const int nops = pres ? 2 : 1;
    const int njobs = (nT + 1) * nops;                       // (chunk, operand) loads; chunk nT holds the last 2 * HALO rows

    // one load job: chunk k of operand op -> ring slot jx % NS.  The chunk's valid rows are the pixels [lo, hi): one contiguous block.
    auto issue_job = [&](int jx) {
        const int k = jx / nops, op = jx - k * nops, s = jx % K::NS;
        const int lo = s_lo[k], hi = s_lo[k + 1];
        const float* src = (op == 0 ? a.in : a.pro_res) + (size_t)lo * C;
        if (hi > lo) bulk_load(smem_u32(sRing + (size_t)s * K::SLOT), src, (uint32_t)(hi - lo) * K::ROWB, full + s);
        else mbar_arrive(full + s);
    };
    int jx_next = 0;
#ifdef LC_TCP_EARLY_LOADER
    if (warp == 18) {
        // Opt-in variant (-DLC_TCP_EARLY_LOADER): the loader does not wait for the CTA's set-up (TMEM allocation, row table, the other barriers: ~1 us) — its
        // warp initialises the barriers its copies complete on and the pixel offset of every chunk (one lane each), waits for the predecessor grid and has the
        // weights + a ring-full of chunks in flight before the CTA-wide barrier below.  Timed alone on cold inputs it wins (8.89 -> 8.44 us at C = 16,
        // 7.71 -> 7.38 at C = 32); inside the step, where programmatic dependent launch already hides the set-up under the predecessor's tail, the forward pass
        // is 8 us SLOWER with it (354 -> 362 us over 20 launches, profiles/r3l_step_ab.txt; 368 us when the refills recomputed the offsets), so it is off.
        if (lane < 1 + K::NS) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + lane)), "r"(1));          // wbar, full[]
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (lane <= nT + 1) s_lo[lane] = tcp_pixels_below<W>(Qbase + (lane * 128 < R ? lane * 128 : R), total);
        __syncwarp();
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (lane == 0) {
            bulk_load(smem_u32(sB), a.wtc, (uint32_t)K::B_BYTES, wbar);
            for (; jx_next < njobs && jx_next < K::NS; ++jx_next) issue_job(jx_next);
        }
    }
    constexpr int BAR0 = 1 + K::NS;              // warp 1 initialises the rest
#else
    constexpr int BAR0 = 0;
#endif
    if (warp == 1 && lane < (K::NBAR - BAR0) / 2 + 1) {       // <= 34 / 19 barriers: two per lane, one fence per lane (a single thread initialising them all is ~1 us)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = BAR0 + lane * 2 + h;
            if (i < K::NBAR) {
                const bool many = (i >= 1 + K::NS && i < 1 + 2 * K::NS) || (i >= 1 + 2 * K::NS && i < 1 + 2 * K::NS + K::TMAX + 1);      // empty[], staged[]
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + i)), "r"(many ? K::NARRIVE : 1));
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, K::TMEM_COLS);
    // row table (the only divisions of the kernel) and the pixel offset of every chunk
    for (int r = tid; r < R; r += K::NTHREADS) {
        const int Q = Qbase + r;
        s_rowsrc[r] = tcp_valid<W>(Q, total) ? tcp_pixels_below<W>(Q, total) : -1;
    }
#ifndef LC_TCP_EARLY_LOADER
    if (tid <= nT + 1) s_lo[tid] = tcp_pixels_below<W>(Qbase + (tid * 128 < R ? tid * 128 : R), total);
#endif
    // Programmatic dependent launch: nothing above reads what a predecessor writes
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) LC_PSTAMP(1);

    const bool lazy = a.pro_lazy.partial != nullptr;
    const bool pro = a.pro_scale != nullptr || lazy;
    bool ok = true;
    const bool stats = a.stat.partial != nullptr;

    if (warp == 18) {
        // ---------------------------------------------------------------- loader: weights, a ring-full of chunks at once, then refills as slots are released
        if (lane == 0) {
#ifndef LC_TCP_EARLY_LOADER
            bulk_load(smem_u32(sB), a.wtc, (uint32_t)K::B_BYTES, wbar);
            for (; jx_next < njobs && jx_next < K::NS; ++jx_next) issue_job(jx_next);
#endif
            for (; jx_next < njobs; ++jx_next) {
                const int s = jx_next % K::NS, u = jx_next / K::NS;
                ok = mbar_wait(empty + s, (uint32_t)((u - 1) & 1)) && ok;
                issue_job(jx_next);
            }
        }
    } else if (warp >= 16) {
        // ---------------------------------------------------------------- MMA issuers: warp 16 even tiles, warp 17 odd tiles
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(K::N);
            // descriptors = one 64-bit base + compile-time offsets (all shared-memory addresses are < 256 KB: the 14-bit start field cannot carry)
            const uint64_t a0 = make_desc(0, K::PLANE, 128) | (uint64_t)(smem_u32(sA) >> 4), b0 = make_desc(0, K::N * 16, 128) | (uint64_t)(smem_u32(sB) >> 4);
            const int par = warp - 16;
            ok = mbar_wait(wbar, 0) && ok;
            ok = mbar_wait(staged, 0) && ok;
#pragma unroll
            for (int k = 0; k < K::TMAX; ++k) {
                if (k < nT && (k & 1) == par) {
                    ok = mbar_wait(staged + k, 0) && ok;
                    ok = mbar_wait(staged + k + 1, 0) && ok;          // chunk k + 1 holds tile k's upper halo
                    fence_after_sync();
                    const uint64_t ak = a0 + (uint64_t)(k * 128 + K::HALO);
                    const uint32_t dcol = tmem_base + (uint32_t)(k * K::N);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int kc = 0; kc < C / 8; ++kc) {
                            const int shift = (tap / 3 - 1) * K::WP + (tap % 3 - 1);
                            const uint64_t ad = ak + (uint64_t)(long long)(shift + 2 * kc * (K::PLANE >> 4));
                            const uint64_t bd = b0 + (uint64_t)(tap * (K::BTAP >> 4) + 2 * kc * K::N);
                            mma_tf32(dcol, ad, bd, idesc, (tap | kc) != 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(mdone + k);
                    LC_PSTAMP(30 + k);
                }
            }
        }
    } else if (warp < 8) {
        // ---------------------------------------------------------------- transform: ring slot -> prologue -> TF32 -> planar tile
        const int grp = tid / K::GTHREADS, gt = tid - grp * K::GTHREADS;
        const int j = gt % K::CH, rr0 = gt / K::CH;                   // this thread's 16-byte channel chunk; first row of a pass
        constexpr int RPP = K::GTHREADS / K::CH;                      // rows per pass
        constexpr int NPASS = 128 / RPP;
        // prologue BatchNorm coefficients: only these 256 threads need them (lazy: reduced here from the producer's partial rows while the first chunks
        // fly), so the reduction synchronises on a named barrier and never holds up the loader / issuers
        {
            double* red = reinterpret_cast<double*>(s_red);
            if (lazy) {
                bn_partial_sums_load(a.pro_lazy.partial, a.pro_lazy.nparts, C, red);
                constexpr int CQ = C >> 1, NSL = 256 / CQ;
                tcp_bar256();
                double t = 0.0;
                if (tid < 2 * C) {
                    const int cqi = tid >> 2, kq = tid & 3;
                    for (int q = 0; q < NSL; ++q) t += red[(q * CQ + cqi) * 4 + kq];
                }
                tcp_bar256();
                if (tid < 2 * C) red[tid] = t;
                tcp_bar256();
                if (tid < C) {
                    float sc_, sh_, m_, is_; double var_;
                    bn_affine_from_sums(red[tid], red[C + tid], (double)a.pro_lazy.count, a.pro_lazy.gamma[tid], a.pro_lazy.beta[tid], a.pro_lazy.eps, &sc_, &sh_, &m_, &is_, &var_);
                    s_aff[tid] = sc_; s_aff[C + tid] = sh_;
                }
            } else if (tid < 2 * C) {
                s_aff[tid] = pro ? (tid < C ? a.pro_scale[tid] : a.pro_shift[tid - C]) : (tid < C ? 1.f : 0.f);
            }
            tcp_bar256();
        }
        if (tid == 0) LC_PSTAMP(2);
        const float4 sc = *reinterpret_cast<const float4*>(s_aff + j * 4), sh = *reinterpret_cast<const float4*>(s_aff + C + j * 4);
        for (int k = grp; k <= nT; k += K::NGROUP) {
            const int jin = k * nops, s_in = jin % K::NS;
            ok = mbar_wait(full + s_in, (uint32_t)((jin / K::NS) & 1)) && ok;
            const unsigned char* raw_in = sRing + (size_t)s_in * K::SLOT;
            const unsigned char* raw_res = raw_in;
            int s_res = 0;
            if (pres) {
                s_res = (jin + 1) % K::NS;
                ok = mbar_wait(full + s_res, (uint32_t)(((jin + 1) / K::NS) & 1)) && ok;
                raw_res = sRing + (size_t)s_res * K::SLOT;
            }
            if (gt == 0) LC_PSTAMP(10 + k);
            const int lo = s_lo[k];
#pragma unroll
            for (int i = 0; i < NPASS; ++i) {
                const int r = k * 128 + rr0 + i * RPP;
                if (r < R) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int px = s_rowsrc[r];
                    if (px >= 0) {
                        const size_t off = (size_t)(px - lo) * K::ROWB + j * 16;
                        v = *reinterpret_cast<const float4*>(raw_in + off);
                        if (RES && pres) {       // the previous block's output: relu(bn_b(y) + residual), same operation order as bn_act_fwd_kernel
                            const float4 y = *reinterpret_cast<const float4*>(raw_res + off);
                            v.x = fmaxf(fmaf(v.x, sc.x, sh.x) + y.x, 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y) + y.y, 0.f);
                            v.z = fmaxf(fmaf(v.z, sc.z, sh.z) + y.z, 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w) + y.w, 0.f);
                            if (r >= K::HALO && r < K::HALO + nT * 128)      // rows this CTA owns (halo rows belong to its neighbours)
                                *reinterpret_cast<float4*>(a.pro_out + (size_t)px * C + j * 4) = v;
                        } else if (pro) {
                            v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
                            v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
                            // MODE 2 without a residual operand: the producer's relu(bn(y)) is itself wanted in HBM (the stem's activation = block 0's residual)
                            if (RES && a.pro_out != nullptr && r >= K::HALO && r < K::HALO + nT * 128) *reinterpret_cast<float4*>(a.pro_out + (size_t)px * C + j * 4) = v;
                        }
                        v.x = to_tf32_fast(v.x); v.y = to_tf32_fast(v.y); v.z = to_tf32_fast(v.z); v.w = to_tf32_fast(v.w);
                    }
                    *reinterpret_cast<float4*>(sA + (size_t)j * K::PLANE + (size_t)r * 16) = v;
                }
            }
#ifndef LC_TCP_ELECTED_ARRIVE
            mbar_arrive(empty + s_in);                  // this thread is done reading the slot(s)
            if (pres) mbar_arrive(empty + s_res);
            fence_proxy_async();                        // generic-proxy writes of the planar tile -> visible to the tensor core
            mbar_arrive(staged + k);
#else
            fence_proxy_async();                        // generic-proxy writes of the planar tile -> visible to the tensor core
            __syncwarp();                               // every lane has read its ring rows and fenced its tile rows: lane 0 arrives for the warp
            if (lane == 0) {
                mbar_arrive(empty + s_in);
                if (pres) mbar_arrive(empty + s_res);
                mbar_arrive(staged + k);
            }
#endif
            if (gt == 0) LC_PSTAMP(20 + k);
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 8-15: TMEM lane quarter = warp % 4, tiles k = warp / 4 (mod 2))
        const int quarter = warp & 3, ew = warp - 8;
        const int c = lane % K::CPR;
        unsigned char* stg = smem_raw + K::OFF_EPI + (size_t)ew * K::EPI_WARP;
        auto stage = [&](int p, int cc) -> float4* {      // row p, chunk cc of a 32-row block; 64-byte sub-rows XOR-swizzled against bank conflicts
            return reinterpret_cast<float4*>(stg + (size_t)(cc >> 2) * 2048 + p * 64 + (((cc & 3) ^ ((p >> 1) & 3)) << 4));
        };
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        for (int k = ew >> 2; k < nT; k += 2) {
            const int px = s_rowsrc[K::HALO + k * 128 + quarter * 32 + lane];  // accumulator row == TMEM lane
            const bool valid = px >= 0;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const int nvalid = __popc(vmask), rank = __popc(vmask & ((1u << lane) - 1u));
            const int px_first = __shfl_sync(0xffffffffu, px, vmask ? __ffs(vmask) - 1 : 0);
            ok = mbar_wait(mdone + k, 0) && ok;
            fence_after_sync();
            if (quarter == 0 && lane == 0) LC_PSTAMP(40 + k);
            {
                float v[K::N];
                if (K::N == 16) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(k * K::N), v);
                else tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(k * K::N), v);
                if (valid) {
#pragma unroll
                    for (int cc = 0; cc < K::CPR; ++cc) *stage(rank, cc) = make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int kk = 0; kk < K::CPR; ++kk) {
                const int p = kk * (32 / K::CPR) + lane / K::CPR;
                if (p < nvalid) {
                    float4 x = *stage(p, c);
                    const size_t g = (size_t)(px_first + p) * K::N + c * 4;
                    if (a.addend != nullptr) {
                        const float4 a4 = *reinterpret_cast<const float4*>(a.addend + g);
                        x.x += a4.x; x.y += a4.y; x.z += a4.z; x.w += a4.w;
                    }
                    *reinterpret_cast<float4*>(a.out + g) = x;
                    s1.x += x.x; s1.y += x.y; s1.z += x.z; s1.w += x.w;
                    s2.x = fmaf(x.x, x.x, s2.x); s2.y = fmaf(x.y, x.y, s2.y); s2.z = fmaf(x.z, x.z, s2.z); s2.w = fmaf(x.w, x.w, s2.w);
                }
            }
            __syncwarp();      // the next tile's rows overwrite this warp's staging block
            if (quarter == 0 && lane == 0) LC_PSTAMP(50 + k);
        }
        if (stats) {           // lanes sharing a channel group differ in the bits above log2(CPR): fixed-order butterfly, then one row per quarter
#pragma unroll
            for (int off = K::CPR; off < 32; off <<= 1) {
                s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
                s1.z += __shfl_xor_sync(0xffffffffu, s1.z, off); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, off);
                s2.x += __shfl_xor_sync(0xffffffffu, s2.x, off); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, off);
                s2.z += __shfl_xor_sync(0xffffffffu, s2.z, off); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, off);
            }
            if (lane < K::CPR) {
                float* sp = s_part + (size_t)ew * 2 * K::N + c * 4;
                *reinterpret_cast<float4*>(sp) = s1; *reinterpret_cast<float4*>(sp + K::N) = s2;
            }
        }
    }
    if (!ok && lane == 0 && a.error_flag != nullptr) atomicExch(a.error_flag, 1);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, K::TMEM_COLS);
    if (tid == 0) LC_PSTAMP(60);

    if (stats) {
        if (tid < 2 * K::N) {
            const int stat = tid / K::N, ch = tid % K::N;
            float tsum = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) tsum += s_part[(size_t)(q * 2 + stat) * K::N + ch];
            a.stat.partial[((size_t)blockIdx.x * 2 + stat) * K::N + ch] = tsum;
        }
        if (a.stat.defer) return;       // consumers reduce the partial rows themselves (BnLazy)
        if (last_block_done(a.stat.counter, gridDim.x)) bn_finalize_last_block<K::N, 256>(a.stat, (int)gridDim.x, (double)a.B * W * W, s_red);
    }
}

template <int C, int W, int MODE>
static inline int conv_tcp_launch(const ConvTcArgs& a, int nsm, cudaStream_t st) {
    using K = ConvTcpCfg<C, W>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(conv3x3_tcp_kernel<C, W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES) != cudaSuccess) return LC_ERR_CUDA;
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)conv_tcp_grid(a.B, C, W, nsm)); cfg.blockDim = dim3(K::NTHREADS); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, conv3x3_tcp_kernel<C, W, MODE>, a) != cudaSuccess) return LC_ERR_CUDA;
    return lc_launch_status();
}

}  // namespace tc
}  // namespace lc
