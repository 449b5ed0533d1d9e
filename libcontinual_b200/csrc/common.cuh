// Shared device/host helpers for the libcontinual_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LC_OK 0
#define LC_ERR_INVALID (-22)   // -EINVAL : bad shape / unsupported configuration
#define LC_ERR_CUDA (-5)       // -EIO    : a CUDA runtime call failed

#define LC_CHECK_ARG(cond) \
    do {                   \
        if (!(cond)) return LC_ERR_INVALID; \
    } while (0)

static inline int lc_launch_status() { return cudaGetLastError() == cudaSuccess ? LC_OK : LC_ERR_CUDA; }

namespace lc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Column sums of a [32 lanes][16 values] register tile in 16 shuffles (instead of 16 x 5): each butterfly stage halves the number of live columns
// per lane (keep one half, send the other to the partner).  On return every lane holds in v[0] the sum over the 32 lanes of column (lane >> 1)
// (lanes 2k and 2k+1 both hold column k).  Fixed tree order: deterministic.
__device__ __forceinline__ float warp_colsum16(float (&v)[16]) {
    const unsigned lane = threadIdx.x & 31u;
    {
        const bool up = (lane & 16u) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float keep = up ? v[i + 8] : v[i], send = up ? v[i] : v[i + 8];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = (lane & 8u) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = (lane & 4u) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = up ? v[i + 2] : v[i], send = up ? v[i] : v[i + 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = (lane & 2u) != 0;
        const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// "last block done" election (threadFenceReduction pattern).  `counter` must be zero before the first launch and
// resets itself (atomicInc wraps), so the same counter can be reused by consecutive launches on one stream.
// Call from all threads after the block's partial results have been written to global memory.
__device__ __forceinline__ bool last_block_done(unsigned int* counter, unsigned int nblocks) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        unsigned int prev = atomicInc(counter, nblocks - 1);
        s_last = (prev == nblocks - 1);
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// L2-only load: for data produced by other blocks of the same launch
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }

}  // namespace lc
