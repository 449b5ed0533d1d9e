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
