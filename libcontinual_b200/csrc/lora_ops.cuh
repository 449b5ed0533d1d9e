// Low-rank adapter kernels for the ViT attention blocks (core/model/backbone/transformer.py:199-274 MultiHeadAttention_LoRA,
// :276-357 MultiHeadAttention_SDLoRA; core/model/InfLoRA_opt.py:175-189).
//
// The reference adds the adapters on the WEIGHT side every forward (`k_weight + lora_B_k.weight @ lora_A_k.weight`, :246-254) and lets
// autograd push a dense 768 x 768 weight gradient through `B @ A`.  Here
//   * lora_merge_kernel      builds W' = W + B diag(s) A for the adapted slabs of the fused QKV weight once per step, straight into the
//                            two BF16 operand copies the tcgen05 GEMMs read ([3D][D] for the forward, [D][3D] for the data gradient);
//   * rowouter_partial/reduce computes the adapter gradient in its rank-r form, dB[c][j] = sum_n dY[n][c] * (h A^T)[n][j], directly from
//                            the token gradient dY of the slab and the saved down-projection (h A^T) — 2 n D r FLOP instead of the
//                            reference's 2 n D^2 (+ 2 D^2 r), a deterministic two-stage sum over token chunks.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace lc {

constexpr int kLoraMaxR = 128;     // adapters stacked along the rank axis (SD-LoRA keeps one per task)

struct LoraMergeArgs {
    const float* w;          // [L][3D][D] fp32 master QKV weights
    const float* A;          // [L][ns][R][D]
    const float* B;          // [L][ns][D][R]
    const float* scale;      // nullable [L][ns][R]
    __nv_bfloat16* wb;       // nullable [L][3D][D]
    __nv_bfloat16* wbt;      // nullable [L][D][3D]
    float* w_out;            // nullable [L][3D][D]: fp32 merged weights (merge_weight(), transformer.py:237-242); may alias w
    int slab[3];             // adapted slabs in increasing order (0 = q, 1 = k, 2 = v)
    int ns, D, R;
};

// grid (D/32, D/32, L*ns), block (32, 8): one 32 x 32 tile of one slab of one layer
__global__ void __launch_bounds__(256) lora_merge_kernel(LoraMergeArgs a) {
    __shared__ float sA[kLoraMaxR][33];
    __shared__ float sB[32][kLoraMaxR + 1];
    __shared__ float sT[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int layer = blockIdx.z / a.ns, si = blockIdx.z % a.ns;
    const int D = a.D, R = a.R;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const size_t ls = (size_t)layer * a.ns + si;
    const float* A = a.A + ls * R * D;
    const float* Bm = a.B + ls * D * R;
    const float* sc = a.scale != nullptr ? a.scale + ls * R : nullptr;
    for (int i = tid; i < R * 32; i += 256) { const int j = i >> 5, c = i & 31; sA[j][c] = __ldg(A + (size_t)j * D + c0 + c); }
    for (int i = tid; i < 32 * R; i += 256) { const int r = i / R, j = i % R; sB[r][j] = __ldg(Bm + (size_t)(r0 + r) * R + j) * (sc != nullptr ? __ldg(sc + j) : 1.f); }
    __syncthreads();
    const size_t wbase = ((size_t)layer * 3 + a.slab[si]) * D * D;          // slab rows [slab*D, slab*D + D) of layer
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = ty + 8 * k;
        float acc = 0.f;
        for (int j = 0; j < R; ++j) acc = fmaf(sB[r][j], sA[j][tx], acc);
        const size_t idx = wbase + (size_t)(r0 + r) * D + c0 + tx;
        const float v = a.w[idx] + acc;
        if (a.w_out != nullptr) a.w_out[idx] = v;
        if (a.wb != nullptr) a.wb[idx] = __float2bfloat16_rn(v);
        sT[r][tx] = v;
    }
    if (a.wbt != nullptr) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = ty + 8 * k;                                        // transposed tile: row = input feature, tx runs along the output feature
            a.wbt[((size_t)layer * D + c0 + c) * (3 * (size_t)D) + (size_t)a.slab[si] * D + r0 + tx] = __float2bfloat16_rn(sT[tx][c]);
        }
    }
}

// out[s][c][j] = sum_n X[n][x0 + s*xs + c] * Z[n][z0 + s*zs + j]   (X bf16, Z fp32), D a multiple of 768 columns handled 4 per thread.
// grid (nslab * D/768, nchunk), 192 threads; each CTA sums `rows_per` token rows into partial[chunk][s][c][j].
template <int R>
__global__ void __launch_bounds__(192) rowouter_partial_kernel(const __nv_bfloat16* X, long long ldx, int x0, int xs, int D, const float* Z, int ldz, int z0, int zs, int r_real,
                                                               long long n, int rows_per, float* partial) {
    constexpr int TILE = 32;
    __shared__ float sZ[TILE][R];
    const int cpb = D / 768;
    const int s = blockIdx.x / cpb, col = (blockIdx.x % cpb) * 768 + threadIdx.x * 4;
    const long long n0 = (long long)blockIdx.y * rows_per;
    const long long n1 = n0 + rows_per < n ? n0 + rows_per : n;
    float acc[4][R];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
    const __nv_bfloat16* xp = X + x0 + (size_t)s * xs + col;
    for (long long t0 = n0; t0 < n1; t0 += TILE) {
        const int nt = (int)(n1 - t0 < TILE ? n1 - t0 : TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < TILE * R; i += 192) {
            const int r = i / R, j = i % R;
            sZ[r][j] = (r < nt && j < r_real) ? __ldg(Z + (size_t)(t0 + r) * ldz + z0 + s * zs + j) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int rb = 0; rb < nt; rb += 16) {                                // 16 rows (128 B per thread) in flight: the kernel is HBM-latency bound
            uint2 xv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                xv[u] = rb + u < nt ? __ldg(reinterpret_cast<const uint2*>(xp + (size_t)(t0 + rb + u) * ldx)) : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const float x0f = __uint_as_float(xv[u].x << 16), x1f = __uint_as_float(xv[u].x & 0xffff0000u);
                const float x2f = __uint_as_float(xv[u].y << 16), x3f = __uint_as_float(xv[u].y & 0xffff0000u);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float z = sZ[rb + u][j];                          // rows past nt hold zeros
                    acc[0][j] = fmaf(x0f, z, acc[0][j]); acc[1][j] = fmaf(x1f, z, acc[1][j]);
                    acc[2][j] = fmaf(x2f, z, acc[2][j]); acc[3][j] = fmaf(x3f, z, acc[3][j]);
                }
            }
        }
    }
    float* o = partial + (((size_t)blockIdx.y * (gridDim.x / cpb) + s) * D + col) * r_real;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (j < r_real) o[i * r_real + j] = acc[i][j];
}

// out = scale * sum_chunk partial[chunk] in a fixed order; `transposed`: partial is [s][c][j], out is written as [s][j][c] (a lora_A gradient);
// scale_ptr (nullable) is a device scalar (SD-LoRA's trainable magnitude of the current adapter)
__global__ void __launch_bounds__(256) rowouter_reduce_kernel(const float* partial, long long per_chunk, int nchunk, float* out, int D, int r, int transposed,
                                                              const float* scale_ptr) {
    const float sc = scale_ptr != nullptr ? __ldg(scale_ptr) : 1.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < per_chunk; i += (long long)gridDim.x * 256) {
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += __ldcg(partial + (size_t)c * per_chunk + i);
        long long o = i;
        if (transposed) {
            const long long slab = i / ((long long)D * r), rem = i % ((long long)D * r);
            const int col = (int)(rem / r), j = (int)(rem % r);
            o = slab * D * r + (long long)j * D + col;
        }
        out[o] = s * sc;
    }
}

// colsum partials of the elementwise product of two fp32 matrices [n][ld] over `cols` columns: partial[chunk][j] = sum_{rows of chunk} G[n][j] * Z[n][zoff + j]
// grid (nchunk), 256 threads = 8 warps striding rows; cols <= 128
__global__ void __launch_bounds__(256) coldot_partial_kernel(const float* G, int ldg, const float* Z, int ldz, int zoff, int cols, long long n, int rows_per,
                                                             float* partial) {
    __shared__ float s_acc[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = (long long)blockIdx.x * rows_per, n1 = n0 + rows_per < n ? n0 + rows_per : n;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = n0 + warp; r < n1; r += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            if (j < cols) acc[k] = fmaf(__ldg(G + (size_t)r * ldg + j), __ldg(Z + (size_t)r * ldz + zoff + j), acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) s_acc[warp][lane + 32 * k] = acc[k];
    __syncthreads();
    if (threadIdx.x < cols) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += s_acc[w][threadIdx.x];
        partial[(size_t)blockIdx.x * cols + threadIdx.x] = s;
    }
}
// dmag[i] += w[i*r + jj]-weighted sum: dmag[i] += sum_{jj < r} colw[i*r + jj] * sum_chunk partial[chunk][i*r + jj]   (single block; sequential launches on
// one stream accumulate in a fixed order)
__global__ void __launch_bounds__(128) coldot_finish_kernel(const float* partial, int nchunk, int cols, int r, const float* colw, float* dmag) {
    __shared__ float s_col[128];
    if (threadIdx.x < cols) {
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += partial[(size_t)c * cols + threadIdx.x];
        s_col[threadIdx.x] = s * (colw != nullptr ? colw[threadIdx.x] : 1.f);
    }
    __syncthreads();
    if (threadIdx.x < cols / r) {
        float s = 0.f;
        for (int jj = 0; jj < r; ++jj) s += s_col[threadIdx.x * r + jj];
        dmag[threadIdx.x] += s;
    }
}

}  // namespace lc
