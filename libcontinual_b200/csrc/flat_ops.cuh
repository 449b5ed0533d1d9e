// Flat-arena (multi-tensor) HBM kernels over the parameter / gradient / state arenas.
//   ewc_penalty_grad : compute_ewc + its autograd      (core/model/ewc.py:207-225, used at :100)      5 x 4 B / element
//   fisher_accumulate: fisher[n] += grad^2 * len(y)    (ewc.py:173)                                   3 x 4 B / element
//   fisher_merge     : /num_samples and alpha-EMA      (ewc.py:202-204, 129-131)
//   sgd_momentum     : torch.optim.SGD(momentum, weight_decay) step (trainer.py:606)                  5 x 4 B / element
//   adam             : torch.optim.Adam step                                                          7 x 4 B / element
// Hyper-parameters that change between steps (lr, ...) are read from device memory so that a captured CUDA graph can be
// replayed across scheduler steps.
#pragma once
#include "common.cuh"

namespace lc {

constexpr int kFlatBlocks = kNumSMs * 2;

// grad += lamda * F * (theta - theta_ref);   scal[0] += lamda * sum(F * (theta-theta_ref)^2) / 2 ;  scal[4] = penalty
__global__ void __launch_bounds__(256) ewc_penalty_grad_kernel(const float* theta, const float* theta_ref, const float* fisher, float* grad,
                                                                long long n, const float* hp_lamda, double* partial, unsigned int* counter,
                                                                float* scal) {
    __shared__ double s_red[256];
    const float lam = *hp_lamda;
    double acc = 0.0;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 t = ldg4(theta + i * 4), r = ldg4(theta_ref + i * 4), f = ldg4(fisher + i * 4);
        float4 g = *reinterpret_cast<const float4*>(grad + i * 4);
        const float dx = t.x - r.x, dy = t.y - r.y, dz = t.z - r.z, dw = t.w - r.w;
        g.x = fmaf(lam * f.x, dx, g.x); g.y = fmaf(lam * f.y, dy, g.y); g.z = fmaf(lam * f.z, dz, g.z); g.w = fmaf(lam * f.w, dw, g.w);
        *reinterpret_cast<float4*>(grad + i * 4) = g;
        acc += (double)(f.x * dx * dx) + (double)(f.y * dy * dy) + (double)(f.z * dz * dz) + (double)(f.w * dw * dw);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = n4 * 4; i < n; ++i) {
            const float d = theta[i] - theta_ref[i];
            grad[i] = fmaf(lam * fisher[i], d, grad[i]);
            acc += (double)(fisher[i] * d * d);
        }
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s_red[0];
    if (last_block_done(counter, gridDim.x)) {
        // the per-block partials in parallel (one L2 round trip), then a fixed-order tree: a serial loop over ~300 dependent loads was 2/3 of this kernel
        double t = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(partial + b);
        s_red[threadIdx.x] = t;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const float pen = (float)(0.5 * s_red[0]);
            scal[4] = pen;
            scal[0] += lam * pen;
        }
    }
}

__global__ void __launch_bounds__(256) fisher_accumulate_kernel(float* fisher, const float* grad, long long n, float weight) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = grad[i];
        fisher[i] = fmaf(g * g, weight, fisher[i]);       // fisher += grad.pow(2) * len(y)
    }
}

// f_new = f_new * inv_n ; if f_old != null: f_new = alpha * f_old + (1 - alpha) * f_new
__global__ void __launch_bounds__(256) fisher_merge_kernel(float* f_new, const float* f_old, long long n, float inv_n_is_div /*num_samples*/, float alpha) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = f_new[i] / inv_n_is_div;
        if (f_old != nullptr) v = alpha * f_old[i] + (1.f - alpha) * v;
        f_new[i] = v;
    }
}

// hp = {lr, momentum, weight_decay}.  Rounding mirrors torch.optim.SGD's op sequence:
//   g = fma(wd, p, g) ; m = round(mu*m) + g ; p = fma(-lr, m, p)
__global__ void __launch_bounds__(256) sgd_momentum_kernel(float* p, const float* g, float* m, long long n, const float* hp) {
    const float lr = hp[0], mu = hp[1], wd = hp[2];
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = *reinterpret_cast<const float4*>(p + i * 4);
        const float4 gv = ldg4(g + i * 4);
        float4 mv = *reinterpret_cast<const float4*>(m + i * 4);
        mv.x = __fadd_rn(__fmul_rn(mu, mv.x), fmaf(wd, pv.x, gv.x));
        mv.y = __fadd_rn(__fmul_rn(mu, mv.y), fmaf(wd, pv.y, gv.y));
        mv.z = __fadd_rn(__fmul_rn(mu, mv.z), fmaf(wd, pv.z, gv.z));
        mv.w = __fadd_rn(__fmul_rn(mu, mv.w), fmaf(wd, pv.w, gv.w));
        pv.x = fmaf(-lr, mv.x, pv.x); pv.y = fmaf(-lr, mv.y, pv.y); pv.z = fmaf(-lr, mv.z, pv.z); pv.w = fmaf(-lr, mv.w, pv.w);
        *reinterpret_cast<float4*>(m + i * 4) = mv;
        *reinterpret_cast<float4*>(p + i * 4) = pv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = n4 * 4; i < n; ++i) {
            const float mm = __fadd_rn(__fmul_rn(mu, m[i]), fmaf(wd, p[i], g[i]));
            m[i] = mm;
            p[i] = fmaf(-lr, mm, p[i]);
        }
    }
}

// same update, but elements in [freeze_lo, freeze_hi) belong to a param group with lr = 0 and weight_decay = 0 (LUCIR keeps the old
// classes' embedding fixed: lucir.py:229-240) and are left untouched
__global__ void __launch_bounds__(256) sgd_momentum_frozen_kernel(float* p, const float* g, float* m, long long n, const float* hp, long long freeze_lo,
                                                                   long long freeze_hi) {
    const float lr = hp[0], mu = hp[1], wd = hp[2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i >= freeze_lo && i < freeze_hi) continue;
        const float mm = __fadd_rn(__fmul_rn(mu, m[i]), fmaf(wd, p[i], g[i]));
        m[i] = mm;
        p[i] = fmaf(-lr, mm, p[i]);
    }
}

// hp = {lr, beta1, beta2, eps, weight_decay, bias_corr1 (1-b1^t), bias_corr2 (1-b2^t)}
__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* g, float* m, float* v, long long n, const float* hp) {
    const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], bc1 = hp[5], bc2 = hp[6];
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pv = p[i];
        const float gv = fmaf(wd, pv, g[i]);
        const float mv = fmaf(1.f - b1, gv - m[i], m[i]);          // lerp form used by torch
        const float vv = fmaf(1.f - b2, gv * gv, b2 * v[i]);
        m[i] = mv; v[i] = vv;
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        p[i] = pv - step_size * (mv / denom);
    }
}

// step counter + bias corrections kept on the device, so that a captured step needs no per-step host staging: hp[7] = t (as float), and the betas
// once more as DOUBLES in hp[8..11] (torch.optim.Adam forms `1 - beta ** step` from the Python doubles, not from their fp32 roundings)
//   t += 1 ; hp[5] = 1 - b1^t ; hp[6] = 1 - b2^t      (double pow, rounded to fp32 once)
__global__ void adam_tick_kernel(float* hp) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float t = hp[7] + 1.f;
        const double* betas = reinterpret_cast<const double*>(hp + 8);
        hp[7] = t;
        hp[5] = (float)(1.0 - pow(betas[0], (double)t));
        hp[6] = (float)(1.0 - pow(betas[1], (double)t));
    }
}

// clip_grad_norm_ (l2p.py:104): scale = min(1, max_norm / (||g|| + 1e-6)); two launches: norm partials, then scale
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* g, long long n, double* partial) {
    __shared__ double s_red[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = g[i];
        acc += (double)x * (double)x;
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s_red[0];
}
__global__ void __launch_bounds__(256) clip_scale_kernel(float* g, long long n, const double* partial, int nparts, float max_norm, float* norm_out) {
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < nparts; ++b) t += partial[b];
        const float nrm = (float)sqrt(t);
        const float coef = max_norm / (nrm + 1e-6f);
        s_scale = coef < 1.f ? coef : 1.f;
        if (blockIdx.x == 0 && norm_out != nullptr) *norm_out = nrm;
    }
    __syncthreads();
    const float sc = s_scale;
    if (sc == 1.f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) g[i] *= sc;
}

}  // namespace lc
