// Input pipeline and evaluation meters on the device.
//
// The reference decodes with PIL and runs torchvision transforms per sample on 24 CPU workers, then copies an fp32 batch over an unpinned,
// synchronous H2D (core/data/dataset.py:248-266, core/data/data.py:11-35, core/data/dataloader.py:17-38).  Here the uint8 dataset is resident in HBM
// (CIFAR-100 = 150 MB, ImageNet-R at 224^2 = 4.5 GB: nothing next to 180 GB) and one launch per batch gathers the samples, applies the transform
// chain with the SAME integer / float arithmetic PIL and torchvision use, and writes the fp32 NCHW batch the step consumes:
//   augment_cifar_kernel : RandomCrop(32, padding=4) -> RandomHorizontalFlip -> ColorJitter(brightness) -> ToTensor -> Normalize   (data.py:11-16)
//                          (identity draws = the test transform, data.py:18)
//   resize_*_kernel      : RandomResizedCrop(224) / Resize(+CenterCrop) -> flip -> ToTensor -> Normalize                          (data.py:27-36)
//                          PIL Resample.c: separable triangle filter, support max(1, scale), coefficients normalised per output pixel in double and
//                          rounded to 22-bit fixed point, horizontal pass then vertical pass, uint8 after each — reproduced bit for bit.
// The random draws themselves come from the host (libcontinual_b200/data.py restates torchvision's get_params) as one small int32 / fp32 table per batch.
//   eval_meter_kernel    : per-task #correct / #seen of `Trainer._validate` (trainer.py:616-720) accumulated on the device with integer atomics.
#pragma once
#include "common.cuh"

namespace lc {

struct AugCifarArgs {
    const unsigned char* src;     // [N][H][W][3] uint8
    const long long* idx;         // [B] sample indices into src (nullable: 0..B-1)
    const int* draw;              // [B][4]: dx, dy (crop window origin inside the zero-padded image; pad, pad = centred), flip, unused
    const float* bright;          // [B] brightness factor (nullable: 1.0)
    float* out;                   // [B][3][H][W] fp32
    float mean[3], std[3];
    int B, H, W, pad;
};

__global__ void __launch_bounds__(256) augment_cifar_kernel(AugCifarArgs a) {
    const int hw = a.H * a.W;
    const long long total = (long long)a.B * hw;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / hw), p = (int)(e - (long long)b * hw);
        const int y = p / a.W, x = p - y * a.W;
        const int dx = a.draw[b * 4], dy = a.draw[b * 4 + 1], flip = a.draw[b * 4 + 2];
        const int xs = (flip ? a.W - 1 - x : x) + dx - a.pad, ys = y + dy - a.pad;      // flip acts on the cropped image
        const long long n = a.idx != nullptr ? a.idx[b] : b;
        int v[3] = {0, 0, 0};
        if (xs >= 0 && xs < a.W && ys >= 0 && ys < a.H) {
            const unsigned char* s = a.src + ((n * a.H + ys) * a.W + xs) * 3;
            v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
        }
        const float f = a.bright != nullptr ? a.bright[b] : 1.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int u = v[c];
            if (f != 1.f) {      // PIL ImageEnhance.Brightness = Image.blend(black, img, f): (UINT8)(f * v) with clipping
                const float t = __fmul_rn(f, (float)u);
                u = t >= 255.f ? 255 : (t <= 0.f ? 0 : (int)t);
            }
            const float t = __fdiv_rn((float)u, 255.f);                                        // ToTensor
            a.out[((long long)b * 3 + c) * hw + p] = __fdiv_rn(__fsub_rn(t, a.mean[c]), a.std[c]);   // Normalize
        }
    }
}

// ---- PIL-exact bilinear resize ------------------------------------------------------------------------------------------------------------------
constexpr int kResizeKMax = 24;          // taps per output pixel: supports down-scaling by up to ~11x
constexpr int kPrecisionBits = 32 - 8 - 2;

// draw[b][8] = top, left, h, w (crop box), OH, OW (resized size), oy, ox (origin of the out_h x out_w output window inside the resized image)
struct ResizeArgs {
    const unsigned char* src;     // [N][H][W][3]
    const long long* idx;         // [B] (nullable)
    const int* draw;              // [B][8]
    const int* flip;              // [B] (nullable)
    int* coef;                    // [B][2][OUT][kResizeKMax] fixed-point coefficients; axis 0 = x, 1 = y
    int* cmin;                    // [B][2][OUT][2]: first input index, tap count
    unsigned char* tmp;           // [B][H][OUT][3] horizontal-pass result (rows of the crop box)
    float* out;                   // [B][3][OUT][OUT]
    float mean[3], std[3];
    int B, H, W, OUT;
};

// one thread per (sample, axis, output index inside the window): PIL precompute_coeffs + normalize_coeffs_8bpc in double, no fused multiply-add
__global__ void resize_coeffs_kernel(ResizeArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.B * 2 * a.OUT) return;
    const int o = t % a.OUT, axis = (t / a.OUT) % 2, b = t / (2 * a.OUT);
    const int* d = a.draw + b * 8;
    const int in_size = axis == 0 ? d[3] : d[2], out_size = axis == 0 ? d[5] : d[4];
    const int xx = o + (axis == 0 ? d[7] : d[6]);
    int* k = a.coef + (size_t)t * kResizeKMax;
    int xmin = 0, cnt = 0;
    if (xx < out_size) {
        const double scale = __ddiv_rn((double)in_size, (double)out_size);
        const double fs = scale < 1.0 ? 1.0 : scale;
        const double support = fs;                                 // bilinear support 1.0 * filterscale
        const double ss = __ddiv_rn(1.0, fs);
        const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
        xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
        if (xmax > in_size) xmax = in_size;
        cnt = xmax - xmin;
        if (cnt > kResizeKMax) cnt = kResizeKMax;
        double w[kResizeKMax];
        double ww = 0.0;
        for (int x = 0; x < cnt; ++x) {
            double v = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (v < 0.0) v = -v;
            v = v < 1.0 ? __dsub_rn(1.0, v) : 0.0;
            w[x] = v;
            ww = __dadd_rn(ww, v);
        }
        for (int x = 0; x < cnt; ++x) {
            const double kk = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
            k[x] = (int)__dadd_rn(0.5, __dmul_rn(kk, (double)(1 << kPrecisionBits)));
        }
    }
    a.cmin[(size_t)t * 2] = xmin;
    a.cmin[(size_t)t * 2 + 1] = cnt;
}

__device__ __forceinline__ int clip8_fixed(long long ss) {
    const long long v = ss >> kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : (int)v);
}

// horizontal pass: tmp[b][r][o][c] for every row r of the crop box
__global__ void __launch_bounds__(256) resize_h_kernel(ResizeArgs a) {
    const long long total = (long long)a.B * a.H * a.OUT;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(e % a.OUT), r = (int)((e / a.OUT) % a.H), b = (int)(e / ((long long)a.OUT * a.H));
        const int* d = a.draw + b * 8;
        if (r >= d[2]) continue;
        const size_t t = ((size_t)b * 2 + 0) * a.OUT + o;
        const int xmin = a.cmin[t * 2], cnt = a.cmin[t * 2 + 1];
        const int* k = a.coef + t * kResizeKMax;
        const long long n = a.idx != nullptr ? a.idx[b] : b;
        const unsigned char* s = a.src + ((n * a.H + d[0] + r) * a.W + d[1] + xmin) * 3;
        unsigned char* dst = a.tmp + (((size_t)b * a.H + r) * a.OUT + o) * 3;
        if (d[3] == d[5]) {          // no horizontal resampling (PIL skips the pass)
            dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2];
            continue;
        }
        long long s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < cnt; ++x) {
            const long long kk = k[x];
            s0 += kk * s[x * 3]; s1 += kk * s[x * 3 + 1]; s2 += kk * s[x * 3 + 2];
        }
        dst[0] = (unsigned char)clip8_fixed(s0); dst[1] = (unsigned char)clip8_fixed(s1); dst[2] = (unsigned char)clip8_fixed(s2);
    }
}

// vertical pass + flip + ToTensor + Normalize
__global__ void __launch_bounds__(256) resize_v_kernel(ResizeArgs a) {
    const int oo = a.OUT * a.OUT;
    const long long total = (long long)a.B * oo;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / oo), p = (int)(e - (long long)b * oo);
        const int y = p / a.OUT, x = p - y * a.OUT;
        const int* d = a.draw + b * 8;
        const int xs = (a.flip != nullptr && a.flip[b]) ? a.OUT - 1 - x : x;
        const size_t t = ((size_t)b * 2 + 1) * a.OUT + y;
        const int ymin = a.cmin[t * 2], cnt = a.cmin[t * 2 + 1];
        const int* k = a.coef + t * kResizeKMax;
        const unsigned char* s = a.tmp + (((size_t)b * a.H + ymin) * a.OUT + xs) * 3;
        int v[3];
        if (d[2] == d[4]) {
            v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
        } else {
            long long s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
            for (int r = 0; r < cnt; ++r) {
                const long long kk = k[r];
                const unsigned char* q = s + (size_t)r * a.OUT * 3;
                s0 += kk * q[0]; s1 += kk * q[1]; s2 += kk * q[2];
            }
            v[0] = clip8_fixed(s0); v[1] = clip8_fixed(s1); v[2] = clip8_fixed(s2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float tt = __fdiv_rn((float)v[c], 255.f);
            a.out[((long long)b * 3 + c) * oo + p] = __fdiv_rn(__fsub_rn(tt, a.mean[c]), a.std[c]);
        }
    }
}

// ---- evaluation meter ------------------------------------------------------------------------------------------------------------------------------
// counts[t][0] += #correct, counts[t][1] += #seen for the task t whose class range [bounds[t], bounds[t+1]) holds the label; with ntask == 1 and a fixed
// `task` every sample goes to that row (per-task loaders).  Integer atomics: the result does not depend on the order.
__global__ void __launch_bounds__(256) eval_meter_kernel(const long long* pred, const long long* label, int n, const int* bounds, int ntask, int task,
                                                          unsigned long long* counts) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long y = label[i];
        int t = task;
        if (t < 0) {
            t = -1;
            for (int k = 0; k < ntask; ++k)
                if (y >= bounds[k] && y < bounds[k + 1]) { t = k; break; }
            if (t < 0) continue;
        }
        atomicAdd(counts + 2 * t + 1, 1ull);
        if (pred[i] == y) atomicAdd(counts + 2 * t, 1ull);
    }
}

// Fold one batch's counts into the totals the way `Trainer._validate` does in its per-task mode: `correct_task += int(acc * batch_size)` with
// acc = correct / batch_size as a Python float (trainer.py:644): the double round trip loses one sample now and then (29 / 100 * 100 = 28.999...), and
// the reported accuracy is defined by it.  batch[t] = {#correct, #seen} is consumed (zeroed).
__global__ void eval_fold_kernel(unsigned long long* batch, unsigned long long* total, int ntask, int reference_rounding) {
    const int t = threadIdx.x;
    if (t >= ntask) return;
    const unsigned long long c = batch[2 * t], n = batch[2 * t + 1];
    if (n > 0) {
        unsigned long long add = c;
        if (reference_rounding) add = (unsigned long long)(long long)(__dmul_rn(__ddiv_rn((double)c, (double)n), (double)n));
        total[2 * t] += add;
        total[2 * t + 1] += n;
    }
    batch[2 * t] = 0; batch[2 * t + 1] = 0;
}

}  // namespace lc
