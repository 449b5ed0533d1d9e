// Network-level driver for the CIFAR ResNet backbone + the per-kernel C entry points.  Host code only launches kernels;
// all arithmetic lives in the .cuh kernels.  See include/lc_b200.h for the contract of every exported symbol.
#include "../../include/lc_b200.h"
#include "bn_elem.cuh"
#include "conv_aux.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_tcp.cuh"
#include "conv_s2_tc.cuh"
#include "wgrad_tc.cuh"
#include "flat_ops.cuh"
#include "head_loss.cuh"

#include <algorithm>
#include <cstring>
#include <cstdio>
#include <new>
#include <vector>

using namespace lc;

namespace {

constexpr float kBnEps = 1e-5f;       // nn.BatchNorm2d defaults (resnet.py:296)
constexpr float kBnMomentum = 0.1f;
constexpr int kBnBwdBlocks = 592;

// ---- kernel configurations for the CifarResNet layer shapes ------------------------------------------------------
//                      CIN COUT WO PT CT RS COUT_CTA CHUNK STRIDE DILATE NCHW
using CfgStem = Conv3x3Cfg<3, 16, 32, 4, 8, 2, 16, 3, 1, false, true>;
using CfgS1 = Conv3x3Cfg<16, 16, 32, 4, 8, 2, 16, 16, 1, false, false>;
using CfgS2F = Conv3x3Cfg<16, 32, 16, 2, 8, 1, 32, 16, 2, false, false>;
using CfgS2 = Conv3x3Cfg<32, 32, 16, 4, 8, 1, 32, 16, 1, false, false>;
using CfgS3F = Conv3x3Cfg<32, 64, 8, 2, 8, 1, 32, 8, 2, false, false>;
using CfgS3 = Conv3x3Cfg<64, 64, 8, 2, 8, 1, 32, 16, 1, false, false>;
using CfgD2 = Conv3x3Cfg<32, 16, 32, 4, 8, 2, 16, 8, 1, true, false>;   // dgrad of S2F
using CfgD3 = Conv3x3Cfg<64, 32, 16, 4, 8, 1, 32, 8, 1, true, false>;   // dgrad of S3F

#define LC_CONV_KERNEL(CFG, ...) conv3x3_kernel<__VA_ARGS__>

int launch_conv3x3(int cin, int cout, int wo, int stride, bool dilate, bool nchw, const Conv3x3Args& a, cudaStream_t st) {
    if (nchw && cin == 3 && cout == 16 && wo == 32 && stride == 1 && !dilate)
        return conv_launch<CfgStem>(conv3x3_kernel<3, 16, 32, 4, 8, 2, 16, 3, 1, false, true>, a, st);
    if (nchw) return LC_ERR_INVALID;
    if (!dilate) {
        if (cin == 16 && cout == 16 && wo == 32 && stride == 1) return conv_launch<CfgS1>(conv3x3_kernel<16, 16, 32, 4, 8, 2, 16, 16, 1, false, false>, a, st);
        if (cin == 16 && cout == 32 && wo == 16 && stride == 2) return conv_launch<CfgS2F>(conv3x3_kernel<16, 32, 16, 2, 8, 1, 32, 16, 2, false, false>, a, st);
        if (cin == 32 && cout == 32 && wo == 16 && stride == 1) return conv_launch<CfgS2>(conv3x3_kernel<32, 32, 16, 4, 8, 1, 32, 16, 1, false, false>, a, st);
        if (cin == 32 && cout == 64 && wo == 8 && stride == 2) return conv_launch<CfgS3F>(conv3x3_kernel<32, 64, 8, 2, 8, 1, 32, 8, 2, false, false>, a, st);
        if (cin == 64 && cout == 64 && wo == 8 && stride == 1) return conv_launch<CfgS3>(conv3x3_kernel<64, 64, 8, 2, 8, 1, 32, 16, 1, false, false>, a, st);
    } else {
        if (cin == 32 && cout == 16 && wo == 32) return conv_launch<CfgD2>(conv3x3_kernel<32, 16, 32, 4, 8, 2, 16, 8, 1, true, false>, a, st);
        if (cin == 64 && cout == 32 && wo == 16) return conv_launch<CfgD3>(conv3x3_kernel<64, 32, 16, 4, 8, 1, 32, 8, 1, true, false>, a, st);
    }
    return LC_ERR_INVALID;
}

//                     CIN COUT WO STRIDE CI_CTA KS TILE_H NCHW
using WCfgStem = WgradCfg<3, 16, 32, 1, 4, 16, 16, true>;
using WCfgS1 = WgradCfg<16, 16, 32, 1, 16, 4, 8, false>;
using WCfgS2F = WgradCfg<16, 32, 16, 2, 16, 2, 8, false>;
using WCfgS2 = WgradCfg<32, 32, 16, 1, 32, 1, 8, false>;
using WCfgS3F = WgradCfg<32, 64, 8, 2, 16, 1, 8, false>;
using WCfgS3 = WgradCfg<64, 64, 8, 1, 16, 1, 8, false>;

int launch_wgrad3x3(int cin, int cout, int wo, int stride, bool nchw, const WgradArgs& a, cudaStream_t st) {
    if (nchw && cin == 3 && cout == 16 && wo == 32 && stride == 1) return wgrad_launch<WCfgStem>(wgrad3x3_kernel<3, 16, 32, 1, 4, 16, 16, true>, a, st);
    if (nchw) return LC_ERR_INVALID;
    if (cin == 16 && cout == 16 && wo == 32 && stride == 1) return wgrad_launch<WCfgS1>(wgrad3x3_kernel<16, 16, 32, 1, 16, 4, 8, false>, a, st);
    if (cin == 16 && cout == 32 && wo == 16 && stride == 2) return wgrad_launch<WCfgS2F>(wgrad3x3_kernel<16, 32, 16, 2, 16, 2, 8, false>, a, st);
    if (cin == 32 && cout == 32 && wo == 16 && stride == 1) return wgrad_launch<WCfgS2>(wgrad3x3_kernel<32, 32, 16, 1, 32, 1, 8, false>, a, st);
    if (cin == 32 && cout == 64 && wo == 8 && stride == 2) return wgrad_launch<WCfgS3F>(wgrad3x3_kernel<32, 64, 8, 2, 16, 1, 8, false>, a, st);
    if (cin == 64 && cout == 64 && wo == 8 && stride == 1) return wgrad_launch<WCfgS3>(wgrad3x3_kernel<64, 64, 8, 1, 16, 1, 8, false>, a, st);
    return LC_ERR_INVALID;
}

bool tc_eligible(int cin, int cout, int wo, int stride, int ksize) {
    return ksize == 3 && stride == 1 && cin == cout && ((cin == 16 && wo == 32) || (cin == 32 && wo == 16) || (cin == 64 && wo == 8));
}
// LC_CONV_PERSIST=0 keeps every forward conv on the one-wave kernel (conv_tc.cuh); default: stages 1 and 2 run the persistent, warp-specialised
// kernel (conv_tcp.cuh), whose grid — and therefore the number of BatchNorm partial rows — is conv_tcp_grid()
int conv_persist_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LC_CONV_PERSIST"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
    return v;
}
int device_sms() {
    static int v = 0;
    if (v == 0) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148; }
    return v;
}
// stride-2 first conv of stages 2 / 3 on the tensor cores (forward only); LC_CONV_S2_TC=0 keeps them on the CUDA-core kernel
bool s2_tc_eligible(int cin, int cout, int wo, int stride, int ksize) {
    return ksize == 3 && stride == 2 && cout == 2 * cin && ((cin == 16 && wo == 16) || (cin == 32 && wo == 8));
}
int conv_s2_tc_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LC_CONV_S2_TC"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
    return v;
}
int launch_conv3x3s2_tc(int cin, int wo, const tc::ConvS2Args& a, cudaStream_t st) {
    if (cin == 16 && wo == 16) return tc::conv_s2_tc_launch<16, 16>(a, st);
    if (cin == 32 && wo == 8) return tc::conv_s2_tc_launch<32, 8>(a, st);
    return LC_ERR_INVALID;
}
int launch_dgrad3x3s2_tc(int cin, int wo, const tc::DgradS2Args& a, cudaStream_t st) {
    if (cin == 16 && wo == 16) return tc::dgrad_s2_tc_launch<16, 16>(a, st);
    if (cin == 32 && wo == 8) return tc::dgrad_s2_tc_launch<32, 8>(a, st);
    return LC_ERR_INVALID;
}
bool tcp_eligible(int c, int wo) { return (c == 16 && wo == 32) || (c == 32 && wo == 16); }
// partial rows a forward tensor-core conv of this shape writes (= its grid size)
int conv_tc_fwd_parts(long long batch, int c, int wo, bool persist) {
    if (persist && tcp_eligible(c, wo)) return tc::conv_tcp_grid(batch, c, wo, device_sms());
    const int mrows = 128 * (c == 16 ? 4 : (c == 32 ? 2 : 1));
    return (int)((batch * (wo + 2) * (wo + 2) + mrows - 1) / mrows);
}
int launch_conv3x3_tc(int c, int wo, const tc::ConvTcArgs& a, cudaStream_t st, bool persist = false) {
    if (persist && c == 16 && wo == 32) return tc::conv_tcp_launch<16, 32, 0>(a, device_sms(), st);
    if (persist && c == 32 && wo == 16) return tc::conv_tcp_launch<32, 16, 0>(a, device_sms(), st);
    if (c == 16 && wo == 32) return tc::conv_tc_launch<16, 32>(a, st);
    if (c == 32 && wo == 16) return tc::conv_tc_launch<32, 16>(a, st);
    if (c == 64 && wo == 8) return tc::conv_tc_launch<64, 8>(a, st);
    return LC_ERR_INVALID;
}
// forward variant whose prologue finishes the previous residual block (conv_tc.cuh MODE 2)
int launch_conv3x3_tc_res(int c, int wo, const tc::ConvTcArgs& a, cudaStream_t st, bool persist = false) {
    if (persist && c == 16 && wo == 32) return tc::conv_tcp_launch<16, 32, 2>(a, device_sms(), st);
    if (persist && c == 32 && wo == 16) return tc::conv_tcp_launch<32, 16, 2>(a, device_sms(), st);
    if (c == 16 && wo == 32) return tc::conv_tc_launch<16, 32, 2>(a, st);
    if (c == 32 && wo == 16) return tc::conv_tc_launch<32, 16, 2>(a, st);
    if (c == 64 && wo == 8) return tc::conv_tc_launch<64, 8, 2>(a, st);
    return LC_ERR_INVALID;
}
// data-gradient variant with the fused BatchNorm backward (apply in the prologue, mask + reduction in the epilogue)
int launch_conv3x3_tc_bwd(int c, int wo, const tc::ConvTcArgs& a, cudaStream_t st) {
    if (c == 16 && wo == 32) return tc::conv_tc_launch<16, 32, 1>(a, st);
    if (c == 32 && wo == 16) return tc::conv_tc_launch<32, 16, 1>(a, st);
    if (c == 64 && wo == 8) return tc::conv_tc_launch<64, 8, 1>(a, st);
    return LC_ERR_INVALID;
}

int wgrad_nsplit(int cin, int cout) {
    if (cout == 16) return 256;
    if (cout == 32) return 128;
    return 64;      // cout 64 (CUDA-core kernel: x CIN_SPLIT CTAs)
}
int wgrad_nsplit_tc(int c) {
    // CTAs (= split-K partials) per layer.  Measured inside the step (tools/step_breakdown.py), where these kernels share the SMs with the data-gradient
    // chain — not alone, where more CTAs always win: LC_WGRAD_NSPLIT="n16,n32,n64" overrides for A/B runs.
    static int ns[3] = {0, 0, 0};
    if (ns[0] == 0) {
        ns[0] = 222; ns[1] = 111; ns[2] = 34;      // profiles/r3w_wgrad_nsplit_sweep.txt: 296,148,34 -> 222,111,34 = backward 784 -> 768 us
        const char* e = getenv("LC_WGRAD_NSPLIT");
        int a = 0, b = 0, d = 0;
        if (e != nullptr && sscanf(e, "%d,%d,%d", &a, &b, &d) == 3 && a >= 1 && a <= 296 && b >= 1 && b <= 148 && d >= 1 && d <= 100) { ns[0] = a; ns[1] = b; ns[2] = d; }
    }
    return c == 64 ? ns[2] : (c == 32 ? ns[1] : ns[0]);
}
int launch_wgrad3x3_tc(int c, int wo, const tc::WgradTcArgs& a, int nsplit, cudaStream_t st) {
    if (c == 16 && wo == 32) return tc::wgrad_tc_launch<16, 32>(a, nsplit, st);
    if (c == 32 && wo == 16) return tc::wgrad_tc_launch<32, 16>(a, nsplit, st);
    if (c == 64 && wo == 8) return tc::wgrad_tc_launch<64, 8>(a, nsplit, st);
    return LC_ERR_INVALID;
}

int launch_conv1x1_fwd(int cin, int cout, int wo, const Conv1x1Args& a, cudaStream_t st) {
    const long long npix = (long long)a.B * wo * wo;
    const int grid = (int)((npix + 127) / 128);
    if (cin == 16 && cout == 32 && wo == 16) conv1x1s2_fwd_kernel<16, 32, 16><<<grid, 128, 0, st>>>(a);
    else if (cin == 32 && cout == 64 && wo == 8) conv1x1s2_fwd_kernel<32, 64, 8><<<grid, 128, 0, st>>>(a);
    else return LC_ERR_INVALID;
    return lc_launch_status();
}
int launch_conv1x1_dgrad(int cin, int cout, int wo, const float* dy, const float* w, float* gin, int B, cudaStream_t st) {
    const long long npix = (long long)B * wo * wo;
    const int grid = (int)((npix + 127) / 128);
    if (cin == 16 && cout == 32 && wo == 16) conv1x1s2_dgrad_accum_kernel<16, 32, 16><<<grid, 128, 0, st>>>(dy, w, gin, B);
    else if (cin == 32 && cout == 64 && wo == 8) conv1x1s2_dgrad_accum_kernel<32, 64, 8><<<grid, 128, 0, st>>>(dy, w, gin, B);
    else return LC_ERR_INVALID;
    return lc_launch_status();
}
constexpr int k1x1Split = 64;
int launch_conv1x1_wgrad(int cin, int cout, int wo, const float* in, const float* dy, float* partial, int B, cudaStream_t st) {
    if (cin == 16 && cout == 32 && wo == 16) conv1x1s2_wgrad_kernel<16, 32, 16><<<k1x1Split, 256, 0, st>>>(in, dy, partial, B);
    else if (cin == 32 && cout == 64 && wo == 8) conv1x1s2_wgrad_kernel<32, 64, 8><<<k1x1Split, 256, 0, st>>>(in, dy, partial, B);
    else return LC_ERR_INVALID;
    return lc_launch_status();
}

int launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t st) {
    long long blocks = (a.npix + 127) / 128;
    const int grid = (int)(blocks < kBnBwdBlocks ? blocks : kBnBwdBlocks);
    if (a.C == 16) bn_bwd_reduce_kernel<16><<<grid, 256, 0, st>>>(a);
    else if (a.C == 32) bn_bwd_reduce_kernel<32><<<grid, 256, 0, st>>>(a);
    else if (a.C == 64) bn_bwd_reduce_kernel<64><<<grid, 256, 0, st>>>(a);
    else return LC_ERR_INVALID;
    return lc_launch_status();
}
int launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st) {
    int grid = elem_grid(a.npix * (a.C / 4));
    if (a.blazy.partial != nullptr && grid > 2 * kNumSMs) grid = 2 * kNumSMs;
    bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(a);
    return lc_launch_status();
}
int launch_bn_bwd(const BnBwdArgs& a, cudaStream_t st) {
    if (launch_bn_bwd_reduce(a, st) != LC_OK) return LC_ERR_CUDA;
    return launch_bn_bwd_apply(a, st);
}

int launch_bn_act(const BnActArgs& a, cudaStream_t st) {
    int grid = elem_grid(a.n4);
    if (a.lazy.partial != nullptr && grid > 2 * kNumSMs) grid = 2 * kNumSMs;     // every CTA re-reduces the partial rows: fewer, longer CTAs
    bn_act_fwd_kernel<<<grid, 256, 0, st>>>(a);
    return lc_launch_status();
}

// ---- plan ----------------------------------------------------------------------------------------------------------
struct ConvL {
    int cin, cout, ksize, stride, wo;     // wo = output width
    long long w_off, gamma_off, beta_off; // parameter arena
    long long rstat_off;                  // running-stat arena (mean[C], var[C])
    long long aff_off;                    // workspace: scale, shift, mean, invstd (4*C)
    long long y_off;                      // workspace: raw conv output
    long long wf_off, wd_off, part_off;   // workspace-relative (packed weights / partials)
    long long wtf_off, wtd_off;           // tensor-core packings (-1 when the layer stays on the CUDA-core path)
    long long wts_off = -1, wtsd_off = -1; // stride-2 layers: tensor-core forward / data-gradient packings (conv_s2_tc.cuh); the weight gradient stays on CUDA cores
    long long fpartL_off;                 // workspace: this layer's own forward-statistics partial rows (deferred finalisation), -1 if not tensor-core
    long long bpartL_off;                 // workspace: this layer's BatchNorm-backward partial rows (written by a fused data-gradient epilogue)
    int nsplit;
};
struct BlockL {
    int conv_a, conv_b, conv_d;   // conv indices (conv_d = -1: identity shortcut)
    int stage;                    // 0,1,2
    long long out_off;            // workspace: block output (post ReLU)
};

}  // namespace

struct lc_resnet {
    int depth, in_ch, img, max_batch, nblk;
    std::vector<ConvL> convs;
    std::vector<BlockL> blocks;
    long long n_params = 0, n_rstat = 0, ws_floats = 0;
    // workspace offsets (floats)
    long long off_counters, off_aff, off_fpart, off_bpart, off_coef, off_packed, off_wpart, off_feat, off_dfeat, off_a0, off_G[3], off_T1, off_T2, off_T3;
    long long packed_floats = 0, wpart_floats = 0;
    ConvTabEntry* d_tab = nullptr;
    BnEvalEntry* d_bntab = nullptr;
    BnFinEntry* d_fintab = nullptr;
    int n_deferred = 0, lazy_stats = 1;
    int tab_blocks = 0;
    int launches_fwd = 0, launches_bwd = 0;
    int mode = 0;     // 0: exact fp32 CUDA-core convs; 1: TF32 tcgen05 convs (fwd + dgrad of the stride-1 3x3 layers)
    int last_relu = 1; // 0: the last residual block has no final ReLU (LUCIR's modified_ResNet, resnet.py:472-502)
    // backward runs two chains: the data-gradient chain (BN backward -> dgrad -> ...) on the caller's stream and the weight-gradient kernels, which
    // are leaves of the dependency graph, on `side` — forked / joined with events so that the pair is capturable into one CUDA graph
    long long off_T1b = 0;
    // fused-backward flow (mode 1): per-layer BatchNorm-backward coefficients, a second gradient buffer per stage and a second conv_b data-gradient
    // buffer, so that a weight-gradient kernel on the side chain can still read version k while the main chain already writes version k^1
    // The weight-gradient kernels are leaves, independent of each other: they are dealt round-robin to `nside` side streams so that several run beside
    // the main chain at once, and G / the conv_b data gradient live in rings of kRing versions so that the main chain can run ahead of its readers.
    static constexpr int kRing = 4, kMaxSide = 3;
    long long off_coefL = 0, off_Gr[3][kRing] = {}, off_T2r[kRing] = {};
    cudaEvent_t ev_pub[2 * kRing] = {}, ev_rd[2 * kRing] = {}, ev_joinS[kMaxSide] = {};
    cudaStream_t sides[kMaxSide] = {};
    int nside = 2;
    int debug_skip = 0;
    int persist = 1;               // forward convs of stages 1 / 2 on the persistent kernel (conv_tcp.cuh)
    int fuse_block_out = 1;
    int fused = 1;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_dy[2] = {nullptr, nullptr}, ev_w[2] = {nullptr, nullptr}, ev_dy3 = nullptr, ev_w3 = nullptr, ev_join = nullptr;
    int overlap = 1;
};


#define LC_TRY(expr)                  \
    do {                              \
        int _e = (expr);              \
        if (_e != LC_OK) return _e;   \
        ++launches;                   \
    } while (0)
#define LC_CALL(expr)                 \
    do {                              \
        int _e = (expr);              \
        if (_e != LC_OK) return _e;   \
    } while (0)

// Backward of the tensor-core mode with the BatchNorm backward folded into the convolutions (resnet.py:306-316,382 + autograd): per residual block the
// main chain is TWO launches — the data-gradient conv of conv_b and of conv_a — instead of six.  Each conv's epilogue masks its result with the ReLU
// that precedes it in the forward pass, stores the masked gradient and reduces (sum g, sum g*xhat) for the BatchNorm below; the last CTA leaves the
// coefficients of dy = c0*g + c1*y + c2, which the next data-gradient conv and the weight-gradient kernel evaluate while staging their operands.
// Stage transitions (stride-2 conv_a, 1x1 shortcut) and the stem stay on the CUDA-core kernels and use the stand-alone reduce / apply launches.
static int resnet_backward_fused(lc_resnet* n, const float* x, int batch, const float* params, float* ws, float* grads, cudaStream_t st) {
    constexpr int R = lc_resnet::kRing;
    cudaStream_t sw = n->overlap ? n->side : st;        // CUDA-core weight-gradient kernels (stage transitions, stem)
    int wsel = 0;                                       // round-robin over the side streams for the tensor-core weight gradients
    bool side_used[lc_resnet::kMaxSide] = {false, false, false};
    int launches = 0;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws + n->off_counters);
    int* err_flag = reinterpret_cast<int*>(counters) + 8;
    float* packed = ws + n->off_packed;
    float* wpart = ws + n->off_wpart;
    float* Gv[3][R];
    float* T2v[R];
    for (int r = 0; r < R; ++r) { for (int q = 0; q < 3; ++q) Gv[q][r] = ws + n->off_Gr[q][r]; T2v[r] = ws + n->off_T2r[r]; }
    float* Tdy[2] = {ws + n->off_T1, ws + n->off_T1b};
    float* T3 = ws + n->off_T3;
    int gk = 0, tk = 0, dk = 0;
    bool rd_rec[2 * R] = {}, wrec[2] = {false, false}, w3rec = false;
    bool g_fused = false;                       // the current G is already masked and the coefficients of its BatchNorm are ready

    auto coef_of = [&](int conv) { return ws + n->off_coefL + (long long)conv * 3 * 64; };
    // coef_lazy[i]: the BatchNorm-backward sums of layer i exist only as the partial rows of a fused epilogue; its consumers reduce them (BnBwdLazy)
    const bool lazyb = n->lazy_stats != 0;
    std::vector<char> coef_lazy(n->convs.size(), 0);
    auto blazy_of = [&](int conv_idx, int write_grads) {
        BnBwdLazy z{};
        if (!coef_lazy[conv_idx]) return z;
        const ConvL& c = n->convs[conv_idx];
        const float* aff = ws + c.aff_off;
        const int mrows = 128 * (c.cout == 16 ? 4 : (c.cout == 32 ? 2 : 1));
        z.partial = ws + c.bpartL_off; z.scale = aff; z.mean = aff + 2 * c.cout; z.invstd = aff + 3 * c.cout;
        z.dgamma = grads + c.gamma_off; z.dbeta = grads + c.beta_off;
        z.nparts = (int)(((long long)batch * (c.wo + 2) * (c.wo + 2) + mrows - 1) / mrows);
        z.count = (float)((long long)batch * c.wo * c.wo); z.write_grads = write_grads;
        return z;
    };
    auto bn_args = [&](const ConvL& c, int conv_idx, const float* g, const float* mask_src, int mask_mode) {
        BnBwdArgs a{};
        const float* aff = ws + c.aff_off;
        a.g = g; a.mask_src = mask_src; a.mask_mode = mask_mode; a.y = ws + c.y_off;
        a.scale = aff; a.shift = aff + c.cout; a.mean = aff + 2 * c.cout; a.invstd = aff + 3 * c.cout;
        a.partial = ws + n->off_bpart; a.counter = counters + 1; a.coef = coef_of(conv_idx);
        a.dgamma = grads + c.gamma_off; a.dbeta = grads + c.beta_off;
        a.npix = (long long)batch * c.wo * c.wo; a.C = c.cout;
        return a;
    };
    // slot s: [0, R) = the G ring, [R, 2R) = the ring of conv_b data-gradient buffers
    cudaStream_t swt = st;                      // side stream of the tensor-core weight gradient being issued
    auto acquire = [&](int slot) -> int {       // main chain is about to overwrite the buffer: its last side-chain reader must be done
        if (n->overlap && rd_rec[slot] && cudaStreamWaitEvent(st, n->ev_rd[slot], 0) != cudaSuccess) return LC_ERR_CUDA;
        return LC_OK;
    };
    auto publish = [&](int slot) -> int {       // buffer (and the coefficients that go with it) complete on the main chain: pick the reader's stream
        if (!n->overlap) { swt = st; return LC_OK; }
        swt = n->sides[wsel]; side_used[wsel] = true; wsel = (wsel + 1) % n->nside;
        if (cudaEventRecord(n->ev_pub[slot], st) != cudaSuccess || cudaStreamWaitEvent(swt, n->ev_pub[slot], 0) != cudaSuccess) return LC_ERR_CUDA;
        return LC_OK;
    };
    auto read_done = [&](int slot) -> int {
        if (n->overlap) { if (cudaEventRecord(n->ev_rd[slot], swt) != cudaSuccess) return LC_ERR_CUDA; rd_rec[slot] = true; }
        return LC_OK;
    };
    auto wgrad_tc = [&](const ConvL& c, int conv_idx, const float* in, const float* pro_scale, const float* pro_shift, const float* g) {
        tc::WgradTcArgs w{};
        w.in = in; w.dy = g; w.dy_y = ws + c.y_off; w.dy_coef = coef_of(conv_idx); w.dy_blazy = blazy_of(conv_idx, 1);
        w.partial = wpart + c.part_off; w.B = batch; w.error_flag = err_flag;
        w.pro_scale = pro_scale; w.pro_shift = pro_shift;
        if (n->debug_skip == 1) return LC_OK;            // timing experiments only (LC_RESNET_DEBUG_SKIP=1): main chain without the weight gradients
        return launch_wgrad3x3_tc(c.cin, c.wo, w, wgrad_nsplit_tc(c.cin), swt);
    };
    // fused data-gradient conv: in = c0*g + c1*y + c2 of layer c; the result is the gradient w.r.t. the ReLU output of BatchNorm `below`
    auto dgrad_tc = [&](const ConvL& c, int conv_idx, const float* g, float* out, const float* addend, const ConvL& below, int below_idx,
                        const float* mask_out) {
        tc::ConvTcArgs t{};
        t.in = g; t.pro_y = ws + c.y_off; t.pro_coef = coef_of(conv_idx); t.pro_blazy = blazy_of(conv_idx, 0);
        t.wtc = packed + c.wtd_off; t.out = out; t.addend = addend; t.B = batch;
        t.error_flag = err_flag;
        const float* aff = ws + below.aff_off;
        t.bw.y = ws + below.y_off; t.bw.mask_out = mask_out;
        t.bw.scale = aff; t.bw.shift = aff + below.cout; t.bw.mean = aff + 2 * below.cout; t.bw.invstd = aff + 3 * below.cout;
        t.bw.partial = lazyb ? ws + below.bpartL_off : ws + n->off_fpart; t.bw.defer = lazyb ? 1 : 0;
        t.bw.counter = counters + 0; t.bw.coef = coef_of(below_idx);
        t.bw.dgamma = grads + below.gamma_off; t.bw.dbeta = grads + below.beta_off;
        coef_lazy[below_idx] = lazyb ? 1 : 0;
        if (n->debug_skip == 2) return LC_OK;            // LC_RESNET_DEBUG_SKIP=2: weight-gradient kernels without the data-gradient chain
        return launch_conv3x3_tc_bwd(c.cin, c.wo, t, st);
    };
    // materialised dy for a CUDA-core consumer (stride-2 conv_a, stem): Tdy[dk] <- c0*g + c1*y + c2
    auto apply_to_tdy = [&](const ConvL& c, int conv_idx, const float* g) -> int {
        if (n->overlap && wrec[dk] && cudaStreamWaitEvent(st, n->ev_w[dk], 0) != cudaSuccess) return LC_ERR_CUDA;
        BnBwdArgs a = bn_args(c, conv_idx, g, nullptr, LC_MASK_NONE);
        a.dy = Tdy[dk]; a.blazy = blazy_of(conv_idx, 1);
        if (launch_bn_bwd_apply(a, st) != LC_OK) return LC_ERR_CUDA;
        if (n->overlap && (cudaEventRecord(n->ev_dy[dk], st) != cudaSuccess || cudaStreamWaitEvent(sw, n->ev_dy[dk], 0) != cudaSuccess)) return LC_ERR_CUDA;
        return LC_OK;
    };
    auto tdy_read_done = [&]() -> int {
        if (n->overlap) { if (cudaEventRecord(n->ev_w[dk], sw) != cudaSuccess) return LC_ERR_CUDA; wrec[dk] = true; }
        dk ^= 1;
        return LC_OK;
    };

    for (int bi = (int)n->blocks.size() - 1; bi >= 0; --bi) {
        const BlockL& bl = n->blocks[bi];
        const ConvL& ca = n->convs[bl.conv_a];
        const ConvL& cb = n->convs[bl.conv_b];
        const int s = bl.stage;
        const float* blk_in = bi == 0 ? ws + n->off_a0 : ws + n->blocks[bi - 1].out_off;
        float* G = Gv[s][gk];
        if (!g_fused) {
            // G arrives unmasked (head backward, or the CUDA-core data gradient of a stage transition): stand-alone reduction, which also masks G in place
            const bool no_relu = !n->last_relu && bi == (int)n->blocks.size() - 1;
            BnBwdArgs a = bn_args(cb, bl.conv_b, G, ws + bl.out_off, no_relu ? LC_MASK_NONE : LC_MASK_FROM_OUT);
            a.g_out = G; a.reduce_writes_g = no_relu ? 0 : 1;
            LC_TRY(launch_bn_bwd_reduce(a, st));
        }
        LC_CALL(publish(gk));
        if (bl.conv_d >= 0) {
            const ConvL& cd = n->convs[bl.conv_d];
            if (n->overlap && w3rec && cudaStreamWaitEvent(st, n->ev_w3, 0) != cudaSuccess) return LC_ERR_CUDA;
            BnBwdArgs a = bn_args(cd, bl.conv_d, G, nullptr, LC_MASK_NONE);
            a.dy = T3;
            LC_TRY(launch_bn_bwd_reduce(a, st));
            LC_TRY(launch_bn_bwd_apply(a, st));
            if (n->overlap && (cudaEventRecord(n->ev_dy3, st) != cudaSuccess || cudaStreamWaitEvent(sw, n->ev_dy3, 0) != cudaSuccess)) return LC_ERR_CUDA;
            LC_TRY(launch_conv1x1_wgrad(cd.cin, cd.cout, cd.wo, blk_in, T3, wpart + cd.part_off, batch, sw));
            if (n->overlap) { if (cudaEventRecord(n->ev_w3, sw) != cudaSuccess) return LC_ERR_CUDA; w3rec = true; }
        }
        // conv_b: weight gradient on the side chain (input = relu(bn_a(y_a)) and dy = bn_b backward of G, both evaluated while staging)
        LC_TRY(wgrad_tc(cb, bl.conv_b, ws + ca.y_off, ws + ca.aff_off, ws + ca.aff_off + ca.cout, G));
        LC_CALL(read_done(gk));
        // conv_b data gradient -> T2 = masked d(relu(bn_a)) + the bn_a reduction
        float* T2 = T2v[tk];
        LC_CALL(acquire(R + tk));
        LC_TRY(dgrad_tc(cb, bl.conv_b, G, T2, nullptr, ca, bl.conv_a, nullptr));
        if (ca.stride == 1) {
            LC_CALL(publish(R + tk));
            LC_TRY(wgrad_tc(ca, bl.conv_a, blk_in, nullptr, nullptr, T2));
            LC_CALL(read_done(R + tk));
            // conv_a data gradient + the residual-branch gradient -> next G (masked with the ReLU that produced blk_in) + the reduction of its BatchNorm
            const int below_idx = bi == 0 ? 0 : n->blocks[bi - 1].conv_b;
            const int gn = (gk + 1) % R;
            LC_CALL(acquire(gn));
            LC_TRY(dgrad_tc(ca, bl.conv_a, T2, Gv[s][gn], G, n->convs[below_idx], below_idx, blk_in));
            gk = gn;
            g_fused = true;
        } else {
            const ConvL& cd = n->convs[bl.conv_d];
            LC_CALL(apply_to_tdy(ca, bl.conv_a, T2)); ++launches;
            WgradArgs w{};
            w.in = blk_in; w.dy = Tdy[dk]; w.partial = wpart + ca.part_off; w.B = batch; w.nsplit = ca.nsplit;
            LC_TRY(launch_wgrad3x3(ca.cin, ca.cout, ca.wo, ca.stride, false, w, sw));
            const int gn = (gk + 1) % R;
            float* Gprev = Gv[s - 1][gn];
            LC_CALL(acquire(gn));
            bool shortcut_dgrad_done = false;
            if (ca.wtsd_off >= 0) {       // parity-plane data gradient on the tensor cores
                tc::DgradS2Args a{};
                a.dy = Tdy[dk]; a.wtc = packed + ca.wtsd_off; a.out = Gprev; a.B = batch; a.error_flag = err_flag;
                if (cd.wtsd_off >= 0) { a.dy1 = T3; a.w1 = packed + cd.wtsd_off; shortcut_dgrad_done = true; }      // the shortcut's data gradient rides along
                LC_TRY(launch_dgrad3x3s2_tc(ca.cin, ca.wo, a, st));
            } else {
                Conv3x3Args a{};
                a.in = Tdy[dk]; a.wpack = packed + ca.wd_off; a.B = batch; a.out = Gprev;
                LC_TRY(launch_conv3x3(ca.cout, ca.cin, ca.wo * 2, 1, true, false, a, st));
            }
            if (!shortcut_dgrad_done) LC_TRY(launch_conv1x1_dgrad(cd.cin, cd.cout, cd.wo, T3, params + cd.w_off, Gprev, batch, st));
            LC_CALL(tdy_read_done());
            gk = gn;
            g_fused = false;
        }
        tk = (tk + 1) % R;
    }
    {   // stem: G is masked with the stem ReLU and its coefficients are ready (block 0's conv_a epilogue)
        const ConvL& c = n->convs[0];
        if (!g_fused) return LC_ERR_INVALID;
        LC_CALL(apply_to_tdy(c, 0, Gv[0][gk])); ++launches;
        WgradArgs w{};
        w.in = x; w.dy = Tdy[dk]; w.partial = wpart + c.part_off; w.B = batch; w.nsplit = c.nsplit;
        LC_TRY(launch_wgrad3x3(c.cin, c.cout, c.wo, 1, true, w, sw));
    }
    if (n->overlap) {
        side_used[0] = true;        // sw
        for (int i = 0; i < lc_resnet::kMaxSide; ++i)
            if (side_used[i] && (cudaEventRecord(n->ev_joinS[i], n->sides[i]) != cudaSuccess || cudaStreamWaitEvent(st, n->ev_joinS[i], 0) != cudaSuccess))
                return LC_ERR_CUDA;
    }
    wgrad_reduce_all_kernel<<<n->tab_blocks, 256, 0, st>>>(n->d_tab, (int)n->convs.size(), wpart, grads, n->mode);
    LC_TRY(lc_launch_status());
    n->launches_bwd = launches;
    return LC_OK;
}

extern "C" {

const char* lc_version(void) { return "libcontinual_b200 0.1 (sm_100a)"; }

int lc_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return LC_ERR_CUDA;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return LC_ERR_CUDA;
    return p.major == 10 ? LC_OK : LC_ERR_INVALID;
}

int lc_resnet_create(int depth, int in_ch, int img, int max_batch, lc_resnet** out) {
    LC_CHECK_ARG(out != nullptr && (depth - 2) % 6 == 0 && depth >= 8 && in_ch == 3 && img == 32 && max_batch >= 1 && max_batch <= 4096);
    lc_resnet* n = new (std::nothrow) lc_resnet();
    if (!n) return LC_ERR_INVALID;
    n->depth = depth; n->in_ch = in_ch; n->img = img; n->max_batch = max_batch; n->nblk = (depth - 2) / 6;
    const long long B = max_batch;
    long long po = 0, ro = 0;
    auto add_conv = [&](int cin, int cout, int k, int stride, int wo) {
        ConvL c{};
        c.cin = cin; c.cout = cout; c.ksize = k; c.stride = stride; c.wo = wo;
        c.w_off = po; po += (long long)cout * cin * k * k;
        c.gamma_off = po; po += cout;
        c.beta_off = po; po += cout;
        c.rstat_off = ro; ro += 2 * cout;
        n->convs.push_back(c);
        return (int)n->convs.size() - 1;
    };
    add_conv(in_ch, 16, 3, 1, 32);
    int inpl = 16, w = 32;
    const int planes[3] = {16, 32, 64};
    for (int s = 0; s < 3; ++s) {
        for (int b = 0; b < n->nblk; ++b) {
            const int stride = (b == 0 && s > 0) ? 2 : 1;
            if (stride == 2) w /= 2;
            BlockL bl{};
            bl.stage = s;
            bl.conv_a = add_conv(b == 0 ? inpl : planes[s], planes[s], 3, stride, w);
            bl.conv_b = add_conv(planes[s], planes[s], 3, 1, w);
            bl.conv_d = (b == 0 && s > 0) ? add_conv(inpl, planes[s], 1, stride, w) : -1;
            n->blocks.push_back(bl);
        }
        inpl = planes[s];
    }
    n->n_params = po; n->n_rstat = ro;

    // workspace layout
    long long o = 0;
    auto take = [&](long long cnt) { long long r = o; o += (cnt + 3) / 4 * 4; return r; };
    n->off_counters = take(64);
    n->off_aff = take((long long)n->convs.size() * 4 * 64);
    long long fpart = 0;
    for (auto& c : n->convs) {
        long long parts = c.ksize == 3 ? B * 8 : (B * c.wo * c.wo + 127) / 128;   // >= tiles per image of every 3x3 config
        if (c.ksize == 3) parts = std::max(parts, tc::conv_tc_tiles((int)B, c.wo));
        fpart = std::max(fpart, parts * 2 * c.cout);
    }
    n->off_fpart = take(fpart);
    n->off_bpart = take((long long)kBnBwdBlocks * 2 * 64);
    n->off_coef = take(3 * 64);
    long long pk = 0, wp = 0;
    for (auto& c : n->convs) {
        const long long ne = (long long)c.cout * c.cin * c.ksize * c.ksize;
        c.wf_off = pk; pk += ne;
        if (c.ksize == 3) { c.wd_off = pk; pk += ne; } else c.wd_off = -1;
        if (tc_eligible(c.cin, c.cout, c.wo, c.stride, c.ksize)) { c.wtf_off = pk; pk += ne; c.wtd_off = pk; pk += ne; } else { c.wtf_off = c.wtd_off = -1; }
        if (s2_tc_eligible(c.cin, c.cout, c.wo, c.stride, c.ksize) && conv_s2_tc_enabled()) { c.wts_off = pk; pk += ne; c.wtsd_off = pk; pk += ne; }
        // the 1x1 / stride-2 shortcut of the same blocks: forward fused into conv3x3s2_tc_kernel (its operand is the staged parity plane (0,0))
        if (c.ksize == 1 && c.stride == 2 && conv_s2_tc_enabled() && s2_tc_eligible(c.cin, c.cout, c.wo, 2, 3)) { c.wts_off = pk; pk += ne; c.wtsd_off = pk; pk += ne; }
        c.nsplit = c.ksize == 3 ? wgrad_nsplit(c.cin, c.cout) : k1x1Split;
        c.part_off = wp; wp += ne * std::max(c.nsplit, c.wtf_off >= 0 ? wgrad_nsplit_tc(c.cout) : 0);
    }
    n->packed_floats = pk; n->wpart_floats = wp;
    n->off_packed = take(pk);
    n->off_wpart = take(wp);
    n->off_feat = take(B * 64);
    n->off_dfeat = take(B * 64);
    for (size_t i = 0; i < n->convs.size(); ++i) {
        auto& c = n->convs[i];
        c.aff_off = n->off_aff + (long long)i * 4 * 64;
        c.y_off = take(B * c.wo * c.wo * c.cout);
    }
    n->off_a0 = take(B * 32 * 32 * 16);
    {
        int wcur = 32;
        for (auto& bl : n->blocks) {
            wcur = n->convs[bl.conv_b].wo;
            bl.out_off = take(B * wcur * wcur * n->convs[bl.conv_b].cout);
        }
    }
    n->off_G[0] = take(B * 32 * 32 * 16);
    n->off_G[1] = take(B * 16 * 16 * 32);
    n->off_G[2] = take(B * 8 * 8 * 64);
    n->off_T1 = take(B * 32 * 32 * 16);
    n->off_T2 = take(B * 32 * 32 * 16);
    n->off_T3 = take(B * 16 * 16 * 32);
    n->off_T1b = take(B * 32 * 32 * 16);
    n->off_coefL = take((long long)n->convs.size() * 3 * 64);
    for (int r = 0; r < lc_resnet::kRing; ++r) {
        n->off_Gr[0][r] = r == 0 ? n->off_G[0] : take(B * 32 * 32 * 16);
        n->off_Gr[1][r] = r == 0 ? n->off_G[1] : take(B * 16 * 16 * 32);
        n->off_Gr[2][r] = r == 0 ? n->off_G[2] : take(B * 8 * 8 * 64);
        n->off_T2r[r] = r == 0 ? n->off_T2 : take(B * 32 * 32 * 16);
    }
    for (auto& c : n->convs)
        c.fpartL_off = c.wtf_off >= 0 ? take(tc::conv_tc_tiles((int)B, c.wo) * 2 * c.cout)
                                      : (c.wts_off >= 0 && c.ksize == 3 ? take((long long)tc::conv_s2_tc_grid(B, c.wo) * 2 * c.cout) : -1);
    for (auto& c : n->convs) c.bpartL_off = c.ksize == 3 ? take(tc::conv_tc_tiles((int)B, c.wo) * 2 * c.cout) : -1;
    n->ws_floats = o;
    {
        const char* env = getenv("LC_RESNET_SERIAL");
        n->overlap = (env != nullptr && env[0] == '1') ? 0 : 1;
        const char* envf = getenv("LC_RESNET_UNFUSED");
        n->fused = (envf != nullptr && envf[0] == '1') ? 0 : 1;
        const char* envl = getenv("LC_RESNET_EAGER_STATS");
        n->lazy_stats = (envl != nullptr && envl[0] == '1') ? 0 : 1;
        bool okev = cudaStreamCreateWithFlags(&n->side, cudaStreamNonBlocking) == cudaSuccess;
        cudaEvent_t* evs[7] = {&n->ev_dy[0], &n->ev_dy[1], &n->ev_w[0], &n->ev_w[1], &n->ev_dy3, &n->ev_w3, &n->ev_join};
        for (auto e : evs) okev = okev && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 * lc_resnet::kRing; ++i)
            okev = okev && cudaEventCreateWithFlags(&n->ev_pub[i], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&n->ev_rd[i], cudaEventDisableTiming) == cudaSuccess;
        n->sides[0] = n->side;
        for (int i = 0; i < lc_resnet::kMaxSide; ++i) {
            if (i > 0) okev = okev && cudaStreamCreateWithFlags(&n->sides[i], cudaStreamNonBlocking) == cudaSuccess;
            okev = okev && cudaEventCreateWithFlags(&n->ev_joinS[i], cudaEventDisableTiming) == cudaSuccess;
        }
        const char* envb = getenv("LC_RESNET_NO_BLOCK_FUSE");
        n->fuse_block_out = (envb != nullptr && envb[0] == '1') ? 0 : 1;
        n->persist = conv_persist_enabled();
        const char* envd = getenv("LC_RESNET_DEBUG_SKIP");
        if (envd != nullptr && (envd[0] == '1' || envd[0] == '2')) n->debug_skip = envd[0] - '0';
        const char* envs = getenv("LC_RESNET_SIDE");
        if (envs != nullptr && envs[0] >= '1' && envs[0] <= '3') n->nside = envs[0] - '0';
        if (!okev) { lc_resnet_destroy(n); return LC_ERR_CUDA; }
    }

    // device tables
    std::vector<ConvTabEntry> tab;
    int blk = 0;
    for (auto& c : n->convs) {
        ConvTabEntry t{};
        t.w_off = c.w_off; t.wf_off = c.wf_off; t.wd_off = c.wd_off; t.part_off = c.part_off; t.wtf_off = c.wtf_off; t.wtd_off = c.wtd_off;
        t.wts_off1 = c.wts_off >= 0 ? c.wts_off + 1 : 0; t.wtsd_off1 = c.wtsd_off >= 0 ? c.wtsd_off + 1 : 0;
        t.cout = c.cout; t.cin = c.cin; t.ntap = c.ksize * c.ksize; t.nsplit = c.nsplit; t.blk_begin = blk;
        t.nsplit_tc = c.wtf_off >= 0 ? wgrad_nsplit_tc(c.cout) : c.nsplit;
        blk += (c.cout * c.cin * t.ntap + 255) / 256;
        tab.push_back(t);
    }
    n->tab_blocks = blk;
    std::vector<BnEvalEntry> bt;
    for (auto& c : n->convs) {
        BnEvalEntry e{};
        e.gamma_off = c.gamma_off; e.beta_off = c.beta_off; e.rstat_off = c.rstat_off; e.aff_off = c.aff_off; e.C = c.cout;
        bt.push_back(e);
    }
    std::vector<BnFinEntry> ft;
    for (auto& c : n->convs) {
        if (c.fpartL_off < 0) continue;
        BnFinEntry e{};
        e.part_off = c.fpartL_off; e.gamma_off = c.gamma_off; e.beta_off = c.beta_off; e.rstat_off = c.rstat_off; e.aff_off = c.aff_off; e.C = c.cout;
        e.pp = (c.wo + 2) * (c.wo + 2); e.mrows = 128 * (c.cout == 16 ? 4 : (c.cout == 32 ? 2 : 1)); e.hw = c.wo * c.wo;
        if (c.wts_off >= 0) { e.pp = (c.wo + 1) * (c.wo + 1); e.mrows = 128; }      // conv_s2_tc_grid()
        else if (n->persist && tcp_eligible(c.cout, c.wo)) { e.tmax = c.cout == 16 ? 8 : 3; e.nsm = device_sms(); }
        ft.push_back(e);
    }
    n->n_deferred = (int)ft.size();
    if (cudaMalloc(&n->d_fintab, (ft.size() + 1) * sizeof(BnFinEntry)) != cudaSuccess ||
        cudaMemcpy(n->d_fintab, ft.data(), ft.size() * sizeof(BnFinEntry), cudaMemcpyHostToDevice) != cudaSuccess) {
        lc_resnet_destroy(n);
        return LC_ERR_CUDA;
    }
    if (cudaMalloc(&n->d_tab, tab.size() * sizeof(ConvTabEntry)) != cudaSuccess || cudaMalloc(&n->d_bntab, bt.size() * sizeof(BnEvalEntry)) != cudaSuccess ||
        cudaMemcpy(n->d_tab, tab.data(), tab.size() * sizeof(ConvTabEntry), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(n->d_bntab, bt.data(), bt.size() * sizeof(BnEvalEntry), cudaMemcpyHostToDevice) != cudaSuccess) {
        lc_resnet_destroy(n);
        return LC_ERR_CUDA;
    }
    *out = n;
    return LC_OK;
}

void lc_resnet_destroy(lc_resnet* n) {
    if (!n) return;
    if (n->d_tab) cudaFree(n->d_tab);
    if (n->d_bntab) cudaFree(n->d_bntab);
    if (n->d_fintab) cudaFree(n->d_fintab);
    if (n->side) cudaStreamDestroy(n->side);
    for (cudaEvent_t e : {n->ev_dy[0], n->ev_dy[1], n->ev_w[0], n->ev_w[1], n->ev_dy3, n->ev_w3, n->ev_join}) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2 * lc_resnet::kRing; ++i) { if (n->ev_pub[i]) cudaEventDestroy(n->ev_pub[i]); if (n->ev_rd[i]) cudaEventDestroy(n->ev_rd[i]); }
    for (int i = 0; i < lc_resnet::kMaxSide; ++i) {
        if (i > 0 && n->sides[i]) cudaStreamDestroy(n->sides[i]);
        if (n->ev_joinS[i]) cudaEventDestroy(n->ev_joinS[i]);
    }
    delete n;
}

int lc_resnet_set_mode(lc_resnet* n, int mode) {
    LC_CHECK_ARG(n && (mode == 0 || mode == 1));
    n->mode = mode;
    return LC_OK;
}
int lc_resnet_get_mode(const lc_resnet* n) { return n ? n->mode : LC_ERR_INVALID; }
int lc_resnet_set_last_relu(lc_resnet* n, int last_relu) {
    LC_CHECK_ARG(n && (last_relu == 0 || last_relu == 1));
    n->last_relu = last_relu;
    return LC_OK;
}

long long lc_resnet_param_count(const lc_resnet* n) { return n ? n->n_params : LC_ERR_INVALID; }
long long lc_resnet_rstat_count(const lc_resnet* n) { return n ? n->n_rstat : LC_ERR_INVALID; }
long long lc_resnet_workspace_floats(const lc_resnet* n) { return n ? n->ws_floats : LC_ERR_INVALID; }
int lc_resnet_num_convs(const lc_resnet* n) { return n ? (int)n->convs.size() : LC_ERR_INVALID; }
int lc_resnet_num_launches(const lc_resnet* n, int backward) { return n ? (backward ? n->launches_bwd : n->launches_fwd) : LC_ERR_INVALID; }

int lc_resnet_conv_info(const lc_resnet* n, int idx, long long* w_off, int* cout, int* cin, int* ksize, int* stride, long long* gamma_off,
                        long long* beta_off, long long* rstat_off) {
    LC_CHECK_ARG(n && idx >= 0 && idx < (int)n->convs.size());
    const ConvL& c = n->convs[idx];
    if (w_off) *w_off = c.w_off;
    if (cout) *cout = c.cout;
    if (cin) *cin = c.cin;
    if (ksize) *ksize = c.ksize;
    if (stride) *stride = c.stride;
    if (gamma_off) *gamma_off = c.gamma_off;
    if (beta_off) *beta_off = c.beta_off;
    if (rstat_off) *rstat_off = c.rstat_off;
    return LC_OK;
}

long long lc_resnet_ws_offset(const lc_resnet* n, int what) {
    if (!n) return LC_ERR_INVALID;
    switch (what) {
        case LC_WS_FEAT: return n->off_feat;
        case LC_WS_DFEAT: return n->off_dfeat;
        case LC_WS_GRAD_LAST: return n->off_G[2];
        case LC_WS_FMAP1: return n->blocks[n->nblk - 1].out_off;
        case LC_WS_FMAP2: return n->blocks[2 * n->nblk - 1].out_off;
        case LC_WS_FMAP3: return n->blocks[3 * n->nblk - 1].out_off;
        default: return LC_ERR_INVALID;
    }
}

int lc_resnet_forward(lc_resnet* n, const float* x, int batch, const float* params, float* rstat, float* ws, int train, int update_running,
                      lc_stream_t stream) {
    LC_CHECK_ARG(n && x && params && ws && batch >= 1 && batch <= n->max_batch && (rstat || (train && !update_running)));
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws + n->off_counters);
    int* err_flag = reinterpret_cast<int*>(counters) + 8;
    float* packed = ws + n->off_packed;
    pack_weights_kernel<<<n->tab_blocks, 256, 0, st>>>(n->d_tab, (int)n->convs.size(), params, packed);
    LC_TRY(lc_launch_status());
    if (!train) {
        bn_eval_affine_kernel<<<(int)n->convs.size(), 64, 0, st>>>(n->d_bntab, (int)n->convs.size(), params, rstat, ws, kBnEps);
        LC_TRY(lc_launch_status());
    }
    const bool lazy = train && n->mode == 1 && n->lazy_stats;
    const bool persist = n->persist != 0;
    auto stat_for = [&](const ConvL& c) {
        BnStatArgs s{};
        if (!train) return s;
        s.partial = ws + n->off_fpart; s.counter = counters + 0;
        s.gamma = params + c.gamma_off; s.beta = params + c.beta_off;
        s.running_mean = rstat ? rstat + c.rstat_off : nullptr; s.running_var = rstat ? rstat + c.rstat_off + c.cout : nullptr;
        float* aff = ws + c.aff_off;
        s.scale = aff; s.shift = aff + c.cout; s.mean = aff + 2 * c.cout; s.invstd = aff + 3 * c.cout;
        s.momentum = kBnMomentum; s.eps = kBnEps; s.update_running = (update_running && rstat) ? 1 : 0;
        if (lazy && c.fpartL_off >= 0) { s.partial = ws + c.fpartL_off; s.defer = 1; }
        return s;
    };
    // consumer-side view of a deferred layer's statistics (conv_simt.cuh: BnLazy); .partial == null -> the layer was finalised by its producer
    auto lazy_of = [&](const ConvL& c) {
        BnLazy z{};
        if (!(lazy && c.fpartL_off >= 0)) return z;
        z.partial = ws + c.fpartL_off; z.gamma = params + c.gamma_off; z.beta = params + c.beta_off;
        z.nparts = c.wts_off >= 0 ? tc::conv_s2_tc_grid(batch, c.wo) : conv_tc_fwd_parts(batch, c.cout, c.wo, n->persist != 0);
        z.count = (float)((long long)batch * c.wo * c.wo); z.eps = kBnEps;
        return z;
    };
    // stem
    {
        const ConvL& c = n->convs[0];
        Conv3x3Args a{};
        a.in = x; a.wpack = packed + c.wf_off; a.out = ws + c.y_off; a.stat = stat_for(c); a.B = batch;
        LC_TRY(launch_conv3x3(c.cin, c.cout, c.wo, 1, false, true, a, st));
    }
    // the stem's relu(bn(y)) is evaluated (and stored to a0: it is block 0's residual) by block 0's conv_a prologue when that conv runs the persistent
    // kernel (conv_tcp.cuh MODE 2 without a residual operand); otherwise by bn_act_fwd_kernel
    const bool stem_fused = n->mode == 1 && n->fuse_block_out && n->persist && !n->blocks.empty() &&
                            n->convs[n->blocks[0].conv_a].wtf_off >= 0 && tcp_eligible(n->convs[n->blocks[0].conv_a].cout, n->convs[n->blocks[0].conv_a].wo);
    if (!stem_fused) {
        const ConvL& c = n->convs[0];
        BnActArgs e{};
        e.y = ws + c.y_off; e.scale = ws + c.aff_off; e.shift = ws + c.aff_off + c.cout; e.out = ws + n->off_a0;
        e.n4 = (long long)batch * c.wo * c.wo * c.cout / 4; e.C = c.cout;
        LC_TRY(launch_bn_act(e, st));
    }
    const float* cur = ws + n->off_a0;
    // `pending`: the previous block's output relu(bn_b(y_b) + residual) has not been materialised yet — the next tensor-core conv_a evaluates it in its
    // prologue (and stores it); any other consumer gets it from bn_act_fwd_kernel first
    const BlockL* pend = nullptr;
    const float* pend_res = nullptr;
    const bool fuse_out = n->mode == 1 && n->fuse_block_out;
    auto flush_pending = [&]() -> int {
        if (pend == nullptr) return LC_OK;
        const ConvL& pb = n->convs[pend->conv_b];
        BnActArgs e{};
        e.y = ws + pb.y_off; e.scale = ws + pb.aff_off; e.shift = ws + pb.aff_off + pb.cout; e.out = ws + pend->out_off;
        e.n4 = (long long)batch * pb.wo * pb.wo * pb.cout / 4; e.C = pb.cout; e.lazy = lazy_of(pb); e.res = pend_res;
        pend = nullptr;
        return launch_bn_act(e, st);
    };
    for (size_t bi = 0; bi < n->blocks.size(); ++bi) {
        const BlockL& bl = n->blocks[bi];
        const ConvL& ca = n->convs[bl.conv_a];
        const ConvL& cb = n->convs[bl.conv_b];
        bool shortcut_done = false;
        if (n->mode == 1 && ca.wtf_off >= 0) {
            tc::ConvTcArgs a{};
            a.wtc = packed + ca.wtf_off; a.out = ws + ca.y_off; a.stat = stat_for(ca); a.B = batch; a.error_flag = err_flag;
            if (pend != nullptr) {
                const ConvL& pb = n->convs[pend->conv_b];
                a.in = ws + pb.y_off; a.pro_scale = ws + pb.aff_off; a.pro_shift = ws + pb.aff_off + pb.cout; a.pro_lazy = lazy_of(pb);
                a.pro_res = pend_res; a.pro_out = ws + pend->out_off;
                pend = nullptr;
                LC_TRY(launch_conv3x3_tc_res(ca.cin, ca.wo, a, st, persist));
            } else if (bi == 0 && stem_fused) {
                const ConvL& c0 = n->convs[0];
                a.in = ws + c0.y_off; a.pro_scale = ws + c0.aff_off; a.pro_shift = ws + c0.aff_off + c0.cout; a.pro_out = ws + n->off_a0;
                LC_TRY(launch_conv3x3_tc_res(ca.cin, ca.wo, a, st, true));
            } else {
                a.in = cur;
                LC_TRY(launch_conv3x3_tc(ca.cin, ca.wo, a, st, persist));
            }
        } else if (n->mode == 1 && ca.wts_off >= 0) {       // stage transition: stride-2 conv on the tensor cores (parity planes)
            if (pend != nullptr) LC_TRY(flush_pending());
            tc::ConvS2Args a{};
            a.in = cur; a.wtc = packed + ca.wts_off; a.out = ws + ca.y_off; a.stat = stat_for(ca); a.B = batch; a.error_flag = err_flag;
            // the block's 1x1 shortcut conv rides along (training: only with deferred statistics for conv_a — the two layers must not share the eager
            // partial buffer)
            if (bl.conv_d >= 0 && n->convs[bl.conv_d].wts_off >= 0 && (!train || (lazy && ca.fpartL_off >= 0))) {
                const ConvL& cd = n->convs[bl.conv_d];
                a.w1 = packed + cd.wts_off; a.out1 = ws + cd.y_off; a.stat1 = stat_for(cd);
                shortcut_done = true;
            }
            LC_TRY(launch_conv3x3s2_tc(ca.cin, ca.wo, a, st));
        } else {
            if (pend != nullptr) LC_TRY(flush_pending());
            Conv3x3Args a{};
            a.in = cur; a.wpack = packed + ca.wf_off; a.out = ws + ca.y_off; a.stat = stat_for(ca); a.B = batch;
            LC_TRY(launch_conv3x3(ca.cin, ca.cout, ca.wo, ca.stride, false, false, a, st));
        }
        if (n->mode == 1 && cb.wtf_off >= 0) {
            tc::ConvTcArgs a{};
            a.in = ws + ca.y_off; a.wtc = packed + cb.wtf_off; a.out = ws + cb.y_off; a.stat = stat_for(cb); a.B = batch; a.error_flag = err_flag;
            a.pro_scale = ws + ca.aff_off; a.pro_shift = ws + ca.aff_off + ca.cout; a.pro_lazy = lazy_of(ca);
            LC_TRY(launch_conv3x3_tc(cb.cin, cb.wo, a, st, persist));
        } else {
            Conv3x3Args a{};
            a.in = ws + ca.y_off; a.wpack = packed + cb.wf_off; a.out = ws + cb.y_off; a.stat = stat_for(cb); a.B = batch;
            a.pro_scale = ws + ca.aff_off; a.pro_shift = ws + ca.aff_off + ca.cout;
            LC_TRY(launch_conv3x3(cb.cin, cb.cout, cb.wo, 1, false, false, a, st));
        }
        const bool last = bi + 1 == n->blocks.size();
        // the block output can stay pending when the next block's conv_a is a tensor-core conv (same stage, stride 1) and this block has an identity
        // shortcut and a final ReLU
        const bool defer_out = fuse_out && bl.conv_d < 0 && !last && n->convs[n->blocks[bi + 1].conv_a].wtf_off >= 0;
        if (defer_out) {
            pend = &bl; pend_res = cur;
        } else {
            BnActArgs e{};
            e.y = ws + cb.y_off; e.scale = ws + cb.aff_off; e.shift = ws + cb.aff_off + cb.cout; e.out = ws + bl.out_off;
            e.n4 = (long long)batch * cb.wo * cb.wo * cb.cout / 4; e.C = cb.cout; e.lazy = lazy_of(cb);
            if (bl.conv_d >= 0) {
                const ConvL& cd = n->convs[bl.conv_d];
                if (!shortcut_done) {
                    Conv1x1Args a{};
                    a.in = cur; a.w = packed + cd.wf_off; a.out = ws + cd.y_off; a.stat = stat_for(cd); a.B = batch;
                    LC_TRY(launch_conv1x1_fwd(cd.cin, cd.cout, cd.wo, a, st));
                }
                e.res = ws + cd.y_off; e.res_scale = ws + cd.aff_off; e.res_shift = ws + cd.aff_off + cd.cout;
            } else {
                e.res = cur;
            }
            e.no_relu = (!n->last_relu && last) ? 1 : 0;
            LC_TRY(launch_bn_act(e, st));
        }
        cur = ws + bl.out_off;
    }
    if (lazy && n->n_deferred > 0) {      // scale / shift / mean / invstd of the deferred layers for the backward pass, and their running statistics
        bn_finalize_layers_kernel<<<n->n_deferred, 256, 0, st>>>(n->d_fintab, params, rstat, ws, batch, kBnMomentum, kBnEps, (update_running && rstat) ? 1 : 0);
        LC_TRY(lc_launch_status());
    }
    n->launches_fwd = launches;
    return LC_OK;
}

int lc_resnet_backward(lc_resnet* n, const float* x, int batch, const float* params, float* ws, float* grads, lc_stream_t stream) {
    LC_CHECK_ARG(n && x && params && ws && grads && batch >= 1 && batch <= n->max_batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (n->mode == 1 && n->fused) return resnet_backward_fused(n, x, batch, params, ws, grads, st);
    cudaStream_t sw = n->overlap ? n->side : st;           // weight-gradient chain
    int launches = 0;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws + n->off_counters);
    int* err_flag = reinterpret_cast<int*>(counters) + 8;
    float* packed = ws + n->off_packed;
    float* wpart = ws + n->off_wpart;
    float* Tdy[2] = {ws + n->off_T1, ws + n->off_T1b};     // d(conv output), double-buffered: the weight gradient of layer i reads buffer k while
    float* T2 = ws + n->off_T2;                            // the BN backward of layer i-1 already fills buffer k^1
    float* T3 = ws + n->off_T3;
    int k = 0;
    bool wrec[2] = {false, false}, w3rec = false;

    auto bn_bwd = [&](const ConvL& c, const float* g, const float* mask_src, int mask_mode, float* dy, float* g_out) {
        BnBwdArgs a{};
        const float* aff = ws + c.aff_off;
        a.g = g; a.mask_src = mask_src; a.mask_mode = mask_mode; a.y = ws + c.y_off;
        a.scale = aff; a.shift = aff + c.cout; a.mean = aff + 2 * c.cout; a.invstd = aff + 3 * c.cout;
        a.partial = ws + n->off_bpart; a.counter = counters + 1; a.coef = ws + n->off_coef;
        a.dgamma = grads + c.gamma_off; a.dbeta = grads + c.beta_off; a.dy = dy; a.g_out = g_out;
        a.npix = (long long)batch * c.wo * c.wo; a.C = c.cout;
        return launch_bn_bwd(a, st);
    };
    // main chain: buffer k may be overwritten once the weight-gradient kernel that last read it has finished
    auto dy_acquire = [&]() -> int {
        if (n->overlap && wrec[k] && cudaStreamWaitEvent(st, n->ev_w[k], 0) != cudaSuccess) return LC_ERR_CUDA;
        return LC_OK;
    };
    // buffer k is complete on the main chain: let the weight-gradient chain read it
    auto dy_publish = [&]() -> int {
        if (n->overlap && (cudaEventRecord(n->ev_dy[k], st) != cudaSuccess || cudaStreamWaitEvent(sw, n->ev_dy[k], 0) != cudaSuccess)) return LC_ERR_CUDA;
        return LC_OK;
    };
    auto w_done = [&]() -> int {
        if (n->overlap) { if (cudaEventRecord(n->ev_w[k], sw) != cudaSuccess) return LC_ERR_CUDA; wrec[k] = true; }
        return LC_OK;
    };
    auto wgrad3x3 = [&](const ConvL& c, const float* in, const float* dy, const float* pro_scale, const float* pro_shift, bool nchw) -> int {
        if (n->mode == 1 && c.wtf_off >= 0) {
            tc::WgradTcArgs w{};
            w.in = in; w.dy = dy; w.partial = wpart + c.part_off; w.B = batch; w.error_flag = err_flag; w.pro_scale = pro_scale; w.pro_shift = pro_shift;
            return launch_wgrad3x3_tc(c.cin, c.wo, w, wgrad_nsplit_tc(c.cin), sw);
        }
        WgradArgs w{};
        w.in = in; w.dy = dy; w.partial = wpart + c.part_off; w.B = batch; w.nsplit = c.nsplit; w.pro_scale = pro_scale; w.pro_shift = pro_shift;
        return launch_wgrad3x3(c.cin, c.cout, c.wo, c.stride, nchw, w, sw);
    };

    for (int bi = (int)n->blocks.size() - 1; bi >= 0; --bi) {
        const BlockL& bl = n->blocks[bi];
        const ConvL& ca = n->convs[bl.conv_a];
        const ConvL& cb = n->convs[bl.conv_b];
        float* G = ws + n->off_G[bl.stage];
        const float* blk_in = bi == 0 ? ws + n->off_a0 : ws + n->blocks[bi - 1].out_off;
        const float* blk_out = ws + bl.out_off;
        // bn_b (+ ReLU of the block output): dy = d(y2), G <- masked gradient (the residual-branch gradient)
        const bool no_relu = !n->last_relu && bi == (int)n->blocks.size() - 1;
        LC_TRY(dy_acquire());
        LC_TRY(bn_bwd(cb, G, blk_out, no_relu ? LC_MASK_NONE : LC_MASK_FROM_OUT, Tdy[k], no_relu ? nullptr : G)); ++launches;
        LC_TRY(dy_publish());
        if (bl.conv_d >= 0) {
            const ConvL& cd = n->convs[bl.conv_d];
            if (n->overlap && w3rec && cudaStreamWaitEvent(st, n->ev_w3, 0) != cudaSuccess) return LC_ERR_CUDA;   // the previous reader of T3
            LC_TRY(bn_bwd(cd, G, nullptr, LC_MASK_NONE, T3, nullptr)); ++launches;      // T3 is written at the two stage transitions only
            if (n->overlap && (cudaEventRecord(n->ev_dy3, st) != cudaSuccess || cudaStreamWaitEvent(sw, n->ev_dy3, 0) != cudaSuccess)) return LC_ERR_CUDA;
            LC_TRY(launch_conv1x1_wgrad(cd.cin, cd.cout, cd.wo, blk_in, T3, wpart + cd.part_off, batch, sw));
            if (n->overlap) { if (cudaEventRecord(n->ev_w3, sw) != cudaSuccess) return LC_ERR_CUDA; w3rec = true; }
        }
        // conv_b: weight gradient (input = relu(bn_a(y1)) recomputed on load) on the side chain, data gradient on the main chain
        LC_TRY(wgrad3x3(cb, ws + ca.y_off, Tdy[k], ws + ca.aff_off, ws + ca.aff_off + ca.cout, false));
        LC_TRY(w_done());
        if (n->mode == 1 && cb.wtd_off >= 0) {
            tc::ConvTcArgs a{};
            a.in = Tdy[k]; a.wtc = packed + cb.wtd_off; a.out = T2; a.B = batch; a.error_flag = err_flag;
            LC_TRY(launch_conv3x3_tc(cb.cin, cb.wo, a, st));
        } else {
            Conv3x3Args a{};
            a.in = Tdy[k]; a.wpack = packed + cb.wd_off; a.out = T2; a.B = batch;
            LC_TRY(launch_conv3x3(cb.cout, cb.cin, cb.wo, 1, false, false, a, st));
        }
        k ^= 1;
        // bn_a (+ ReLU): dy = d(y1)
        LC_TRY(dy_acquire());
        LC_TRY(bn_bwd(ca, T2, nullptr, LC_MASK_FROM_BN, Tdy[k], nullptr)); ++launches;
        LC_TRY(dy_publish());
        LC_TRY(wgrad3x3(ca, blk_in, Tdy[k], nullptr, nullptr, false));
        LC_TRY(w_done());
        {
            Conv3x3Args a{};
            a.in = Tdy[k]; a.wpack = packed + ca.wd_off; a.B = batch;
            if (ca.stride == 1 && n->mode == 1 && ca.wtd_off >= 0) {
                tc::ConvTcArgs t{};
                t.in = Tdy[k]; t.wtc = packed + ca.wtd_off; t.out = G; t.addend = G; t.B = batch; t.error_flag = err_flag;
                LC_TRY(launch_conv3x3_tc(ca.cin, ca.wo, t, st));
            } else if (ca.stride == 1) {
                a.out = G; a.addend = G;     // identity shortcut: dX = dgrad + masked G (in place)
                LC_TRY(launch_conv3x3(ca.cout, ca.cin, ca.wo, 1, false, false, a, st));
            } else {
                float* Gprev = ws + n->off_G[bl.stage - 1];
                a.out = Gprev;
                LC_TRY(launch_conv3x3(ca.cout, ca.cin, ca.wo * 2, 1, true, false, a, st));
                const ConvL& cd = n->convs[bl.conv_d];
                LC_TRY(launch_conv1x1_dgrad(cd.cin, cd.cout, cd.wo, T3, params + cd.w_off, Gprev, batch, st));
            }
        }
        k ^= 1;
    }
    {   // stem
        const ConvL& c = n->convs[0];
        float* G = ws + n->off_G[0];
        LC_TRY(dy_acquire());
        LC_TRY(bn_bwd(c, G, ws + n->off_a0, LC_MASK_FROM_OUT, Tdy[k], nullptr)); ++launches;
        LC_TRY(dy_publish());
        WgradArgs w{};
        w.in = x; w.dy = Tdy[k]; w.partial = wpart + c.part_off; w.B = batch; w.nsplit = c.nsplit;
        LC_TRY(launch_wgrad3x3(c.cin, c.cout, c.wo, 1, true, w, sw));
    }
    if (n->overlap && (cudaEventRecord(n->ev_join, sw) != cudaSuccess || cudaStreamWaitEvent(st, n->ev_join, 0) != cudaSuccess)) return LC_ERR_CUDA;
    wgrad_reduce_all_kernel<<<n->tab_blocks, 256, 0, st>>>(n->d_tab, (int)n->convs.size(), wpart, grads, n->mode);
    LC_TRY(lc_launch_status());
    n->launches_bwd = launches;
    return LC_OK;
}

// ---- head / loss ----------------------------------------------------------------------------------------------------
int lc_head_forward(const float* act, int batch, int hw, int feat_dim, const float* W, const float* bias, int ncls, float* feat, float* logits,
                    int ldl, lc_stream_t stream) {
    LC_CHECK_ARG(act && W && feat && logits && batch >= 1 && hw >= 1 && ncls >= 1 && ldl >= ncls && feat_dim == 64);
    avgpool_fc_fwd_kernel<64><<<batch, 64, 0, (cudaStream_t)stream>>>(act, hw, W, bias, ncls, feat, logits, ldl);
    return lc_launch_status();
}

int lc_loss_ce_kd(const float* logits, int ldl, const float* teacher, int ldt, const int64_t* y, int batch, int ce_lo, int ce_hi, int kd_n,
                  float kd_w, float T, int pred_n, float* dlogits, int64_t* pred, float* scal, lc_stream_t stream) {
    LC_CHECK_ARG(logits && y && dlogits && pred && scal && batch >= 1 && ce_lo >= 0 && ce_hi > ce_lo && ce_hi <= ldl && pred_n >= 1 && pred_n <= ldl);
    LC_CHECK_ARG(kd_n == 0 || (teacher && kd_n <= ldl && kd_n <= ldt && T > 0.f));
    LossArgs a{};
    a.logits = logits; a.teacher = teacher; a.y = reinterpret_cast<const long long*>(y); a.dlogits = dlogits;
    a.pred = reinterpret_cast<long long*>(pred); a.scal = scal; a.B = batch; a.ldl = ldl; a.ldt = ldt; a.ncols = ldl;
    a.ce_lo = ce_lo; a.ce_hi = ce_hi; a.kd_n = kd_n; a.pred_n = pred_n; a.kd_w = kd_w; a.T = T;
    ce_kd_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_loss_ce_masked(const float* logits, int ldl, const int64_t* y, int batch, int lo, int hi, const float* extra, float extra_coeff, float* dlogits,
                      int64_t* pred, float* scal, lc_stream_t stream) {
    LC_CHECK_ARG(logits && y && dlogits && pred && scal && batch >= 1 && lo >= 0 && hi > lo && hi <= ldl);
    LossArgs a{};
    a.logits = logits; a.y = reinterpret_cast<const long long*>(y); a.dlogits = dlogits; a.pred = reinterpret_cast<long long*>(pred); a.scal = scal;
    a.B = batch; a.ldl = ldl; a.ncols = ldl; a.ce_lo = lo; a.ce_hi = hi; a.pred_lo = lo; a.pred_n = hi; a.extra = extra; a.extra_coeff = extra_coeff;
    ce_kd_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
    return lc_launch_status();
}

int lc_head_backward(const float* dlogits, int ldl, const float* feat, const float* W, int ncls, int batch, int feat_dim, float* dW, float* db,
                     float* dfeat, float* gact, int hw, lc_stream_t stream) {
    LC_CHECK_ARG(dlogits && feat && W && dW && dfeat && ncls >= 1 && batch >= 1 && feat_dim == 64 && ldl >= ncls);
    head_bwd_kernel<64><<<ncls + batch, 64, 0, (cudaStream_t)stream>>>(dlogits, ldl, feat, W, ncls, batch, dW, db, dfeat, gact, hw);
    return lc_launch_status();
}

int lc_avgpool_forward(const float* act, int batch, int hw, int feat_dim, float* feat, lc_stream_t stream) {
    LC_CHECK_ARG(act && feat && batch >= 1 && hw >= 1 && feat_dim == 64);
    avgpool_fwd_kernel<64><<<batch, 64, 0, (cudaStream_t)stream>>>(act, hw, feat);
    return lc_launch_status();
}
int lc_avgpool_backward(const float* dfeat, int batch, int hw, int feat_dim, float* gact, lc_stream_t stream) {
    LC_CHECK_ARG(dfeat && gact && batch >= 1 && hw >= 1 && feat_dim == 64);
    avgpool_bwd_kernel<64><<<batch, 64, 0, (cudaStream_t)stream>>>(dfeat, hw, gact);
    return lc_launch_status();
}

// ---- flat arena ops ---------------------------------------------------------------------------------------------------
int lc_ewc_penalty_grad(const float* theta, const float* theta_ref, const float* fisher, float* grad, long long n, const float* hp_lamda,
                        float* scratch, uint32_t* counter, float* scal, lc_stream_t stream) {
    LC_CHECK_ARG(theta && theta_ref && fisher && grad && n > 0 && hp_lamda && scratch && counter && scal);
    LC_CHECK_ARG(((uintptr_t)theta % 16 == 0) && ((uintptr_t)theta_ref % 16 == 0) && ((uintptr_t)fisher % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)scratch % 8 == 0));
    ewc_penalty_grad_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(theta, theta_ref, fisher, grad, n, hp_lamda, reinterpret_cast<double*>(scratch), counter, scal);
    return lc_launch_status();
}
int lc_fisher_accumulate(float* fisher, const float* grad, long long n, float weight, lc_stream_t stream) {
    LC_CHECK_ARG(fisher && grad && n > 0);
    fisher_accumulate_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(fisher, grad, n, weight);
    return lc_launch_status();
}
int lc_fisher_merge(float* f_new, const float* f_old, long long n, float num_samples, float alpha, lc_stream_t stream) {
    LC_CHECK_ARG(f_new && n > 0 && num_samples > 0.f);
    fisher_merge_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(f_new, f_old, n, num_samples, alpha);
    return lc_launch_status();
}
int lc_sgd_momentum(float* p, const float* g, float* m, long long n, const float* hp, lc_stream_t stream) {
    LC_CHECK_ARG(p && g && m && hp && n > 0);
    if (((uintptr_t)p % 16) || ((uintptr_t)g % 16) || ((uintptr_t)m % 16)) {
        // a range that does not start on a 16-byte boundary (e.g. the bias slice of a 10-class task head inside a flat arena): scalar kernel
        sgd_momentum_frozen_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, n, hp, 0, 0);
        return lc_launch_status();
    }
    sgd_momentum_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, n, hp);
    return lc_launch_status();
}
int lc_sgd_momentum_frozen(float* p, const float* g, float* m, long long n, const float* hp, long long freeze_lo, long long freeze_hi,
                           lc_stream_t stream) {
    LC_CHECK_ARG(p && g && m && hp && n > 0 && freeze_lo >= 0 && freeze_hi >= freeze_lo && freeze_hi <= n);
    sgd_momentum_frozen_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, n, hp, freeze_lo, freeze_hi);
    return lc_launch_status();
}
int lc_adam(float* p, const float* g, float* m, float* v, long long n, const float* hp, lc_stream_t stream) {
    LC_CHECK_ARG(p && g && m && v && hp && n > 0);
    adam_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hp);
    return lc_launch_status();
}
int lc_adam_tick(float* hp, lc_stream_t stream) {
    LC_CHECK_ARG(hp != nullptr && ((uintptr_t)hp % 8 == 0));
    adam_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hp);
    return lc_launch_status();
}
int lc_clip_grad_norm(float* g, long long n, float max_norm, float* scratch, float* norm_out, lc_stream_t stream) {
    LC_CHECK_ARG(g && n > 0 && scratch && ((uintptr_t)scratch % 8 == 0));
    double* part = reinterpret_cast<double*>(scratch);
    sqnorm_partial_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(g, n, part);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    clip_scale_kernel<<<kFlatBlocks, 256, 0, (cudaStream_t)stream>>>(g, n, part, kFlatBlocks, max_norm, norm_out);
    return lc_launch_status();
}

// ---- per-kernel entry points ------------------------------------------------------------------------------------------
// scratch layout of the per-kernel entry points (floats): [0,64) election counters (caller-zeroed) | [64,96) one table
// entry | [96, ...) packed weights / partial sums
static_assert(sizeof(ConvTabEntry) <= 32 * sizeof(float), "table slot");
constexpr int kOpData = 96;
static inline long long round4(long long v) { return (v + 3) / 4 * 4; }

long long lc_conv_scratch_floats(int batch, int cin, int cout, int width_out) {
    const long long ne = (long long)cout * cin * 9;
    const int cm = cout > cin ? cout : cin;
    long long parts = (long long)batch * 8 * 2 * cm;
    const long long p1 = ((long long)batch * width_out * width_out + 127) / 128 * 2 * cm;
    if (p1 > parts) parts = p1;
    return kOpData + round4(2 * ne) + round4(parts) + ne * 256 + 64;
}

static void fill_stat(BnStatArgs& s, const float* gamma, const float* beta, float* rstat, float* stat_out, int C, float* partial, unsigned int* counter) {
    s.partial = partial; s.counter = counter; s.gamma = gamma; s.beta = beta;
    s.running_mean = rstat; s.running_var = rstat ? rstat + C : nullptr;
    s.scale = stat_out; s.shift = stat_out + C; s.mean = stat_out + 2 * C; s.invstd = stat_out + 3 * C;
    s.momentum = kBnMomentum; s.eps = kBnEps; s.update_running = rstat ? 1 : 0;
}

int lc_conv3x3(const float* in, const float* w_oihw, float* out, int batch, int cin, int cout, int width_out, int stride, int mode, int in_nchw,
               const float* pro_scale, const float* pro_shift, const float* addend, const float* gamma, const float* beta, float* rstat,
               float* stat_out, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(in && w_oihw && out && scratch && batch >= 1 && (mode == 0 || mode == 1) && ((uintptr_t)scratch % 16 == 0));
    cudaStream_t st = (cudaStream_t)stream;
    const long long ne = (long long)cout * cin * 9;
    ConvTabEntry t{};
    t.w_off = 0; t.wf_off = 0; t.wd_off = ne; t.wtf_off = -1; t.wtd_off = -1; t.cout = cout; t.cin = cin; t.ntap = 9; t.nsplit = 0; t.blk_begin = 0;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* packed = scratch + kOpData;
    float* partial = packed + round4(2 * ne);
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    pack_weights_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, w_oihw, packed);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    Conv3x3Args a{};
    a.in = in; a.out = out; a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.addend = addend; a.B = batch;
    if (mode == 0) {
        a.wpack = packed;
        if (stat_out) { LC_CHECK_ARG(gamma && beta); fill_stat(a.stat, gamma, beta, rstat, stat_out, cout, partial, reinterpret_cast<unsigned int*>(scratch)); }
        return launch_conv3x3(cin, cout, width_out, stride, false, in_nchw != 0, a, st);
    }
    // data gradient of a (cin -> cout, stride) conv: `in` is dy [B][wo][wo][cout]; out is [B][wo*stride][wo*stride][cin]
    a.wpack = packed + ne;
    return launch_conv3x3(cout, cin, width_out * stride, 1, stride == 2, false, a, st);
}

// Single launch of the conv kernel on pre-packed weights ([cin][9][cout], as left at scratch+96 by lc_conv3x3): no packing,
// no statistics.  Used by bench.py to time the dominant kernel in isolation.
int lc_conv3x3_packed(const float* in, const float* wpack, float* out, int batch, int cin, int cout, int width_out, int stride,
                      const float* pro_scale, const float* pro_shift, lc_stream_t stream) {
    LC_CHECK_ARG(in && wpack && out && batch >= 1);
    Conv3x3Args a{};
    a.in = in; a.wpack = wpack; a.out = out; a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.B = batch;
    return launch_conv3x3(cin, cout, width_out, stride, false, false, a, (cudaStream_t)stream);
}

// Tensor-core (tcgen05 kind::tf32) version of lc_conv3x3 for the square stride-1 layers: (c, width) in {(16,32),(32,16),(64,8)}.
// mode 0 forward, 1 data gradient.  scratch as for lc_conv3x3; scratch[8] (int) receives 1 if the MMA barrier timed out.
int lc_conv3x3_tc(const float* in, const float* w_oihw, float* out, int batch, int c, int width, int mode, const float* pro_scale,
                  const float* pro_shift, const float* addend, const float* gamma, const float* beta, float* rstat, float* stat_out,
                  float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(in && w_oihw && out && scratch && batch >= 1 && (mode == 0 || mode == 1) && ((uintptr_t)scratch % 16 == 0));
    LC_CHECK_ARG(tc_eligible(c, c, width, 1, 3));
    cudaStream_t st = (cudaStream_t)stream;
    const long long ne = (long long)c * c * 9;
    ConvTabEntry t{};
    t.w_off = 0; t.wf_off = 0; t.wd_off = ne; t.wtf_off = 2 * ne; t.wtd_off = 3 * ne; t.cout = c; t.cin = c; t.ntap = 9; t.blk_begin = 0;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* packed = scratch + kOpData;
    float* partial = packed + round4(4 * ne);
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    pack_weights_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, w_oihw, packed);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    tc::ConvTcArgs a{};
    a.in = in; a.out = out; a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.addend = addend; a.B = batch;
    a.error_flag = reinterpret_cast<int*>(scratch) + 8;
    a.wtc = packed + (mode == 0 ? 2 * ne : 3 * ne);
    if (mode == 0 && stat_out) { LC_CHECK_ARG(gamma && beta); fill_stat(a.stat, gamma, beta, rstat, stat_out, c, partial, reinterpret_cast<unsigned int*>(scratch)); }
    return launch_conv3x3_tc(c, width, a, st, mode == 0 && conv_persist_enabled());
}
// Stride-2 tensor-core forward conv (cin -> 2*cin, output width wo) with optional BatchNorm statistics of the output; scratch as lc_conv3x3
// (>= lc_conv_scratch_floats(batch, cin, 2*cin, wo), first 64 words zero); scratch word 8 (int) is set to 1 if the MMA completion barrier timed out.
int lc_conv3x3s2_tc(const float* in, const float* w_oihw, float* out, int batch, int cin, int width_out, const float* gamma, const float* beta,
                    float* rstat, float* stat_out, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(in && w_oihw && out && scratch && batch >= 1 && ((uintptr_t)scratch % 16 == 0) && s2_tc_eligible(cin, 2 * cin, width_out, 2, 3));
    cudaStream_t st = (cudaStream_t)stream;
    const int cout = 2 * cin;
    const long long ne = (long long)cout * cin * 9;
    ConvTabEntry t{};
    t.w_off = 0; t.wf_off = 0; t.wd_off = -1; t.wtf_off = -1; t.wtd_off = -1; t.wts_off1 = ne + 1; t.cout = cout; t.cin = cin; t.ntap = 9; t.blk_begin = 0;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* packed = scratch + kOpData;
    float* partial = packed + round4(2 * ne);
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    pack_weights_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, w_oihw, packed);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    tc::ConvS2Args a{};
    a.in = in; a.wtc = packed + ne; a.out = out; a.B = batch; a.error_flag = reinterpret_cast<int*>(scratch) + 8;
    if (stat_out) { LC_CHECK_ARG(gamma && beta); fill_stat(a.stat, gamma, beta, rstat, stat_out, cout, partial, reinterpret_cast<unsigned int*>(scratch)); }
    return launch_conv3x3s2_tc(cin, width_out, a, st);
}
// Data gradient of the same conv: dy [B][wo][wo][2*cin] -> dx [B][2*wo][2*wo][cin] (every element written).  scratch as lc_conv3x3s2_tc.
int lc_conv3x3s2_dgrad_tc(const float* dy, const float* w_oihw, float* dx, int batch, int cin, int width_out, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(dy && w_oihw && dx && scratch && batch >= 1 && ((uintptr_t)scratch % 16 == 0) && s2_tc_eligible(cin, 2 * cin, width_out, 2, 3));
    cudaStream_t st = (cudaStream_t)stream;
    const int cout = 2 * cin;
    const long long ne = (long long)cout * cin * 9;
    ConvTabEntry t{};
    t.w_off = 0; t.wf_off = 0; t.wd_off = -1; t.wtf_off = -1; t.wtd_off = -1; t.wts_off1 = ne + 1; t.wtsd_off1 = 2 * ne + 1; t.cout = cout; t.cin = cin; t.ntap = 9;
    t.blk_begin = 0;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* packed = scratch + kOpData;      // [wf | wts | wtsd]: 3 * ne floats (the scratch holds >= 2 * ne + ne * 256)
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    pack_weights_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, w_oihw, packed);
    if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
    tc::DgradS2Args a{};
    a.dy = dy; a.wtc = packed + 2 * ne; a.out = dx; a.B = batch; a.error_flag = reinterpret_cast<int*>(scratch) + 8;
    return launch_dgrad3x3s2_tc(cin, width_out, a, st);
}
// One launch of the tensor-core conv on pre-packed TF32 weights ([9][c/4][c][4], as left at scratch+96+2*9*c*c by lc_conv3x3_tc):
// no packing, no statistics.  Used by bench.py to time the dominant kernel in isolation.
int lc_conv3x3_tc_packed(const float* in, const float* wtc, float* out, int batch, int c, int width, const float* pro_scale,
                         const float* pro_shift, int* error_flag, lc_stream_t stream) {
    LC_CHECK_ARG(in && wtc && out && batch >= 1 && tc_eligible(c, c, width, 1, 3));
    tc::ConvTcArgs a{};
    a.in = in; a.wtc = wtc; a.out = out; a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.B = batch; a.error_flag = error_flag;
    return launch_conv3x3_tc(c, width, a, (cudaStream_t)stream, conv_persist_enabled() != 0);
}
long long lc_conv_tc_scratch_floats(int batch, int c, int width) {
    const long long ne = (long long)c * c * 9;
    return kOpData + round4(4 * ne) + tc::conv_tc_tiles(batch, width) * 2 * c + 64;
}

int lc_conv3x3_wgrad(const float* in, const float* dy, float* dw, int batch, int cin, int cout, int width_out, int stride, int in_nchw,
                     const float* pro_scale, const float* pro_shift, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(in && dy && dw && scratch && batch >= 1 && ((uintptr_t)scratch % 16 == 0));
    cudaStream_t st = (cudaStream_t)stream;
    const long long ne = (long long)cout * cin * 9;
    const int nsplit = wgrad_nsplit(cin, cout);
    ConvTabEntry t{};
    t.w_off = 0; t.part_off = 0; t.cout = cout; t.cin = cin; t.ntap = 9; t.nsplit = nsplit; t.blk_begin = 0; t.wd_off = -1; t.wtf_off = -1; t.wtd_off = -1;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* partial = scratch + kOpData;
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    WgradArgs w{};
    w.in = in; w.dy = dy; w.partial = partial; w.pro_scale = pro_scale; w.pro_shift = pro_shift; w.B = batch; w.nsplit = nsplit;
    int e = launch_wgrad3x3(cin, cout, width_out, stride, in_nchw != 0, w, st);
    if (e != LC_OK) return e;
    wgrad_reduce_all_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, partial, dw, 0);
    return lc_launch_status();
}

// Tensor-core (tcgen05 kind::tf32, MN-major operands) weight gradient of the square stride-1 layers; same contract as
// lc_conv3x3_wgrad.  scratch word 8 (int) receives 1 if an MMA barrier timed out.
int lc_conv3x3_wgrad_tc(const float* in, const float* dy, float* dw, int batch, int c, int width, const float* pro_scale,
                        const float* pro_shift, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(in && dy && dw && scratch && batch >= 1 && ((uintptr_t)scratch % 16 == 0) && tc_eligible(c, c, width, 1, 3));
    cudaStream_t st = (cudaStream_t)stream;
    const long long ne = (long long)c * c * 9;
    const int nsplit = wgrad_nsplit_tc(c);
    ConvTabEntry t{};
    t.w_off = 0; t.part_off = 0; t.cout = c; t.cin = c; t.ntap = 9; t.nsplit = nsplit; t.nsplit_tc = nsplit; t.blk_begin = 0; t.wd_off = -1; t.wtf_off = 0; t.wtd_off = -1;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    float* partial = scratch + kOpData;
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    tc::WgradTcArgs w{};
    w.in = in; w.dy = dy; w.partial = partial; w.pro_scale = pro_scale; w.pro_shift = pro_shift; w.B = batch;
    w.error_flag = reinterpret_cast<int*>(scratch) + 8;
    int e = launch_wgrad3x3_tc(c, width, w, nsplit, st);
    if (e != LC_OK) return e;
    wgrad_reduce_all_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, partial, dw, 1);
    return lc_launch_status();
}

int lc_conv1x1s2(const float* a_, const float* b_, float* out, int batch, int cin, int cout, int width_out, int mode, const float* gamma,
                 const float* beta, float* rstat, float* stat_out, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(a_ && b_ && out && scratch && batch >= 1 && mode >= 0 && mode <= 2 && ((uintptr_t)scratch % 16 == 0));
    cudaStream_t st = (cudaStream_t)stream;
    const long long ne = (long long)cout * cin;
    ConvTabEntry t{};
    t.w_off = 0; t.wf_off = 0; t.wd_off = -1; t.wtf_off = -1; t.wtd_off = -1; t.part_off = 0; t.cout = cout; t.cin = cin; t.ntap = 1; t.nsplit = k1x1Split; t.blk_begin = 0;
    ConvTabEntry* d_t = reinterpret_cast<ConvTabEntry*>(scratch + 64);
    if (cudaMemcpyAsync(d_t, &t, sizeof(t), cudaMemcpyHostToDevice, st) != cudaSuccess) return LC_ERR_CUDA;
    float* buf = scratch + kOpData;
    if (mode == 0) {          // a_ = in, b_ = W [cout][cin]
        pack_weights_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, b_, buf);
        if (lc_launch_status() != LC_OK) return LC_ERR_CUDA;
        Conv1x1Args a{};
        a.in = a_; a.w = buf; a.out = out; a.B = batch;
        if (stat_out) { LC_CHECK_ARG(gamma && beta); fill_stat(a.stat, gamma, beta, rstat, stat_out, cout, buf + round4(ne), reinterpret_cast<unsigned int*>(scratch)); }
        return launch_conv1x1_fwd(cin, cout, width_out, a, st);
    }
    if (mode == 1)            // a_ = dy, b_ = W; accumulate into out
        return launch_conv1x1_dgrad(cin, cout, width_out, a_, b_, out, batch, st);
    // mode 2: a_ = in, b_ = dy -> out = dW
    int e = launch_conv1x1_wgrad(cin, cout, width_out, a_, b_, buf, batch, st);
    if (e != LC_OK) return e;
    wgrad_reduce_all_kernel<<<(int)((ne + 255) / 256), 256, 0, st>>>(d_t, 1, buf, out, 0);
    return lc_launch_status();
}

int lc_bn_act_forward(const float* y, const float* scale, const float* shift, const float* res, const float* res_scale, const float* res_shift,
                      float* out, long long npix, int C, lc_stream_t stream) {
    LC_CHECK_ARG(y && scale && shift && out && npix > 0 && C % 4 == 0);
    BnActArgs e{};
    e.y = y; e.scale = scale; e.shift = shift; e.res = res; e.res_scale = res_scale; e.res_shift = res_shift; e.out = out; e.n4 = npix * C / 4; e.C = C;
    return launch_bn_act(e, (cudaStream_t)stream);
}

int lc_bn_backward(const float* g, const float* mask_src, int mask_mode, const float* y, const float* stat, float* dy, float* g_out,
                   float* dgamma, float* dbeta, long long npix, int C, float* scratch, lc_stream_t stream) {
    LC_CHECK_ARG(g && y && stat && dy && dgamma && dbeta && scratch && npix > 0 && (mask_mode != LC_MASK_FROM_OUT || mask_src));
    BnBwdArgs a{};
    a.g = g; a.mask_src = mask_src; a.mask_mode = mask_mode; a.y = y;
    a.scale = stat; a.shift = stat + C; a.mean = stat + 2 * C; a.invstd = stat + 3 * C;
    a.counter = reinterpret_cast<unsigned int*>(scratch); a.coef = scratch + 64; a.partial = scratch + 64 + 3 * C + (4 - (3 * C) % 4) % 4;
    a.dgamma = dgamma; a.dbeta = dbeta; a.dy = dy; a.g_out = g_out; a.npix = npix; a.C = C;
    return launch_bn_bwd(a, (cudaStream_t)stream);
}

}  // extern "C"
