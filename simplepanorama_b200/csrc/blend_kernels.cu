// blend_kernels.cu -- the reference's multiband blend as fused separable-filter kernels (sm_100a).
//
// Replaces blnd::multi_blend (reference src/math/_blending.cpp:186-252), the per-pixel loops of
// imgm::elementwiseOperation (src/math/_img_manipulation.cpp:31-84) and the x255/convertTo tail of
// stitch_parameters::blend (src/classes/_panorama.cpp:242-249).
//
// What the reference computes (NOT a decimated Laplacian pyramid -- see SURVEY.md section 0.3):
//   for band i in 0..B-1:  s_i = sqrt(2(B-i)+1)*sigma, kernel size 2*ceil(3 sigma)+1 for EVERY band
//     for image j:         G_i = GaussianBlur(I_j, s_i),  W_i = GaussianBlur(mask_cut_j, s_i)/255
//                          band = (i==B-1) ? I - G_i : (i>0 ? G_i - G_{i+1} : G_0)
//                          W_i = 0 where validity mask != 255
//                          color += band*W_i ; alpha += W_i          (in the tile's canvas ROI)
//   out = color / clamp(alpha) / float(255/B)        [integer division]      (-> x255 -> u8)
// BORDER_REFLECT applies at the borders of each warped TILE.
//
// Kernels: the marching-strip kernel of blend_march.cuh (radius 21, i.e. sigma = 7: the reference default and every
// BASELINE config) and a block-tiled kernel for any other radius <= 32 (also the on-device cross-check of the first).
// Band algebra, weights, validity zeroing and the sum over bands happen in registers / shared memory; the only HBM
// traffic is the u8 inputs (5 B/tile-px, halo re-reads come from L2) and one float4 read-modify-write of the canvas
// accumulator per tile pixel.  No level ever round-trips to HBM.
//
// This stage is bound by the FP32 FMA pipe, not by HBM (4 ch x B x 2 x 43 MACs per tile pixel against ~37 B): see
// DESIGN.md.  The Gaussian taps travel in the kernel parameters (constant bank), one set per launch: contexts share no
// mutable device state.
#include "spano_internal.h"
#include <cmath>
#include <cstring>
#include <cstdio>
#include <mutex>

namespace {

constexpr int MAXB = SPANO_MAX_BANDS;
constexpr int MAXR = SPANO_BLUR_RADIUS_MAX;

// Tap slots of the marching kernel (see blend_march.cuh): c_taps[s + b][k] = tap of band b at distance k from the
// centre; c_tap2[s + b][d] = {T[d], T[d-1]} with T[d] = tap at |d - 21| for d in 0..42 and T[-1] = T[43] = 0 (the tap
// pair that output rows o, o+1 of the vertical pass apply to input row o + d).
// A slot is a run of `bands` consecutive rows of the two tables (slot = index of its first row).
constexpr int TAP_ROWS = 120;   // 120 * (96 + 352) B = 53.8 KB of the 64 KB constant space
constexpr int TAP_SLOTS = 32;
__constant__ __align__(16) float c_taps[TAP_ROWS][24];
__constant__ __align__(16) float2 c_tap2[TAP_ROWS][44];

struct BlendParams {
    const uint8_t *tile;  size_t tile_step;
    const uint8_t *cut;   size_t cut_step;
    const uint8_t *valid; size_t valid_step;
    int w, h;          // tile extent
    int ty_begin, ty_end; // tile rows to produce
    int wx0, wx1;      // tile columns to produce: the part of the tile inside the accumulator's columns (reads go 21 px beyond)
    float4 *acc;       // canvas accumulator rows [row0, ...), pitch canvas_w
    int canvas_w;
    int ax, ay;        // tile corner relative to acc origin (canvas x, canvas y - row0)
    int radius;
    const int *plan = nullptr; // marching kernel: plan made beforehand (launch_blend_plan), else made at launch
    float taps[MAXB][MAXR + 1]; // taps[b][k] = tap of band b at distance k from the centre (generic kernel)
};

// cv::borderInterpolate(p, len, BORDER_REFLECT)
__device__ __forceinline__ int reflect_idx(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// channel 0 = mask_cut, 1..3 = B,G,R of the tile
__device__ __forceinline__ float load_channel(const BlendParams &P, int ch, int x, int y)
{
    if (ch == 0) return (float)__ldg(P.cut + (size_t)y * P.cut_step + x);
    return (float)__ldg(P.tile + (size_t)y * P.tile_step + (size_t)x * 3 + (ch - 1));
}

constexpr int FR = 21;           // radius of the marching kernel

#include "blend_march.cuh"
#include "blend_ws.cuh"

// ---------------------------------------------------------------------------------------------
// Generic path: any radius <= 32 (other sigma values).  Same arithmetic, runtime loops,
// 32x32 block, one sigma at a time.  Also the on-device cross-check of the fast kernel.
// ---------------------------------------------------------------------------------------------
constexpr int GB = 32;
constexpr int GTHREADS = 256;

template <int B>
__global__ void __launch_bounds__(GTHREADS) blend_generic_kernel(const __grid_constant__ BlendParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = P.radius;
    const int IW = GB + 2 * R;           // staged width / height
    const int IP = IW + 1;               // pitch
    float *s_in = reinterpret_cast<float *>(smem_raw);   // [IW][IP]
    float *s_tmp = s_in + (size_t)IW * IP;               // [IW][GB]

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * GB;
    const int ty0 = P.ty_begin + blockIdx.y * GB;
    const int cx = tid & 31, cy = tid >> 5; // my pixels: (cx, cy + 8 j), j = 0..3

    float wgt[B][4], contrib[3][4], wsum[4], cur[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ok[j] = (tx0 + cx >= P.wx0) && (tx0 + cx < P.wx1) && (ty0 + cy + 8 * j < P.ty_end);
        wsum[j] = 0.f;
        cur[j] = 0.f;
        contrib[0][j] = contrib[1][j] = contrib[2][j] = 0.f;
    }

    for (int ch = 0; ch < 4; ++ch) {
        __syncthreads();
        for (int i = tid; i < IW * IW; i += GTHREADS) {
            const int r = i / IW, c = i - r * IW;
            s_in[r * IP + c] = load_channel(P, ch, reflect_idx(tx0 - R + c, P.w), reflect_idx(ty0 - R + r, P.h));
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < B; ++b) {
            for (int i = tid; i < IW * GB; i += GTHREADS) {
                const int r = i >> 5, c = i & 31;
                const float *p = s_in + r * IP + c + R;
                float a = P.taps[b][0] * p[0];
                for (int k = 1; k <= R; ++k) a = fmaf(P.taps[b][k], p[-k] + p[k], a);
                s_tmp[r * GB + c] = a;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ry = cy + 8 * j + R;
                const float *p = s_tmp + ry * GB + cx;
                float a = 0.f;
                for (int k = -R; k <= R; ++k) a = fmaf(P.taps[b][k < 0 ? -k : k], p[k * GB], a);
                if (ch == 0) {
                    bool keep = false;
                    if (ok[j]) keep = __ldg(P.valid + (size_t)(ty0 + cy + 8 * j) * P.valid_step + (tx0 + cx)) == 255;
                    const float wv = keep ? a * (float)(1.0 / 255.0) : 0.f;
                    wgt[b][j] = wv;
                    wsum[j] += wv;
                } else {
                    float *cc = contrib[ch - 1];
                    if (b == 0) { if (B > 1) cc[j] = a * wgt[0][j]; }
                    else if (b >= 2) cc[j] = fmaf(cur[j] - a, wgt[b - 1][j], cc[j]);
                    if (b == B - 1) {
                        const float I = s_in[ry * IP + cx + R];
                        cc[j] = fmaf(I - a, wgt[B - 1][j], cc[j]);
                    }
                    cur[j] = a;
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!ok[j]) continue;
        const int x = tx0 + cx, y = ty0 + cy + 8 * j;
        float4 *q = P.acc + (size_t)(P.ay + y) * P.canvas_w + (P.ax + x);
        float4 t = *q;
        t.x += contrib[0][j]; t.y += contrib[1][j]; t.z += contrib[2][j]; t.w += wsum[j];
        *q = t;
    }
}

// out = color / clamp(alpha) / float(255/B)  [ -> *255 -> u8 ]
__global__ void normalise_kernel(const float4 *acc, int canvas_w, int rows, float inv_div, int out_kind, void *out,
                                 size_t out_step, int col0, int col1)
{
    const int x = col0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= col1 || y >= rows) return;
    const float4 t = acc[(size_t)y * canvas_w + x];
    float d = t.w;
    d = copysignf(fmaxf(fabsf(d), 1e-6f), d);
    const float s = __fdiv_rn(1.f, d);
    const float c0 = __fmul_rn(__fmul_rn(t.x, s), inv_div);
    const float c1 = __fmul_rn(__fmul_rn(t.y, s), inv_div);
    const float c2 = __fmul_rn(__fmul_rn(t.z, s), inv_div);
    if (out_kind == SPANO_OUT_F32) {
        float *o = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(out) + (size_t)y * out_step) + 3 * (size_t)x;
        o[0] = c0; o[1] = c1; o[2] = c2;
    } else {
        unsigned char *o = reinterpret_cast<unsigned char *>(out) + (size_t)y * out_step + 3 * (size_t)x;
        o[0] = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(c0, 255.f))));
        o[1] = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(c1, 255.f))));
        o[2] = (unsigned char)min(255, max(0, __float2int_rn(__fmul_rn(c2, 255.f))));
    }
}

// ---- FP32 pipe microbenchmark (roofline denominator of the blend kernels) ----
struct PeakTaps { float t[8][16]; };
template <int VARIANT>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *sink, int iters, float seed, const __grid_constant__ PeakTaps T)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float m = 1.0000001f * seed, c = 1e-9f;
    for (int it = 0; it < iters; ++it) {
        if (VARIANT == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        } else if (VARIANT == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], T.t[i & 7][i], a[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 d;
                asm volatile("{ .reg .b64 ra, rb, rc, rd;\n"
                             "  mov.b64 ra, {%2, %3};\n"
                             "  mov.b64 rb, {%4, %4};\n"
                             "  mov.b64 rc, {%5, %5};\n"
                             "  fma.rn.f32x2 rd, ra, rb, rc;\n"
                             "  mov.b64 {%0, %1}, rd; }\n"
                             : "=f"(d.x), "=f"(d.y)
                             : "f"(a[i]), "f"(a[i + 1]), "f"(m), "f"(c));
                a[i] = d.x;
                a[i + 1] = d.y;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

struct PreviewCut {   // mask_cut at preview scale, to be up-scaled into Q.cut where the plan reads it
    const uint8_t *small;
    int sw, sh;
    size_t sstep;
};

// activity + plan kernels of one tile launch into `plan` (device, march::PlanView::ints(strips, sms) ints)
template <int SW>
int make_plan(spano_ctx *ctx, const BlendParams &Q, int sms, int *plan, const PreviewCut *pv = nullptr)
{
    const int strips = (Q.w + SW - 1) / SW;
    if (strips > 2048 || sms > 1024) return spano_fail(ctx, SPANO_E_LIMIT, "tile wider than %d px", 2048 * SW);
    const march::PlanView V(strips, sms);
    if (!ctx->blend_stats) {
        SPANO_CUDA(ctx, cudaMalloc((void **)&ctx->blend_stats, 2 * sizeof(unsigned long long)));
        SPANO_CUDA(ctx, cudaMemsetAsync(ctx->blend_stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
        ctx->owned.push_back(ctx->blend_stats);
    }
    int *ymin = plan + V.ymin(), *ymax = plan + V.ymax();
    march::plan_init_kernel<<<(strips + 255) / 256, 256, 0, ctx->stream>>>(ymin, ymax, strips);
    const int dense = ctx->opt_blend_dense;
    const int ra = std::max(0, Q.ty_begin - march::R), rb = std::min(Q.h, Q.ty_end + march::R);
    int extra = 0;
    if (!dense) {
        if (pv) {
            const int k = launch_resize_activity(ctx, pv->small, pv->sw, pv->sh, pv->sstep, Q.w, Q.h, SW, ra, rb, ymin, ymax);
            if (k < 0) return k;
            extra += k - 1;
        } else {
            dim3 ag((Q.w + 511) / 512, (rb - ra + 63) / 64);
            march::activity_kernel<SW><<<ag, 256, 0, ctx->stream>>>(Q.cut, Q.cut_step, Q.w, ra, rb, ymin, ymax);
        }
    }
    march::plan_kernel<SW><<<1, 256, 0, ctx->stream>>>(plan, strips, sms, Q.ty_begin, Q.ty_end, dense, ctx->blend_stats, Q.wx0, Q.wx1, pv ? Q.h : 0);
    SPANO_CUDA(ctx, cudaGetLastError());
    if (pv) {
        const int k = launch_resize_mask_rows(ctx, pv->small, pv->sw, pv->sh, pv->sstep, const_cast<uint8_t *>(Q.cut), Q.w, Q.h, Q.cut_step, SW, ymin, ymax);
        if (k < 0) return k;
        extra += k;
    }
    ctx->launches += (dense ? 2 : 3) + extra;
    return 0;
}

template <int B, int SW, int VT, bool UNIFORM>
int launch_ws(spano_ctx *ctx, const march::Params &P, int sms)
{
    using W = march::WsCfg<B, SW, VT>;
    static bool configured[64] = {false};
    const int dev = ctx->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(march::blend_ws_kernel<B, SW, VT, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::SMEM);
        if (e != cudaSuccess) return spano_fail(ctx, SPANO_E_CUDA, "cudaFuncSetAttribute(blend_ws<%d,%d>, %zu B): %s", B, VT, W::SMEM, cudaGetErrorString(e));
        configured[dev] = true;
    }
    march::blend_ws_kernel<B, SW, VT, UNIFORM><<<sms, W::THREADS, W::SMEM, ctx->stream>>>(P);
    return 0;
}

template <int B>
int launch_march(spano_ctx *ctx, const BlendParams &Q, int sms)
{
    constexpr int SW = (B <= 6) ? 32 : 16;
    using C = march::Cfg<B, SW>;
    // 0 default (warp-specialised, 8 V warps; 12 for B >= 9), 2 the 8-warp marching kernel, 3 warp-specialised with 12 V warps,
    // 4 warp-specialised without the setmaxnreg register split
    const int mode = ctx->opt_blend_kernel;
    static bool configured[64] = {false};
    int dev = ctx->device & 63;
    if (mode == 2 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(march::blend_march_kernel<B, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return spano_fail(ctx, SPANO_E_CUDA, "cudaFuncSetAttribute(blend_march<%d>, %zu B): %s", B, C::SMEM, cudaGetErrorString(e));
        configured[dev] = true;
    }
    march::Params P;
    P.slot = ctx->tap_slot;
    P.tile = Q.tile;  P.tile_step = Q.tile_step;
    P.cut = Q.cut;  P.cut_step = Q.cut_step;
    P.valid = Q.valid;  P.valid_step = Q.valid_step;
    P.w = Q.w;  P.h = Q.h;
    P.ty_begin = Q.ty_begin;  P.ty_end = Q.ty_end;
    P.wx0 = Q.wx0;  P.wx1 = Q.wx1;
    P.acc = Q.acc;  P.canvas_w = Q.canvas_w;  P.ax = Q.ax;  P.ay = Q.ay;
    // sparsity plan (activity of mask_cut per strip -> active output rows -> even split over the CTAs), unless the
    // caller already made it (fused path: on the auxiliary stream, one image ahead)
    if (Q.plan) {
        P.plan = Q.plan;
    } else {
        int *plan = nullptr;
        const int strips = (Q.w + SW - 1) / SW;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_BLENDPLAN, march::PlanView::ints(strips, sms) * sizeof(int), (void **)&plan)) return rc;
        if (int rc = make_plan<SW>(ctx, Q, sms, plan)) return rc;
        P.plan = plan;
    }
    if (mode == 2) {
        march::blend_march_kernel<B, SW><<<sms, march::THREADS, C::SMEM, ctx->stream>>>(P);
        return 0;
    }
    if constexpr (SW == 32) {
        if (mode == 3) return launch_ws<B, SW, 384, false>(ctx, P, sms);
    }
    if (mode == 4) return launch_ws<B, SW, 256, true>(ctx, P, sms);   // no setmaxnreg: measured 2x slower, kept for A/B
    if constexpr (B >= 9) return launch_ws<B, SW, 384, false>(ctx, P, sms);   // 6 sigma groups x 2 instead of 4 x 3
    else return launch_ws<B, SW, 256, false>(ctx, P, sms);
}

template <int B>
int launch_generic(spano_ctx *ctx, const BlendParams &P, dim3 grid)
{
    static bool configured[64] = {false};
    const int IW = GB + 2 * P.radius;
    const size_t smem = ((size_t)IW * (IW + 1) + (size_t)IW * GB) * sizeof(float);
    int dev = ctx->device & 63;
    if (!configured[dev]) {
        const int IWm = GB + 2 * MAXR;
        const size_t smem_max = ((size_t)IWm * (IWm + 1) + (size_t)IWm * GB) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(blend_generic_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return spano_fail(ctx, SPANO_E_CUDA, "cudaFuncSetAttribute(blend_generic<%d>): %s", B, cudaGetErrorString(e));
        configured[dev] = true;
    }
    blend_generic_kernel<B><<<grid, GTHREADS, smem, ctx->stream>>>(P);
    return 0;
}

} // namespace

// Tap slots of the marching kernel: write-once per (device, bands, sigma).  A slot is filled under the mutex with a
// synchronous copy that is waited for before the slot is published, and is never rewritten afterwards, so kernels on any
// stream of any context can read it without further ordering.  When the tables of a device are full (more than
// TAP_ROWS bands in total over the distinct (bands, sigma) pairs of one process) the caller falls back to the generic
// kernel, whose taps travel in its launch parameters.
struct TapSlots {
    std::mutex mu;
    int n = 0, rows = 0;
    int bands[TAP_SLOTS], row0[TAP_SLOTS];
    double sigma[TAP_SLOTS];
};
static TapSlots g_tap_slots[64];

static int tap_slot(spano_ctx *ctx, int bands, double sigma, int *slot)
{
    TapSlots &T = g_tap_slots[ctx->device & 63];
    std::lock_guard<std::mutex> lk(T.mu);
    *slot = -1;
    for (int i = 0; i < T.n; ++i)
        if (T.bands[i] == bands && T.sigma[i] == sigma) { *slot = T.row0[i]; return 0; }
    if (T.n == TAP_SLOTS || T.rows + bands > TAP_ROWS) return 0;
    const int i = T.n, r0 = T.rows;
    float taps[MAXB][24] = {};
    float2 pairs[MAXB][44] = {};
    for (int b = 0; b < bands; ++b) {
        for (int k = 0; k <= FR; ++k) taps[b][k] = ctx->taps[b][k];
        for (int d = 0; d <= 2 * FR + 1; ++d) {
            const float hi = d <= 2 * FR ? ctx->taps[b][d < FR ? FR - d : d - FR] : 0.f;
            const float lo = d >= 1 ? ctx->taps[b][(d - 1) < FR ? FR - (d - 1) : (d - 1) - FR] : 0.f;
            pairs[b][d] = make_float2(hi, lo);
        }
    }
    SPANO_CUDA(ctx, cudaMemcpyToSymbol(c_taps, taps, (size_t)bands * sizeof(taps[0]), (size_t)r0 * sizeof(taps[0]), cudaMemcpyHostToDevice));
    SPANO_CUDA(ctx, cudaMemcpyToSymbol(c_tap2, pairs, (size_t)bands * sizeof(pairs[0]), (size_t)r0 * sizeof(pairs[0]), cudaMemcpyHostToDevice));
    SPANO_CUDA(ctx, cudaStreamSynchronize(cudaStreamLegacy));   // the copies (pageable source, legacy stream) have landed
    T.bands[i] = bands;  T.sigma[i] = sigma;  T.row0[i] = r0;
    T.n = i + 1;
    T.rows = r0 + bands;
    *slot = r0;
    return 0;
}

// Gaussian taps of every band (cv::getGaussianKernel, see projector_host.cpp), kept on the host in the context and
// copied into the parameters of every blend launch.
int launch_blend_setup(spano_ctx *ctx, int bands, double sigma)
{
    if (bands < 1 || bands > MAXB) return spano_fail(ctx, SPANO_E_INVALID, "bands %d not in [1,%d]", bands, MAXB);
    if (!(sigma > 0)) return spano_fail(ctx, SPANO_E_INVALID, "sigma must be > 0");
    const int radius = (int)std::ceil(3 * sigma);
    if (radius < 1 || radius > MAXR)
        return spano_fail(ctx, SPANO_E_LIMIT, "blur radius ceil(3*sigma)=%d exceeds %d", radius, MAXR);
    if (ctx->tap_bands == bands && ctx->tap_sigma == sigma) return radius;
    ctx->tap_bands = 0;
    float full[2 * MAXR + 1];
    const int n = 2 * radius + 1;
    memset(ctx->taps, 0, sizeof(ctx->taps));
    for (int i = 0; i < bands; ++i) {
        const double sb = std::sqrt((double)(2 * (bands - i) + 1)) * sigma;
        spano_host_gaussian_taps(n, sb, full);
        for (int k = 0; k <= radius; ++k) ctx->taps[i][k] = full[radius + k];
    }
    ctx->tap_slot = -1;
    if (radius == FR)
        if (int rc = tap_slot(ctx, bands, sigma, &ctx->tap_slot)) return rc;
    ctx->tap_bands = bands;
    ctx->tap_sigma = sigma;
    return radius;
}

int launch_blend_clear(spano_ctx *ctx, float4 *acc, int canvas_w, int rows)
{
    SPANO_CUDA(ctx, cudaMemsetAsync(acc, 0, (size_t)canvas_w * rows * sizeof(float4), ctx->stream));
    return 0;
}

// ctx->opt_blend_kernel: 0 = default (warp-specialised marching kernel when the radius is 21), 1 = always the
// generic-radius kernel, 2 = the 8-warp marching kernel (same arithmetic and order as the default: bit-identical),
// 3 = warp-specialised with 12 instead of 8 V warps on 32-column strips (measured slower: 1.85 vs 1.64 ms per dense tile)
static bool march_path(const spano_ctx *ctx, int radius) { return radius == FR && ctx->opt_blend_kernel != 1 && ctx->tap_slot >= 0; }

size_t blend_plan_bytes(spano_ctx *ctx, int w, int bands, int radius)
{
    if (!march_path(ctx, radius) || w <= 0) return 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int SW = bands <= 6 ? 32 : 16;
    return march::PlanView::ints((w + SW - 1) / SW, sms) * sizeof(int);
}

int launch_blend_plan(spano_ctx *ctx, const BlendTile &t, int bands, int radius, int row0, int row1, int *plan, int canvas_w)
{
    int ty_begin = row0 - t.cy, ty_end = row1 - t.cy;
    if (ty_begin < 0) ty_begin = 0;
    if (ty_end > t.h) ty_end = t.h;
    if (ty_end <= ty_begin || t.w <= 0 || !march_path(ctx, radius)) return 0;
    BlendParams P;
    P.cut = t.cut;  P.cut_step = t.cut_step;
    P.w = t.w;  P.h = t.h;
    P.ty_begin = ty_begin;  P.ty_end = ty_end;
    P.wx0 = std::max(0, -t.cx);  P.wx1 = std::min(t.w, canvas_w - t.cx);
    if (P.wx1 <= P.wx0) return 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int rc = bands <= 6 ? make_plan<32>(ctx, P, sms, plan) : make_plan<16>(ctx, P, sms, plan);
    return rc ? rc : 3;
}

int launch_blend_plan_preview(spano_ctx *ctx, const BlendTile &t, const uint8_t *small, int sw, int sh, size_t sstep, int bands, int radius,
                              int row0, int row1, int *plan, int canvas_w)
{
    int ty_begin = row0 - t.cy, ty_end = row1 - t.cy;
    if (ty_begin < 0) ty_begin = 0;
    if (ty_end > t.h) ty_end = t.h;
    if (ty_end <= ty_begin || t.w <= 0 || !plan || !march_path(ctx, radius)) return 0;
    BlendParams P;
    P.cut = t.cut;  P.cut_step = t.cut_step;
    P.w = t.w;  P.h = t.h;
    P.ty_begin = ty_begin;  P.ty_end = ty_end;
    P.wx0 = std::max(0, -t.cx);  P.wx1 = std::min(t.w, canvas_w - t.cx);
    if (P.wx1 <= P.wx0) return 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const PreviewCut pv{small, sw, sh, sstep};
    const int rc = bands <= 6 ? make_plan<32>(ctx, P, sms, plan, &pv) : make_plan<16>(ctx, P, sms, plan, &pv);
    return rc ? rc : 5;
}

int launch_blend_tile(spano_ctx *ctx, const BlendTile &t, int bands, int radius, float4 *acc, int canvas_w, int row0,
                      int row1, const int *plan)
{
    // tile rows that fall into canvas rows [row0,row1)
    int ty_begin = row0 - t.cy, ty_end = row1 - t.cy;
    if (ty_begin < 0) ty_begin = 0;
    if (ty_end > t.h) ty_end = t.h;
    if (ty_end <= ty_begin || t.w <= 0) return 0;
    BlendParams P;
    P.tile = t.tile;  P.tile_step = t.tile_step;
    P.cut = t.cut;  P.cut_step = t.cut_step;
    P.valid = t.valid;  P.valid_step = t.valid_step;
    P.w = t.w;  P.h = t.h;
    P.ty_begin = ty_begin;  P.ty_end = ty_end;
    // a tile may stick out of the accumulator's columns (column bands: the accumulator is a column range of the canvas)
    P.wx0 = std::max(0, -t.cx);  P.wx1 = std::min(t.w, canvas_w - t.cx);
    if (P.wx1 <= P.wx0) return 0;
    P.acc = acc;  P.canvas_w = canvas_w;
    P.ax = t.cx;  P.ay = t.cy - row0;
    P.radius = radius;
    P.plan = plan;
    int rc = 0;
    if (ctx->tap_bands != bands || !(ctx->tap_sigma > 0)) return spano_fail(ctx, SPANO_E_INVALID, "blend launched without launch_blend_setup");
    if (march_path(ctx, radius)) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        switch (bands) {
#define CASE(B) case B: rc = launch_march<B>(ctx, P, sms); break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: return spano_fail(ctx, SPANO_E_INVALID, "bands %d unsupported", bands);
        }
    } else {
        memcpy(P.taps, ctx->taps, sizeof(P.taps));
        dim3 grid((t.w + GB - 1) / GB, (ty_end - ty_begin + GB - 1) / GB);
        switch (bands) {
#define CASE(B) case B: rc = launch_generic<B>(ctx, P, grid); break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: return spano_fail(ctx, SPANO_E_INVALID, "bands %d unsupported", bands);
        }
    }
    if (rc) return rc;
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_normalise(spano_ctx *ctx, const float4 *acc, int canvas_w, int rows, int bands, int out_kind, void *out,
                     size_t out_step, int col0, int col1)
{
    if (col1 < 0) col1 = canvas_w;
    if (rows <= 0 || col1 <= col0) return 0;
    const float divisor = (float)(255 / bands); // integer division, as in the reference
    const float inv_div = (float)(1.0 / (double)divisor);
    dim3 block(256), grid((col1 - col0 + 255) / 256, rows);
    normalise_kernel<<<grid, block, 0, ctx->stream>>>(acc, canvas_w, rows, inv_div, out_kind, out, out_step, col0, col1);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_fp32_peak(spano_ctx *ctx, int variant, double *tflops)
{
    float *sink = nullptr;
    int rc = spano_reserve(ctx, spano_ctx::BUF_MISC, 256, (void **)&sink);
    if (rc) return rc;
    int sms = 0;
    SPANO_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int iters = 20000, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    SPANO_CUDA(ctx, cudaEventCreate(&e0));
    SPANO_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    PeakTaps T;
    for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 16; ++k) T.t[i][k] = 1.0f / (float)(17 + i + k);
    for (int rep = 0; rep < 4; ++rep) {
        SPANO_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        if (variant == 0) fp32_peak_kernel<0><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 1.0f, T);
        else if (variant == 1) fp32_peak_kernel<1><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 1.0f, T);
        else fp32_peak_kernel<2><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 1.0f, T);
        SPANO_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        SPANO_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        SPANO_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += 4;
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}
