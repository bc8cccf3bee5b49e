// capi.cu -- C ABI of the compositing library (include/spano.h) and the host orchestration of
// the fused path.  No exceptions cross this boundary; every entry point returns a status code.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <algorithm>

#include "spano_internal.h"

// ---------------------------------------------------------------------------------------------
// context plumbing
// ---------------------------------------------------------------------------------------------
int spano_fail(spano_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int spano_reserve(spano_ctx *ctx, int which, size_t bytes, void **out)
{
    DeviceBuffer &b = ctx->buf[which];
    if (b.bytes < bytes) {
        if (b.ptr) {
            // the old buffer may still be in use by work queued on the stream
            cudaStreamSynchronize(ctx->stream);
            cudaFree(b.ptr);
            b.ptr = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.ptr, want);
        if (e != cudaSuccess) {
            b.ptr = nullptr;
            return spano_fail(ctx, SPANO_E_NOMEM, "cudaMalloc(%zu B) failed: %s", want, cudaGetErrorString(e));
        }
        b.bytes = want;
    }
    *out = b.ptr;
    return 0;
}

namespace {

// serialises the calls on one context and makes the context's device current for the duration of the call (the
// caller's current device -- torch reads it with cudaGetDevice -- is restored on the way out)
struct Guard {
    spano_ctx *c;
    int prev = -1;
    explicit Guard(spano_ctx *ctx) : c(ctx)
    {
        c->mu.lock();
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device) cudaSetDevice(c->device);
    }
    ~Guard()
    {
        if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
        c->mu.unlock();
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StageTimer {
    spano_ctx *ctx;
    int stage;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    StageTimer(spano_ctx *c, int s) : ctx(c), stage(s)
    {
        if (!ctx->timers_on) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
    }
    void stop(int launches)
    {
        if (!ctx->timers_on) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->pending_events.push_back({stage, {e0, e1}});
        ctx->stage_launches[stage] += launches;
    }
};

void drain_timers(spano_ctx *ctx)
{
    for (auto &p : ctx->pending_events) {
        float ms = 0.f;
        cudaEventSynchronize(p.second.second);
        cudaEventElapsedTime(&ms, p.second.first, p.second.second);
        ctx->stage_ms[p.first] += ms;
        cudaEventDestroy(p.second.first);
        cudaEventDestroy(p.second.second);
    }
    ctx->pending_events.clear();
}

int check_image_args(spano_ctx *ctx, const void *p, int w, int h, size_t step, int cn, const char *what)
{
    if (!p) return spano_fail(ctx, SPANO_E_INVALID, "%s: null pointer", what);
    if (w <= 0 || h <= 0) return spano_fail(ctx, SPANO_E_INVALID, "%s: empty image %dx%d", what, w, h);
    if (step < (size_t)w * cn) return spano_fail(ctx, SPANO_E_INVALID, "%s: step %zu < row bytes %zu", what, step, (size_t)w * cn);
    return 0;
}

// cv::remap asserts every dimension < SHRT_MAX
int check_remap_limits(spano_ctx *ctx, int sw, int sh, int dw, int dh)
{
    if (sw >= 32767 || sh >= 32767 || dw >= 32767 || dh >= 32767)
        return spano_fail(ctx, SPANO_E_LIMIT, "remap dimension >= 32767 (src %dx%d dst %dx%d): the reference's cv::remap rejects this", sw, sh, dw, dh);
    return 0;
}

int valid_proj(spano_ctx *ctx, int proj, float scale)
{
    if (proj < SPANO_SPHERICAL || proj > SPANO_STEREOGRAPHIC) return spano_fail(ctx, SPANO_E_INVALID, "unknown projection %d", proj);
    if (!(scale > 0.f)) return spano_fail(ctx, SPANO_E_INVALID, "scale must be > 0");
    return 0;
}

// warp + dark flags + validity mask of one full tile, all on device
int dev_warp_tile(spano_ctx *ctx, const SpanoProjector &P, const uint8_t *d_src, int src_w, int src_h, size_t src_step,
                  double gain, int tl_x, int tl_y, int w, int h, uint8_t *d_tile, size_t tile_step, uint8_t *d_mask,
                  size_t mask_step)
{
    uint8_t *dark = nullptr;
    size_t dark_step = 0;
    if (d_mask) {
        dark_step = align_up((size_t)w, 16);
        int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, dark_step * h, (void **)&dark);
        if (rc) return rc;
    }
    StageTimer t0(ctx, 0);
    int n = launch_warp(ctx, P, d_src, src_w, src_h, src_step, gain, tl_x, tl_y, w, h, 0, h, d_tile, tile_step, dark, dark_step);
    if (n < 0) return n;
    t0.stop(n);
    if (d_mask) {
        StageTimer t1(ctx, 1);
        n = launch_valid_mask(ctx, dark, w, h, dark_step, 3, d_mask, mask_step);
        if (n < 0) return n;
        t1.stop(n);
    }
    return 0;
}

int dev_multiblend(spano_ctx *ctx, int n, const BlendTile *tiles, int canvas_w, int canvas_h, int bands, double sigma,
                   int row0, int row1, int out_kind, void *d_out, size_t out_step)
{
    if (row0 < 0) row0 = 0;
    if (row1 > canvas_h) row1 = canvas_h;
    const int rows = row1 - row0;
    if (rows <= 0) return 0;
    int radius = launch_blend_setup(ctx, bands, sigma);
    if (radius < 0) return radius;
    float4 *acc = nullptr;
    int rc = spano_reserve(ctx, spano_ctx::BUF_ACC, (size_t)canvas_w * rows * sizeof(float4), (void **)&acc);
    if (rc) return rc;
    StageTimer t2(ctx, 2);
    rc = launch_blend_clear(ctx, acc, canvas_w, rows);
    if (rc) return rc;
    int launches = 0;
    for (int j = 0; j < n; ++j) {
        int k = launch_blend_tile(ctx, tiles[j], bands, radius, acc, canvas_w, row0, row1);
        if (k < 0) return k;
        launches += k;
    }
    t2.stop(launches);
    StageTimer t3(ctx, 3);
    int k = launch_normalise(ctx, acc, canvas_w, rows, bands, out_kind, d_out, out_step);
    if (k < 0) return k;
    t3.stop(k);
    return 0;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" int spano_version(void) { return SPANO_VERSION; }

extern "C" int spano_create(spano_ctx **out, int device)
{
    if (!out) return SPANO_E_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return SPANO_E_NODEVICE;
    if (device < 0 || device >= count) return SPANO_E_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SPANO_E_CUDA;
    if (prop.major != 10) return SPANO_E_NODEVICE; // built for sm_100a only
    spano_ctx *ctx = new (std::nothrow) spano_ctx();
    if (!ctx) return SPANO_E_NOMEM;
    ctx->device = device;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    // the context's own stream gets the highest priority, so that a blend's large CTAs (main stream) are placed before the small
    // CTAs of the next image's warp / mask kernels (auxiliary stream, default priority) when both become runnable at once
    int prio_lo = 0, prio_hi = 0;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    if (ok) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    ok = ok && cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (!ok) {
        delete ctx;
        return SPANO_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SPANO_OK;
}

extern "C" void spano_destroy(spano_ctx *ctx)
{
    if (!ctx) return;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drain_timers(ctx);
    for (auto &b : ctx->buf)
        if (b.ptr) cudaFree(b.ptr);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->d2h_stream) {
        cudaStreamSynchronize(ctx->d2h_stream);
        cudaStreamDestroy(ctx->d2h_stream);
    }
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(ctx->ev_warped[b]); cudaEventDestroy(ctx->ev_blended[b]); }
        cudaEventDestroy(ctx->ev_start2);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(ctx->ev_copied[b]); cudaEventDestroy(ctx->ev_free[b]); }
        cudaEventDestroy(ctx->ev_start);
    }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    delete ctx;
}

extern "C" const char *spano_last_error(spano_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int spano_set_stream(spano_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SPANO_OK;
}

extern "C" int spano_sync(spano_ctx *ctx)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" long long spano_launch_count(spano_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int spano_set_option(spano_ctx *ctx, int option, int value)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    switch (option) {
    case SPANO_OPT_BLEND_DENSE: ctx->opt_blend_dense = value != 0; return SPANO_OK;
    case SPANO_OPT_FLAG_WAIT: ctx->opt_flag_wait = value != 0; return SPANO_OK;
    case SPANO_OPT_WARP_KERNEL: ctx->opt_warp_kernel = value != 0; return SPANO_OK;
    case SPANO_OPT_BLEND_KERNEL:
        if (value < 0 || value > 4) return spano_fail(ctx, SPANO_E_INVALID, "SPANO_OPT_BLEND_KERNEL: value %d not in [0,4]", value);
        ctx->opt_blend_kernel = value;
        return SPANO_OK;
    default: return spano_fail(ctx, SPANO_E_INVALID, "unknown option %d", option);
    }
}

extern "C" int spano_timers_enable(spano_ctx *ctx, int on)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    ctx->timers_on = on != 0;
    return SPANO_OK;
}

extern "C" int spano_blend_stats(spano_ctx *ctx, unsigned long long *processed_px, unsigned long long *offered_px, int reset)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    unsigned long long v[2] = {0, 0};
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->blend_stats) {
        SPANO_CUDA(ctx, cudaMemcpy(v, ctx->blend_stats, sizeof(v), cudaMemcpyDeviceToHost));
        if (reset) SPANO_CUDA(ctx, cudaMemset(ctx->blend_stats, 0, sizeof(v)));
    }
    if (processed_px) *processed_px = v[0];
    if (offered_px) *offered_px = v[1];
    return SPANO_OK;
}

extern "C" int spano_timers_reset(spano_ctx *ctx)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    drain_timers(ctx);
    for (int i = 0; i < 4; ++i) { ctx->stage_ms[i] = 0.f; ctx->stage_launches[i] = 0; }
    return SPANO_OK;
}

extern "C" int spano_timers_read(spano_ctx *ctx, float ms[4], long long launches[4])
{
    if (!ctx || !ms) return SPANO_E_INVALID;
    Guard g(ctx);
    drain_timers(ctx);
    for (int i = 0; i < 4; ++i) {
        ms[i] = ctx->stage_ms[i];
        if (launches) launches[i] = ctx->stage_launches[i];
    }
    return SPANO_OK;
}

extern "C" int spano_fp32_peak(spano_ctx *ctx, int variant, double *tflops)
{
    if (!ctx || !tflops) return SPANO_E_INVALID;
    Guard g(ctx);
    return launch_fp32_peak(ctx, variant, tflops);
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
extern "C" int spano_warp_roi(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], int src_w,
                              int src_h, int *tl_x, int *tl_y, int *dst_w, int *dst_h)
{
    // pure host arithmetic: ctx may be NULL (then no error message is recorded)
    if (!K || !R || !tl_x || !tl_y || !dst_w || !dst_h) return spano_fail(ctx, SPANO_E_INVALID, "spano_warp_roi: null argument");
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    if (src_w <= 0 || src_h <= 0) return spano_fail(ctx, SPANO_E_INVALID, "spano_warp_roi: empty source %dx%d", src_w, src_h);
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, K, R);
    int roi[4];
    spano_host_roi(&P, src_w, src_h, roi);
    *tl_x = roi[0];
    *tl_y = roi[1];
    // RotationWarperBase::warp: dst.create(roi.height + 1, roi.width + 1), roi = Rect(tl, br)
    const long long w = (long long)roi[2] - roi[0] + 1, h = (long long)roi[3] - roi[1] + 1;
    if (w <= 0 || h <= 0 || w > 0x7fffffffLL || h > 0x7fffffffLL)
        return spano_fail(ctx, SPANO_E_LIMIT, "degenerate result ROI (%d,%d)-(%d,%d)", roi[0], roi[1], roi[2], roi[3]);
    *dst_w = (int)w;
    *dst_h = (int)h;
    return SPANO_OK;
}

extern "C" int spano_pan_dimension(int n, const int *tl_x, const int *tl_y, const int *w, const int *h, int *canvas_w,
                                   int *canvas_h, int *min_x, int *min_y)
{
    if (n <= 0 || !tl_x || !tl_y || !w || !h) return SPANO_E_INVALID;
    int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
    for (int i = 0; i < n; ++i) {
        x0 = std::min(x0, tl_x[i]);
        y0 = std::min(y0, tl_y[i]);
        x1 = std::max(x1, tl_x[i] + w[i]);
        y1 = std::max(y1, tl_y[i] + h[i]);
    }
    if (canvas_w) *canvas_w = x1 - x0;
    if (canvas_h) *canvas_h = y1 - y0;
    if (min_x) *min_x = x0;
    if (min_y) *min_y = y0;
    return SPANO_OK;
}

// ---------------------------------------------------------------------------------------------
// device-pointer stage entry points
// ---------------------------------------------------------------------------------------------
extern "C" int spano_dev_warp(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9],
                              const uint8_t *src_bgr, int src_w, int src_h, size_t src_step, double gain, int tl_x,
                              int tl_y, int dst_w, int dst_h, uint8_t *dst_bgr, size_t dst_step,
                              uint8_t *dst_valid_mask, size_t mask_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (!K || !R) return spano_fail(ctx, SPANO_E_INVALID, "spano_dev_warp: null K/R");
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    if (int rc = check_image_args(ctx, src_bgr, src_w, src_h, src_step, 3, "source")) return rc;
    if (int rc = check_image_args(ctx, dst_bgr, dst_w, dst_h, dst_step, 3, "destination")) return rc;
    if (dst_valid_mask && mask_step < (size_t)dst_w) return spano_fail(ctx, SPANO_E_INVALID, "mask step too small");
    if (!(gain > 0.0)) return spano_fail(ctx, SPANO_E_INVALID, "gain must be > 0");
    if (int rc = check_remap_limits(ctx, src_w, src_h, dst_w, dst_h)) return rc;
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, K, R);
    return dev_warp_tile(ctx, P, src_bgr, src_w, src_h, src_step, gain, tl_x, tl_y, dst_w, dst_h, dst_bgr, dst_step,
                         dst_valid_mask, mask_step);
}

extern "C" int spano_dev_tile_mask(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9],
                                   const uint8_t *src_bgr, int src_w, int src_h, size_t src_step, int tl_x, int tl_y, int w,
                                   int h, uint8_t *valid_mask, size_t mask_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (!K || !R) return spano_fail(ctx, SPANO_E_INVALID, "spano_dev_tile_mask: null K/R");
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    if (int rc = check_image_args(ctx, src_bgr, src_w, src_h, src_step, 3, "source")) return rc;
    if (int rc = check_image_args(ctx, valid_mask, w, h, mask_step, 1, "mask")) return rc;
    if (int rc = check_remap_limits(ctx, src_w, src_h, w, h)) return rc;
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, K, R);
    return dev_warp_tile(ctx, P, src_bgr, src_w, src_h, src_step, 1.0, tl_x, tl_y, w, h, nullptr, 0, valid_mask, mask_step);
}

extern "C" int spano_dev_multiblend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                                    const uint8_t *const *masks, const size_t *mask_steps,
                                    const uint8_t *const *masks_orig, const size_t *orig_steps, const int *tl_x,
                                    const int *tl_y, const int *w, const int *h, int bands, double sigma, int row0,
                                    int row1, int out_kind, void *out, size_t out_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !tiles || !tile_steps || !masks || !mask_steps || !masks_orig || !orig_steps || !tl_x || !tl_y || !w || !h || !out)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_dev_multiblend: null/empty argument");
    if (out_kind != SPANO_OUT_F32 && out_kind != SPANO_OUT_U8) return spano_fail(ctx, SPANO_E_INVALID, "unknown out_kind %d", out_kind);
    int cw, chh, mx, my;
    spano_pan_dimension(n, tl_x, tl_y, w, h, &cw, &chh, &mx, &my);
    if (out_step < (size_t)cw * (out_kind == SPANO_OUT_F32 ? 12 : 3)) return spano_fail(ctx, SPANO_E_INVALID, "out_step too small");
    std::vector<BlendTile> bt(n);
    for (int j = 0; j < n; ++j) {
        if (int rc = check_image_args(ctx, tiles[j], w[j], h[j], tile_steps[j], 3, "tile")) return rc;
        if (int rc = check_image_args(ctx, masks[j], w[j], h[j], mask_steps[j], 1, "mask_cut")) return rc;
        if (int rc = check_image_args(ctx, masks_orig[j], w[j], h[j], orig_steps[j], 1, "mask_orig")) return rc;
        bt[j] = BlendTile{tiles[j], tile_steps[j], masks[j], mask_steps[j], masks_orig[j], orig_steps[j], w[j], h[j], tl_x[j] - mx, tl_y[j] - my};
    }
    return dev_multiblend(ctx, n, bt.data(), cw, chh, bands, sigma, row0, row1, out_kind, out, out_step);
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------
extern "C" int spano_warp(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9],
                          const uint8_t *src_bgr, int src_w, int src_h, size_t src_step, double gain, uint8_t *dst_bgr,
                          size_t dst_step, uint8_t *dst_valid_mask, size_t mask_step)
{
    if (!ctx) return SPANO_E_INVALID;
    int tl_x, tl_y, w, h;
    if (int rc = spano_warp_roi(ctx, proj, scale, K, R, src_w, src_h, &tl_x, &tl_y, &w, &h)) return rc;
    Guard g(ctx);
    if (int rc = check_image_args(ctx, src_bgr, src_w, src_h, src_step, 3, "source")) return rc;
    if (int rc = check_image_args(ctx, dst_bgr, w, h, dst_step, 3, "destination")) return rc;
    if (dst_valid_mask && mask_step < (size_t)w) return spano_fail(ctx, SPANO_E_INVALID, "mask step too small");
    if (!(gain > 0.0)) return spano_fail(ctx, SPANO_E_INVALID, "gain must be > 0");
    if (int rc = check_remap_limits(ctx, src_w, src_h, w, h)) return rc;
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, K, R);
    const size_t s_step = align_up((size_t)src_w * 3, 16), t_step = align_up((size_t)w * 3, 16), m_step = align_up((size_t)w, 16);
    uint8_t *d_src, *d_tile, *d_mask = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, s_step * src_h + 16, (void **)&d_src)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * h, (void **)&d_tile)) return rc;
    if (dst_valid_mask)
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILEMASK, m_step * h, (void **)&d_mask)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_src, s_step, src_bgr, src_step, (size_t)src_w * 3, src_h, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = dev_warp_tile(ctx, P, d_src, src_w, src_h, s_step, gain, tl_x, tl_y, w, h, d_tile, t_step, d_mask, m_step)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(dst_bgr, dst_step, d_tile, t_step, (size_t)w * 3, h, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst_valid_mask)
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(dst_valid_mask, mask_step, d_mask, m_step, (size_t)w, h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_build_maps(spano_ctx *ctx, int proj, float scale, const float K[9], const float R[9], int tl_x,
                                int tl_y, int w, int h, float *xmap, float *ymap)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (!K || !R || !xmap || !ymap) return spano_fail(ctx, SPANO_E_INVALID, "spano_build_maps: null argument");
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    if (w <= 0 || h <= 0) return spano_fail(ctx, SPANO_E_INVALID, "spano_build_maps: empty map %dx%d", w, h);
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, K, R);
    float *d_maps;
    const size_t n = (size_t)w * h;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_MISC, 2 * n * sizeof(float), (void **)&d_maps)) return rc;
    int k = launch_warp(ctx, P, nullptr, 1, 1, 0, 1.0, tl_x, tl_y, w, h, 0, h, nullptr, 0, nullptr, 0, d_maps, d_maps + n);
    if (k < 0) return k;
    SPANO_CUDA(ctx, cudaMemcpyAsync(xmap, d_maps, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaMemcpyAsync(ymap, d_maps + n, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_remap(spano_ctx *ctx, const uint8_t *src_bgr, int src_w, int src_h, size_t src_step,
                           const float *xmap, const float *ymap, int dst_w, int dst_h, uint8_t *dst_bgr, size_t dst_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (!xmap || !ymap) return spano_fail(ctx, SPANO_E_INVALID, "spano_remap: null map");
    if (int rc = check_image_args(ctx, src_bgr, src_w, src_h, src_step, 3, "source")) return rc;
    if (int rc = check_image_args(ctx, dst_bgr, dst_w, dst_h, dst_step, 3, "destination")) return rc;
    if (int rc = check_remap_limits(ctx, src_w, src_h, dst_w, dst_h)) return rc;
    const size_t s_step = align_up((size_t)src_w * 3, 16), t_step = align_up((size_t)dst_w * 3, 16), n = (size_t)dst_w * dst_h;
    uint8_t *d_src, *d_tile;
    float *d_maps;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, s_step * src_h + 16, (void **)&d_src)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * dst_h, (void **)&d_tile)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_MISC, 2 * n * sizeof(float), (void **)&d_maps)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_src, s_step, src_bgr, src_step, (size_t)src_w * 3, src_h, cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaMemcpyAsync(d_maps, xmap, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaMemcpyAsync(d_maps + n, ymap, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int k = launch_remap(ctx, d_src, src_w, src_h, s_step, d_maps, d_maps + n, dst_w, dst_h, d_tile, t_step);
    if (k < 0) return k;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(dst_bgr, dst_step, d_tile, t_step, (size_t)dst_w * 3, dst_h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_surrounding_mask(spano_ctx *ctx, const uint8_t *bgr, int w, int h, size_t step, int erode_iters,
                                      uint8_t *mask, size_t mask_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_image_args(ctx, bgr, w, h, step, 3, "image")) return rc;
    if (int rc = check_image_args(ctx, mask, w, h, mask_step, 1, "mask")) return rc;
    const size_t t_step = align_up((size_t)w * 3, 16), m_step = align_up((size_t)w, 16);
    uint8_t *d_img, *d_dark, *d_mask;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * h, (void **)&d_img)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, m_step * h, (void **)&d_dark)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILEMASK, m_step * h, (void **)&d_mask)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_img, t_step, bgr, step, (size_t)w * 3, h, cudaMemcpyHostToDevice, ctx->stream));
    int n = launch_dark_flags(ctx, d_img, w, h, t_step, d_dark, m_step);
    if (n < 0) return n;
    n = launch_valid_mask(ctx, d_dark, w, h, m_step, erode_iters, d_mask, m_step);
    if (n < 0) return n;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(mask, mask_step, d_mask, m_step, (size_t)w, h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_resize_mask(spano_ctx *ctx, const uint8_t *src, int src_w, int src_h, size_t src_step, uint8_t *dst,
                                 int dst_w, int dst_h, size_t dst_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_image_args(ctx, src, src_w, src_h, src_step, 1, "source mask")) return rc;
    if (int rc = check_image_args(ctx, dst, dst_w, dst_h, dst_step, 1, "destination mask")) return rc;
    const size_t s_step = align_up((size_t)src_w, 16), d_step = align_up((size_t)dst_w, 16);
    uint8_t *d_src, *d_dst;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTSMALL, s_step * src_h, (void **)&d_src)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILEMASK, d_step * dst_h, (void **)&d_dst)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_src, s_step, src, src_step, (size_t)src_w, src_h, cudaMemcpyHostToDevice, ctx->stream));
    int k = launch_resize_mask(ctx, d_src, src_w, src_h, s_step, d_dst, dst_w, dst_h, d_step);
    if (k < 0) return k;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_step, d_dst, d_step, (size_t)dst_w, dst_h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_adjust_intensity(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, const float *field,
                                      int field_w, int field_h, size_t field_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_image_args(ctx, bgr, w, h, step, 3, "image")) return rc;
    if (int rc = check_image_args(ctx, field, field_w, field_h, field_step, 4, "intensity field")) return rc;
    const size_t t_step = align_up((size_t)w * 3, 16), f_step = align_up((size_t)field_w * 4, 16);
    uint8_t *d_img;
    float *d_field;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * h, (void **)&d_img)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_FIELD, f_step * field_h, (void **)&d_field)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_img, t_step, bgr, step, (size_t)w * 3, h, cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_field, f_step, field, field_step, (size_t)field_w * 4, field_h, cudaMemcpyHostToDevice, ctx->stream));
    int k = launch_adjust_intensity(ctx, d_img, w, h, t_step, d_field, field_w, field_h, f_step / 4);
    if (k < 0) return k;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(bgr, step, d_img, t_step, (size_t)w * 3, h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_apply_gain(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, double gain)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_image_args(ctx, bgr, w, h, step, 3, "image")) return rc;
    if (!(gain > 0.0)) return spano_fail(ctx, SPANO_E_INVALID, "gain must be > 0");
    const size_t t_step = align_up((size_t)w * 3, 16);
    uint8_t *d_img;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * h, (void **)&d_img)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_img, t_step, bgr, step, (size_t)w * 3, h, cudaMemcpyHostToDevice, ctx->stream));
    int n = launch_gain(ctx, d_img, w, h, t_step, gain);
    if (n < 0) return n;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(bgr, step, d_img, t_step, (size_t)w * 3, h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_disk_reproj_size(spano_ctx *ctx, int n, const int *tl_x, const int *tl_y, const int *w, const int *h,
                                      int ansatz_x, int ansatz_y, float radius, int quadratic, int *out_tl_x,
                                      int *out_tl_y, int *out_w, int *out_h)
{
    if (n <= 0 || !tl_x || !tl_y || !w || !h || !out_tl_x || !out_tl_y || !out_w || !out_h)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_disk_reproj_size: null/empty argument");
    for (int i = 0; i < n; ++i)
        if (w[i] <= 0 || h[i] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "tile %d is empty", i);
    SpanoDiskParams P;
    std::vector<int> ox(n), oy(n);
    if (spano_disk_plan(n, tl_x, tl_y, w, h, ansatz_x, ansatz_y, radius, quadratic, &P, ox.data(), oy.data(), out_tl_x,
                        out_tl_y, out_w, out_h))
        return spano_fail(ctx, SPANO_E_LIMIT, "degenerate tile in disk_reproj (no border samples)");
    return SPANO_OK;
}

extern "C" int spano_disk_reproj(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                                 const int *tl_x, const int *tl_y, const int *w, const int *h, int ansatz_x, int ansatz_y,
                                 float radius, int quadratic, uint8_t *const *out_tiles, const size_t *out_steps,
                                 uint8_t *const *out_masks, const size_t *out_mask_steps)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !tiles || !tile_steps || !tl_x || !tl_y || !w || !h || !out_tiles || !out_steps)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_disk_reproj: null/empty argument");
    for (int i = 0; i < n; ++i)
        if (int rc = check_image_args(ctx, tiles[i], w[i], h[i], tile_steps[i], 3, "tile")) return rc;
    SpanoDiskParams P;
    std::vector<int> ox(n), oy(n), nx(n), ny(n), nw(n), nh(n);
    if (spano_disk_plan(n, tl_x, tl_y, w, h, ansatz_x, ansatz_y, radius, quadratic, &P, ox.data(), oy.data(), nx.data(),
                        ny.data(), nw.data(), nh.data()))
        return spano_fail(ctx, SPANO_E_LIMIT, "degenerate tile in disk_reproj (no border samples)");
    for (int i = 0; i < n; ++i) {
        if (int rc = check_image_args(ctx, out_tiles[i], nw[i], nh[i], out_steps[i], 3, "output tile")) return rc;
        if (int rc = check_remap_limits(ctx, w[i], h[i], nw[i], nh[i])) return rc;
        const size_t s_step = align_up((size_t)w[i] * 3, 16), t_step = align_up((size_t)nw[i] * 3, 16), m_step = align_up((size_t)nw[i], 16);
        uint8_t *d_src, *d_tile, *d_dark, *d_mask;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, s_step * h[i] + 16, (void **)&d_src)) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, t_step * nh[i], (void **)&d_tile)) return rc;
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_src, s_step, tiles[i], tile_steps[i], (size_t)w[i] * 3, h[i], cudaMemcpyHostToDevice, ctx->stream));
        int k = launch_disk_gather(ctx, P, d_src, w[i], h[i], s_step, ox[i], oy[i], d_tile, nw[i], nh[i], t_step, nx[i], ny[i]);
        if (k < 0) return k;
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(out_tiles[i], out_steps[i], d_tile, t_step, (size_t)nw[i] * 3, nh[i], cudaMemcpyDeviceToHost, ctx->stream));
        if (out_masks && out_masks[i]) {
            if (!out_mask_steps || out_mask_steps[i] < (size_t)nw[i]) return spano_fail(ctx, SPANO_E_INVALID, "output mask step too small");
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, m_step * nh[i], (void **)&d_dark)) return rc;
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILEMASK, m_step * nh[i], (void **)&d_mask)) return rc;
            k = launch_dark_flags(ctx, d_tile, nw[i], nh[i], t_step, d_dark, m_step);
            if (k < 0) return k;
            k = launch_valid_mask(ctx, d_dark, nw[i], nh[i], m_step, 3, d_mask, m_step);
            if (k < 0) return k;
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(out_masks[i], out_mask_steps[i], d_mask, m_step, (size_t)nw[i], nh[i], cudaMemcpyDeviceToHost, ctx->stream));
        }
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // staging buffers are reused by the next tile
    }
    return SPANO_OK;
}

extern "C" int spano_multiblend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                                const uint8_t *const *masks, const size_t *mask_steps,
                                const uint8_t *const *masks_orig, const size_t *orig_steps, const int *tl_x,
                                const int *tl_y, const int *w, const int *h, int bands, double sigma, int out_kind,
                                void *out, size_t out_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !tiles || !tile_steps || !masks || !mask_steps || !masks_orig || !orig_steps || !tl_x || !tl_y || !w || !h || !out)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_multiblend: null/empty argument (the reference throws \"Input consistency!\")");
    if (out_kind != SPANO_OUT_F32 && out_kind != SPANO_OUT_U8) return spano_fail(ctx, SPANO_E_INVALID, "unknown out_kind %d", out_kind);
    int cw, chh, mx, my;
    spano_pan_dimension(n, tl_x, tl_y, w, h, &cw, &chh, &mx, &my);
    const size_t px_bytes = out_kind == SPANO_OUT_F32 ? 12 : 3;
    if (out_step < (size_t)cw * px_bytes) return spano_fail(ctx, SPANO_E_INVALID, "out_step too small");
    // upload every tile + masks into one arena
    size_t total = 0;
    std::vector<size_t> off_t(n), off_c(n), off_v(n), st_t(n), st_m(n);
    for (int j = 0; j < n; ++j) {
        if (int rc = check_image_args(ctx, tiles[j], w[j], h[j], tile_steps[j], 3, "tile")) return rc;
        if (int rc = check_image_args(ctx, masks[j], w[j], h[j], mask_steps[j], 1, "mask_cut")) return rc;
        if (int rc = check_image_args(ctx, masks_orig[j], w[j], h[j], orig_steps[j], 1, "mask_orig")) return rc;
        st_t[j] = align_up((size_t)w[j] * 3, 16);
        st_m[j] = align_up((size_t)w[j], 16);
        off_t[j] = total; total += st_t[j] * h[j];
        off_c[j] = total; total += st_m[j] * h[j];
        off_v[j] = total; total += st_m[j] * h[j];
    }
    uint8_t *arena;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, total, (void **)&arena)) return rc;
    std::vector<BlendTile> bt(n);
    for (int j = 0; j < n; ++j) {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_t[j], st_t[j], tiles[j], tile_steps[j], (size_t)w[j] * 3, h[j], cudaMemcpyHostToDevice, ctx->stream));
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_c[j], st_m[j], masks[j], mask_steps[j], (size_t)w[j], h[j], cudaMemcpyHostToDevice, ctx->stream));
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_v[j], st_m[j], masks_orig[j], orig_steps[j], (size_t)w[j], h[j], cudaMemcpyHostToDevice, ctx->stream));
        bt[j] = BlendTile{arena + off_t[j], st_t[j], arena + off_c[j], st_m[j], arena + off_v[j], st_m[j], w[j], h[j], tl_x[j] - mx, tl_y[j] - my};
    }
    const size_t o_step = align_up((size_t)cw * px_bytes, 16);
    uint8_t *d_out;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_CANVAS, o_step * chh, (void **)&d_out)) return rc;
    if (int rc = dev_multiblend(ctx, n, bt.data(), cw, chh, bands, sigma, 0, chh, out_kind, d_out, o_step)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(out, out_step, d_out, o_step, (size_t)cw * px_bytes, chh, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

// ---------------------------------------------------------------------------------------------
// fused path (return_full): sources -> canvas
// ---------------------------------------------------------------------------------------------
namespace {

int composite_impl(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *im_in, int bands, double sigma,
                   int row0, int row1, uint8_t *canvas, size_t canvas_step, bool host, const spano_center_fix *fix = nullptr)
{
    if (n <= 0 || !im_in || !canvas) return spano_fail(ctx, SPANO_E_INVALID, "spano_composite: null/empty argument");
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    std::vector<int> tlx(n), tly(n), ww(n), hh(n);
    for (int j = 0; j < n; ++j) { tlx[j] = im_in[j].tl_x; tly[j] = im_in[j].tl_y; ww[j] = im_in[j].w; hh[j] = im_in[j].h; }
    // Little-planet centre fix (sten_proj::disk_reproj between the warp and the gain, _panorama.cpp:292-311): every tile is
    // re-projected radially, which gives it a new corner (relative to the canvas centre) and a new size; from here on
    // `im` describes the tiles as the BLEND sees them, `orig` as the WARP produces them.
    const spano_image_desc *orig = im_in;
    std::vector<spano_image_desc> fixed;
    SpanoDiskParams DP = {};
    std::vector<int> org_x, org_y;
    size_t pre_max = 0;
    if (fix) {
        for (int j = 0; j < n; ++j)
            if (ww[j] <= 0 || hh[j] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "tile %d is empty", j);
        org_x.resize(n); org_y.resize(n);
        std::vector<int> nx(n), ny(n), nw(n), nh(n);
        if (spano_disk_plan(n, tlx.data(), tly.data(), ww.data(), hh.data(), fix->ansatz_x, fix->ansatz_y, fix->radius, fix->quadratic, &DP,
                            org_x.data(), org_y.data(), nx.data(), ny.data(), nw.data(), nh.data()))
            return spano_fail(ctx, SPANO_E_LIMIT, "degenerate tile in disk_reproj (no border samples)");
        fixed.assign(im_in, im_in + n);
        for (int j = 0; j < n; ++j) {
            if (int rc = check_remap_limits(ctx, ww[j], hh[j], nw[j], nh[j])) return rc;
            pre_max = std::max(pre_max, align_up(align_up((size_t)ww[j] * 3, 16) * hh[j], 256));
            fixed[j].tl_x = tlx[j] = nx[j];  fixed[j].tl_y = tly[j] = ny[j];
            fixed[j].w = ww[j] = nw[j];      fixed[j].h = hh[j] = nh[j];
            if (fixed[j].valid_mask) return spano_fail(ctx, SPANO_E_INVALID, "a precomputed validity mask cannot be combined with the centre fix");
        }
    }
    const spano_image_desc *im = fix ? fixed.data() : im_in;
    int cw, chh, mx, my;
    spano_pan_dimension(n, tlx.data(), tly.data(), ww.data(), hh.data(), &cw, &chh, &mx, &my);
    if (row0 < 0) row0 = 0;
    if (row1 > chh) row1 = chh;
    if (row1 <= row0) return spano_fail(ctx, SPANO_E_INVALID, "empty row band [%d,%d)", row0, row1);
    if (canvas_step < (size_t)cw * 3) return spano_fail(ctx, SPANO_E_INVALID, "canvas_step too small");
    const int rows = row1 - row0;

    // Tiles that touch canvas rows [row0,row1) take part.  BORDER_REFLECT is resolved inside each tile,
    // so a band needs the full extent of exactly those tiles and nothing from neighbouring bands.
    std::vector<int> use;
    size_t src_max = 0, tile_max = 0, mask_max = 0, small_max = 0, field_max = 0, plan_max = 0;
    bool any_small = false;
    for (int j = 0; j < n; ++j) {
        if (int rc = check_image_args(ctx, im[j].src_bgr, im[j].src_w, im[j].src_h, im[j].src_step, 3, "source")) return rc;
        const bool small = im[j].mask_cut_w > 0 || im[j].mask_cut_h > 0;   // preview-scale mask, resized on the device
        if (int rc = check_image_args(ctx, im[j].mask_cut, small ? im[j].mask_cut_w : im[j].w, small ? im[j].mask_cut_h : im[j].h,
                                      im[j].mask_cut_step, 1, "mask_cut"))
            return rc;
        if (small) {
            any_small = true;
            small_max = std::max(small_max, align_up(align_up((size_t)im[j].mask_cut_w, 16) * im[j].mask_cut_h, 256));
        }
        if (im[j].intensity) {
            if (int rc = check_image_args(ctx, im[j].intensity, im[j].intensity_w, im[j].intensity_h, im[j].intensity_step, 4, "intensity field"))
                return rc;
            field_max = std::max(field_max, align_up(align_up((size_t)im[j].intensity_w * 4, 16) * im[j].intensity_h, 256));
        }
        if (!(im[j].gain > 0.0)) return spano_fail(ctx, SPANO_E_INVALID, "gain[%d] must be > 0", j);
        if (int rc = check_remap_limits(ctx, im[j].src_w, im[j].src_h, orig[j].w, orig[j].h)) return rc;
        const int cy = im[j].tl_y - my;
        if (cy + im[j].h <= row0 || cy >= row1) continue;
        use.push_back(j);
        src_max = std::max(src_max, align_up((size_t)im[j].src_w * 3, 16) * im[j].src_h + 16);
        tile_max = std::max(tile_max, align_up(align_up((size_t)im[j].w * 3, 16) * im[j].h, 256));
        mask_max = std::max(mask_max, align_up(align_up((size_t)im[j].w, 16) * im[j].h, 256));
    }
    const int radius = launch_blend_setup(ctx, bands, sigma);
    if (radius < 0) return radius;
    for (int j : use) plan_max = std::max(plan_max, blend_plan_bytes(ctx, im[j].w, bands, radius));
    int *d_plan[2] = {nullptr, nullptr};
    if (plan_max) {
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_BLENDPLAN, plan_max, (void **)&d_plan[0])) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_BLENDPLAN2, plan_max, (void **)&d_plan[1])) return rc;
    }
    float4 *acc = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_ACC, (size_t)cw * rows * sizeof(float4), (void **)&acc)) return rc;
    uint8_t *d_pre = nullptr;   // centre fix: the un-gained warp of the current image, before its radial re-projection
    if (fix && pre_max)
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_PRETILE, pre_max, (void **)&d_pre)) return rc;
    uint8_t *d_tile = nullptr, *d_valid = nullptr, *d_srcbuf[2] = {nullptr, nullptr}, *d_cutbuf[2] = {nullptr, nullptr};
    uint8_t *d_cutsmall[2] = {nullptr, nullptr};
    float *d_field[2] = {nullptr, nullptr};
    if (!use.empty()) {
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, 2 * tile_max, (void **)&d_tile)) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILEMASK, 2 * mask_max, (void **)&d_valid)) return rc;
        if (host) {
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, src_max, (void **)&d_srcbuf[0])) return rc;
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC2, src_max, (void **)&d_srcbuf[1])) return rc;
        }
        if (host || any_small) {
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTMASK, mask_max, (void **)&d_cutbuf[0])) return rc;
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUT2, mask_max, (void **)&d_cutbuf[1])) return rc;
        }
        if (host && any_small) {
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTSMALL, small_max, (void **)&d_cutsmall[0])) return rc;
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTSMALL2, small_max, (void **)&d_cutsmall[1])) return rc;
        }
        if (host && field_max) {
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_FIELD, field_max, (void **)&d_field[0])) return rc;
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_FIELD2, field_max, (void **)&d_field[1])) return rc;
        }
    }
    uint8_t *d_canvas = canvas;
    size_t d_step = canvas_step;
    if (host) {
        d_step = align_up((size_t)cw * 3, 16);
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_CANVAS, d_step * rows, (void **)&d_canvas)) return rc;
        if (!ctx->copy_stream) {
            SPANO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
                SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming));
            }
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        }
        // uploads must not overtake work already queued on the compute stream that still reads the staging buffers
        SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
        SPANO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_start, 0));
    }
    // host path: image i+1 is uploaded on the copy stream while image i is warped and blended
    auto issue_copy = [&](int idx) -> int {
        const int j = use[idx], b = idx & 1;
        cudaStream_t cs = ctx->copy_stream;
        if (idx >= 2) SPANO_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_free[b], 0));
        const size_t row_bytes = (size_t)im[j].src_w * 3, dpitch = align_up(row_bytes, 16);
        if (im[j].src_step == row_bytes && dpitch == row_bytes)   // contiguous on both sides: one linear copy
            SPANO_CUDA(ctx, cudaMemcpyAsync(d_srcbuf[b], im[j].src_bgr, row_bytes * im[j].src_h, cudaMemcpyHostToDevice, cs));
        else
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_srcbuf[b], dpitch, im[j].src_bgr, im[j].src_step, row_bytes, im[j].src_h, cudaMemcpyHostToDevice, cs));
        if (im[j].mask_cut_w > 0 || im[j].mask_cut_h > 0)   // preview-scale mask: resized on the device after the upload
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_cutsmall[b], align_up((size_t)im[j].mask_cut_w, 16), im[j].mask_cut, im[j].mask_cut_step,
                                              (size_t)im[j].mask_cut_w, im[j].mask_cut_h, cudaMemcpyHostToDevice, cs));
        else
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_cutbuf[b], align_up((size_t)im[j].w, 16), im[j].mask_cut, im[j].mask_cut_step,
                                              (size_t)im[j].w, im[j].h, cudaMemcpyHostToDevice, cs));
        SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], cs));
        return 0;
    };
    // Warp + validity mask of image i+1 run on an auxiliary stream while image i is blended on the main
    // stream (two tile buffers): a blend CTA leaves 16 K registers and 1.4 KB of shared memory of its SM free, room
    // for one small CTA, so every kernel of that chain is kept within 64 registers x 256 threads and without
    // shared memory of its own (one that does not fit waits for the blend to end and stalls the chain behind it).
    if (!ctx->aux_stream) {
        SPANO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_warped[b], cudaEventDisableTiming));
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_blended[b], cudaEventDisableTiming));
        }
        SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start2, cudaEventDisableTiming));
    }
    cudaStream_t main_stream = ctx->stream, aux = ctx->aux_stream;
    SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_start2, main_stream));
    SPANO_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->ev_start2, 0));
    {   // the accumulator clear (pure HBM writes) runs while the auxiliary stream already warps the first image
        StageTimer t(ctx, 2);
        if (int rc = launch_blend_clear(ctx, acc, cw, rows)) return rc;
        t.stop(0);
    }
    struct StreamSwap {
        spano_ctx *c; cudaStream_t keep;
        StreamSwap(spano_ctx *ctx_, cudaStream_t s) : c(ctx_), keep(ctx_->stream) { c->stream = s; }
        ~StreamSwap() { c->stream = keep; }
    };
    // host path: for every canvas column the last image (in processing order) whose tile covers it
    struct ColRun { int c0, c1, idx; };
    std::vector<ColRun> runs;
    std::vector<cudaEvent_t> flush_events;
    // (device path too: the finished columns are normalised on the auxiliary stream while the next images blend)
    if (!use.empty()) {
        if (host && !ctx->d2h_stream) SPANO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        std::vector<int> last(cw, 0);   // uncovered columns are final from the start: flushed after the first image
        for (int idx = 0; idx < (int)use.size(); ++idx) {
            const int j = use[idx];
            const int c0 = std::max(0, im[j].tl_x - mx), c1 = std::min(cw, im[j].tl_x - mx + im[j].w);
            for (int c = c0; c < c1; ++c) last[c] = idx;
        }
        for (int c = 0; c < cw;) {
            int e = c + 1;
            while (e < cw && last[e] == last[c]) ++e;
            runs.push_back(ColRun{c, e, last[c]});
            c = e;
        }
    }
    if (host && !use.empty())
        if (int rc = issue_copy(0)) return rc;
    for (int idx = 0; idx < (int)use.size(); ++idx) {
        const int j = use[idx], b = idx & 1;
        if (host && idx + 1 < (int)use.size())
            if (int rc = issue_copy(idx + 1)) return rc;
        SpanoProjector P;
        spano_host_set_camera(&P, proj, scale, im[j].K, im[j].R);
        const uint8_t *src = im[j].src_bgr, *cut = im[j].mask_cut;
        size_t s_step = im[j].src_step, c_step = im[j].mask_cut_step;
        const size_t t_step = align_up((size_t)im[j].w * 3, 16), m_step = align_up((size_t)im[j].w, 16);
        uint8_t *tile_b = d_tile + (size_t)b * tile_max, *valid_b = d_valid + (size_t)b * mask_max;
        const uint8_t *valid = valid_b;
        size_t v_step = m_step;
        {   // ---- auxiliary stream: warp (+ mask) into tile buffer b ----
            StreamSwap sw(ctx, aux);
            if (idx >= 2) SPANO_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->ev_blended[b], 0));
            if (host) {
                SPANO_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->ev_copied[b], 0));
                src = d_srcbuf[b];  s_step = align_up((size_t)im[j].src_w * 3, 16);
                cut = d_cutbuf[b];  c_step = m_step;
            }
            bool planned = false;
            if (im[j].mask_cut_w > 0 || im[j].mask_cut_h > 0) {
                // mask_cut is at preview scale: cv::resize(.., tile size) with the 8-bit INTER_LINEAR arithmetic.  With a
                // sparsity plan the plan comes FIRST, from the preview mask, and only the rows of the tile-size mask that
                // the plan's pieces read are produced (most of a seam mask is zero and never looked at).
                const uint8_t *small = host ? d_cutsmall[b] : im[j].mask_cut;
                const size_t small_step = host ? align_up((size_t)im[j].mask_cut_w, 16) : im[j].mask_cut_step;
                if (d_plan[b]) {
                    const BlendTile pt{nullptr, 0, d_cutbuf[b], m_step, nullptr, 0, im[j].w, im[j].h, im[j].tl_x - mx, im[j].tl_y - my};
                    int kk = launch_blend_plan_preview(ctx, pt, small, im[j].mask_cut_w, im[j].mask_cut_h, small_step, bands, radius, row0, row1,
                                                       d_plan[b], cw);
                    if (kk < 0) return kk;
                    planned = kk > 0;
                }
                if (!planned) {
                    int kk = launch_resize_mask(ctx, small, im[j].mask_cut_w, im[j].mask_cut_h, small_step, d_cutbuf[b], im[j].w, im[j].h, m_step);
                    if (kk < 0) return kk;
                }
                cut = d_cutbuf[b];  c_step = m_step;
            }
            if (!host && im[j].valid_mask) {
                // mask supplied (computed elsewhere for the whole tile): warp only the rows this band reads --
                // its own rows plus the blur radius, which BORDER_REFLECT keeps inside the tile
                if (im[j].valid_mask_step < (size_t)im[j].w) return spano_fail(ctx, SPANO_E_INVALID, "valid_mask step too small");
                valid = im[j].valid_mask;
                v_step = im[j].valid_mask_step;
                const int cy = im[j].tl_y - my;
                int r0 = std::max(0, row0 - cy) - radius, r1 = std::min(im[j].h, row1 - cy) + radius;
                if (im[j].h < 4 * radius) { r0 = 0; r1 = im[j].h; }     // several reflections possible: keep it simple
                r0 = std::max(0, r0);
                r1 = std::min(im[j].h, r1);
                StageTimer t0(ctx, 0);
                int kk = launch_warp(ctx, P, src, im[j].src_w, im[j].src_h, s_step, im[j].gain, im[j].tl_x, im[j].tl_y, im[j].w, im[j].h,
                                     r0, r1, tile_b, t_step, nullptr, 0);
                if (kk < 0) return kk;
                t0.stop(kk);
            } else if (fix) {
                // warp (no gain, no mask) -> radial gather into the new tile -> dark flags + validity mask -> gain
                const size_t p_step = align_up((size_t)orig[j].w * 3, 16);
                StageTimer t0(ctx, 0);
                int kk = launch_warp(ctx, P, src, orig[j].src_w, orig[j].src_h, s_step, 1.0, orig[j].tl_x, orig[j].tl_y, orig[j].w, orig[j].h, 0,
                                     orig[j].h, d_pre, p_step, nullptr, 0);
                if (kk < 0) return kk;
                int k2 = launch_disk_gather(ctx, DP, d_pre, orig[j].w, orig[j].h, p_step, org_x[j], org_y[j], tile_b, im[j].w, im[j].h, t_step,
                                            im[j].tl_x, im[j].tl_y);
                if (k2 < 0) return k2;
                t0.stop(kk + k2);
                uint8_t *dark = nullptr;
                if (int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, m_step * im[j].h, (void **)&dark)) return rc;
                StageTimer t1(ctx, 1);
                int k3 = launch_dark_flags(ctx, tile_b, im[j].w, im[j].h, t_step, dark, m_step);
                if (k3 < 0) return k3;
                int k4 = launch_valid_mask(ctx, dark, im[j].w, im[j].h, m_step, 3, valid_b, m_step);
                if (k4 < 0) return k4;
                t1.stop(k3 + k4);
                int k5 = launch_gain(ctx, tile_b, im[j].w, im[j].h, t_step, im[j].gain);
                if (k5 < 0) return k5;
            } else if (int rc = dev_warp_tile(ctx, P, src, im[j].src_w, im[j].src_h, s_step, im[j].gain, im[j].tl_x, im[j].tl_y, im[j].w,
                                              im[j].h, tile_b, t_step, valid_b, m_step))
                return rc;
            if (im[j].intensity) {
                // test::adjust_intensity on the gained tile (conf.blend_intensity), before the blend
                const float *fld = im[j].intensity;
                size_t fpitch = im[j].intensity_step / 4;
                if (host) {
                    const size_t f_step = align_up((size_t)im[j].intensity_w * 4, 16);
                    SPANO_CUDA(ctx, cudaMemcpy2DAsync(d_field[b], f_step, im[j].intensity, im[j].intensity_step, (size_t)im[j].intensity_w * 4,
                                                      im[j].intensity_h, cudaMemcpyHostToDevice, aux));
                    fld = d_field[b];
                    fpitch = f_step / 4;
                }
                int kk = launch_adjust_intensity(ctx, tile_b, im[j].w, im[j].h, t_step, fld, im[j].intensity_w, im[j].intensity_h, fpitch);
                if (kk < 0) return kk;
            }
            if (d_plan[b] && !planned) {
                // the blend's sparsity plan (activity of mask_cut -> pieces) is made here, one image ahead, so that
                // the blends follow each other back to back on the main stream
                const BlendTile pt{tile_b, t_step, cut, c_step, valid, v_step, im[j].w, im[j].h, im[j].tl_x - mx, im[j].tl_y - my};
                int kk = launch_blend_plan(ctx, pt, bands, radius, row0, row1, d_plan[b], cw);
                if (kk < 0) return kk;
            }
            SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_warped[b], aux));
        }
        // ---- main stream: blend tile b into the accumulator ----
        SPANO_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_warped[b], 0));
        StageTimer t2(ctx, 2);
        const BlendTile bt{tile_b, t_step, cut, c_step, valid, v_step, im[j].w, im[j].h, im[j].tl_x - mx, im[j].tl_y - my};
        int k = launch_blend_tile(ctx, bt, bands, radius, acc, cw, row0, row1, d_plan[b]);
        if (k < 0) return k;
        t2.stop(k);
        SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_blended[b], main_stream));
        if (host) SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_free[b], main_stream));
        // canvas columns that no later image touches are final: normalise them now (host path: and download them on
        // the download stream) while the remaining images are still being blended
        for (const ColRun &r : runs) {
            if (r.idx != idx) continue;
            StageTimer t3(ctx, 3);
            int kn = launch_normalise(ctx, acc, cw, rows, bands, SPANO_OUT_U8, d_canvas, d_step, r.c0, r.c1);
            if (kn < 0) return kn;
            t3.stop(kn);
            if (!host) continue;
            cudaEvent_t e;
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            flush_events.push_back(e);
            SPANO_CUDA(ctx, cudaEventRecord(e, main_stream));
            SPANO_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, e, 0));
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(canvas + (size_t)r.c0 * 3, canvas_step, d_canvas + (size_t)r.c0 * 3, d_step,
                                              (size_t)(r.c1 - r.c0) * 3, rows, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
    }
    if (use.empty()) {
        StageTimer t3(ctx, 3);
        int k = launch_normalise(ctx, acc, cw, rows, bands, SPANO_OUT_U8, d_canvas, d_step);
        if (k < 0) return k;
        t3.stop(k);
        if (host)
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(canvas, canvas_step, d_canvas, d_step, (size_t)cw * 3, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (host) {
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
        for (cudaEvent_t e : flush_events) cudaEventDestroy(e);
    }
    return SPANO_OK;
}

} // namespace

extern "C" int spano_composite(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                               double sigma, int row0, int row1, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return composite_impl(ctx, proj, scale, n, images, bands, sigma, row0, row1, canvas, canvas_step, true);
}

extern "C" int spano_dev_composite(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images,
                                   int bands, double sigma, int row0, int row1, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return composite_impl(ctx, proj, scale, n, images, bands, sigma, row0, row1, canvas, canvas_step, false);
}

namespace {
int check_fix(spano_ctx *ctx, int proj, const spano_center_fix *fix)
{
    if (!fix) return 0;
    if (proj != SPANO_STEREOGRAPHIC) return spano_fail(ctx, SPANO_E_INVALID, "the centre fix belongs to the stereographic projection (conf.proj == STEREOGRAPHIC)");
    if (!(fix->radius > 0.f)) return spano_fail(ctx, SPANO_E_INVALID, "centre fix: radius must be > 0 (estimate_circle found no circle: pass NULL)");
    return 0;
}
} // namespace

extern "C" int spano_composite_fixed(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                                     double sigma, const spano_center_fix *fix, int row0, int row1, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_fix(ctx, proj, fix)) return rc;
    return composite_impl(ctx, proj, scale, n, images, bands, sigma, row0, row1, canvas, canvas_step, true, fix);
}

extern "C" int spano_dev_composite_fixed(spano_ctx *ctx, int proj, float scale, int n, const spano_image_desc *images, int bands,
                                         double sigma, const spano_center_fix *fix, int row0, int row1, uint8_t *canvas,
                                         size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_fix(ctx, proj, fix)) return rc;
    return composite_impl(ctx, proj, scale, n, images, bands, sigma, row0, row1, canvas, canvas_step, false, fix);
}

// ---------------------------------------------------------------------------------------------
// tile-sharded multi-GPU path: peer memory, owner-side warp + scatter, band-side incremental blend
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "spano.h promises a 64-byte handle");

extern "C" int spano_peer_alloc(spano_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64])
{
    if (!ctx || !dptr || !handle || bytes == 0) return ctx ? spano_fail(ctx, SPANO_E_INVALID, "spano_peer_alloc: bad argument") : SPANO_E_INVALID;
    Guard g(ctx);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return spano_fail(ctx, SPANO_E_NOMEM, "cudaMalloc(%zu B) failed: %s", bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return spano_fail(ctx, SPANO_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, 64);
    *dptr = p;
    return SPANO_OK;
}

extern "C" int spano_peer_open(spano_ctx *ctx, const unsigned char handle[64], void **dptr)
{
    if (!ctx || !dptr || !handle) return ctx ? spano_fail(ctx, SPANO_E_INVALID, "spano_peer_open: bad argument") : SPANO_E_INVALID;
    Guard g(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    SPANO_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return SPANO_OK;
}

extern "C" int spano_peer_close(spano_ctx *ctx, void *dptr)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (dptr) SPANO_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return SPANO_OK;
}

extern "C" int spano_peer_free(spano_ctx *ctx, void *dptr)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (dptr) {
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SPANO_CUDA(ctx, cudaFree(dptr));
    }
    return SPANO_OK;
}

namespace {

int warp_scatter_impl(spano_ctx *ctx, int proj, float scale, const spano_image_desc *im, int n_slices, const spano_slice *slices,
                      bool host)
{
    if (!im || n_slices < 0 || (n_slices > 0 && !slices)) return spano_fail(ctx, SPANO_E_INVALID, "spano_warp_scatter: null argument");
    if (n_slices > SPANO_MAX_SLICES) return spano_fail(ctx, SPANO_E_INVALID, "spano_warp_scatter: at most %d slices per tile", SPANO_MAX_SLICES);
    if (int rc = valid_proj(ctx, proj, scale)) return rc;
    if (int rc = check_image_args(ctx, im->src_bgr, im->src_w, im->src_h, im->src_step, 3, "source")) return rc;
    if (im->w <= 0 || im->h <= 0) return spano_fail(ctx, SPANO_E_INVALID, "empty tile %dx%d", im->w, im->h);
    if (!(im->gain > 0.0)) return spano_fail(ctx, SPANO_E_INVALID, "gain must be > 0");
    if (im->intensity) return spano_fail(ctx, SPANO_E_INVALID, "the intensity field is not supported on the tile-sharded path");
    if (int rc = check_remap_limits(ctx, im->src_w, im->src_h, im->w, im->h)) return rc;
    SpanoScatter st, sm;
    for (int d = 0; d < n_slices; ++d) {
        const spano_slice &s = slices[d];
        if (s.row0 < 0 || s.row1 > im->h || s.row1 <= s.row0) return spano_fail(ctx, SPANO_E_INVALID, "slice %d: rows [%d,%d) outside the tile (h=%d)", d, s.row0, s.row1, im->h);
        int c0 = 0, c1 = im->w;
        if (s.col1 > s.col0) {   // the slice holds a column range of the tile (32-pixel granularity: one word of the mask kernel)
            c0 = s.col0;  c1 = s.col1;
            if (c0 < 0 || c1 > im->w || (c0 & 31) || ((c1 & 31) && c1 != im->w))
                return spano_fail(ctx, SPANO_E_INVALID, "slice %d: columns [%d,%d) must lie inside the tile (w=%d) on multiples of 32", d, c0, c1, im->w);
        }
        if (!s.tile || !s.valid || s.tile_step < (size_t)(c1 - c0) * 3 || s.valid_step < (size_t)(c1 - c0))
            return spano_fail(ctx, SPANO_E_INVALID, "slice %d: null pointer or step too small", d);
        st.row0[d] = sm.row0[d] = s.row0;
        st.row1[d] = sm.row1[d] = s.row1;
        st.col0[d] = sm.col0[d] = c0;
        st.col1[d] = sm.col1[d] = c1;
        st.base[d] = s.tile - (size_t)s.row0 * s.tile_step - (size_t)c0 * 3;   st.step[d] = s.tile_step;
        sm.base[d] = s.valid - (size_t)s.row0 * s.valid_step - (size_t)c0;     sm.step[d] = s.valid_step;
    }
    st.n = sm.n = n_slices;
    if (n_slices == 0) return SPANO_OK;
    const uint8_t *src = im->src_bgr;
    size_t s_step = im->src_step;
    if (host) {
        uint8_t *stage = nullptr;
        s_step = align_up((size_t)im->src_w * 3, 16);
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, s_step * im->src_h + 16, (void **)&stage)) return rc;
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(stage, s_step, im->src_bgr, im->src_step, (size_t)im->src_w * 3, im->src_h, cudaMemcpyHostToDevice, ctx->stream));
        src = stage;
    }
    uint8_t *dark = nullptr;
    const size_t dark_step = align_up((size_t)im->w, 16);
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, dark_step * im->h, (void **)&dark)) return rc;
    SpanoProjector P;
    spano_host_set_camera(&P, proj, scale, im->K, im->R);
    StageTimer t0(ctx, 0);
    int n = launch_warp(ctx, P, src, im->src_w, im->src_h, s_step, im->gain, im->tl_x, im->tl_y, im->w, im->h, 0, im->h, nullptr, 0, dark,
                        dark_step, nullptr, nullptr, &st);
    if (n < 0) return n;
    t0.stop(n);
    StageTimer t1(ctx, 1);
    n = launch_valid_mask(ctx, dark, im->w, im->h, dark_step, 3, nullptr, 0, &sm);
    if (n < 0) return n;
    t1.stop(n);
    return SPANO_OK;
}

// mask_cut of image `im` at tile size for the tile rows [need0, need1): returns the VIRTUAL address of row 0 and the
// step.  Preview-scale masks are up-scaled (host ones are first looked up among those staged at begin, else uploaded),
// tile-sized host masks are uploaded, tile-sized device masks are used in place.  Work goes to ctx->stream; the
// resized rows land in buffer `buf` at byte offset `buf_off`.
// `defer` (optional): when the mask is at preview scale it is NOT up-scaled here; *defer receives the (device) preview mask
// and the caller up-scales what it needs (launch_blend_plan_preview, or launch_resize_mask for the rows [need0, need1)).
struct PreviewRef { const uint8_t *small = nullptr; int sw = 0, sh = 0; size_t sstep = 0; };
int blend_resolve_cut(spano_ctx *ctx, const spano_image_desc *im, int need0, int need1, bool host, int buf, size_t buf_off,
                      const uint8_t **cut_v, size_t *cut_step, PreviewRef *defer = nullptr)
{
    spano_ctx::BlendSession &S = ctx->bs;
    const bool small = im->mask_cut_w > 0 || im->mask_cut_h > 0;
    if (int rc = check_image_args(ctx, im->mask_cut, small ? im->mask_cut_w : im->w, small ? im->mask_cut_h : im->h, im->mask_cut_step, 1, "mask_cut"))
        return rc;
    const size_t m_step = align_up((size_t)im->w, 16);
    if (!small && !host) {
        *cut_v = im->mask_cut;
        *cut_step = im->mask_cut_step;
        return 0;
    }
    uint8_t *cutbuf = (uint8_t *)ctx->buf[buf].ptr;
    if (buf == spano_ctx::BUF_CUTMASK) {
        if (int rc = spano_reserve(ctx, buf, m_step * (need1 - need0) + 256, (void **)&cutbuf)) return rc;
    }
    cutbuf += buf_off;
    uint8_t *cut_base = cutbuf - (size_t)need0 * m_step;
    if (small) {
        const uint8_t *sm = im->mask_cut;
        size_t sm_step = im->mask_cut_step;
        const spano_ctx::BlendSession::Staged *hit = nullptr;
        if (host)
            for (const auto &st : S.staged)
                if (st.host == im->mask_cut) { hit = &st; break; }
        if (hit) {
            sm = hit->dev;
            sm_step = hit->step;
        } else if (host) {
            uint8_t *stage = nullptr;
            sm_step = align_up((size_t)im->mask_cut_w, 16);
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTSMALL2, sm_step * im->mask_cut_h + 256, (void **)&stage)) return rc;
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(stage, sm_step, im->mask_cut, im->mask_cut_step, (size_t)im->mask_cut_w, im->mask_cut_h,
                                              cudaMemcpyHostToDevice, ctx->stream));
            sm = stage;
        }
        if (defer) {
            defer->small = sm;  defer->sw = im->mask_cut_w;  defer->sh = im->mask_cut_h;  defer->sstep = sm_step;
        } else {
            int k = launch_resize_mask(ctx, sm, im->mask_cut_w, im->mask_cut_h, sm_step, cut_base, im->w, im->h, m_step, need0, need1);
            if (k < 0) return k;
        }
    } else {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(cutbuf, m_step, im->mask_cut + (size_t)need0 * im->mask_cut_step, im->mask_cut_step, (size_t)im->w,
                                          need1 - need0, cudaMemcpyHostToDevice, ctx->stream));
    }
    *cut_v = cut_base;
    *cut_step = m_step;
    return 0;
}

// tile rows of `im` that band [S.row0, S.row1) produces and reads; false when the tile does not touch the band
bool blend_rows(const spano_ctx::BlendSession &S, const spano_image_desc *im, int *first, int *last, int *need0, int *need1)
{
    const int cy = im->tl_y - S.my;
    *first = std::max(0, S.row0 - cy);
    *last = std::min(im->h, S.row1 - cy);
    if (*last <= *first) return false;
    const int R = S.radius;
    *need0 = (im->h < 4 * R) ? 0 : std::max(0, *first - R);
    *need1 = (im->h < 4 * R) ? im->h : std::min(im->h, *last + R);
    return true;
}

int blend_prepare_impl(spano_ctx *ctx, int n, const spano_image_desc *images, const spano_slice *slices, bool host)
{
    spano_ctx::BlendSession &S = ctx->bs;
    if (!S.open) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_prepare without spano_dev_blend_begin");
    if (n < 0 || (n > 0 && (!images || !slices))) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_prepare: null argument");
    S.prepared.clear();
    // layout of the two arenas
    std::vector<size_t> off_c(n, 0), off_p(n, 0);
    std::vector<char> take(n, 0);
    size_t cut_total = 0, plan_total = 0;
    for (int j = 0; j < n; ++j) {
        const spano_image_desc *im = images + j;
        int first, last, need0, need1;
        if (im->w <= 0 || im->h <= 0 || slices[j].row1 <= slices[j].row0 || !blend_rows(S, im, &first, &last, &need0, &need1)) continue;
        const bool small = im->mask_cut_w > 0 || im->mask_cut_h > 0;
        take[j] = 1;
        if (small || host) {
            off_c[j] = cut_total;
            cut_total += align_up(align_up((size_t)im->w, 16) * (need1 - need0), 256);
        }
        off_p[j] = plan_total;
        plan_total += align_up(blend_plan_bytes(ctx, im->w, S.bands, S.radius), 256);
    }
    uint8_t *cut_arena = nullptr, *plan_arena = nullptr;
    if (cut_total)
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_PREP_CUT, cut_total, (void **)&cut_arena)) return rc;
    if (plan_total)
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_PREP_PLAN, plan_total, (void **)&plan_arena)) return rc;
    if (!ctx->aux_stream) {
        SPANO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_warped[b], cudaEventDisableTiming));
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_blended[b], cudaEventDisableTiming));
        }
        SPANO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start2, cudaEventDisableTiming));
    }
    // the auxiliary stream starts after everything queued so far (begin's uploads, the previous step's blends)
    SPANO_CUDA(ctx, cudaEventRecord(ctx->ev_start2, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_start2, 0));
    struct StreamSwap {
        spano_ctx *c; cudaStream_t keep;
        StreamSwap(spano_ctx *ctx_, cudaStream_t s) : c(ctx_), keep(ctx_->stream) { c->stream = s; }
        ~StreamSwap() { c->stream = keep; }
    } sw(ctx, ctx->aux_stream);
    size_t ev_used = 0;
    for (int j = 0; j < n; ++j) {
        if (!take[j]) continue;
        const spano_image_desc *im = images + j;
        int first, last, need0, need1;
        blend_rows(S, im, &first, &last, &need0, &need1);
        const uint8_t *cut_v = nullptr;
        size_t c_step = 0;
        PreviewRef pv;
        if (int rc = blend_resolve_cut(ctx, im, need0, need1, host, spano_ctx::BUF_PREP_CUT, off_c[j], &cut_v, &c_step, &pv)) return rc;
        int *plan = nullptr;
        const bool want_plan = blend_plan_bytes(ctx, im->w, S.bands, S.radius) != 0;
        if (want_plan) plan = reinterpret_cast<int *>(plan_arena + off_p[j]);
        const BlendTile pt{nullptr, 0, cut_v, c_step, nullptr, 0, im->w, im->h, im->tl_x - S.mx, im->tl_y - S.my};
        bool planned = false;
        if (pv.small && want_plan) {   // plan from the preview mask; only the rows / strips the plan reads are up-scaled
            int k = launch_blend_plan_preview(ctx, pt, pv.small, pv.sw, pv.sh, pv.sstep, S.bands, S.radius, S.row0, S.row1, plan, S.cw);
            if (k < 0) return k;
            planned = k > 0;
        }
        if (pv.small && !planned) {    // no plan applies: all the rows the band reads
            int k = launch_resize_mask(ctx, pv.small, pv.sw, pv.sh, pv.sstep, const_cast<uint8_t *>(cut_v), im->w, im->h, c_step, need0, need1);
            if (k < 0) return k;
        }
        if (want_plan && !planned) {
            int k = launch_blend_plan(ctx, pt, S.bands, S.radius, S.row0, S.row1, plan, S.cw);
            if (k < 0) return k;
        }
        if (ev_used == ctx->event_pool.size()) {
            cudaEvent_t e;
            SPANO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->event_pool.push_back(e);
        }
        cudaEvent_t ready = ctx->event_pool[ev_used++];
        SPANO_CUDA(ctx, cudaEventRecord(ready, ctx->stream));
        S.prepared.push_back({im, cut_v, c_step, plan, ready});
    }
    return SPANO_OK;
}

// host variant: normalise + download the canvas column runs whose last image is `idx` (idx < 0: everything left)
int blend_flush_runs(spano_ctx *ctx, int idx)
{
    spano_ctx::BlendSession &S = ctx->bs;
    const int rows = S.row1 - S.row0;
    for (auto &r : S.runs) {
        if (r.flushed || (idx >= 0 && r.idx != idx)) continue;
        r.flushed = true;
        StageTimer t3(ctx, 3);
        int kn = launch_normalise(ctx, S.acc, S.cw, rows, S.bands, SPANO_OUT_U8, S.d_canvas, S.d_step, r.c0, r.c1);
        if (kn < 0) return kn;
        t3.stop(kn);
        cudaEvent_t e;
        SPANO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        S.events.push_back(e);
        SPANO_CUDA(ctx, cudaEventRecord(e, ctx->stream));
        SPANO_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, e, 0));
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(S.h_canvas + (size_t)r.c0 * 3, S.h_step, S.d_canvas + (size_t)r.c0 * 3, S.d_step, (size_t)(r.c1 - r.c0) * 3,
                                          rows, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    return SPANO_OK;
}

int blend_add_impl(spano_ctx *ctx, const spano_image_desc *im, const spano_slice *slice, bool host)
{
    spano_ctx::BlendSession &S = ctx->bs;
    if (!S.open) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_add without spano_dev_blend_begin");
    if (!im || !slice) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_add: null argument");
    if (im->w <= 0 || im->h <= 0) return spano_fail(ctx, SPANO_E_INVALID, "empty tile %dx%d", im->w, im->h);
    const int cy = im->tl_y - S.my;
    const int first = std::max(0, S.row0 - cy), last = std::min(im->h, S.row1 - cy);
    if (last <= first) return SPANO_OK;   // the tile does not touch this band
    const int R = S.radius;
    const int need0 = (im->h < 4 * R) ? 0 : std::max(0, first - R), need1 = (im->h < 4 * R) ? im->h : std::min(im->h, last + R);
    if (slice->row0 > need0 || slice->row1 < need1)
        return spano_fail(ctx, SPANO_E_INVALID, "slice rows [%d,%d) do not cover the rows the band reads [%d,%d)", slice->row0, slice->row1, need0, need1);
    // columns: the slice may hold a column range of the tile; the band reads the strips (32 columns) that intersect its own
    // columns, plus the blur radius (BORDER_REFLECT stays inside the tile, as for the rows)
    int sc0 = 0, sc1 = im->w;
    if (slice->col1 > slice->col0) { sc0 = slice->col0;  sc1 = slice->col1; }
    {
        const int cx = im->tl_x - S.mx;
        const int wx0 = std::max(0, -cx), wx1 = std::min(im->w, S.cw - cx);
        if (wx1 <= wx0) return SPANO_OK;   // the tile does not touch this band's columns
        const int needc0 = (im->w < 4 * R) ? 0 : std::max(0, (wx0 & ~31) - R), needc1 = (im->w < 4 * R) ? im->w : std::min(im->w, ((wx1 + 31) & ~31) + R);
        if (sc0 > needc0 || sc1 < needc1)
            return spano_fail(ctx, SPANO_E_INVALID, "slice columns [%d,%d) do not cover the columns the band reads [%d,%d)", sc0, sc1, needc0, needc1);
    }
    if (!slice->tile || !slice->valid || slice->tile_step < (size_t)(sc1 - sc0) * 3 || slice->valid_step < (size_t)(sc1 - sc0))
        return spano_fail(ctx, SPANO_E_INVALID, "slice: null pointer or step too small");
    const uint8_t *cut_v = nullptr;   // virtual address of mask_cut row 0 at tile size
    size_t c_step = 0;
    const int *plan = nullptr;
    for (const auto &pr : S.prepared)
        if (pr.im == im) {               // prepared ahead of time (spano_*_blend_prepare)
            SPANO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pr.ready, 0));
            cut_v = pr.cut_v;  c_step = pr.cut_step;  plan = pr.plan;
            break;
        }
    if (!cut_v) {
        if (int rc = blend_resolve_cut(ctx, im, need0, need1, host, spano_ctx::BUF_CUTMASK, 0, &cut_v, &c_step)) return rc;
    }
    StageTimer t2(ctx, 2);
    const BlendTile bt{slice->tile - (size_t)slice->row0 * slice->tile_step - (size_t)sc0 * 3, slice->tile_step, cut_v, c_step,
                       slice->valid - (size_t)slice->row0 * slice->valid_step - (size_t)sc0, slice->valid_step, im->w, im->h, im->tl_x - S.mx, cy};
    int k = launch_blend_tile(ctx, bt, S.bands, S.radius, S.acc, S.cw, S.row0, S.row1, plan);
    if (k < 0) return k;
    t2.stop(k);
    if (host && S.h_canvas && im >= S.images && im < S.images + S.n_images)
        return blend_flush_runs(ctx, (int)(im - S.images));
    return SPANO_OK;
}

int blend_finish_impl(spano_ctx *ctx, uint8_t *canvas, size_t canvas_step, bool host)
{
    spano_ctx::BlendSession &S = ctx->bs;
    if (!S.open) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_finish without spano_dev_blend_begin");
    if (!canvas || canvas_step < (size_t)S.cw * 3) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_finish: null canvas or step too small");
    const int rows = S.row1 - S.row0;
    if (host && S.h_canvas) {
        // the canvas was announced at begin: most columns are already on their way; flush the rest and wait
        if (canvas != S.h_canvas || canvas_step != S.h_step) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_finish: not the canvas announced at spano_blend_begin");
        if (int rc = blend_flush_runs(ctx, -1)) return rc;
        S.open = false;
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
        for (cudaEvent_t e : S.events) cudaEventDestroy(e);
        S.events.clear();
        return SPANO_OK;
    }
    uint8_t *d_canvas = canvas;
    size_t d_step = canvas_step;
    if (host) {
        d_step = align_up((size_t)S.cw * 3, 16);
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_CANVAS, d_step * rows, (void **)&d_canvas)) return rc;
    }
    StageTimer t3(ctx, 3);
    int k = launch_normalise(ctx, S.acc, S.cw, rows, S.bands, SPANO_OUT_U8, d_canvas, d_step);
    if (k < 0) return k;
    t3.stop(k);
    S.open = false;
    if (host) {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(canvas, canvas_step, d_canvas, d_step, (size_t)S.cw * 3, rows, cudaMemcpyDeviceToHost, ctx->stream));
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SPANO_OK;
}

} // namespace

extern "C" int spano_dev_warp_scatter(spano_ctx *ctx, int proj, float scale, const spano_image_desc *im, int n_slices, const spano_slice *slices)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return warp_scatter_impl(ctx, proj, scale, im, n_slices, slices, false);
}

extern "C" int spano_warp_scatter(spano_ctx *ctx, int proj, float scale, const spano_image_desc *im, int n_slices, const spano_slice *slices)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return warp_scatter_impl(ctx, proj, scale, im, n_slices, slices, true);
}

namespace {
int blend_begin_impl(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma);
}

extern "C" int spano_dev_blend_begin(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_begin_impl(ctx, canvas_w, min_x, min_y, row0, row1, bands, sigma);
}

namespace {
// host variant of begin: uploads the preview-scale masks of the tiles that touch the band right away (ahead of the owners'
// large source uploads) and, when the destination canvas is announced, prepares the early column download
int blend_begin_host_extras(spano_ctx *ctx, int n, const spano_image_desc *images, uint8_t *canvas, size_t canvas_step)
{
    spano_ctx::BlendSession &S = ctx->bs;
    // preview-scale masks of the tiles that touch this band: one staging arena, uploaded now
    size_t total = 0;
    std::vector<size_t> off(n, (size_t)-1);
    for (int j = 0; j < n; ++j) {
        const spano_image_desc &im = images[j];
        if (!(im.mask_cut_w > 0 || im.mask_cut_h > 0) || !im.mask_cut) continue;
        const int cy = im.tl_y - S.my;
        if (std::min(im.h, S.row1 - cy) <= std::max(0, S.row0 - cy)) continue;
        if (int rc = check_image_args(ctx, im.mask_cut, im.mask_cut_w, im.mask_cut_h, im.mask_cut_step, 1, "mask_cut")) { S.open = false; return rc; }
        off[j] = total;
        total += align_up(align_up((size_t)im.mask_cut_w, 16) * im.mask_cut_h, 256);
    }
    if (total) {
        uint8_t *arena = nullptr;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_CUTSMALL, total, (void **)&arena)) { S.open = false; return rc; }
        for (int j = 0; j < n; ++j) {
            if (off[j] == (size_t)-1) continue;
            const spano_image_desc &im = images[j];
            const size_t st = align_up((size_t)im.mask_cut_w, 16);
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off[j], st, im.mask_cut, im.mask_cut_step, (size_t)im.mask_cut_w, im.mask_cut_h,
                                              cudaMemcpyHostToDevice, ctx->stream));
            S.staged.push_back({im.mask_cut, arena + off[j], st});
        }
    }
    if (canvas && n > 0) {
        if (canvas_step < (size_t)S.cw * 3) { S.open = false; return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_begin: canvas_step too small"); }
        const int rows = S.row1 - S.row0;
        S.d_step = align_up((size_t)S.cw * 3, 16);
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_CANVAS, S.d_step * rows, (void **)&S.d_canvas)) { S.open = false; return rc; }
        if (!ctx->d2h_stream) SPANO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        S.images = images;  S.n_images = n;  S.h_canvas = canvas;  S.h_step = canvas_step;
        // for every canvas column the last announced image (in array order) whose tile covers it inside this band;
        // columns nobody covers are flushed at finish
        std::vector<int> last(S.cw, -1);
        for (int j = 0; j < n; ++j) {
            const int cy = images[j].tl_y - S.my;
            if (std::min(images[j].h, S.row1 - cy) <= std::max(0, S.row0 - cy)) continue;
            const int c0 = std::max(0, images[j].tl_x - S.mx), c1 = std::min(S.cw, images[j].tl_x - S.mx + images[j].w);
            for (int c = c0; c < c1; ++c) last[c] = j;
        }
        for (int c = 0; c < S.cw;) {
            int e = c + 1;
            while (e < S.cw && last[e] == last[c]) ++e;
            S.runs.push_back({c, e, last[c], false});
            c = e;
        }
    }
    return SPANO_OK;
}
} // namespace

extern "C" int spano_blend_begin(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma,
                                 int n, const spano_image_desc *images, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n < 0 || (n > 0 && !images)) return spano_fail(ctx, SPANO_E_INVALID, "spano_blend_begin: null images");
    if (int rc = blend_begin_impl(ctx, canvas_w, min_x, min_y, row0, row1, bands, sigma)) return rc;
    return blend_begin_host_extras(ctx, n, images, canvas, canvas_step);
}

namespace {
int blend_begin_impl(spano_ctx *ctx, int canvas_w, int min_x, int min_y, int row0, int row1, int bands, double sigma)
{
    if (canvas_w <= 0 || row1 <= row0 || row0 < 0) return spano_fail(ctx, SPANO_E_INVALID, "spano_dev_blend_begin: empty band [%d,%d) x %d", row0, row1, canvas_w);
    const int radius = launch_blend_setup(ctx, bands, sigma);
    if (radius < 0) return radius;
    spano_ctx::BlendSession &S = ctx->bs;
    S.cw = canvas_w;  S.mx = min_x;  S.my = min_y;  S.row0 = row0;  S.row1 = row1;  S.bands = bands;  S.radius = radius;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_ACC, (size_t)canvas_w * (row1 - row0) * sizeof(float4), (void **)&S.acc)) return rc;
    StageTimer t(ctx, 2);
    if (int rc = launch_blend_clear(ctx, S.acc, canvas_w, row1 - row0)) return rc;
    t.stop(0);
    S.staged.clear();
    S.prepared.clear();
    S.runs.clear();
    S.images = nullptr;  S.n_images = 0;  S.h_canvas = nullptr;  S.d_canvas = nullptr;
    S.open = true;
    return SPANO_OK;
}
} // namespace

extern "C" int spano_dev_blend_prepare(spano_ctx *ctx, int n, const spano_image_desc *images, const spano_slice *slices)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_prepare_impl(ctx, n, images, slices, false);
}

extern "C" int spano_blend_prepare(spano_ctx *ctx, int n, const spano_image_desc *images, const spano_slice *slices)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_prepare_impl(ctx, n, images, slices, true);
}

extern "C" int spano_dev_blend_add(spano_ctx *ctx, const spano_image_desc *im, const spano_slice *slice)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_add_impl(ctx, im, slice, false);
}

extern "C" int spano_blend_add(spano_ctx *ctx, const spano_image_desc *im, const spano_slice *slice)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_add_impl(ctx, im, slice, true);
}

extern "C" int spano_dev_blend_finish(spano_ctx *ctx, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_finish_impl(ctx, canvas, canvas_step, false);
}

extern "C" int spano_blend_finish(spano_ctx *ctx, uint8_t *canvas, size_t canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return blend_finish_impl(ctx, canvas, canvas_step, true);
}

// ---------------------------------------------------------------------------------------------
// one step of the tile-sharded path: owner side and band side of a rank, ordered by readiness flags
// ---------------------------------------------------------------------------------------------
namespace {

// cuStreamWaitValue32 through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef int (*stream_wait32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
stream_wait32_fn driver_stream_wait32()
{
    static stream_wait32_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<stream_wait32_fn>(p);
    }();
    return fn;
}

// the context's stream waits until *flag >= value (wrap-safe), without occupying an SM
int stream_wait_flag(spano_ctx *ctx, const uint32_t *flag, uint32_t value)
{
    stream_wait32_fn fn = ctx->opt_flag_wait ? nullptr : driver_stream_wait32();
    if (fn) {
        const int rc = fn(ctx->stream, (unsigned long long)(uintptr_t)flag, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (rc != 0) return spano_fail(ctx, SPANO_E_CUDA, "cuStreamWaitValue32 failed: CUresult %d", rc);
        return 0;
    }
    if (!ctx->opt_flag_wait) return spano_fail(ctx, SPANO_E_CUDA, "cuStreamWaitValue32 is not available from this driver");
    const int k = launch_flag_wait_kernel(ctx, flag, value);
    return k < 0 ? k : 0;
}

int check_shard_plan(spano_ctx *ctx, const spano_shard_plan *P, unsigned step)
{
    if (!P || P->world <= 0 || P->rank < 0 || P->rank >= P->world || P->n <= 0 || !P->images || !P->owner || !P->order || !P->slices || !P->flags)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step: null / inconsistent plan");
    if (step == 0) return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step: steps are numbered from 1");
    for (int k = 0; k < P->world; ++k)
        if (!P->flags[k]) return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step: flag block of rank %d is null", k);
    return 0;
}

} // namespace

extern "C" int spano_shard_step_owner(spano_ctx *ctx, const spano_shard_plan *P, unsigned step, int host)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_shard_plan(ctx, P, step)) return rc;
    const int n = P->n, W = P->world;
    bool any = false;
    for (int j = 0; j < n && !any; ++j) any = P->owner[j] == P->rank;
    if (!any) return SPANO_OK;
    // Size the scratch buffers for the largest owned image up front: growing one later would cudaFree the old buffer,
    // and cudaFree waits for the whole device -- including a band stream of this process that may already sit in a flag
    // wait which only this call's kernels can satisfy.
    {
        size_t src = 0, dark = 0, labels = 0, bits = 0, tables = 0;
        for (int j = 0; j < n; ++j) {
            if (P->owner[j] != P->rank) continue;
            const spano_image_desc &im = P->images[j];
            if (im.w <= 0 || im.h <= 0 || im.src_w <= 0 || im.src_h <= 0) continue;
            src = std::max(src, align_up((size_t)im.src_w * 3, 16) * im.src_h + 16);
            dark = std::max(dark, align_up((size_t)im.w, 16) * im.h);
            labels = std::max(labels, ((size_t)im.w * im.h + 1) * sizeof(uint32_t));
            bits = std::max(bits, (size_t)((im.w + 31) / 32) * im.h * sizeof(uint32_t));
            tables = std::max(tables, (6 * (size_t)((im.w + 3) & ~3) + 4 * (size_t)im.h) * sizeof(float));
        }
        void *dummy;
        if (host)
            if (int rc = spano_reserve(ctx, spano_ctx::BUF_SRC, src, &dummy)) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_DARK, dark, &dummy)) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_LABELS, labels, &dummy)) return rc;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_MASK0, bits, &dummy)) return rc;
        if (int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_TABLES, spano_ctx::BUF_TABLES_AUX), tables, &dummy)) return rc;
    }
    // every band has finished reading the arena this step overwrites (the previous step's, or with two alternating sets of
    // arenas the one before)
    const unsigned lag = P->done_lag >= 2 ? 2u : 1u;
    if (step > lag)
        for (int k = 0; k < W; ++k)
            if (int rc = stream_wait_flag(ctx, P->flags[P->rank] + n + k, step - lag)) return rc;
    std::vector<spano_slice> sl;
    std::vector<uint32_t *> targets;
    for (int t = 0; t < n; ++t) {
        const int j = P->order[t];
        if (j < 0 || j >= n) return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step_owner: order[%d] = %d", t, j);
        if (P->owner[j] != P->rank) continue;
        sl.clear();
        targets.clear();
        for (int k = 0; k < W; ++k) {
            const spano_slice &s = P->slices[(size_t)k * n + j];
            if (s.row1 <= s.row0) continue;
            sl.push_back(s);
            targets.push_back(P->flags[k] + j);
        }
        if (sl.empty()) continue;
        if (int rc = warp_scatter_impl(ctx, P->proj, P->scale, P->images + j, (int)sl.size(), sl.data(), host != 0)) return rc;
        const int k = launch_flag_signal(ctx, targets.data(), (int)targets.size(), step);
        if (k < 0) return k;
    }
    return SPANO_OK;
}

extern "C" int spano_shard_step_band(spano_ctx *ctx, const spano_shard_plan *P, unsigned step, int host, uint8_t *host_canvas,
                                     size_t host_canvas_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (int rc = check_shard_plan(ctx, P, step)) return rc;
    const int n = P->n, W = P->world;
    std::vector<uint32_t *> done_targets;
    for (int r = 0; r < W; ++r) done_targets.push_back(P->flags[r] + n + P->rank);
    if (P->row1 <= P->row0 || P->canvas_w <= 0) {   // an empty band still tells the owners that it is "done"
        const int k = launch_flag_signal(ctx, done_targets.data(), W, step);
        return k < 0 ? k : SPANO_OK;
    }
    if (host) {
        if (!host_canvas) return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step_band: host variant without a host canvas");
    } else if (!P->canvas) {
        return spano_fail(ctx, SPANO_E_INVALID, "spano_shard_step_band: null canvas");
    }
    const spano_slice *mine = P->slices + (size_t)P->rank * n;
    if (int rc = blend_begin_impl(ctx, P->canvas_w, P->min_x, P->min_y, P->row0, P->row1, P->bands, P->sigma)) return rc;
    if (host)
        if (int rc = blend_begin_host_extras(ctx, n, P->images, host_canvas, host_canvas_step)) return rc;
    if (int rc = blend_prepare_impl(ctx, n, P->images, mine, host != 0)) { ctx->bs.open = false; return rc; }
    for (int j = 0; j < n; ++j) {
        if (mine[j].row1 <= mine[j].row0) continue;
        if (int rc = stream_wait_flag(ctx, P->flags[P->rank] + j, step)) { ctx->bs.open = false; return rc; }
        if (int rc = blend_add_impl(ctx, P->images + j, mine + j, host != 0)) { ctx->bs.open = false; return rc; }
    }
    // the arena is not read after the last blend: tell the owners before the normalise / download tail
    {
        const int k = launch_flag_signal(ctx, done_targets.data(), W, step);
        if (k < 0) { ctx->bs.open = false; return k; }
    }
    if (host) return blend_finish_impl(ctx, host_canvas, host_canvas_step, true);
    return blend_finish_impl(ctx, P->canvas, P->canvas_step, false);
}

// ---------------------------------------------------------------------------------------------
// seam search by distance: cv::distanceTransform (5x5 chamfer) and dcut::dist_cut, host buffers
// ---------------------------------------------------------------------------------------------
namespace {

// uploads n masks into one arena, runs the distance transforms; returns device pointers (masks, dists) and pitches
int upload_and_transform(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *mask_steps, const int *w, const int *h,
                         std::vector<const uint8_t *> &d_masks, std::vector<size_t> &m_steps, std::vector<float *> &d_dist,
                         std::vector<size_t> &d_steps, size_t extra_bytes, uint8_t **extra)
{
    size_t total = 0;
    std::vector<size_t> off_m(n), off_d(n);
    d_masks.resize(n); m_steps.resize(n); d_dist.resize(n); d_steps.resize(n);
    for (int k = 0; k < n; ++k) {
        if (int rc = check_image_args(ctx, masks[k], w[k], h[k], mask_steps[k], 1, "mask")) return rc;
        m_steps[k] = align_up((size_t)w[k], 16);
        d_steps[k] = align_up((size_t)w[k], 4);
        off_m[k] = total;  total += align_up(m_steps[k] * h[k], 256);
        off_d[k] = total;  total += align_up(d_steps[k] * h[k] * sizeof(float), 256);
    }
    const size_t off_extra = total;
    total += extra_bytes;
    uint8_t *arena = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_ARENA, total, (void **)&arena)) return rc;
    for (int k = 0; k < n; ++k) {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_m[k], m_steps[k], masks[k], mask_steps[k], (size_t)w[k], h[k], cudaMemcpyHostToDevice, ctx->stream));
        d_masks[k] = arena + off_m[k];
        d_dist[k] = reinterpret_cast<float *>(arena + off_d[k]);
    }
    if (extra) *extra = arena + off_extra;
    int rc = launch_distance_transform(ctx, n, d_masks.data(), m_steps.data(), w, h, d_dist.data(), d_steps.data());
    return rc < 0 ? rc : 0;
}

} // namespace

extern "C" int spano_distance_transform(spano_ctx *ctx, const uint8_t *mask, int w, int h, size_t step, float *dist, size_t dist_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (!dist || dist_step < (size_t)std::max(w, 0) * sizeof(float)) return spano_fail(ctx, SPANO_E_INVALID, "spano_distance_transform: null output or step too small");
    std::vector<const uint8_t *> dm; std::vector<size_t> ms, ds; std::vector<float *> dd;
    if (int rc = upload_and_transform(ctx, 1, &mask, &step, &w, &h, dm, ms, dd, ds, 0, nullptr)) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(dist, dist_step, dd[0], ds[0] * sizeof(float), (size_t)w * sizeof(float), h, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

extern "C" int spano_dist_cut(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *mask_steps, const int *tl_x,
                              const int *tl_y, const int *w, const int *h, uint8_t *const *cut, const size_t *cut_steps)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !masks || !mask_steps || !tl_x || !tl_y || !w || !h || !cut || !cut_steps)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_dist_cut: null/empty argument");
    size_t cut_bytes = 0;
    std::vector<size_t> off_c(n), c_steps(n);
    for (int k = 0; k < n; ++k) {
        if (w[k] <= 0 || h[k] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "spano_dist_cut: empty mask %d", k);
        if (!cut[k] || cut_steps[k] < (size_t)w[k]) return spano_fail(ctx, SPANO_E_INVALID, "spano_dist_cut: output %d null or step too small", k);
        c_steps[k] = align_up((size_t)w[k], 16);
        off_c[k] = cut_bytes;
        cut_bytes += align_up(c_steps[k] * h[k], 256);
    }
    std::vector<const uint8_t *> dm; std::vector<size_t> ms, ds; std::vector<float *> dd;
    uint8_t *d_cut_arena = nullptr;
    if (int rc = upload_and_transform(ctx, n, masks, mask_steps, w, h, dm, ms, dd, ds, cut_bytes, &d_cut_arena)) return rc;
    std::vector<uint8_t *> d_cut(n);
    for (int k = 0; k < n; ++k) d_cut[k] = d_cut_arena + off_c[k];
    std::vector<const float *> cd(dd.begin(), dd.end());
    int rc = launch_dist_cut(ctx, n, dm.data(), ms.data(), cd.data(), ds.data(), tl_x, tl_y, w, h, d_cut.data(), c_steps.data());
    if (rc < 0) return rc;
    for (int k = 0; k < n; ++k)
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(cut[k], cut_steps[k], d_cut[k], c_steps[k], (size_t)w[k], h[k], cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

// ---------------------------------------------------------------------------------------------
// test::equalizeIntensities, host buffers
// ---------------------------------------------------------------------------------------------
extern "C" int spano_equalize_intensities_size(int w, int h, float ratio, int *field_w, int *field_h)
{
    if (w <= 0 || h <= 0 || !(ratio > 0.f) || !field_w || !field_h) return SPANO_E_INVALID;
    spano_equalize_field_size(w, h, ratio, field_w, field_h);
    return (*field_w > 0 && *field_h > 0) ? SPANO_OK : SPANO_E_INVALID;
}

extern "C" int spano_equalize_intensities(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps,
                                          const uint8_t *const *masks, const size_t *mask_steps, const int *tl_x, const int *tl_y,
                                          const int *w, const int *h, float ratio, float *const *fields, const size_t *field_steps)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !tiles || !tile_steps || !masks || !mask_steps || !tl_x || !tl_y || !w || !h || !fields || !field_steps)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_equalize_intensities: null/empty argument");
    if (!(ratio > 0.f) || ratio > 1.f) return spano_fail(ctx, SPANO_E_INVALID, "spano_equalize_intensities: ratio must be in (0, 1]");
    size_t total = 0;
    std::vector<size_t> off_t(n), off_m(n), off_f(n), ts(n), ms(n), fp(n);
    std::vector<int> fw(n), fh(n);
    for (int k = 0; k < n; ++k) {
        if (int rc = check_image_args(ctx, tiles[k], w[k], h[k], tile_steps[k], 3, "tile")) return rc;
        if (int rc = check_image_args(ctx, masks[k], w[k], h[k], mask_steps[k], 1, "mask")) return rc;
        spano_equalize_field_size(w[k], h[k], ratio, &fw[k], &fh[k]);
        if (fw[k] <= 0 || fh[k] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "image %d is too small for ratio %g", k, (double)ratio);
        if (!fields[k] || field_steps[k] < (size_t)fw[k] * sizeof(float)) return spano_fail(ctx, SPANO_E_INVALID, "field %d: null pointer or step too small", k);
        ts[k] = align_up((size_t)w[k] * 3, 16);
        ms[k] = align_up((size_t)w[k], 16);
        fp[k] = align_up((size_t)fw[k], 4);
        off_t[k] = total;  total += align_up(ts[k] * h[k], 256);
        off_m[k] = total;  total += align_up(ms[k] * h[k], 256);
        off_f[k] = total;  total += align_up(fp[k] * fh[k] * sizeof(float), 256);
    }
    uint8_t *arena = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_TILE, total, (void **)&arena)) return rc;
    std::vector<const uint8_t *> dt(n), dm(n);
    std::vector<float *> df(n);
    for (int k = 0; k < n; ++k) {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_t[k], ts[k], tiles[k], tile_steps[k], (size_t)w[k] * 3, h[k], cudaMemcpyHostToDevice, ctx->stream));
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_m[k], ms[k], masks[k], mask_steps[k], (size_t)w[k], h[k], cudaMemcpyHostToDevice, ctx->stream));
        dt[k] = arena + off_t[k];  dm[k] = arena + off_m[k];  df[k] = reinterpret_cast<float *>(arena + off_f[k]);
    }
    int rc = launch_equalize_intensities(ctx, n, dt.data(), ts.data(), dm.data(), ms.data(), tl_x, tl_y, w, h, ratio, df.data(), fp.data());
    if (rc < 0) return rc;
    for (int k = 0; k < n; ++k)
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(fields[k], field_steps[k], df[k], fp[k] * sizeof(float), (size_t)fw[k] * sizeof(float), fh[k],
                                          cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

// ---------------------------------------------------------------------------------------------
// blnd::simple_blend / blnd::no_blend, host buffers
// ---------------------------------------------------------------------------------------------
namespace {

int simple_or_no_blend(spano_ctx *ctx, bool simple, int n, const uint8_t *const *tiles, const size_t *tile_steps, const uint8_t *const *masks,
                       const size_t *mask_steps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *out, size_t out_step)
{
    if (n <= 0 || !tiles || !tile_steps || !masks || !mask_steps || !tl_x || !tl_y || !w || !h || !out)
        return spano_fail(ctx, SPANO_E_INVALID, "Input consistency!");
    int cw, chh, mx, my;
    spano_pan_dimension(n, tl_x, tl_y, w, h, &cw, &chh, &mx, &my);
    if (out_step < (size_t)cw * 3) return spano_fail(ctx, SPANO_E_INVALID, "out_step too small");
    // tiles + canvas (+ accumulator) in one arena after the masks / distance maps
    size_t extra = 0;
    std::vector<size_t> off_t(n), t_steps(n);
    for (int k = 0; k < n; ++k) {
        if (int rc = check_image_args(ctx, tiles[k], w[k], h[k], tile_steps[k], 3, "tile")) return rc;
        t_steps[k] = align_up((size_t)w[k] * 3, 16);
        off_t[k] = extra;
        extra += align_up(t_steps[k] * h[k], 256);
    }
    const size_t o_step = align_up((size_t)cw * 3, 16);
    const size_t off_out = extra;
    extra += align_up(o_step * chh, 256);
    const size_t off_acc = extra;
    if (simple) extra += (size_t)cw * chh * sizeof(float4);
    std::vector<const uint8_t *> dm; std::vector<size_t> ms, ds; std::vector<float *> dd;
    uint8_t *arena = nullptr;
    if (simple) {
        if (int rc = upload_and_transform(ctx, n, masks, mask_steps, w, h, dm, ms, dd, ds, extra, &arena)) return rc;
    } else {
        // masks only (no distance transform)
        size_t total = 0;
        std::vector<size_t> off_m(n);
        dm.resize(n); ms.resize(n);
        for (int k = 0; k < n; ++k) {
            if (int rc = check_image_args(ctx, masks[k], w[k], h[k], mask_steps[k], 1, "mask")) return rc;
            ms[k] = align_up((size_t)w[k], 16);
            off_m[k] = total;
            total += align_up(ms[k] * h[k], 256);
        }
        uint8_t *base = nullptr;
        if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_ARENA, total + extra, (void **)&base)) return rc;
        for (int k = 0; k < n; ++k) {
            SPANO_CUDA(ctx, cudaMemcpy2DAsync(base + off_m[k], ms[k], masks[k], mask_steps[k], (size_t)w[k], h[k], cudaMemcpyHostToDevice, ctx->stream));
            dm[k] = base + off_m[k];
        }
        arena = base + total;
    }
    std::vector<const uint8_t *> d_tiles(n);
    std::vector<int> ax(n), ay(n);
    for (int k = 0; k < n; ++k) {
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(arena + off_t[k], t_steps[k], tiles[k], tile_steps[k], (size_t)w[k] * 3, h[k], cudaMemcpyHostToDevice, ctx->stream));
        d_tiles[k] = arena + off_t[k];
        ax[k] = tl_x[k] - mx;
        ay[k] = tl_y[k] - my;
    }
    uint8_t *d_out = arena + off_out;
    int rc;
    if (simple) {
        std::vector<const float *> cd(dd.begin(), dd.end());
        rc = launch_simple_blend(ctx, n, d_tiles.data(), t_steps.data(), cd.data(), ds.data(), ax.data(), ay.data(), w, h,
                                 reinterpret_cast<float4 *>(arena + off_acc), cw, chh, d_out, o_step);
    } else {
        rc = launch_no_blend(ctx, n, d_tiles.data(), t_steps.data(), dm.data(), ms.data(), ax.data(), ay.data(), w, h, cw, chh, d_out, o_step);
    }
    if (rc < 0) return rc;
    SPANO_CUDA(ctx, cudaMemcpy2DAsync(out, out_step, d_out, o_step, (size_t)cw * 3, chh, cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPANO_OK;
}

} // namespace

extern "C" int spano_simple_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const uint8_t *const *masks,
                                  const size_t *mask_steps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *out,
                                  size_t out_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return simple_or_no_blend(ctx, true, n, tiles, tile_steps, masks, mask_steps, tl_x, tl_y, w, h, out, out_step);
}

extern "C" int spano_no_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const uint8_t *const *masks,
                              const size_t *mask_steps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *out,
                              size_t out_step)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    return simple_or_no_blend(ctx, false, n, tiles, tile_steps, masks, mask_steps, tl_x, tl_y, w, h, out, out_step);
}

// ---------------------------------------------------------------------------------------------
// gain::get_overlapp_intensity, host buffers
// ---------------------------------------------------------------------------------------------
extern "C" int spano_overlap_intensity(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tile_steps, const int *tl_x,
                                       const int *tl_y, const int *w, const int *h, const double *adj, spano_overlap_info *out, int *n_out)
{
    if (!ctx) return SPANO_E_INVALID;
    Guard g(ctx);
    if (n <= 0 || !tiles || !tile_steps || !tl_x || !tl_y || !w || !h || !adj || !out || !n_out)
        return spano_fail(ctx, SPANO_E_INVALID, "spano_overlap_intensity: null/empty argument");
    size_t total = 0;
    std::vector<size_t> off_t(n), off_g(n), off_m(n), off_d(n), ts(n), gs(n);
    for (int k = 0; k < n; ++k) {
        if (int rc = check_image_args(ctx, tiles[k], w[k], h[k], tile_steps[k], 3, "tile")) return rc;
        ts[k] = align_up((size_t)w[k] * 3, 16);
        gs[k] = align_up((size_t)w[k], 16);
        off_t[k] = total;  total += align_up(ts[k] * h[k], 256);
        off_g[k] = total;  total += align_up(gs[k] * h[k], 256);
        off_m[k] = total;  total += align_up(gs[k] * h[k], 256);
        off_d[k] = total;  total += align_up(gs[k] * h[k], 256);
    }
    const int max_pairs = n * (n + 1) / 2;
    const size_t off_acc = total;
    total += (size_t)max_pairs * 3 * sizeof(unsigned long long);
    uint8_t *arena = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_ARENA, total, (void **)&arena)) return rc;
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(arena + off_acc);
    SPANO_CUDA(ctx, cudaMemsetAsync(acc, 0, (size_t)max_pairs * 3 * sizeof(unsigned long long), ctx->stream));
    for (int k = 0; k < n; ++k) {
        uint8_t *t = arena + off_t[k];
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(t, ts[k], tiles[k], tile_steps[k], (size_t)w[k] * 3, h[k], cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = launch_gray(ctx, t, ts[k], w[k], h[k], arena + off_g[k], gs[k]); rc < 0) return rc;
        if (int rc = launch_dark_flags(ctx, t, w[k], h[k], ts[k], arena + off_d[k], gs[k]); rc < 0) return rc;
        if (int rc = launch_valid_mask(ctx, arena + off_d[k], w[k], h[k], gs[k], 0, arena + off_m[k], gs[k]); rc < 0) return rc;
    }
    int count = 0;
    std::vector<int> slot;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            if (!(adj[(size_t)i * n + j] + (i == j ? 1.0 : 0.0) > 0)) continue;
            out[count].i = i;  out[count].j = j;  out[count].area = out[count].I_i = out[count].I_j = 0.0;
            const int x0 = std::max(tl_x[i], tl_x[j]), y0 = std::max(tl_y[i], tl_y[j]);
            const int x1 = std::min(tl_x[i] + w[i], tl_x[j] + w[j]), y1 = std::min(tl_y[i] + h[i], tl_y[j] + h[j]);
            if (x1 > x0 && y1 > y0)
                if (int rc = launch_overlap_sums(ctx, arena + off_g[i], gs[i], arena + off_m[i], gs[i], arena + off_g[j], gs[j], arena + off_m[j],
                                                 gs[j], x0 - tl_x[i], y0 - tl_y[i], x0 - tl_x[j], y0 - tl_y[j], x1 - x0, y1 - y0, acc + 3 * (size_t)count);
                    rc < 0)
                    return rc;
            ++count;
        }
    std::vector<unsigned long long> host((size_t)std::max(1, count) * 3);
    SPANO_CUDA(ctx, cudaMemcpyAsync(host.data(), acc, (size_t)count * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < count; ++k) {
        out[k].area = (double)host[3 * (size_t)k];
        out[k].I_i = (double)host[3 * (size_t)k + 1];
        out[k].I_j = (double)host[3 * (size_t)k + 2];
    }
    *n_out = count;
    return SPANO_OK;
}
