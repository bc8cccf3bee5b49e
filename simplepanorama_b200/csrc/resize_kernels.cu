// resize_kernels.cu -- up-scaling of the preview-scale seam masks to tile size (sm_100a).
//
// Replaces  cv::resize(mask_cut[i], blend_dat.msks_cut[i], blend_dat.imgs[i].size(), cv::INTER_CUBIC)
// in stitch_parameters::return_full (reference src/classes/_panorama.cpp:329-335).  The fourth argument
// is cv::resize's `fx` slot, so the interpolation is the DEFAULT, INTER_LINEAR.
// For CV_8UC1 OpenCV's linear resize is fixed point: per axis two 11-bit coefficients
// (INTER_RESIZE_COEF_SCALE = 2048) from a float fraction, horizontal pass in int, vertical pass
//   dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2 .
// The source coordinate is (d + 0.5) * scale - 0.5 in double, cast to float, floor'ed, clamped.
// Bit-exact against cv2 4.13 (tests/golden/kernels.npz and tests/test_gpu_resize.py).
// Integer work; 1 B written per tile pixel, the small source stays in L2.
#include "spano_internal.h"

namespace {

struct AxisEntry {
    int ofs;      // first source index (second = min(ofs + 1, len - 1)), clamped rows for y
    short c0, c1; // 11-bit coefficients
};

// one entry per destination column (axis 0) and row (axis 1)
__global__ void resize_tables_kernel(int sw, int sh, int dw, int dh, AxisEntry *xt, AxisEntry *yt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dw) {
        const double scale = (double)sw / dw;
        float f = (float)((i + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= sw - 1) { f = 0.f; s = sw - 1; }
        AxisEntry e;
        e.ofs = s;
        e.c0 = (short)__float2int_rn((1.f - f) * 2048.f);
        e.c1 = (short)__float2int_rn(f * 2048.f);
        xt[i] = e;
    }
    if (i < dh) {
        const double scale = (double)sh / dh;
        float f = (float)((i + 0.5) * scale - 0.5);
        const int s = (int)floorf(f);
        f -= (float)s;
        AxisEntry e;
        e.ofs = s;   // rows s and s+1 are clamped to [0, sh-1] when they are read (OpenCV keeps the fraction)
        e.c0 = (short)__float2int_rn((1.f - f) * 2048.f);
        e.c1 = (short)__float2int_rn(f * 2048.f);
        yt[i] = e;
    }
}

// First / last result row in [ra, rb) that has a non-zero source tap, per strip of strip_w result columns: block =
// (strip, 1024 rows), thread = 4 rows.  The source taps of a strip are the columns xt[first].ofs .. xt[last].ofs + 1 of the
// two source rows of the result row.  atomicMin / atomicMax into ymin / ymax (one per warp that found something).
__global__ void resize_activity_kernel(const uint8_t *src, int sw, int sh, size_t sstep, const AxisEntry *xt, const AxisEntry *yt, int dw,
                                       int strip_w, int ra, int rb, int *ymin, int *ymax)
{
    const int s = blockIdx.x;
    const int c0 = s * strip_w, c1 = min(dw, c0 + strip_w) - 1;
    const int sx0 = xt[c0].ofs, sx1 = min(xt[c1].ofs + 1, sw - 1);
    int lo = 0x7fffffff, hi = -1;
    for (int k = 0; k < 4; ++k) {
        const int dy = ra + blockIdx.y * 1024 + k * 256 + threadIdx.x;
        if (dy >= rb) break;
        const int o = yt[dy].ofs;
        const int y0 = min(max(o, 0), sh - 1), y1 = min(max(o + 1, 0), sh - 1);
        const uint8_t *r0 = src + (size_t)y0 * sstep, *r1 = src + (size_t)y1 * sstep;
        uint32_t any = 0;
        for (int x = sx0; x <= sx1; ++x) any |= (uint32_t)__ldg(r0 + x) | (uint32_t)__ldg(r1 + x);
        if (any) { lo = min(lo, dy); hi = max(hi, dy); }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(ymin + s, lo); atomicMax(ymax + s, hi); }
}

// One thread = 4 consecutive destination pixels (one 32-bit store) of RESIZE_ROWS consecutive rows: the x-table
// entries are loaded once and stay in registers, the rows are independent of each other (their loads overlap), and
// the y-table entry of a row is the same for the whole block.  Seam masks are mostly 0 (SURVEY.md appendix A:
// dist_cut / graph_cut assign every canvas pixel to one image) and a destination pixel whose four source taps are 0
// is 0 whatever the coefficients, so such groups skip the arithmetic.
constexpr int RESIZE_ROWS = 16;

// At most 56 registers (x 256 threads = 14 K): this is the first kernel of the auxiliary-stream work of an image in the fused
// path, and an SM that runs a blend CTA of the previous image has 16 K registers left.  At 80 registers (what ptxas takes with
// the row loop unrolled) the CTA does not fit, and the whole warp / mask chain behind it waits until the blend has ended.
// need0 / need1 (optional): per strip of strip_w result columns the result rows [need0[s], need1[s]) that anybody will read;
// the rest is not produced (a thread's 4 columns lie in one strip: strip_w is a multiple of 4).
__global__ void __maxnreg__(56) resize_linear_u8_kernel(const uint8_t *src, int sw, int sh, size_t sstep, const AxisEntry *xt,
                                                               const AxisEntry *yt, uint8_t *dst, int dw, int dh, size_t dstep,
                                                               int row_begin, int row_end, int strip_w, const int *need0, const int *need1)
{
    const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (dx0 >= dw) return;
    // this thread's rows: the block's RESIZE_ROWS rows of the launch's range, cut down to what the strip needs
    int y_first = row_begin + blockIdx.y * RESIZE_ROWS, y_end = min(y_first + RESIZE_ROWS, row_end);
    if (need0) {
        const int s = dx0 / strip_w;
        y_first = max(y_first, __ldg(need0 + s));
        y_end = min(y_end, __ldg(need1 + s));
    }
    if (y_end <= y_first) return;
    AxisEntry ex[4];
    int x1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ex[i] = xt[min(dx0 + i, dw - 1)];
        x1[i] = min(ex[i].ofs + 1, sw - 1);
    }
    const bool vec = dx0 + 4 <= dw && ((((uintptr_t)dst) | dstep) & 3) == 0;
    {
        // whole-thread early out: the source taps of all RESIZE_ROWS x 4 pixels lie in a small rectangle (both tables
        // are monotonic); if every byte of it is 0 -- most of a seam mask -- the pixels are 0
        const int y_last = y_end - 1;
        const int sy0 = min(max(yt[y_first].ofs, 0), sh - 1), sy1 = min(max(yt[y_last].ofs + 1, 0), sh - 1);
        const int sx0 = ex[0].ofs, sx1 = x1[3];
        uint32_t any = 1;
        if ((sy1 - sy0 + 1) * (sx1 - sx0 + 1) <= 64) {   // (an up-scaling: a handful of bytes; skip the test when shrinking)
            any = 0;
            for (int y = sy0; y <= sy1; ++y) {
                const uint8_t *rp = src + (size_t)y * sstep;
                for (int x = sx0; x <= sx1; ++x) any |= __ldg(rp + x);
            }
        }
        if (!any) {
            for (int dy = y_first; dy <= y_last; ++dy) {
                uint8_t *d = dst + (size_t)dy * dstep + dx0;
                if (vec) *reinterpret_cast<uint32_t *>(d) = 0u;
                else
                    for (int i = 0; i < 4 && dx0 + i < dw; ++i) d[i] = 0;
            }
            return;
        }
    }
#pragma unroll 1
    for (int dy = y_first; dy < y_end; ++dy) {
        const AxisEntry ey = yt[dy];
        const int y0 = min(max(ey.ofs, 0), sh - 1), y1 = min(max(ey.ofs + 1, 0), sh - 1);
        const uint8_t *r0p = src + (size_t)y0 * sstep, *r1p = src + (size_t)y1 * sstep;
        int a0[4], a1[4], b0[4], b1[4];
        int any = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a0[i] = __ldg(r0p + ex[i].ofs);  a1[i] = __ldg(r0p + x1[i]);
            b0[i] = __ldg(r1p + ex[i].ofs);  b1[i] = __ldg(r1p + x1[i]);
            any |= a0[i] | a1[i] | b0[i] | b1[i];
        }
        uint32_t packed = 0;
        if (any) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h0 = a0[i] * ex[i].c0 + a1[i] * ex[i].c1;
                const int h1 = b0[i] * ex[i].c0 + b1[i] * ex[i].c1;
                const int v = ((((int)ey.c0 * (h0 >> 4)) >> 16) + (((int)ey.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
                packed |= (uint32_t)min(255, max(0, v)) << (8 * i);
            }
        }
        uint8_t *d = dst + (size_t)dy * dstep + dx0;
        if (vec) *reinterpret_cast<uint32_t *>(d) = packed;
        else
            for (int i = 0; i < 4 && dx0 + i < dw; ++i) d[i] = (uint8_t)(packed >> (8 * i));
    }
}

// ---- test::adjust_intensity (reference src/test/_test.cpp:110-122): float-bilinear up-scaling of the
// per-image intensity-correction field (cv::resize CV_32FC1 INTER_LINEAR) fused with
//   tile <- sat_u8(rint(((float(v) * float(1/255)) * (1.f / clamp(field))) * 255.f))
// One pass over the tile, in place; the field itself is never materialised at tile size.
struct AxisEntryF {
    int ofs;
    float f;
};

__global__ void intensity_tables_kernel(int sw, int sh, int dw, int dh, AxisEntryF *xt, AxisEntryF *yt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dw) {
        const double scale = (double)sw / dw;
        float f = (float)((i + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= sw - 1) { f = 0.f; s = sw - 1; }
        xt[i] = AxisEntryF{s, f};
    }
    if (i < dh) {
        const double scale = (double)sh / dh;
        float f = (float)((i + 0.5) * scale - 0.5);
        const int s = (int)floorf(f);
        f -= (float)s;
        yt[i] = AxisEntryF{s, f};
    }
}

__global__ void adjust_intensity_kernel(uint8_t *bgr, int w, int h, size_t step, const float *field, int fw, int fh,
                                        size_t fpitch, const AxisEntryF *xt, const AxisEntryF *yt)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const AxisEntryF ex = xt[x], ey = yt[y];
    const int y0 = min(max(ey.ofs, 0), fh - 1), y1 = min(max(ey.ofs + 1, 0), fh - 1);
    const int x0 = ex.ofs, x1 = min(ex.ofs + 1, fw - 1);
    const float *p0 = field + (size_t)y0 * fpitch, *p1 = field + (size_t)y1 * fpitch;
    const float ax1 = ex.f, ax0 = __fsub_rn(1.f, ex.f), ay1 = ey.f, ay0 = __fsub_rn(1.f, ey.f);
    const float r0 = __fadd_rn(__fmul_rn(__ldg(p0 + x0), ax0), __fmul_rn(__ldg(p0 + x1), ax1));
    const float r1 = __fadd_rn(__fmul_rn(__ldg(p1 + x0), ax0), __fmul_rn(__ldg(p1 + x1), ax1));
    float d = __fadd_rn(__fmul_rn(r0, ay0), __fmul_rn(r1, ay1));
    d = copysignf(fmaxf(fabsf(d), 1e-6f), d);
    const float s = __fdiv_rn(1.f, d);
    uint8_t *p = bgr + (size_t)y * step + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = __fmul_rn(__fmul_rn((float)p[c], (float)(1.0 / 255.0)), s);
        p[c] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(v, 255.f))));
    }
}

} // namespace

int launch_adjust_intensity(spano_ctx *ctx, uint8_t *bgr, int w, int h, size_t step, const float *field, int fw, int fh,
                            size_t fpitch_elems)
{
    if (w <= 0 || h <= 0 || fw <= 0 || fh <= 0) return spano_fail(ctx, SPANO_E_INVALID, "adjust_intensity: empty image");
    AxisEntryF *tab = nullptr;
    int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_RESIZE, spano_ctx::BUF_RESIZE_AUX), (size_t)(w + h) * sizeof(AxisEntryF), (void **)&tab);
    if (rc) return rc;
    const int n = w > h ? w : h;
    intensity_tables_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(fw, fh, w, h, tab, tab + w);
    dim3 block(256), grid((w + 255) / 256, h);
    adjust_intensity_kernel<<<grid, block, 0, ctx->stream>>>(bgr, w, h, step, field, fw, fh, fpitch_elems, tab, tab + w);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return 2;
}

int launch_resize_mask(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh,
                       size_t dstep, int row_begin, int row_end)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return spano_fail(ctx, SPANO_E_INVALID, "resize: empty image");
    if (row_end < 0 || row_end > dh) row_end = dh;
    if (row_begin < 0) row_begin = 0;
    if (row_end <= row_begin) return 0;
    AxisEntry *tab = nullptr;
    int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_RESIZE, spano_ctx::BUF_RESIZE_AUX), (size_t)(dw + dh) * sizeof(AxisEntry), (void **)&tab);
    if (rc) return rc;
    const int n = dw > dh ? dw : dh;
    resize_tables_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(sw, sh, dw, dh, tab, tab + dw);
    dim3 block(256), grid((dw + 1023) / 1024, (row_end - row_begin + RESIZE_ROWS - 1) / RESIZE_ROWS);
    resize_linear_u8_kernel<<<grid, block, 0, ctx->stream>>>(src, sw, sh, sstep, tab, tab + dw, dst, dw, dh, dstep, row_begin, row_end, 4, nullptr, nullptr);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return 2;
}

int launch_resize_activity(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, int dw, int dh, int strip_w, int ra, int rb,
                           int *ymin, int *ymax)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || strip_w <= 0) return spano_fail(ctx, SPANO_E_INVALID, "resize: empty image");
    ra = std::max(ra, 0);
    rb = std::min(rb, dh);
    AxisEntry *tab = nullptr;
    int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_RESIZE, spano_ctx::BUF_RESIZE_AUX), (size_t)(dw + dh) * sizeof(AxisEntry), (void **)&tab);
    if (rc) return rc;
    const int n = dw > dh ? dw : dh;
    resize_tables_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(sw, sh, dw, dh, tab, tab + dw);
    if (rb > ra) {
        const int strips = (dw + strip_w - 1) / strip_w;
        dim3 grid(strips, (rb - ra + 1023) / 1024);
        resize_activity_kernel<<<grid, 256, 0, ctx->stream>>>(src, sw, sh, sstep, tab, tab + dw, dw, strip_w, ra, rb, ymin, ymax);
    }
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return 2;
}

// (the tables of launch_resize_activity for the same geometry are still in the table buffer of this stream: the two calls are
// made back to back by make_plan; they are rebuilt here all the same -- 2.5 us -- so that the call stands on its own)
int launch_resize_mask_rows(spano_ctx *ctx, const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep,
                            int strip_w, const int *need0, const int *need1)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return spano_fail(ctx, SPANO_E_INVALID, "resize: empty image");
    if (strip_w <= 0 || (strip_w & 3) || !need0 || !need1) return spano_fail(ctx, SPANO_E_INVALID, "resize: bad strip description");
    AxisEntry *tab = nullptr;
    int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_RESIZE, spano_ctx::BUF_RESIZE_AUX), (size_t)(dw + dh) * sizeof(AxisEntry), (void **)&tab);
    if (rc) return rc;
    const int n = dw > dh ? dw : dh;
    resize_tables_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(sw, sh, dw, dh, tab, tab + dw);
    dim3 block(256), grid((dw + 1023) / 1024, (dh + RESIZE_ROWS - 1) / RESIZE_ROWS);
    resize_linear_u8_kernel<<<grid, block, 0, ctx->stream>>>(src, sw, sh, sstep, tab, tab + dw, dst, dw, dh, dstep, 0, dh, strip_w, need0, need1);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return 2;
}
