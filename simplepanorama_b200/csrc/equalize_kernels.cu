// equalize_kernels.cu -- test::equalizeIntensities (reference src/test/_test.cpp:9-106), the solve behind the
// intensity-correction fields that test::adjust_intensity (resize_kernels.cu) divides the tiles by (sm_100a).
//
// Runs once per panorama at PREVIEW scale inside stitch_parameters::set_config (src/classes/_panorama.cpp:131-133); per
// image i (tile T_i, validity mask M_i, both preview size):
//   D_i   = cv::distanceTransform(M_i, DIST_L2, 5) / 255                   (dcut::distance_transform; dist_kernels.cu)
//   m_i   = resize(M_i, ratio), t_i = resize(T_i, ratio), d_i = resize(D_i, ratio)          cv::resize(.., INTER_LINEAR)
//   g_i   = float(gray(t_i)) / 255 where m_i != 0, else 0;   q_i = g_i * d_i
//   over the overlaps with every other image j (ascending j), where m_i != 0:  Q_i = q_i + sum q_j,  A_i = d_i + sum d_j
//   F_i   = GaussianBlur_13x13,sigma=7,REFLECT( g_i / (Q_i / (A_i + 1e-5) + 1e-5) + (255 - m_i) / 255 )
// cv::resize(src, dst, Size(), ratio, ratio, INTER_LINEAR): the scale is 1 / ratio exactly (not src / dst), and for ratio = 0.5
// -- the reference's only value -- OpenCV replaces INTER_LINEAR by the 2x2 AREA average ("INTER_LINEAR && is_area_fast &&
// iscale == 2 -> INTER_AREA"), for odd sizes too: a cell that sticks out of the source averages the pixels it has
// ((float)sum / count, rounded half to even for 8-bit).  Other ratios take OpenCV's fixed-point (8-bit) / float linear scheme.
// cv::divide gives 0 where the divisor is 0.
// Preview-scale data (a few hundred kB per image): nothing here is performance critical; exactness is (integer stages
// bit-exact, the float field within 1e-5 relative of OpenCV's).
#include <cmath>
#include <vector>

#include "spano_internal.h"

namespace {

struct Axis {
    int ofs;
    float f;      // fraction (float path)
    short c0, c1; // 11-bit coefficients (8-bit path)
};

// one entry per destination index: source offset + coefficients, as cv::resize computes them for INTER_LINEAR
__global__ void axis_kernel(int slen, int dlen, double scale, Axis *t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dlen) return;
    float f = (float)((i + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= slen - 1) { f = 0.f; s = slen - 1; }
    Axis e;
    e.ofs = s;
    e.f = f;
    e.c0 = (short)__float2int_rn((1.f - f) * 2048.f);
    e.c1 = (short)__float2int_rn(f * 2048.f);
    t[i] = e;
}

template <int CN>
__global__ void resize_u8_kernel(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep,
                                 const Axis *xt, const Axis *yt, int area2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    uint8_t *d = dst + (size_t)y * dstep + (size_t)x * CN;
    if (area2) {
        const int nx = min(2, sw - 2 * x), ny = min(2, sh - 2 * y);      // pixels of the 2x2 cell inside the source
        if (nx <= 0 || ny <= 0) {
#pragma unroll
            for (int c = 0; c < CN; ++c) d[c] = 0;
            return;
        }
        const uint8_t *r0 = src + (size_t)(2 * y) * sstep + (size_t)(2 * x) * CN;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            int sum = 0;
            for (int yy = 0; yy < ny; ++yy)
                for (int xx = 0; xx < nx; ++xx) sum += r0[(size_t)yy * sstep + xx * CN + c];
            d[c] = (nx * ny == 4) ? (uint8_t)((sum + 2) >> 2) : (uint8_t)min(255, max(0, __float2int_rn(__fdiv_rn((float)sum, (float)(nx * ny)))));
        }
        return;
    }
    const Axis ex = xt[x], ey = yt[y];
    const int x1 = min(ex.ofs + 1, sw - 1), y1 = min(ey.ofs + 1, sh - 1);
    const uint8_t *r0 = src + (size_t)ey.ofs * sstep, *r1 = src + (size_t)y1 * sstep;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        const int h0 = r0[ex.ofs * CN + c] * ex.c0 + r0[x1 * CN + c] * ex.c1;
        const int h1 = r1[ex.ofs * CN + c] * ex.c0 + r1[x1 * CN + c] * ex.c1;
        const int v = ((((int)ey.c0 * (h0 >> 4)) >> 16) + (((int)ey.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        d[c] = (uint8_t)min(255, max(0, v));
    }
}

// D/255 resized (the distance map is scaled before the resize, as in the reference)
__global__ void resize_f32_kernel(const float *src, int sw, int sh, size_t spitch, float *dst, int dw, int dh, size_t dpitch,
                                  const Axis *xt, const Axis *yt, int area2, float pre_scale)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    if (area2) {
        const int nx = min(2, sw - 2 * x), ny = min(2, sh - 2 * y);
        float s = 0.f;
        if (nx > 0 && ny > 0) {
            const float *r0 = src + (size_t)(2 * y) * spitch + 2 * x;
            for (int yy = 0; yy < ny; ++yy)
                for (int xx = 0; xx < nx; ++xx) s = __fadd_rn(s, __fmul_rn(r0[(size_t)yy * spitch + xx], pre_scale));
            s = (nx * ny == 4) ? __fmul_rn(s, 0.25f) : __fdiv_rn(s, (float)(nx * ny));
        }
        dst[(size_t)y * dpitch + x] = s;
        return;
    }
    const Axis ex = xt[x], ey = yt[y];
    const int x1 = min(ex.ofs + 1, sw - 1), y1 = min(ey.ofs + 1, sh - 1);
    const float *r0 = src + (size_t)ey.ofs * spitch, *r1 = src + (size_t)y1 * spitch;
    const float ax1 = ex.f, ax0 = __fsub_rn(1.f, ex.f), ay1 = ey.f, ay0 = __fsub_rn(1.f, ey.f);
    const float h0 = __fadd_rn(__fmul_rn(__fmul_rn(r0[ex.ofs], pre_scale), ax0), __fmul_rn(__fmul_rn(r0[x1], pre_scale), ax1));
    const float h1 = __fadd_rn(__fmul_rn(__fmul_rn(r1[ex.ofs], pre_scale), ax0), __fmul_rn(__fmul_rn(r1[x1], pre_scale), ax1));
    dst[(size_t)y * dpitch + x] = __fadd_rn(__fmul_rn(h0, ay0), __fmul_rn(h1, ay1));
}

// g = float(gray(t)) / 255 where m != 0 else 0;  q = g * d
__global__ void intensity_kernel(const uint8_t *t, size_t tstep, const uint8_t *m, size_t mstep, const float *d, float *g, float *q, int w,
                                 int h, size_t pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = t + (size_t)y * tstep + (size_t)x * 3;
    const int gray = (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + (1 << 14)) >> 15;
    const float v = m[(size_t)y * mstep + x] ? __fmul_rn((float)gray, (float)(1.0 / 255.0)) : 0.f;
    g[(size_t)y * pitch + x] = v;
    q[(size_t)y * pitch + x] = __fmul_rn(v, d[(size_t)y * pitch + x]);
}

struct EqImage {
    const float *d, *q, *g;   // resized distance map, g * d, masked gray (pitch in floats)
    const uint8_t *m;         // resized mask
    size_t pitch, mstep;
    int x, y, w, h;           // scaled ROI on the preview canvas (util::scaleRect); w, h may differ from the field size by 1
    int fw, fh;               // size of the fields
};

// the accumulation over the overlaps (ascending j, float adds in that order) and the per-pixel algebra
__global__ void equalize_kernel(const EqImage *im, int n, int i, float *out, size_t opitch)
{
    const EqImage I = im[i];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= I.fw || y >= I.fh) return;
    const size_t at = (size_t)y * I.pitch + x;
    float Q = I.q[at], A = I.d[at];
    const uint8_t mv = I.m[(size_t)y * I.mstep + x];
    if (mv) {
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const EqImage J = im[j];
            // overlap = roi_i & roi_j on the canvas; local coordinates of this pixel in both images
            const int ox0 = max(I.x, J.x), oy0 = max(I.y, J.y);
            const int ox1 = min(I.x + I.w, J.x + J.w), oy1 = min(I.y + I.h, J.y + J.h);
            if (ox1 <= ox0 || oy1 <= oy0) continue;
            const int cx = I.x + x, cy = I.y + y;
            if (cx < ox0 || cx >= ox1 || cy < oy0 || cy >= oy1) continue;
            const int jx = cx - J.x, jy = cy - J.y;
            if (jx >= J.fw || jy >= J.fh) continue;
            Q = __fadd_rn(Q, J.q[(size_t)jy * J.pitch + jx]);
            A = __fadd_rn(A, J.d[(size_t)jy * J.pitch + jx]);
        }
    }
    const float eps = 0.00001f;
    A = __fadd_rn(A, eps);
    float t = (A != 0.f) ? __fdiv_rn(Q, A) : 0.f;
    t = __fadd_rn(t, eps);
    t = (t != 0.f) ? __fdiv_rn(I.g[at], t) : 0.f;
    t = __fadd_rn(t, __fmul_rn((float)(255 - mv), (float)(1.0 / 255.0)));
    out[(size_t)y * opitch + x] = t;
}

struct Taps13 { float t[13]; };

__device__ __forceinline__ int reflect(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// one pass of the separable 13-tap filter (BORDER_REFLECT); dir = 0 rows, 1 columns
__global__ void blur13_kernel(const float *src, float *dst, int w, int h, size_t pitch, Taps13 T, int dir)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const int xx = dir ? x : reflect(x + k - 6, w), yy = dir ? reflect(y + k - 6, h) : y;
        a = __fadd_rn(a, __fmul_rn(src[(size_t)yy * pitch + xx], T.t[k]));
    }
    dst[(size_t)y * pitch + x] = a;
}

inline size_t al(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cv_round(double v) { return (int)std::nearbyint(v); }   // cvRound: round half to even

} // namespace

// size of the field of a w x h preview tile: cv::resize(src, dst, Size(), ratio, ratio) -> Size(cvRound(w * ratio), cvRound(h * ratio))
void spano_equalize_field_size(int w, int h, float ratio, int *fw, int *fh)
{
    *fw = cv_round((double)w * (double)ratio);
    *fh = cv_round((double)h * (double)ratio);
}

// tiles / masks: DEVICE buffers at preview scale (8UC3 / 8UC1); out[i]: device float buffers of the field sizes
int launch_equalize_intensities(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const uint8_t *const *masks,
                                const size_t *msteps, const int *tl_x, const int *tl_y, const int *w, const int *h, float ratio,
                                float *const *out, const size_t *opitch)
{
    int min_x = INT32_MAX, min_y = INT32_MAX;
    for (int i = 0; i < n; ++i) { min_x = std::min(min_x, tl_x[i]); min_y = std::min(min_y, tl_y[i]); }
    // arena: per image D (full size), then the resized mask / tile / d / g / q / pre-blur + axis tables
    std::vector<size_t> offD(n), offM(n), offT(n), offd(n), offg(n), offq(n), offX(n), offY(n), dpitch(n), fpitch(n);
    std::vector<int> fw(n), fh(n);
    size_t total = 0, tmp_max = 0;
    for (int i = 0; i < n; ++i) {
        spano_equalize_field_size(w[i], h[i], ratio, &fw[i], &fh[i]);
        if (fw[i] <= 0 || fh[i] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "equalizeIntensities: image %d is too small for ratio %g", i, (double)ratio);
        dpitch[i] = al((size_t)w[i], 4);
        fpitch[i] = al((size_t)fw[i], 4);
        offD[i] = total;  total += al(dpitch[i] * h[i] * sizeof(float), 256);
        offM[i] = total;  total += al(al((size_t)fw[i], 16) * fh[i], 256);
        offT[i] = total;  total += al(al((size_t)fw[i] * 3, 16) * fh[i], 256);
        offd[i] = total;  total += al(fpitch[i] * fh[i] * sizeof(float), 256);
        offg[i] = total;  total += al(fpitch[i] * fh[i] * sizeof(float), 256);
        offq[i] = total;  total += al(fpitch[i] * fh[i] * sizeof(float), 256);
        offX[i] = total;  total += al((size_t)fw[i] * sizeof(Axis), 256);
        offY[i] = total;  total += al((size_t)fh[i] * sizeof(Axis), 256);
        tmp_max = std::max(tmp_max, al(fpitch[i] * fh[i] * sizeof(float), 256));
    }
    const size_t offTmp = total;
    total += 2 * tmp_max;
    const size_t offDesc = total;
    total += al((size_t)n * sizeof(EqImage), 256);
    uint8_t *arena = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_EQUALIZE, total, (void **)&arena)) return rc;
    // distance transforms of all masks
    std::vector<float *> dist(n);
    std::vector<size_t> dsteps(n);
    for (int i = 0; i < n; ++i) { dist[i] = reinterpret_cast<float *>(arena + offD[i]); dsteps[i] = dpitch[i]; }
    if (int rc = launch_distance_transform(ctx, n, masks, msteps, w, h, dist.data(), dsteps.data()); rc < 0) return rc;
    int launches = 0;
    std::vector<EqImage> desc(n);
    for (int i = 0; i < n; ++i) {
        Axis *xt = reinterpret_cast<Axis *>(arena + offX[i]), *yt = reinterpret_cast<Axis *>(arena + offY[i]);
        const double inv_scale = 1.0 / (double)ratio;    // cv::resize with fx, fy given: scale = 1 / fx, whatever the rounded size is
        axis_kernel<<<(fw[i] + 255) / 256, 256, 0, ctx->stream>>>(w[i], fw[i], inv_scale, xt);
        axis_kernel<<<(fh[i] + 255) / 256, 256, 0, ctx->stream>>>(h[i], fh[i], inv_scale, yt);
        // cv::resize switches INTER_LINEAR to the 2x2 area average when the scale is exactly 2 in both directions
        const int area2 = (ratio == 0.5f);
        dim3 b(128), g((fw[i] + 127) / 128, fh[i]);
        uint8_t *m = arena + offM[i], *t = arena + offT[i];
        float *d = reinterpret_cast<float *>(arena + offd[i]), *gg = reinterpret_cast<float *>(arena + offg[i]),
              *q = reinterpret_cast<float *>(arena + offq[i]);
        const size_t ms = al((size_t)fw[i], 16), ts = al((size_t)fw[i] * 3, 16);
        resize_u8_kernel<1><<<g, b, 0, ctx->stream>>>(masks[i], w[i], h[i], msteps[i], m, fw[i], fh[i], ms, xt, yt, area2);
        resize_u8_kernel<3><<<g, b, 0, ctx->stream>>>(tiles[i], w[i], h[i], tsteps[i], t, fw[i], fh[i], ts, xt, yt, area2);
        resize_f32_kernel<<<g, b, 0, ctx->stream>>>(dist[i], w[i], h[i], dpitch[i], d, fw[i], fh[i], fpitch[i], xt, yt, area2, (float)(1.0 / 255.0));
        intensity_kernel<<<g, b, 0, ctx->stream>>>(t, ts, m, ms, d, gg, q, fw[i], fh[i], fpitch[i]);
        launches += 6;
        // roi = scaleRect(Rect(corner - min, mask size), cols_small / cols, rows_small / rows)   (std::round)
        const double C = (double)fw[i] / w[i], R = (double)fh[i] / h[i];
        EqImage &e = desc[i];
        e.d = d;  e.q = q;  e.g = gg;  e.m = m;  e.pitch = fpitch[i];  e.mstep = ms;
        e.x = (int)std::round((tl_x[i] - min_x) * C);  e.y = (int)std::round((tl_y[i] - min_y) * R);
        e.w = (int)std::round(w[i] * C);  e.h = (int)std::round(h[i] * R);
        e.fw = fw[i];  e.fh = fh[i];
    }
    EqImage *d_desc = reinterpret_cast<EqImage *>(arena + offDesc);
    SPANO_CUDA(ctx, cudaMemcpyAsync(d_desc, desc.data(), (size_t)n * sizeof(EqImage), cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // (`desc` is a pageable stack vector)
    // cv::getGaussianKernel(13, 7, CV_32F)
    Taps13 T;
    {
        float taps[13];
        spano_host_gaussian_taps(13, 7.0, taps);
        for (int k = 0; k < 13; ++k) T.t[k] = taps[k];
    }
    float *tmp0 = reinterpret_cast<float *>(arena + offTmp), *tmp1 = reinterpret_cast<float *>(arena + offTmp + tmp_max);
    for (int i = 0; i < n; ++i) {
        dim3 b(128), g((fw[i] + 127) / 128, fh[i]);
        equalize_kernel<<<g, b, 0, ctx->stream>>>(d_desc, n, i, tmp0, fpitch[i]);
        blur13_kernel<<<g, b, 0, ctx->stream>>>(tmp0, tmp1, fw[i], fh[i], fpitch[i], T, 0);
        blur13_kernel<<<g, b, 0, ctx->stream>>>(tmp1, tmp0, fw[i], fh[i], fpitch[i], T, 1);
        SPANO_CUDA(ctx, cudaMemcpy2DAsync(out[i], opitch[i] * sizeof(float), tmp0, fpitch[i] * sizeof(float), (size_t)fw[i] * sizeof(float), fh[i],
                                          cudaMemcpyDeviceToDevice, ctx->stream));
        launches += 3;
    }
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += launches;
    return launches;
}
