// dist_kernels.cu -- chamfer distance transform and the distance-based seam cut (sm_100a).
//
// Replaces, at preview scale,
//   cv::distanceTransform(mask, dist, cv::DIST_L2, cv::DIST_MASK_5, CV_32F)   (reference src/math/_distance_cut.cpp:63,
//                                                                              src/math/_blending.cpp:110)
//   dcut::distance_transform + dcut::dist_cut                                 (src/math/_distance_cut.cpp:7-73)
//
// Arithmetic contract.  The OpenCV build the reference is pinned to here (cv2 4.13 + IPP) evaluates the 5x5 chamfer
// transform as the classic two raster passes in FLOAT32 with the metrics a = 1, b = 1.4, c = 2.1969:
//   forward  (top-left -> bottom-right)  t = min over {(-2,-1),(-2,+1),(-1,-2),(-1,+2)} + c, {(-1,-1),(-1,+1)} + b,
//                                                     {(-1,0),(0,-1)} + a;   t = 0 where the mask is 0
//   backward (bottom-right -> top-left)  the mirrored mask, applied to the forward result
// with FLT_MAX outside the image (FLT_MAX + metric == FLT_MAX).  Float addition does not associate, so the value of
// a pixel is the rounding history of one particular shortest path and a relaxation to a fixed point could differ in
// the last bit: the kernel therefore reproduces the DATA FLOW of the two passes exactly and only reorders
// independent pixels -- pixel (y, x) of the forward pass depends on (y, x-1), (y-1, x-2..x+2) and (y-2, x+-1), all of
// which lie on earlier anti-diagonals  x + 3 y = const, so a wavefront over t = x + 3 y runs every pixel of a
// diagonal in parallel (one CTA per image, one barrier per diagonal).  Bit-exact against cv2 (tests/golden/dist.npz).
#include "spano_internal.h"
#include <cfloat>

namespace {

constexpr float CH_A = 1.0f, CH_B = 1.4f, CH_C = 2.1969f;

struct DtJob {
    const uint8_t *src;
    size_t sstep;
    int w, h;
    float *tmp;   // (h + 4) x (w + 4) floats, 2-pixel frame
    float *dst;
    size_t dstep; // floats
};

__global__ void __launch_bounds__(1024) chamfer_dt_kernel(const DtJob *jobs)
{
    const DtJob J = jobs[blockIdx.x];
    const int w = J.w, h = J.h, P = w + 4;
    float *T = J.tmp + 2 * P + 2;   // T[y * P + x], valid for y, x in [-2, h + 1] x [-2, w + 1]
    // frame
    for (int i = threadIdx.x; i < (h + 4) * P; i += blockDim.x) {
        const int y = i / P - 2, x = i % P - 2;
        if (y < 0 || y >= h || x < 0 || x >= w) J.tmp[i] = FLT_MAX;
    }
    __syncthreads();
    const int steps = (w - 1) + 3 * (h - 1);
    // forward pass
    for (int t = 0; t <= steps; ++t) {
        const int y_lo = max(0, (t - (w - 1) + 2) / 3), y_hi = min(h - 1, t / 3);
        for (int y = y_lo + threadIdx.x; y <= y_hi; y += blockDim.x) {
            const int x = t - 3 * y;
            float v = 0.f;
            if (J.src[(size_t)y * J.sstep + x]) {
                const float *p = T + y * P + x;
                v = __fadd_rn(p[-2 * P - 1], CH_C);
                v = fminf(v, __fadd_rn(p[-2 * P + 1], CH_C));
                v = fminf(v, __fadd_rn(p[-P - 2], CH_C));
                v = fminf(v, __fadd_rn(p[-P - 1], CH_B));
                v = fminf(v, __fadd_rn(p[-P], CH_A));
                v = fminf(v, __fadd_rn(p[-P + 1], CH_B));
                v = fminf(v, __fadd_rn(p[-P + 2], CH_C));
                v = fminf(v, __fadd_rn(p[-1], CH_A));
            }
            T[y * P + x] = v;
        }
        __syncthreads();
    }
    // backward pass: the same wavefront on the point-mirrored image
    for (int t = 0; t <= steps; ++t) {
        const int y_lo = max(0, (t - (w - 1) + 2) / 3), y_hi = min(h - 1, t / 3);
        for (int ym = y_lo + threadIdx.x; ym <= y_hi; ym += blockDim.x) {
            const int y = h - 1 - ym, x = w - 1 - (t - 3 * ym);
            float *p = T + y * P + x;
            float v = *p;
            v = fminf(v, __fadd_rn(p[2 * P + 1], CH_C));
            v = fminf(v, __fadd_rn(p[2 * P - 1], CH_C));
            v = fminf(v, __fadd_rn(p[P + 2], CH_C));
            v = fminf(v, __fadd_rn(p[P + 1], CH_B));
            v = fminf(v, __fadd_rn(p[P], CH_A));
            v = fminf(v, __fadd_rn(p[P - 1], CH_B));
            v = fminf(v, __fadd_rn(p[P - 2], CH_C));
            v = fminf(v, __fadd_rn(p[1], CH_A));
            *p = v;
            J.dst[(size_t)y * J.dstep + x] = v;
        }
        __syncthreads();
    }
}

// dcut::dist_cut for one image i: a pixel keeps its mask value unless some overlapping image j has a strictly larger
// (scaled) distance there.  `others` lists the overlapping images.
struct CutOther {
    const float *dist;   // DT of image j (unscaled)
    size_t dstep;        // floats
    int dx, dy;          // pixel (x, y) of image i is pixel (x + dx, y + dy) of image j
    int w, h;
};

__global__ void dist_cut_kernel(const uint8_t *mask, size_t mstep, const float *dist, size_t dstep, int w, int h,
                                const CutOther *others, int n_others, uint8_t *cut, size_t cstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const float s = (float)(1.0 / 255.0);   // `transformed / 255` is a MatExpr: multiplication by (float)(1/255.)
    const float di = __fmul_rn(dist[(size_t)y * dstep + x], s);
    uint8_t v = mask[(size_t)y * mstep + x];
    for (int k = 0; k < n_others; ++k) {
        const CutOther o = others[k];
        const int xj = x + o.dx, yj = y + o.dy;
        if (xj < 0 || yj < 0 || xj >= o.w || yj >= o.h) continue;
        const float dj = __fmul_rn(o.dist[(size_t)yj * o.dstep + xj], s);
        if (di < dj) v = 0;                 // threshold(-(Di - Dj), 0, 1, BINARY) == 1  <=>  Di < Dj
    }
    cut[(size_t)y * cstep + x] = v;
}

// ---- blnd::simple_blend (reference src/math/_blending.cpp:83-153) and blnd::no_blend (:157-182) -----------------
// min / max of a distance map (non-negative floats order like their bit patterns)
__global__ void minmax_kernel(const float *dist, size_t dstep, int w, int h, uint32_t *mm)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    uint32_t lo = 0xffffffffu, hi = 0u;
    if (x < w && y < h) lo = hi = __float_as_uint(dist[(size_t)y * dstep + x]);
    for (int d = 16; d; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

// one image of simple_blend: alpha = normalize(dist, 0, 1, NORM_MINMAX); over-composite into acc {B,G,R,alpha}
__global__ void simple_accumulate_kernel(const uint8_t *tile, size_t tstep, const float *dist, size_t dstep, int w, int h,
                                         const uint32_t *mm, float4 *acc, int canvas_w, int ax, int ay)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    // cv::normalize(NORM_MINMAX, 0, 1): scale and shift in double, applied as floats (convertTo)
    const float smin = __uint_as_float(mm[0]), smax = __uint_as_float(mm[1]);
    const double range = (double)smax - (double)smin;
    const double scale = range > 2.220446049250313e-16 ? 1.0 / range : 0.0;
    const float a = (float)scale, b = (float)(0.0 - (double)smin * scale);
    const float m = __fadd_rn(__fmul_rn(dist[(size_t)y * dstep + x], a), b);
    float4 *p = acc + (size_t)(ay + y) * canvas_w + (ax + x);
    float4 c = *p;
    const float om = __fsub_rn(1.0f, c.w);
    const uint8_t *t = tile + (size_t)y * tstep + (size_t)x * 3;
    const float inv255 = (float)(1.0 / 255.0);
    c.x = __fadd_rn(c.x, __fmul_rn(__fmul_rn(__fmul_rn((float)t[0], inv255), m), om));
    c.y = __fadd_rn(c.y, __fmul_rn(__fmul_rn(__fmul_rn((float)t[1], inv255), m), om));
    c.z = __fadd_rn(c.z, __fmul_rn(__fmul_rn(__fmul_rn((float)t[2], inv255), m), om));
    c.w = __fadd_rn(c.w, __fmul_rn(m, om));
    *p = c;
}

__global__ void simple_finish_kernel(const float4 *acc, int canvas_w, int canvas_h, uint8_t *out, size_t ostep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= canvas_w || y >= canvas_h) return;
    const float4 c = acc[(size_t)y * canvas_w + x];
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (c.w > 0.f) {
        const float ia = __fdiv_rn(1.f, c.w);   // Vec3f / float == Vec3f * (1.f / a)
        r0 = __fmul_rn(c.x, ia); r1 = __fmul_rn(c.y, ia); r2 = __fmul_rn(c.z, ia);
    }
    uint8_t *o = out + (size_t)y * ostep + (size_t)x * 3;
    o[0] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(r0, 255.f))));
    o[1] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(r1, 255.f))));
    o[2] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn(r2, 255.f))));
}

// images[i].copyTo(panorama(roi), masks[i]): later images overwrite earlier ones where their mask is non-zero
__global__ void no_blend_kernel(const uint8_t *tile, size_t tstep, const uint8_t *mask, size_t mstep, int w, int h, uint8_t *out,
                                size_t ostep, int ax, int ay)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h || !mask[(size_t)y * mstep + x]) return;
    const uint8_t *t = tile + (size_t)y * tstep + (size_t)x * 3;
    uint8_t *o = out + (size_t)(ay + y) * ostep + (size_t)(ax + x) * 3;
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}

// ---- gain::get_overlapp_intensity (reference src/math/_gain_compensation.cpp:7-75) ----------------------------
// cv::cvtColor(BGR2GRAY) on 8 bit: (3735 B + 19235 G + 9798 R + 2^14) >> 15
__global__ void gray_kernel(const uint8_t *bgr, size_t step, int w, int h, uint8_t *gray, size_t gstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = bgr + (size_t)y * step + (size_t)x * 3;
    gray[(size_t)y * gstep + x] = (uint8_t)((3735u * p[0] + 19235u * p[1] + 9798u * p[2] + (1u << 14)) >> 15);
}

// one pair: over the overlap rectangle, count the pixels valid in both masks and sum both gray images there
// acc[0] = area, acc[1] = sum gray_i, acc[2] = sum gray_j (exact integers)
__global__ void overlap_sums_kernel(const uint8_t *gi, size_t gis, const uint8_t *mi, size_t mis, const uint8_t *gj, size_t gjs,
                                    const uint8_t *mj, size_t mjs, int xi, int yi, int xj, int yj, int ow, int oh,
                                    unsigned long long *acc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    unsigned int a = 0, si = 0, sj = 0;
    if (x < ow && y < oh && mi[(size_t)(yi + y) * mis + xi + x] && mj[(size_t)(yj + y) * mjs + xj + x]) {
        a = 1;
        si = gi[(size_t)(yi + y) * gis + xi + x];
        sj = gj[(size_t)(yj + y) * gjs + xj + x];
    }
    for (int d = 16; d; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        si += __shfl_xor_sync(0xffffffffu, si, d);
        sj += __shfl_xor_sync(0xffffffffu, sj, d);
    }
    if ((threadIdx.x & 31) == 0 && a) {
        atomicAdd(acc, (unsigned long long)a);
        atomicAdd(acc + 1, (unsigned long long)si);
        atomicAdd(acc + 2, (unsigned long long)sj);
    }
}

} // namespace

int launch_gray(spano_ctx *ctx, const uint8_t *bgr, size_t step, int w, int h, uint8_t *gray, size_t gstep)
{
    dim3 block(256), grid((w + 255) / 256, h);
    gray_kernel<<<grid, block, 0, ctx->stream>>>(bgr, step, w, h, gray, gstep);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_overlap_sums(spano_ctx *ctx, const uint8_t *gi, size_t gis, const uint8_t *mi, size_t mis, const uint8_t *gj, size_t gjs,
                        const uint8_t *mj, size_t mjs, int xi, int yi, int xj, int yj, int ow, int oh, unsigned long long *acc)
{
    dim3 block(256), grid((ow + 255) / 256, oh);
    overlap_sums_kernel<<<grid, block, 0, ctx->stream>>>(gi, gis, mi, mis, gj, gjs, mj, mjs, xi, yi, xj, yj, ow, oh, acc);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

// blnd::simple_blend on device buffers: tiles/masks/dist per image (dist = distance transforms already computed),
// acc = canvas_w x canvas_h float4 scratch, out = 8UC3 canvas.  Images are composited in order.
int launch_simple_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const float *const *dist,
                        const size_t *dsteps, const int *ax, const int *ay, const int *w, const int *h, float4 *acc, int canvas_w,
                        int canvas_h, uint8_t *out, size_t ostep)
{
    uint32_t *mm = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_JOBS2, (size_t)std::max(1, n) * 2 * sizeof(uint32_t), (void **)&mm)) return rc;
    std::vector<uint32_t> init(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { init[2 * i] = 0xffffffffu; init[2 * i + 1] = 0u; }
    SPANO_CUDA(ctx, cudaMemcpyAsync(mm, init.data(), init.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SPANO_CUDA(ctx, cudaMemsetAsync(acc, 0, (size_t)canvas_w * canvas_h * sizeof(float4), ctx->stream));
    for (int i = 0; i < n; ++i) {
        dim3 block(256), grid((w[i] + 255) / 256, h[i]);
        minmax_kernel<<<grid, block, 0, ctx->stream>>>(dist[i], dsteps[i], w[i], h[i], mm + 2 * i);
        simple_accumulate_kernel<<<grid, block, 0, ctx->stream>>>(tiles[i], tsteps[i], dist[i], dsteps[i], w[i], h[i], mm + 2 * i, acc,
                                                                   canvas_w, ax[i], ay[i]);
    }
    dim3 block(256), grid((canvas_w + 255) / 256, canvas_h);
    simple_finish_kernel<<<grid, block, 0, ctx->stream>>>(acc, canvas_w, canvas_h, out, ostep);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2 * n + 1;
    return 2 * n + 1;
}

int launch_no_blend(spano_ctx *ctx, int n, const uint8_t *const *tiles, const size_t *tsteps, const uint8_t *const *masks,
                    const size_t *msteps, const int *ax, const int *ay, const int *w, const int *h, int canvas_w, int canvas_h,
                    uint8_t *out, size_t ostep)
{
    SPANO_CUDA(ctx, cudaMemset2DAsync(out, ostep, 0, (size_t)canvas_w * 3, canvas_h, ctx->stream));
    for (int i = 0; i < n; ++i) {
        dim3 block(256), grid((w[i] + 255) / 256, h[i]);
        no_blend_kernel<<<grid, block, 0, ctx->stream>>>(tiles[i], tsteps[i], masks[i], msteps[i], w[i], h[i], out, ostep, ax[i], ay[i]);
    }
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += n;
    return n;
}

// n images on the device: masks[k] (w[k] x h[k], step mstep[k]) -> dist[k] (float, pitch dstep[k] floats)
int launch_distance_transform(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *msteps, const int *w, const int *h,
                              float *const *dist, const size_t *dsteps)
{
    if (n <= 0) return 0;
    size_t tmp_floats = 0;
    for (int k = 0; k < n; ++k) {
        if (w[k] <= 0 || h[k] <= 0) return spano_fail(ctx, SPANO_E_INVALID, "distance transform: empty image");
        tmp_floats += (size_t)(w[k] + 4) * (h[k] + 4);
    }
    float *tmp = nullptr;
    DtJob *d_jobs = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_TMP, tmp_floats * sizeof(float), (void **)&tmp)) return rc;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_JOBS, (size_t)n * sizeof(DtJob), (void **)&d_jobs)) return rc;
    std::vector<DtJob> jobs(n);
    size_t off = 0;
    for (int k = 0; k < n; ++k) {
        jobs[k] = DtJob{masks[k], msteps[k], w[k], h[k], tmp + off, dist[k], dsteps[k]};
        off += (size_t)(w[k] + 4) * (h[k] + 4);
    }
    SPANO_CUDA(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)n * sizeof(DtJob), cudaMemcpyHostToDevice, ctx->stream));
    SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `jobs` is a local
    chamfer_dt_kernel<<<n, 1024, 0, ctx->stream>>>(d_jobs);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_dist_cut(spano_ctx *ctx, int n, const uint8_t *const *masks, const size_t *msteps, const float *const *dist,
                    const size_t *dsteps, const int *tl_x, const int *tl_y, const int *w, const int *h, uint8_t *const *cut,
                    const size_t *csteps)
{
    if (n <= 0) return 0;
    std::vector<CutOther> table;
    std::vector<int> first(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            const int x0 = std::max(tl_x[i], tl_x[j]), y0 = std::max(tl_y[i], tl_y[j]);
            const int x1 = std::min(tl_x[i] + w[i], tl_x[j] + w[j]), y1 = std::min(tl_y[i] + h[i], tl_y[j] + h[j]);
            if (x1 <= x0 || y1 <= y0) continue;   // cv::Rect & cv::Rect is empty
            table.push_back(CutOther{dist[j], dsteps[j], tl_x[i] - tl_x[j], tl_y[i] - tl_y[j], w[j], h[j]});
        }
        first[i + 1] = (int)table.size();
    }
    CutOther *d_table = nullptr;
    if (int rc = spano_reserve(ctx, spano_ctx::BUF_DT_JOBS2, std::max<size_t>(1, table.size()) * sizeof(CutOther), (void **)&d_table)) return rc;
    if (!table.empty()) {
        SPANO_CUDA(ctx, cudaMemcpyAsync(d_table, table.data(), table.size() * sizeof(CutOther), cudaMemcpyHostToDevice, ctx->stream));
        SPANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < n; ++i) {
        dim3 block(256), grid((w[i] + 255) / 256, h[i]);
        dist_cut_kernel<<<grid, block, 0, ctx->stream>>>(masks[i], msteps[i], dist[i], dsteps[i], w[i], h[i], d_table + first[i],
                                                         first[i + 1] - first[i], cut[i], csteps[i]);
    }
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += n;
    return n;
}
