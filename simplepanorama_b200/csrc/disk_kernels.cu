// disk_kernels.cu -- stereographic "little planet" centre fix: radial re-projection of warped tiles.
//
// Replaces sten_proj::disk_reproj (reference src/math/_projection.cpp:193-294) together with
// sten_proj::get_bounding_box / create_border (:87-190) and util::RadialNormalizer
// (src/system/_util.h:172-200, _util.cpp:603-625).
//
// Host part (spano_disk_plan): canvas-centre coordinates, normaliser (centre = circle centre,
// scale = 1 / farthest tile corner), normalised radius, and for every tile the bounding box of its
// forward-stretched border (1000 sample points) -- a few thousand scalar operations, done with the
// host libm in the reference's expression order.
// Device part (disk_gather_kernel): for every pixel of the new tile the inverse radial law
//     r_src = r^p (p - rho) + rho      (p = 2 quadratic, 1 linear)
// in polar coordinates about the centre, de-normalised and ROUNDED TO INTEGER source coordinates
// (RadialNormalizer::denormalizePoint returns cv::Point), so the reference's
// cv::remap(INTER_CUBIC, BORDER_CONSTANT) is an exact pixel gather.
// The reference calls sqrt/atan2/cos/sin unqualified on floats, i.e. the double overloads, and
// stores the results in float members; the kernel does the same (double math, float stores), which
// makes the integer coordinates reproducible bit for bit.
// HBM traffic: 3 B gathered + 3 B written per new-tile pixel (+ the validity-mask pass afterwards).
#include <cmath>
#include <climits>
#include <algorithm>

#include "spano_internal.h"

namespace {

__host__ __device__ inline void norm_point(const SpanoDiskParams &p, int x, int y, float &fx, float &fy)
{
    fx = ((float)x - p.cx) * p.scale;
    fy = ((float)y - p.cy) * p.scale;
}

__host__ __device__ inline void denorm_point(const SpanoDiskParams &p, float fx, float fy, int &x, int &y)
{
#ifdef __CUDA_ARCH__
    x = (int)(__fadd_rn(__fadd_rn(__fdiv_rn(fx, p.scale), p.cx), 0.5f));
    y = (int)(__fadd_rn(__fadd_rn(__fdiv_rn(fy, p.scale), p.cy), 0.5f));
#else
    x = (int)((fx / p.scale) + p.cx + 0.5f);
    y = (int)((fy / p.scale) + p.cy + 0.5f);
#endif
}

__global__ void disk_gather_kernel(const SpanoDiskParams P, const uint8_t *src, int sw, int sh, size_t sstep, int ox, int oy,
                                   uint8_t *dst, int dw, int dh, size_t dstep, int dx0, int dy0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    float fx, fy;
    fx = __fmul_rn(__fsub_rn((float)(x + dx0), P.cx), P.scale);
    fy = __fmul_rn(__fsub_rn((float)(y + dy0), P.cy), P.scale);
    float r = (float)sqrt((double)__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
    const float phi = (float)atan2((double)fy, (double)fx);
    float e;
    int sub;
    if (P.quadratic) { e = __fmul_rn(r, r); sub = 2; }
    else { e = r; sub = 1; }
    r = __fadd_rn(__fmul_rn(e, __fsub_rn((float)sub, P.radius_n)), P.radius_n);
    double sn, cs;
    sincos((double)phi, &sn, &cs);
    fx = (float)((double)r * cs);
    fy = (float)((double)r * sn);
    int qx, qy;
    denorm_point(P, fx, fy, qx, qy);
    qx -= ox;
    qy -= oy;
    uint8_t *D = dst + (size_t)y * dstep + (size_t)x * 3;
    if (qx >= 0 && qx < sw && qy >= 0 && qy < sh) {
        const uint8_t *S = src + (size_t)qy * sstep + (size_t)qx * 3;
        D[0] = __ldg(S); D[1] = __ldg(S + 1); D[2] = __ldg(S + 2);
    } else {
        D[0] = D[1] = D[2] = 0;
    }
}

} // namespace

// Geometry of the centre fix.  In: tile corners/sizes, circle (ansatz, radius) in canvas pixel
// coordinates as sten_proj::estimate_circle returns them.  Out: params, the centre-relative original
// corners (org) and the new centre-relative corners / sizes of every tile.
int spano_disk_plan(int n, const int *tl_x, const int *tl_y, const int *w, const int *h, int ansatz_x, int ansatz_y,
                    float radius, int quadratic, SpanoDiskParams *P, int *org_x, int *org_y, int *new_x, int *new_y,
                    int *new_w, int *new_h)
{
    int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
    for (int i = 0; i < n; ++i) {
        x0 = std::min(x0, tl_x[i]);
        y0 = std::min(y0, tl_y[i]);
        x1 = std::max(x1, tl_x[i] + w[i]);
        y1 = std::max(y1, tl_y[i] + h[i]);
    }
    const int W = x1 - x0, H = y1 - y0;
    const int sx = W / 2 + 1, sy = H / 2 + 1;     // integer halves, as in the reference
    const int ax = ansatz_x - sx, ay = ansatz_y - sy;
    P->cx = (float)ax;
    P->cy = (float)ay;
    P->quadratic = quadratic ? 1 : 0;
    float far = 0.f;
    for (int i = 0; i < n; ++i) {
        org_x[i] = tl_x[i] - (x0 + sx);
        org_y[i] = tl_y[i] - (y0 + sy);
        for (int k = 0; k < 4; ++k) {
            const float dx = (float)(org_x[i] + ((k == 1 || k == 2) ? w[i] : 0)) - P->cx;
            const float dy = (float)(org_y[i] + (k >= 2 ? h[i] : 0)) - P->cy;
            far = std::max(far, std::sqrt(dx * dx + dy * dy));
        }
    }
    P->scale = (far == 0.0f) ? 1.0f : 1.0f / far;
    {
        float fx, fy;
        norm_point(*P, ax, ay + (int)radius, fx, fy);
        P->radius_n = (float)std::sqrt((double)(fx * fx + fy * fy));
    }
    const int N = 1000; // sten_proj::precision
    for (int i = 0; i < n; ++i) {
        const int bx = org_x[i], by = org_y[i], bw = w[i] + 1, bh = h[i] + 1; // boundingRect of the 4 corners
        const float ppu = (float)N / (2 * (bw + bh));
        const int cnt_h = (int)(bw * ppu), cnt_v = (int)(bh * ppu);
        int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
        for (int side = 0; side < 4; ++side) {
            const bool horiz = (side & 1) == 0;
            const int cnt = horiz ? cnt_h : cnt_v;
            const float step = (float)(horiz ? bw : bh) / (cnt + 1);
            for (int k = 1; k <= cnt; ++k) {
                const int d = (int)(k * step);
                int qx = side == 0 ? bx + d : side == 1 ? bx + bw : side == 2 ? bx + bw - d : bx;
                int qy = side == 0 ? by : side == 1 ? by + d : side == 2 ? by + bh : by + bh - d;
                float fx, fy;
                norm_point(*P, qx, qy, fx, fy);
                float r = (float)std::sqrt((double)(fx * fx + fy * fy));
                const float phi = (float)std::atan2((double)fy, (double)fx);
                const float e = quadratic ? r * r : r;
                if (e > P->radius_n) r = (e - P->radius_n) / (1 - P->radius_n);
                fx = (float)((double)r * std::cos((double)phi));
                fy = (float)((double)r * std::sin((double)phi));
                denorm_point(*P, fx, fy, qx, qy);
                mnx = std::min(mnx, qx); mny = std::min(mny, qy);
                mxx = std::max(mxx, qx); mxy = std::max(mxy, qy);
            }
        }
        if (mnx > mxx) return -1; // no border samples (degenerate tile)
        new_x[i] = mnx;
        new_y[i] = mny;
        new_w[i] = mxx - mnx + 1;
        new_h[i] = mxy - mny + 1;
    }
    return 0;
}

int launch_disk_gather(spano_ctx *ctx, const SpanoDiskParams &P, const uint8_t *src, int sw, int sh, size_t sstep, int ox,
                       int oy, uint8_t *dst, int dw, int dh, size_t dstep, int dx0, int dy0)
{
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    disk_gather_kernel<<<grid, block, 0, ctx->stream>>>(P, src, sw, sh, sstep, ox, oy, dst, dw, dh, dstep, dx0, dy0);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}
