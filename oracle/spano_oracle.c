/*
 * spano_oracle.c -- CPU restatement of SimplePanorama's compositing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (simplepanorama_b200/,
 * include/) may link, import or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * What it restates (file:line relative to the upstream reference tree):
 *   - proj::{spherical,cylindrical,sten}_proj::project      src/math/_projection.cpp:27-84,297-324
 *       -> cv::detail::{Spherical,Cylindrical,Stereographic}Warper::warp (OpenCV stitching,
 *          NOT vendored by the reference; CMakeLists.txt:19 find_package(OpenCV), unpinned).
 *          Restated from OpenCV's published algorithm (warpers_inl.hpp: setCameraParams,
 *          mapForward/mapBackward, detectResultRoi[ByBorder], buildMaps, warp) and
 *          imgproc remap (INTER_LINEAR, 8-bit fixed point, INTER_BITS=5, 15-bit weights).
 *   - blnd::createSurroundingMask + cv::erode(3 iters)      src/math/_blending.cpp:278-324,
 *                                                           src/math/_projection.cpp:441-443
 *   - gain application  imgs[i] / gain[i]                    src/classes/_panorama.cpp:321-327
 *   - blnd::multi_blend                                      src/math/_blending.cpp:186-252
 *   - imgm::elementwiseOperation                             src/math/_img_manipulation.cpp:31-84
 *   - util::get_pan_dimension                                src/system/_util.cpp:204-231
 *   - stitch_parameters::blend (MULTI_BLEND branch)          src/classes/_panorama.cpp:242-249
 *   - sten_proj::disk_reproj / get_bounding_box              src/math/_projection.cpp:132-294
 *   - cv::resize of mask_cut, test::adjust_intensity         src/classes/_panorama.cpp:329-335, src/test/_test.cpp:110-122
 *   - dcut::distance_transform / dcut::dist_cut              src/math/_distance_cut.cpp:7-73
 *   - blnd::simple_blend / blnd::no_blend                    src/math/_blending.cpp:83-182
 *   - gain::get_overlapp_intensity                           src/math/_gain_compensation.cpp:7-75
 *
 * PARITY PIN: the reference has no tests or golden vectors.  This
 * restatement is pinned against OpenCV 4.13.0 (python cv2, the only OpenCV in the build
 * image) by oracle/gen_golden.py -> tests/golden/ and tests/test_oracle_*.py:
 * maps/ROI/remap/gray/mask/erode are bit-exact against cv2; Gaussian passes are
 * accumulated in double here and agree with cv2's float32 passes to <= 1e-6 relative.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; contraction must stay off so
 * float expressions round exactly like OpenCV's SSE3-baseline build).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

#define ORC_PI 3.1415926535897932384626433832795

enum { ORC_SPHERICAL = 0, ORC_CYLINDRICAL = 1, ORC_STEREOGRAPHIC = 2 };

typedef struct {
    int kind;
    float scale;
    float k[9], rinv[9], r_kinv[9], k_rinv[9];
} orc_projector;

/* ---- ProjectorBase::setCameraParams (OpenCV warpers.cpp) -------------------------
 * K, R are CV_32F 3x3.  Rinv = R^T; R_Kinv = R * K.inv(); K_Rinv = K * Rinv.
 * cv::invert on a 3x3 CV_32F matrix evaluates cofactors and determinant in double and
 * stores float; the 3x3 float products go through cv::gemm's small-matrix branch, which
 * sums a[0]*b[0] + a[1]*b[1] + a[2]*b[2] in float, left to right.                      */
static void mat3_mul_f32(const float *a, const float *b, float *d)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float t = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
            d[i * 3 + j] = t;
        }
}

static void mat3_inv_f32(const float *s, float *d)
{
#define S(y, x) ((double)s[(y) * 3 + (x)])
    double det = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) -
                 S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
                 S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    if (det == 0.0) { memset(d, 0, 9 * sizeof(float)); return; }
    double id = 1.0 / det;
    float t[9];
    t[0] = (float)((S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * id);
    t[1] = (float)((S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * id);
    t[2] = (float)((S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * id);
    t[3] = (float)((S(1, 2) * S(2, 0) - S(1, 0) * S(2, 2)) * id);
    t[4] = (float)((S(0, 0) * S(2, 2) - S(0, 2) * S(2, 0)) * id);
    t[5] = (float)((S(0, 2) * S(1, 0) - S(0, 0) * S(1, 2)) * id);
    t[6] = (float)((S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0)) * id);
    t[7] = (float)((S(0, 1) * S(2, 0) - S(0, 0) * S(2, 1)) * id);
    t[8] = (float)((S(0, 0) * S(1, 1) - S(0, 1) * S(1, 0)) * id);
    memcpy(d, t, sizeof(t));
#undef S
}

void orc_set_camera(orc_projector *p, int kind, float scale, const float *K, const float *R)
{
    p->kind = kind;
    p->scale = scale;
    memcpy(p->k, K, 9 * sizeof(float));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) p->rinv[i * 3 + j] = R[j * 3 + i];
    float kinv[9];
    mat3_inv_f32(K, kinv);
    mat3_mul_f32(R, kinv, p->r_kinv);
    mat3_mul_f32(K, p->rinv, p->k_rinv);
}

/* export the derived matrices so the product's host code can be compared to them */
void orc_camera_mats(int kind, float scale, const float *K, const float *R, float *r_kinv, float *k_rinv)
{
    orc_projector p;
    orc_set_camera(&p, kind, scale, K, R);
    memcpy(r_kinv, p.r_kinv, sizeof(p.r_kinv));
    memcpy(k_rinv, p.k_rinv, sizeof(p.k_rinv));
}

/* ---- mapForward / mapBackward (OpenCV warpers_inl.hpp) --------------------------- */
static void map_forward(const orc_projector *p, float x, float y, float *u, float *v)
{
    const float *m = p->r_kinv;
    float x_ = m[0] * x + m[1] * y + m[2];
    float y_ = m[3] * x + m[4] * y + m[5];
    float z_ = m[6] * x + m[7] * y + m[8];
    if (p->kind == ORC_SPHERICAL) {
        *u = p->scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = p->scale * ((float)ORC_PI - acosf(w == w ? w : 0));
    } else if (p->kind == ORC_CYLINDRICAL) {
        *u = p->scale * atan2f(x_, z_);
        *v = p->scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    } else {
        float u_ = atan2f(x_, z_);
        float v_ = (float)ORC_PI - acosf(y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_));
        float r = sinf(v_) / (1 - cosf(v_));
        *u = p->scale * r * cosf(u_);
        *v = p->scale * r * sinf(u_);
    }
}

static void map_backward(const orc_projector *p, float u, float v, float *x, float *y)
{
    const float *m = p->k_rinv;
    float x_, y_, z_;
    u /= p->scale;
    v /= p->scale;
    if (p->kind == ORC_SPHERICAL) {
        float sinv = sinf((float)ORC_PI - v);
        x_ = sinv * sinf(u);
        y_ = cosf((float)ORC_PI - v);
        z_ = sinv * cosf(u);
    } else if (p->kind == ORC_CYLINDRICAL) {
        x_ = sinf(u);
        y_ = v;
        z_ = cosf(u);
    } else {
        float u_ = atan2f(v, u);
        float r = sqrtf(u * u + v * v);
        float v_ = 2 * atanf(1.f / r);
        float sinv = sinf((float)ORC_PI - v_);
        x_ = sinv * sinf(u_);
        y_ = cosf((float)ORC_PI - v_);
        z_ = sinv * cosf(u_);
    }
    float z;
    *x = m[0] * x_ + m[1] * y_ + m[2] * z_;
    *y = m[3] * x_ + m[4] * y_ + m[5] * z_;
    z = m[6] * x_ + m[7] * y_ + m[8] * z_;
    if (z > 0) { *x /= z; *y /= z; }
    else *x = *y = -1;
}

/* ---- detectResultRoi / detectResultRoiByBorder / SphericalWarper override -------- */
static void roi_acc(const orc_projector *p, float x, float y, float *tl_u, float *tl_v, float *br_u, float *br_v)
{
    float u, v;
    map_forward(p, x, y, &u, &v);
    /* std::min / std::max semantics: (b < a) ? b : a  -- NaN in u/v never replaces */
    *tl_u = (u < *tl_u) ? u : *tl_u;  *tl_v = (v < *tl_v) ? v : *tl_v;
    *br_u = (*br_u < u) ? u : *br_u;  *br_v = (*br_v < v) ? v : *br_v;
}

/* out[4] = tl_x, tl_y, br_x, br_y (inclusive ROI as in OpenCV; the warped tile is +1) */
void orc_warp_roi(int kind, float scale, const float *K, const float *R, int src_w, int src_h, int *out)
{
    orc_projector p;
    orc_set_camera(&p, kind, scale, K, R);
    float tl_u = FLT_MAX, tl_v = FLT_MAX, br_u = -FLT_MAX, br_v = -FLT_MAX;
    if (kind == ORC_STEREOGRAPHIC) {
        for (int y = 0; y < src_h; y++)
            for (int x = 0; x < src_w; x++) roi_acc(&p, (float)x, (float)y, &tl_u, &tl_v, &br_u, &br_v);
    } else {
        for (int i = 0; i < src_w; i++) {
            roi_acc(&p, (float)i, 0.f, &tl_u, &tl_v, &br_u, &br_v);
            roi_acc(&p, (float)i, (float)(src_h - 1), &tl_u, &tl_v, &br_u, &br_v);
        }
        for (int i = 0; i < src_h; i++) {
            roi_acc(&p, 0.f, (float)i, &tl_u, &tl_v, &br_u, &br_v);
            roi_acc(&p, (float)(src_w - 1), (float)i, &tl_u, &tl_v, &br_u, &br_v);
        }
    }
    int tlx = (int)tl_u, tly = (int)tl_v, brx = (int)br_u, bry = (int)br_v;
    if (kind == ORC_SPHERICAL) {
        float tl_uf = (float)tlx, tl_vf = (float)tly, br_uf = (float)brx, br_vf = (float)bry;
        for (int pass = 0; pass < 2; pass++) {
            float x = p.rinv[1];
            float y = pass == 0 ? p.rinv[4] : -p.rinv[4];
            float z = p.rinv[7];
            if (y > 0.f) {
                float x_ = (p.k[0] * x + p.k[1] * y) / z + p.k[2];
                float y_ = p.k[4] * y / z + p.k[5];
                if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
                    float pole = pass == 0 ? (float)(ORC_PI * p.scale) : 0.f;
                    tl_uf = fminf(tl_uf, 0.f); tl_vf = fminf(tl_vf, pole);
                    br_uf = fmaxf(br_uf, 0.f); br_vf = fmaxf(br_vf, pole);
                }
            }
        }
        tlx = (int)tl_uf; tly = (int)tl_vf; brx = (int)br_uf; bry = (int)br_vf;
    }
    out[0] = tlx; out[1] = tly; out[2] = brx; out[3] = bry;
}

/* buildMaps: xmap/ymap are (h x w) float, h = br_y-tl_y+1, w = br_x-tl_x+1 */
void orc_build_maps(int kind, float scale, const float *K, const float *R, int tl_x, int tl_y, int w, int h,
                    float *xmap, float *ymap)
{
    orc_projector p;
    orc_set_camera(&p, kind, scale, K, R);
    for (int v = 0; v < h; v++)
        for (int u = 0; u < w; u++)
            map_backward(&p, (float)(u + tl_x), (float)(v + tl_y), &xmap[(size_t)v * w + u], &ymap[(size_t)v * w + u]);
}

/* ---- cv::remap(INTER_LINEAR, BORDER_CONSTANT(0)) on 8UC3 with float maps -----------
 * imgproc/imgwarp.cpp: sx = cvRound(x*32), sy = cvRound(y*32); integer part saturated to
 * short; weights (32-fy)(32-fx)*32 ... (exact 15-bit table); D = (sum + 2^14) >> 15.
 * cvRound on x86 = cvtss2si: round-half-even, 0x80000000 on overflow/NaN.              */
static int cv_round_f32(float v)
{
    if (!(v > -2147483648.f && v < 2147483648.f)) return INT_MIN;
    return (int)lrintf(v);
}
static int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void orc_remap_linear_u8c3(const uint8_t *src, int sw, int sh, size_t sstep, const float *xmap, const float *ymap,
                           int dw, int dh, uint8_t *dst, size_t dstep)
{
    for (int dy = 0; dy < dh; dy++)
        for (int dx = 0; dx < dw; dx++) {
            int fx = cv_round_f32(xmap[(size_t)dy * dw + dx] * 32.f);
            int fy = cv_round_f32(ymap[(size_t)dy * dw + dx] * 32.f);
            int sx = sat_short(fx >> 5), sy = sat_short(fy >> 5);
            int ax = fx & 31, ay = fy & 31;
            int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32;
            int w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
            uint8_t *D = dst + (size_t)dy * dstep + (size_t)dx * 3;
            for (int c = 0; c < 3; c++) {
                int v00 = 0, v01 = 0, v10 = 0, v11 = 0;
                if (sy >= 0 && sy < sh) {
                    if (sx >= 0 && sx < sw) v00 = src[(size_t)sy * sstep + (size_t)sx * 3 + c];
                    if (sx + 1 >= 0 && sx + 1 < sw) v01 = src[(size_t)sy * sstep + (size_t)(sx + 1) * 3 + c];
                }
                if (sy + 1 >= 0 && sy + 1 < sh) {
                    if (sx >= 0 && sx < sw) v10 = src[(size_t)(sy + 1) * sstep + (size_t)sx * 3 + c];
                    if (sx + 1 >= 0 && sx + 1 < sw) v11 = src[(size_t)(sy + 1) * sstep + (size_t)(sx + 1) * 3 + c];
                }
                D[c] = (uint8_t)((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15);
            }
        }
}

/* ---- createSurroundingMask(img, invert=true, thresh=1) + erode(3x3, iterations=3) ---
 * gray = (3735*B + 19235*G + 9798*R + 2^14) >> 15  (cvtColor BGR2GRAY, 8-bit)
 * thresh = gray <= 1; flood from every border pixel (4-connected, exact value);
 * mask = NOT(border-connected dark region); three 3x3 erosions with the default
 * morphology border (+inf) == one 7x7 min with out-of-image ignored.                  */
void orc_gray_u8(const uint8_t *bgr, int w, int h, size_t step, uint8_t *gray)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = bgr + (size_t)y * step + (size_t)x * 3;
            gray[(size_t)y * w + x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + (1 << 14)) >> 15);
        }
}

void orc_surrounding_mask(const uint8_t *bgr, int w, int h, size_t step, int erode_iters, uint8_t *mask)
{
    size_t n = (size_t)w * h;
    uint8_t *dark = (uint8_t *)malloc(n);
    orc_gray_u8(bgr, w, h, step, dark);
    for (size_t i = 0; i < n; i++) dark[i] = dark[i] <= 1 ? 1 : 0;
    /* BFS flood from border dark pixels; dark==2 marks "outside" */
    int *stack = (int *)malloc(n * sizeof(int));
    size_t sp = 0;
#define PUSH(xx, yy) do { size_t q_ = (size_t)(yy) * w + (xx); if (dark[q_] == 1) { dark[q_] = 2; stack[sp++] = (int)q_; } } while (0)
    for (int x = 0; x < w; x++) { PUSH(x, 0); PUSH(x, h - 1); }
    for (int y = 0; y < h; y++) { PUSH(0, y); PUSH(w - 1, y); }
    while (sp) {
        int q = stack[--sp];
        int x = q % w, y = q / w;
        if (x > 0) PUSH(x - 1, y);
        if (x < w - 1) PUSH(x + 1, y);
        if (y > 0) PUSH(x, y - 1);
        if (y < h - 1) PUSH(x, y + 1);
    }
#undef PUSH
    free(stack);
    uint8_t *m0 = (uint8_t *)malloc(n);
    for (size_t i = 0; i < n; i++) m0[i] = dark[i] == 2 ? 0 : 255;
    free(dark);
    /* erode_iters x (3x3 min, outside ignored) */
    uint8_t *a = m0, *b = mask;
    uint8_t *tmp = (uint8_t *)malloc(n);
    for (int it = 0; it < erode_iters; it++) {
        b = (it == erode_iters - 1) ? mask : ((a == tmp) ? m0 : tmp);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                uint8_t mn = 255;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        int xx = x + dx, yy = y + dy;
                        if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
                        uint8_t v = a[(size_t)yy * w + xx];
                        if (v < mn) mn = v;
                    }
                b[(size_t)y * w + x] = mn;
            }
        a = b;
    }
    if (erode_iters == 0) memcpy(mask, m0, n);
    free(tmp);
    free(m0);
}

/* ---- gain: tile / g on CV_8UC3 == convertTo(alpha = 1/g): rint(float(v)*float(1/g)), saturated */
void orc_apply_gain_u8(uint8_t *data, size_t n, double gain)
{
    float a = (float)(1.0 / gain);
    for (size_t i = 0; i < n; i++) {
        float f = (float)data[i] * a;
        int r = cv_round_f32(f);
        data[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

/* ---- cv::getGaussianKernel(n, sigma, CV_32F): exp(-x^2/(2 sigma^2)) in double, normalised, -> float */
void orc_gaussian_taps(int n, double sigma, float *taps)
{
    int n2 = (n - 1) / 2;
    double *v = (double *)malloc((n2 + 1) * sizeof(double));
    double scale2x = -0.5 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0; i < n2; i++) {
        double x = (double)(i - n2);
        v[i] = exp(scale2x * x * x);
        sum += v[i];
    }
    sum = sum * 2.0 + 1.0;
    double mul1 = 1.0 / sum;
    for (int i = 0; i < n2; i++) {
        float t = (float)(v[i] * mul1);
        taps[i] = t;
        taps[n - 1 - i] = t;
    }
    taps[n2] = (float)mul1;
    free(v);
}

/* cv::borderInterpolate(p, len, BORDER_REFLECT): fedcba|abcdefgh|hgfedcb */
static int reflect_idx(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* cv::GaussianBlur(float, ksize n x n, sigma, BORDER_REFLECT) == sepFilter2D(row taps, col taps).
 * cn interleaved channels.  Row pass result is rounded to float (OpenCV's intermediate buffer
 * is CV_32F); each pass accumulates in double here (the oracle is the "true" value both the
 * float32 CPU path and the CUDA path have to stay within 1e-5 relative of).             */
void orc_gaussian_blur_f32(const float *src, int w, int h, int cn, int n, double sigma, float *dst)
{
    float *taps = (float *)malloc(n * sizeof(float));
    orc_gaussian_taps(n, sigma, taps);
    int r = n / 2;
    float *tmp = (float *)malloc((size_t)w * h * cn * sizeof(float));
    int *ix = (int *)malloc((size_t)(w + 2 * r) * sizeof(int));
    for (int x = -r; x < w + r; x++) ix[x + r] = reflect_idx(x, w);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float *s = src + (size_t)y * w * cn;
        float *t = tmp + (size_t)y * w * cn;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < cn; c++) {
                double acc = 0.0;
                for (int k = 0; k < n; k++) acc += (double)taps[k] * (double)s[(size_t)ix[x + k] * cn + c];
                t[(size_t)x * cn + c] = (float)acc;
            }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float *d = dst + (size_t)y * w * cn;
        for (size_t i = 0; i < (size_t)w * cn; i++) {
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += (double)taps[k] * (double)tmp[(size_t)reflect_idx(y + k - r, h) * w * cn + i];
            d[i] = (float)acc;
        }
    }
    free(ix);
    free(tmp);
    free(taps);
}

/* ---- util::get_pan_dimension: out[6] = width, height, min_x, min_y, max_x, max_y --- */
void orc_pan_dimension(int n, const int *tl_x, const int *tl_y, const int *w, const int *h, int *out)
{
    int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
    for (int i = 0; i < n; i++) {
        if (tl_x[i] < min_x) min_x = tl_x[i];
        if (tl_y[i] < min_y) min_y = tl_y[i];
        if (tl_x[i] + w[i] > max_x) max_x = tl_x[i] + w[i];
        if (tl_y[i] + h[i] > max_y) max_y = tl_y[i] + h[i];
    }
    out[0] = max_x - min_x; out[1] = max_y - min_y; out[2] = min_x; out[3] = min_y; out[4] = max_x; out[5] = max_y;
}

/* ---- blnd::multi_blend (src/math/_blending.cpp:186-252) ---------------------------
 * tiles[j]: 8UC3 (packed rows, step = 3*w[j]); masks[j] (mask_cut, 0..255) and
 * masks_orig[j] ({0,255}) 8UC1 packed.  out: canvas float32 x3 (height x width x 3).
 * Loop structure, band algebra, weight zeroing, the integer `255 / bands` divisor and the
 * alpha clamp follow the reference line by line.                                         */
void orc_multi_blend(int n, const uint8_t *const *tiles, const uint8_t *const *masks, const uint8_t *const *masks_orig,
                     const int *tl_x, const int *tl_y, const int *w, const int *h, int bands, double sigma, float *out)
{
    int dim[6];
    orc_pan_dimension(n, tl_x, tl_y, w, h, dim);
    int W = dim[0], H = dim[1];
    size_t N = (size_t)W * H;
    float *acc_c = out; /* accumulate in place */
    float *acc_a = (float *)calloc(N, sizeof(float));
    memset(acc_c, 0, N * 3 * sizeof(float));
    int ksize = 2 * (int)ceil(3 * sigma) + 1;

    for (int i = 0; i < bands; i++) {
        double sigma_band = sqrt(2 * (bands - i) + 1) * sigma;
        for (int j = 0; j < n; j++) {
            size_t tn = (size_t)w[j] * h[j];
            float *I = (float *)malloc(tn * 3 * sizeof(float));
            float *It = (float *)malloc(tn * 3 * sizeof(float));
            float *Wc = (float *)malloc(tn * sizeof(float));
            float *Wb = (float *)malloc(tn * sizeof(float));
            for (size_t q = 0; q < tn * 3; q++) I[q] = (float)tiles[j][q];
            for (size_t q = 0; q < tn; q++) Wc[q] = (float)masks[j][q];
            orc_gaussian_blur_f32(I, w[j], h[j], 3, ksize, sigma_band, It);
            orc_gaussian_blur_f32(Wc, w[j], h[j], 1, ksize, sigma_band, Wb);
            const float inv255 = (float)(1.0 / 255.0);
            for (size_t q = 0; q < tn; q++) Wb[q] = Wb[q] * inv255;
            if (i == bands - 1) {
                for (size_t q = 0; q < tn * 3; q++) It[q] = I[q] - It[q];
            } else if (i > 0) {
                double sigma_prev = sqrt(2 * (bands - i - 1) + 1) * sigma;
                float *P = (float *)malloc(tn * 3 * sizeof(float));
                orc_gaussian_blur_f32(I, w[j], h[j], 3, ksize, sigma_prev, P);
                for (size_t q = 0; q < tn * 3; q++) It[q] = It[q] - P[q];
                free(P);
            }
            for (size_t q = 0; q < tn; q++)
                if (masks_orig[j][q] != 255) Wb[q] = 0.f; /* setTo(0, ~mask_orig) */
            int cx = tl_x[j] - dim[2], cy = tl_y[j] - dim[3];
            for (int y = 0; y < h[j]; y++)
                for (int x = 0; x < w[j]; x++) {
                    size_t q = (size_t)y * w[j] + x;
                    size_t c = (size_t)(cy + y) * W + (cx + x);
                    float wv = Wb[q];
                    acc_c[c * 3 + 0] = acc_c[c * 3 + 0] + It[q * 3 + 0] * wv;
                    acc_c[c * 3 + 1] = acc_c[c * 3 + 1] + It[q * 3 + 1] * wv;
                    acc_c[c * 3 + 2] = acc_c[c * 3 + 2] + It[q * 3 + 2] * wv;
                    acc_a[c] = acc_a[c] + wv;
                }
            free(I); free(It); free(Wc); free(Wb);
        }
    }
    const float divisor = (float)(255 / bands); /* integer division, as in the reference */
    const float inv_div = (float)(1.0 / (double)divisor);
    for (size_t c = 0; c < N; c++) {
        float d = acc_a[c];
        d = copysignf(fmaxf(fabsf(d), 1e-6f), d);
        float s = 1.f / d; /* cv::Vec3f / float multiplies by 1.f/alpha */
        for (int k = 0; k < 3; k++) acc_c[c * 3 + k] = (acc_c[c * 3 + k] * s) * inv_div;
    }
    free(acc_a);
}

/* ---- stitch_parameters::blend, MULTI_BLEND tail: blend*255 -> convertTo(CV_8UC3) ---- */
void orc_blend_to_u8(const float *blend, size_t n, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) {
        float f = blend[i] * 255.f;
        int r = cv_round_f32(f);
        out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

/* ---- cv::resize(8UC1, INTER_LINEAR) as the reference effectively calls it for mask_cut
 * (src/classes/_panorama.cpp:333 passes INTER_CUBIC in the fx slot; interpolation stays
 * INTER_LINEAR).  8-bit linear resize is fixed point: 11-bit coefficients
 * (INTER_RESIZE_COEF_SCALE = 2048) per axis, result = (sum of products + 2^21) >> 22.     */
void orc_resize_linear_u8c1(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh)
{
    double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    int *xofs = (int *)malloc(dw * sizeof(int)), *yofs = (int *)malloc(dh * sizeof(int));
    short *ialpha = (short *)malloc(dw * 2 * sizeof(short)), *ibeta = (short *)malloc(dh * 2 * sizeof(short));
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[dx * 2] = (short)lrintf((1.f - fx) * 2048.f);
        ialpha[dx * 2 + 1] = (short)lrintf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[dy * 2] = (short)lrintf((1.f - fy) * 2048.f);
        ibeta[dy * 2 + 1] = (short)lrintf(fy * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;
        sy0 = sy0 < 0 ? 0 : (sy0 >= sh ? sh - 1 : sy0);
        sy1 = sy1 < 0 ? 0 : (sy1 >= sh ? sh - 1 : sy1);
        const uint8_t *S0 = src + (size_t)sy0 * sw, *S1 = src + (size_t)sy1 * sw;
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sx;
            int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
            int r0 = S0[sx] * a0 + S0[sx1] * a1;
            int r1 = S1[sx] * a0 + S1[sx1] * a1;
            /* VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>:
               (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2 */
            int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * dw + dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(yofs); free(ialpha); free(ibeta);
}

/* ==== sten_proj::disk_reproj (src/math/_projection.cpp:193-294) ===========================
 * Little-planet centre fix: every tile is resampled through a radial stretch about the
 * estimated circle centre so that the annulus fills the hole.  Restated line by line:
 *   util::get_pan_dimension, the shift of `ansatz` and the corners to canvas-centre coordinates
 *   (:196-210), util::RadialNormalizer::computeParameters (src/system/_util.cpp:603-625),
 *   the normalised radius (:214-218), sten_proj::get_bounding_box + create_border (:87-190),
 *   the per-pixel inverse radial map with integer-rounded source coordinates (:229-269;
 *   RadialNormalizer::denormalizePoint returns cv::Point, src/system/_util.h:191-195), and
 *   cv::remap(INTER_CUBIC, BORDER_CONSTANT) on integer maps == exact pixel gather (:278).
 * Unqualified sqrt/atan2/cos/sin on float arguments resolve to the double overloads with
 * libstdc++'s <cmath> (checked with g++ 13 for the reference's include set); results are
 * stored to float members, as in struct polar / struct kartesian (_projection.h:26-35).
 * PARITY PIN: the gather is pinned against cv2.remap(INTER_CUBIC) on integer maps
 * (tests/golden/kernels.npz); the geometry is the reference's own code and has no golden
 * vectors upstream: "parity unpinned" beyond this restatement.                              */
typedef struct {
    float cx, cy, scale;  /* RadialNormalizer */
    float radius_n;       /* normalised circle radius */
    int quadratic;
} orc_disk_params;

static void disk_norm_point(const orc_disk_params *p, int x, int y, float *fx, float *fy)
{
    *fx = ((float)x - p->cx) * p->scale;
    *fy = ((float)y - p->cy) * p->scale;
}

static void disk_denorm_point(const orc_disk_params *p, float fx, float fy, int *x, int *y)
{
    *x = (int)((fx / p->scale) + p->cx + 0.5f);
    *y = (int)((fy / p->scale) + p->cy + 0.5f);
}

static void disk_polar(float kx, float ky, float *r, float *phi)
{
    *r = (float)sqrt((double)(kx * kx + ky * ky));
    *phi = (float)atan2((double)ky, (double)kx);
}

static void disk_cart(float r, float phi, float *kx, float *ky)
{
    *kx = (float)((double)r * cos((double)phi));
    *ky = (float)((double)r * sin((double)phi));
}

/* in/out: tl_x, tl_y (corners; become the transformed, centre-relative corners), out sizes;
 * org_x/org_y receive the centre-relative original corners (org_bbox.x/y). */
void orc_disk_reproj_plan(int n, int *tl_x, int *tl_y, const int *w, const int *h, int ansatz_x, int ansatz_y, float radius,
                          int quadratic, int *out_w, int *out_h, int *org_x, int *org_y, orc_disk_params *P)
{
    int dim[6];
    orc_pan_dimension(n, tl_x, tl_y, w, h, dim);
    const int W = dim[0], H = dim[1], min_x = dim[2], min_y = dim[3];
    ansatz_x = ansatz_x - (int)(W / 2 + 1);
    ansatz_y = ansatz_y - (int)(H / 2 + 1);
    float maxd = 0.f;
    P->cx = (float)ansatz_x;
    P->cy = (float)ansatz_y;
    for (int i = 0; i < n; i++) {
        tl_x[i] -= min_x + (int)(W / 2 + 1);
        tl_y[i] -= min_y + (int)(H / 2 + 1);
        const int px[4] = {tl_x[i], tl_x[i] + w[i], tl_x[i] + w[i], tl_x[i]};
        const int py[4] = {tl_y[i], tl_y[i], tl_y[i] + h[i], tl_y[i] + h[i]};
        for (int k = 0; k < 4; k++) {
            float dx = (float)px[k] - P->cx, dy = (float)py[k] - P->cy;
            float d = sqrtf(dx * dx + dy * dy);
            if (d > maxd) maxd = d;
        }
    }
    P->scale = (maxd == 0.0f) ? 1.0f : 1.0f / maxd;
    P->quadratic = quadratic;
    {
        float fx, fy;
        disk_norm_point(P, ansatz_x, ansatz_y + (int)radius, &fx, &fy);
        P->radius_n = (float)sqrt((double)(fx * fx + fy * fy));
    }
    const int N = 1000; /* sten_proj::precision */
    for (int i = 0; i < n; i++) {
        /* boundingRect of the 4 integer corners is (cols+1) x (rows+1); x,y reset to the corner */
        const int bx = tl_x[i], by = tl_y[i], bw = w[i] + 1, bh = h[i] + 1;
        org_x[i] = bx;
        org_y[i] = by;
        const int perimeter = 2 * (bw + bh);
        const float ppu = (float)N / perimeter;
        const int top = (int)(bw * ppu), right = (int)(bh * ppu);
        int minx = INT_MAX, miny = INT_MAX, maxx = INT_MIN, maxy = INT_MIN;
        for (int side = 0; side < 4; side++) {
            const int cnt = (side % 2 == 0) ? top : right;
            const float step = (side % 2 == 0) ? (float)bw / (cnt + 1) : (float)bh / (cnt + 1);
            for (int k = 1; k <= cnt; k++) {
                int qx, qy;
                const int d = (int)(k * step);
                if (side == 0) { qx = bx + d; qy = by; }
                else if (side == 1) { qx = bx + bw; qy = by + d; }
                else if (side == 2) { qx = bx + bw - d; qy = by + bh; }
                else { qx = bx; qy = by + bh - d; }
                float fx, fy, r, phi;
                disk_norm_point(P, qx, qy, &fx, &fy);
                disk_polar(fx, fy, &r, &phi);
                float e = quadratic ? r * r : r;
                if (e > P->radius_n) r = (e - P->radius_n) / (1 - P->radius_n);
                disk_cart(r, phi, &fx, &fy);
                disk_denorm_point(P, fx, fy, &qx, &qy);
                if (qx < minx) minx = qx;
                if (qy < miny) miny = qy;
                if (qx > maxx) maxx = qx;
                if (qy > maxy) maxy = qy;
            }
        }
        tl_x[i] = minx;
        tl_y[i] = miny;
        out_w[i] = maxx - minx + 1;
        out_h[i] = maxy - miny + 1;
    }
}

/* one tile: dst (dw x dh at centre-relative corner (dx0,dy0)) gathers from src (sw x sh at (ox,oy)) */
void orc_disk_reproj_tile(const orc_disk_params *P, const uint8_t *src, int sw, int sh, size_t sstep, int ox, int oy,
                          uint8_t *dst, int dw, int dh, size_t dstep, int dx0, int dy0)
{
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            float fx, fy, r, phi;
            disk_norm_point(P, x + dx0, y + dy0, &fx, &fy);
            disk_polar(fx, fy, &r, &phi);
            float e;
            int sub;
            if (P->quadratic) { e = r * r; sub = 2; }
            else { e = r; sub = 1; }
            r = e * (sub - P->radius_n) + P->radius_n;
            disk_cart(r, phi, &fx, &fy);
            int qx, qy;
            disk_denorm_point(P, fx, fy, &qx, &qy);
            qx -= ox;
            qy -= oy;
            uint8_t *D = dst + (size_t)y * dstep + (size_t)x * 3;
            if (qx >= 0 && qx < sw && qy >= 0 && qy < sh) {
                const uint8_t *S = src + (size_t)qy * sstep + (size_t)qx * 3;
                D[0] = S[0]; D[1] = S[1]; D[2] = S[2];
            } else {
                D[0] = D[1] = D[2] = 0;
            }
        }
}

/* ==== test::adjust_intensity (src/test/_test.cpp:110-122; SURVEY section 8f "next" #1) ==========
 * per image: field = cv::resize(intensities[i], tile size, INTER_LINEAR)   (CV_32FC1, float bilinear)
 *            tile  = convertTo(CV_32FC3, 1/255) ; elementwiseOperation(DIVIDE by field, clamped) ;
 *            convertTo(CV_8UC3, 255)
 * i.e. per channel  u8 <- sat(rint(((float(v) * float(1/255)) * (1.f / clamp(field))) * 255.f)).
 * OpenCV's float linear resize: source coordinate (d + 0.5) * scale - 0.5 in double -> float, floor,
 * fraction kept in float, horizontal pass S[sx]*(1-fx) + S[sx+1]*fx, vertical pass r0*(1-fy) + r1*fy.
 * PARITY PIN: cv2.resize(float32) + the same arithmetic in numpy (tests/golden/intensity.npz); the
 * interpolated field agrees to float rounding (OpenCV's SIMD path may fuse one multiply-add), the
 * 8-bit result within 1 LSB.                                                                       */
void orc_resize_linear_f32c1(const float *src, int sw, int sh, float *dst, int dw, int dh)
{
    double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        int sy0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy), sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
        for (int dx = 0; dx < dw; dx++) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
            int sx1 = sx + 1 < sw ? sx + 1 : sx;
            float r0 = src[(size_t)sy0 * sw + sx] * (1.f - fx) + src[(size_t)sy0 * sw + sx1] * fx;
            float r1 = src[(size_t)sy1 * sw + sx] * (1.f - fx) + src[(size_t)sy1 * sw + sx1] * fx;
            dst[(size_t)dy * dw + dx] = r0 * (1.f - fy) + r1 * fy;
        }
    }
}

void orc_adjust_intensity(uint8_t *bgr, int w, int h, size_t step, const float *field, int fw, int fh)
{
    float *f = (float *)malloc((size_t)w * h * sizeof(float));
    orc_resize_linear_f32c1(field, fw, fh, f, w, h);
    const float inv255 = (float)(1.0 / 255.0);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float d = f[(size_t)y * w + x];
            d = copysignf(fmaxf(fabsf(d), 1e-6f), d);
            float s = 1.f / d;
            uint8_t *p = bgr + (size_t)y * step + (size_t)x * 3;
            for (int c = 0; c < 3; c++) {
                float v = ((float)p[c] * inv255) * s;
                int r = cv_round_f32(v * 255.f);
                p[c] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    free(f);
}

/* ---------------------------------------------------------------------------------------------
 * cv::distanceTransform(src, dst, DIST_L2, DIST_MASK_5, CV_32F) and dcut::dist_cut
 * (reference src/math/_distance_cut.cpp:7-73; also used by blnd::simple_blend, src/math/_blending.cpp:110).
 * OpenCV is not vendored by the reference; the build it is pinned to here (cv2 4.13.0 with IPP) evaluates the
 * 5x5 chamfer transform as the two raster passes of Borgefors' algorithm in float32 with the metrics
 * (1, 1.4, 2.1969) and FLT_MAX outside the image -- pinned by tests/golden/dist.npz (oracle/gen_golden_dist.py):
 * bit for bit on every vector except the one with distances beyond 32 px, where on a thin band of pixels (float
 * ties of `left neighbour + 1`) IPP's unpublished evaluation order ends one ulp above the two-pass minimum; the
 * tests bound that to <= 1 ulp on < 1 % of the pixels.  (OpenCV's own non-IPP fallback works in 16.16 fixed point
 * and differs from this build by ~1e-5 absolute; the pin decides.)
 * --------------------------------------------------------------------------------------------- */
void orc_distance_transform(const uint8_t *src, int w, int h, size_t step, float *dst)
{
    const float A = 1.0f, B = 1.4f, Cc = 2.1969f, INF = 3.402823466e+38f;
    const int P = w + 4;
    float *tmp = (float *)malloc((size_t)P * (h + 4) * sizeof(float));
    for (size_t i = 0; i < (size_t)P * (h + 4); ++i) tmp[i] = INF;
    float *T = tmp + 2 * P + 2;
#define MINF(a, b) ((b) < (a) ? (b) : (a))
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float *p = T + (size_t)y * P + x;
            if (!src[(size_t)y * step + x]) { *p = 0.f; continue; }
            float v = p[-2 * P - 1] + Cc, t;
            t = p[-2 * P + 1] + Cc; v = MINF(v, t);
            t = p[-P - 2] + Cc;     v = MINF(v, t);
            t = p[-P - 1] + B;      v = MINF(v, t);
            t = p[-P] + A;          v = MINF(v, t);
            t = p[-P + 1] + B;      v = MINF(v, t);
            t = p[-P + 2] + Cc;     v = MINF(v, t);
            t = p[-1] + A;          v = MINF(v, t);
            *p = v;
        }
    for (int y = h - 1; y >= 0; --y)
        for (int x = w - 1; x >= 0; --x) {
            float *p = T + (size_t)y * P + x;
            float v = *p, t;
            t = p[2 * P + 1] + Cc; v = MINF(v, t);
            t = p[2 * P - 1] + Cc; v = MINF(v, t);
            t = p[P + 2] + Cc;     v = MINF(v, t);
            t = p[P + 1] + B;      v = MINF(v, t);
            t = p[P] + A;          v = MINF(v, t);
            t = p[P - 1] + B;      v = MINF(v, t);
            t = p[P - 2] + Cc;     v = MINF(v, t);
            t = p[1] + A;          v = MINF(v, t);
            *p = v;
            dst[(size_t)y * w + x] = v;
        }
#undef MINF
    free(tmp);
}

/* dcut::dist_cut, src/math/_distance_cut.cpp:7-51: masks[i] contiguous w[i] x h[i]; cut[i] likewise.
 * `transformed / 255` is a cv::MatExpr, i.e. a multiplication by (float)(1/255.). */
void orc_dist_cut(int n, const uint8_t *const *masks, const int *tl_x, const int *tl_y, const int *w, const int *h,
                  uint8_t *const *cut)
{
    const float s = (float)(1.0 / 255.0);
    float **D = (float **)malloc((size_t)n * sizeof(float *));
    for (int i = 0; i < n; ++i) {
        D[i] = (float *)malloc((size_t)w[i] * h[i] * sizeof(float));
        orc_distance_transform(masks[i], w[i], h[i], (size_t)w[i], D[i]);
        for (size_t k = 0; k < (size_t)w[i] * h[i]; ++k) D[i][k] = D[i][k] * s;
    }
    for (int i = 0; i < n; ++i) {
        memcpy(cut[i], masks[i], (size_t)w[i] * h[i]);
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            const int x0 = tl_x[i] > tl_x[j] ? tl_x[i] : tl_x[j], y0 = tl_y[i] > tl_y[j] ? tl_y[i] : tl_y[j];
            const int xa = tl_x[i] + w[i], xb = tl_x[j] + w[j], ya = tl_y[i] + h[i], yb = tl_y[j] + h[j];
            const int x1 = xa < xb ? xa : xb, y1 = ya < yb ? ya : yb;
            if (x1 <= x0 || y1 <= y0) continue;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    const float diff = D[i][(size_t)(y - tl_y[i]) * w[i] + (x - tl_x[i])] - D[j][(size_t)(y - tl_y[j]) * w[j] + (x - tl_x[j])];
                    if (-diff > 0.f) cut[i][(size_t)(y - tl_y[i]) * w[i] + (x - tl_x[i])] = 0;   /* threshold(-diff,0,1) -> 1 - 1 = 0 */
                }
        }
    }
    for (int i = 0; i < n; ++i) free(D[i]);
    free(D);
}

/* ---------------------------------------------------------------------------------------------
 * blnd::simple_blend (src/math/_blending.cpp:83-153) and blnd::no_blend (:157-182): the SIMPLE_BLEND and NO_BLEND
 * branches of stitch_parameters::blend (src/classes/_panorama.cpp:220-240).
 * simple_blend: per image, in order: alpha = normalize(distanceTransform(mask), 0, 1, NORM_MINMAX);
 *   color += (img/255 * alpha) * (1 - acc_alpha);  acc_alpha += alpha * (1 - acc_alpha);
 *   result = acc_alpha > 0 ? color * (1.f / acc_alpha) : 0;  convertTo(CV_8UC3, 255).
 * cv::normalize(NORM_MINMAX): scale = 1/(max - min) in double (0 when max - min <= DBL_EPSILON), shift = -min*scale,
 * then convertTo(float(scale), float(shift)).  Tiles/masks contiguous; out = canvas_w x canvas_h x 3 bytes.
 * --------------------------------------------------------------------------------------------- */
void orc_simple_blend(int n, const uint8_t *const *tiles, const uint8_t *const *masks, const int *tl_x, const int *tl_y,
                      const int *w, const int *h, uint8_t *out)
{
    int dim[6];
    orc_pan_dimension(n, tl_x, tl_y, w, h, dim);
    const int W = dim[0], H = dim[1], mx = dim[2], my = dim[3];
    float *col = (float *)calloc((size_t)W * H * 3, sizeof(float));
    float *alp = (float *)calloc((size_t)W * H, sizeof(float));
    for (int i = 0; i < n; ++i) {
        const size_t np = (size_t)w[i] * h[i];
        float *dt = (float *)malloc(np * sizeof(float));
        orc_distance_transform(masks[i], w[i], h[i], (size_t)w[i], dt);
        float smin = dt[0], smax = dt[0];
        for (size_t k = 1; k < np; ++k) { if (dt[k] < smin) smin = dt[k]; if (dt[k] > smax) smax = dt[k]; }
        const double range = (double)smax - (double)smin;
        const double scale = range > DBL_EPSILON ? 1.0 / range : 0.0;
        const float a = (float)scale, b = (float)(0.0 - (double)smin * scale);
        const float inv255 = (float)(1.0 / 255.0);
        for (int y = 0; y < h[i]; ++y)
            for (int x = 0; x < w[i]; ++x) {
                const size_t k = (size_t)y * w[i] + x;
                const size_t c = (size_t)(y + tl_y[i] - my) * W + (x + tl_x[i] - mx);
                const float m = dt[k] * a + b;
                const float om = 1.0f - alp[c];
                for (int ch = 0; ch < 3; ++ch) {
                    const float v = (float)tiles[i][k * 3 + ch] * inv255;
                    col[c * 3 + ch] += (v * m) * om;
                }
                alp[c] += m * om;
            }
        free(dt);
    }
    for (size_t c = 0; c < (size_t)W * H; ++c) {
        const float a = alp[c];
        for (int ch = 0; ch < 3; ++ch) {
            float v = 0.f;
            if (a > 0) v = col[c * 3 + ch] * (1.f / a);
            double r = nearbyint((double)(v * 255.0f));
            out[c * 3 + ch] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
    }
    free(col);
    free(alp);
}

void orc_no_blend(int n, const uint8_t *const *tiles, const uint8_t *const *masks, const int *tl_x, const int *tl_y,
                  const int *w, const int *h, uint8_t *out)
{
    int dim[6];
    orc_pan_dimension(n, tl_x, tl_y, w, h, dim);
    const int W = dim[0], H = dim[1], mx = dim[2], my = dim[3];
    memset(out, 0, (size_t)W * H * 3);
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < h[i]; ++y)
            for (int x = 0; x < w[i]; ++x) {
                const size_t k = (size_t)y * w[i] + x;
                if (!masks[i][k]) continue;   /* Mat::copyTo(dst, mask): where mask != 0 */
                const size_t c = (size_t)(y + tl_y[i] - my) * W + (x + tl_x[i] - mx);
                out[c * 3] = tiles[i][k * 3]; out[c * 3 + 1] = tiles[i][k * 3 + 1]; out[c * 3 + 2] = tiles[i][k * 3 + 2];
            }
}

/* ---------------------------------------------------------------------------------------------
 * gain::get_overlapp_intensity (src/math/_gain_compensation.cpp:7-75): for every pair i <= j with (adj + I)(i,j) > 0:
 * area = countNonZero(mask_i & mask_j) over the overlap rectangle, I_i / I_j = sums of the 8-bit gray images there;
 * masks = createSurroundingMask(img, true, 1) (no erosion).  out[k*5] = {i, j, area, I_i, I_j}; returns the count.
 * --------------------------------------------------------------------------------------------- */
int orc_overlap_intensity(int n, const uint8_t *const *tiles, const int *tl_x, const int *tl_y, const int *w, const int *h,
                          const double *adj, double *out)
{
    uint8_t **gray = (uint8_t **)malloc((size_t)n * sizeof(uint8_t *)), **mask = (uint8_t **)malloc((size_t)n * sizeof(uint8_t *));
    for (int i = 0; i < n; ++i) {
        gray[i] = (uint8_t *)malloc((size_t)w[i] * h[i]);
        mask[i] = (uint8_t *)malloc((size_t)w[i] * h[i]);
        orc_gray_u8(tiles[i], w[i], h[i], (size_t)w[i] * 3, gray[i]);
        orc_surrounding_mask(tiles[i], w[i], h[i], (size_t)w[i] * 3, 0, mask[i]);
    }
    int count = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            if (!(adj[(size_t)i * n + j] + (i == j ? 1.0 : 0.0) > 0)) continue;
            double area = 0, si = 0, sj = 0;
            const int x0 = tl_x[i] > tl_x[j] ? tl_x[i] : tl_x[j], y0 = tl_y[i] > tl_y[j] ? tl_y[i] : tl_y[j];
            const int xa = tl_x[i] + w[i], xb = tl_x[j] + w[j], ya = tl_y[i] + h[i], yb = tl_y[j] + h[j];
            const int x1 = xa < xb ? xa : xb, y1 = ya < yb ? ya : yb;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    const size_t ki = (size_t)(y - tl_y[i]) * w[i] + (x - tl_x[i]), kj = (size_t)(y - tl_y[j]) * w[j] + (x - tl_x[j]);
                    if (mask[i][ki] & mask[j][kj]) { area += 1; si += gray[i][ki]; sj += gray[j][kj]; }
                }
            out[count * 5] = i; out[count * 5 + 1] = j; out[count * 5 + 2] = area; out[count * 5 + 3] = si; out[count * 5 + 4] = sj;
            ++count;
        }
    for (int i = 0; i < n; ++i) { free(gray[i]); free(mask[i]); }
    free(gray); free(mask);
    return count;
}
