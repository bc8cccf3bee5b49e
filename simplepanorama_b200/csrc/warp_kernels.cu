// warp_kernels.cu -- fused inverse-projection remap + bilinear sampling + dark flag + exposure gain (sm_100a).
//
// Replaces, per warped tile, the reference's
//   cv::detail::{Spherical,Cylindrical,Stereographic}Warper::warp   (src/math/_projection.cpp:51,81,321)
//     = buildMaps (projector mapBackward per destination pixel) + cv::remap(INTER_LINEAR, BORDER_CONSTANT)
//   the gray<=1 test at the head of blnd::createSurroundingMask      (src/math/_blending.cpp:283-287)
//   `imgs[i] / gain[i]` on CV_8UC3                                   (src/classes/_panorama.cpp:321-327)
// in ONE pass: the float maps are never materialised, the warped tile is written once.
//
// Arithmetic contract (what makes the result match OpenCV):
//   * mapBackward in float32 with IEEE mul/add/div (no FMA contraction: __fmul_rn/__fadd_rn/__fdiv_rn)
//     and sinf/cosf/atan2f/atanf evaluated with the host libm's own arithmetic (glibc_trig.cuh: bit-identical
//     to glibc for every float, checked exhaustively), so the maps equal cv2's buildMaps bit for bit.
//   * sampling is OpenCV's 8-bit fixed point: sx = cvRound(32 x), weights (32-fy)(32-fx)*32,
//     D = (sum + 2^14) >> 15.  Evaluated separably with exact integers:
//     h = (32-fx) p0 + fx p1 (one dp4a per channel and row), D = ((32-fy) h_top + fy h_bot + 512) >> 10.
//   * gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15 on the UN-gained sample; dark = gray <= 1.
//   * gain: rint(float(v) * float(1/g)) saturated (cv::Mat::convertTo with alpha).
//
// HBM traffic per tile pixel: 3 B source read (each source pixel is touched ~once, neighbours hit
// L1/L2) + 3 B tile write + 1 B dark-flag write = 7 B.
// For spherical / cylindrical the trigonometry is separable (u depends on the column, v on the
// row): sin/cos are evaluated once per tile column / row into small tables, so the per-pixel work
// is 9 mul/add + 2 div + the integer sampler.
#include <cuda.h>   // CUtensorMap and its enums only: the encode entry point is looked up through the runtime
#include "spano_internal.h"
#include "glibc_trig.cuh"

namespace {

constexpr float kPiF = 3.14159265358979323846f; // (float)CV_PI

struct WarpParams {
    float m[9]; // k_rinv
    float scale;
    int tl_x, tl_y;
    int dst_w, dst_h;
    int row_begin, row_end;
    const uint8_t *src;
    int src_w, src_h;
    size_t src_step;
    uint8_t *dst;
    size_t dst_step;
    uint8_t *dark;
    size_t dark_step;
    float inv_gain;
    int apply_gain;
    int src_aligned8; // src pointer and step are multiples of 4: aligned 32-bit window loads allowed
    const float *col, *row;   // tables (see warp_tables_kernel)
    int wp;                   // column-array pitch
    SpanoScatter sc; // n > 0: the tile rows go to these (possibly peer-GPU) slices instead of `dst`
};

// Per-column / per-row tables.  The projector's trigonometry is separable (u depends on the column, v on the
// row), and so are most of the products of  (x,y,z) = k_rinv * (x_,y_,z_)  when they are rounded one by one as
// OpenCV's scalar code does (no contraction):
//   cylindrical: x_ = sin u, y_ = v, z_ = cos u      ->  x = (m0 x_ + m1 y_) + m2 z_  with every product a pure
//                column or row quantity: 6 column arrays (m0 sin u, m2 cos u, m3.., m5.., m6.., m8..) and 3 row
//                arrays (m1 v, m4 v, m7 v); 6 FADD per pixel.
//   spherical:   x_ = sin(pi-v) sin u, y_ = cos(pi-v), z_ = sin(pi-v) cos u: column arrays sin u, cos u; row arrays
//                sin(pi-v), m1 y_, m4 y_, m7 y_.
// Layout: NCOL column arrays of pitch wp = align4(w) floats (float4 loads of a thread's 4 columns), then NROW row
// arrays of pitch h.
template <int KIND> struct TableShape { };
template <> struct TableShape<SPANO_CYLINDRICAL> { static constexpr int NCOL = 6, NROW = 3; };
template <> struct TableShape<SPANO_SPHERICAL> { static constexpr int NCOL = 2, NROW = 4; };
template <> struct TableShape<SPANO_STEREOGRAPHIC> { static constexpr int NCOL = 0, NROW = 0; };

struct Mat9 { float m[9]; };

template <int KIND>
__global__ void warp_tables_kernel(float scale, int tl_x, int tl_y, int w, int h, int wp, Mat9 M, float *col, float *row)
{
    const float *m = M.m;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w) {
        const float u = __fdiv_rn((float)(i + tl_x), scale);
        const float su = gtrig::sinf_glibc(u), cu = gtrig::cosf_glibc(u);
        if (KIND == SPANO_CYLINDRICAL) {
            col[i] = __fmul_rn(m[0], su);           col[wp + i] = __fmul_rn(m[2], cu);
            col[2 * wp + i] = __fmul_rn(m[3], su);  col[3 * wp + i] = __fmul_rn(m[5], cu);
            col[4 * wp + i] = __fmul_rn(m[6], su);  col[5 * wp + i] = __fmul_rn(m[8], cu);
        } else {
            col[i] = su;
            col[wp + i] = cu;
        }
    }
    if (i < h) {
        const float v = __fdiv_rn((float)(i + tl_y), scale);
        if (KIND == SPANO_CYLINDRICAL) {
            row[i] = __fmul_rn(m[1], v);  row[h + i] = __fmul_rn(m[4], v);  row[2 * h + i] = __fmul_rn(m[7], v);
        } else {
            const float t = __fsub_rn(kPiF, v);
            const float y_ = gtrig::cosf_glibc(t);
            row[i] = gtrig::sinf_glibc(t);
            row[h + i] = __fmul_rn(m[1], y_);  row[2 * h + i] = __fmul_rn(m[4], y_);  row[3 * h + i] = __fmul_rn(m[7], y_);
        }
    }
}

// cvRound(32*v) with x86 cvtss2si semantics: INT_MIN for NaN / out of range
__device__ __forceinline__ int fixed_coord(float v)
{
    float s = v * 32.f;
    if (!(s > -2147483648.f && s < 2147483648.f)) return (int)0x80000000;
    return __float2int_rn(s);
}

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

// Bilinear sample of one destination pixel -> packed 0x00RRGGBB (un-gained).
__device__ __forceinline__ uint32_t sample_bilinear(const WarpParams &P, float x, float y)
{
    // cvRound(32 x) is INT_MIN on x86 for NaN / |32 x| >= 2^31, which lands outside every image:
    // any such coordinate makes the whole sample the border constant
    if (!(fabsf(x) < 67108864.f) || !(fabsf(y) < 67108864.f)) return 0u;
    const int fx = __float2int_rn(x * 32.f), fy = __float2int_rn(y * 32.f);
    // (OpenCV saturates the integer part to short; images are < 32767 px, so saturated and unsaturated
    // values fail the same range tests below)
    const int sx = fx >> 5, sy = fy >> 5;
    const int ax = fx & 31, ay = fy & 31;
    const uint32_t wx = (uint32_t)(32 - ax) | ((uint32_t)ax << 24); // bytes {32-ax,0,0,ax}
    uint32_t hb[2], hg[2], hr[2];
    if ((unsigned)sx < (unsigned)(P.src_w - 1) && (unsigned)sy < (unsigned)(P.src_h - 1)) {
        const int o = 3 * sx;
        if (P.src_aligned8) {
            // the 6 bytes {B0 G0 R0 B1 G1 R1} start at byte t of three aligned 32-bit words; straight-line code:
            // two funnel shifts align bytes o..o+7, two more give the G- and R-phased words
            const int t8 = (o & 3) * 8;
            const uint32_t *w0 = reinterpret_cast<const uint32_t *>(P.src + (size_t)sy * P.src_step + (o & ~3));
            const uint32_t *w1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(w0) + P.src_step);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t *w = r ? w1 : w0;
                const uint32_t a = __ldg(w), b = __ldg(w + 1);
                const uint32_t c = (t8 == 24) ? __ldg(w + 2) : 0u;   // third word only when the window crosses into it
                const uint32_t lo = __funnelshift_r(a, b, t8), hi = __funnelshift_r(b, c, t8);
                hb[r] = __dp4a(lo, wx, 0u);
                hg[r] = __dp4a(__funnelshift_r(lo, hi, 8), wx, 0u);
                hr[r] = __dp4a(__funnelshift_r(lo, hi, 16), wx, 0u);
            }
        } else {
            const uint8_t *row = P.src + (size_t)sy * P.src_step;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint64_t win = 0;
#pragma unroll
                for (int b = 0; b < 6; ++b) win |= (uint64_t)__ldg(row + o + b) << (8 * b);
                hb[r] = __dp4a((uint32_t)win, wx, 0u);
                hg[r] = __dp4a((uint32_t)(win >> 8), wx, 0u);
                hr[r] = __dp4a((uint32_t)(win >> 16), wx, 0u);
                row += P.src_step;
            }
        }
    } else {
        // BORDER_CONSTANT(0): every tap outside the source contributes 0
        if (sx >= P.src_w || sx + 1 < 0 || sy >= P.src_h || sy + 1 < 0) return 0u;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int yy = sy + r;
            uint32_t p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
            if (yy >= 0 && yy < P.src_h) {
                const uint8_t *row = P.src + (size_t)yy * P.src_step;
                if (sx >= 0 && sx < P.src_w) {
                    p0[0] = __ldg(row + 3 * sx); p0[1] = __ldg(row + 3 * sx + 1); p0[2] = __ldg(row + 3 * sx + 2);
                }
                if (sx + 1 >= 0 && sx + 1 < P.src_w) {
                    p1[0] = __ldg(row + 3 * sx + 3); p1[1] = __ldg(row + 3 * sx + 4); p1[2] = __ldg(row + 3 * sx + 5);
                }
            }
            hb[r] = (32 - ax) * p0[0] + ax * p1[0];
            hg[r] = (32 - ax) * p0[1] + ax * p1[1];
            hr[r] = (32 - ax) * p0[2] + ax * p1[2];
        }
    }
    const uint32_t wy0 = 32 - ay, wy1 = ay;
    const uint32_t B = (wy0 * hb[0] + wy1 * hb[1] + 512u) >> 10;
    const uint32_t G = (wy0 * hg[0] + wy1 * hg[1] + 512u) >> 10;
    const uint32_t R = (wy0 * hr[0] + wy1 * hr[1] + 512u) >> 10;
    return B | (G << 8) | (R << 16);
}

// ---- TMA-staged variant ---------------------------------------------------------------------------------------------
// The source footprint of a block of BLK_W x BLK_H destination pixels is a small rectangle of the source image (the
// projections are smooth and the panorama scale equals the focal length, so the footprint is about the size of the
// block).  One thread asks the TMA unit for a fixed-size box around it (cp.async.bulk.tensor.2d, source viewed as a 2-D
// tensor of 32-bit words, out-of-bounds words zero-filled) and the block samples from shared memory: 32-bit shared
// addresses instead of 64-bit global ones, no L1 tag traffic, one bulk request per block instead of ~8 k loads.
// Correctness never depends on the box: a tap outside it (or outside the image interior) takes the global path.
constexpr int BLK_W = 64, BLK_H = 16;             // destination pixels per block (256 threads x 4 px)
constexpr int BOX_W = 72, BOX_H = 32;             // staged box: 72 words (288 B = 96 px) x 32 rows = 9216 B
constexpr int BOX_PITCH = BOX_W * 4;

struct alignas(64) TmaDesc { unsigned long long opaque[16]; };
static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "tensor map size");

struct StagedBox {
    const uint8_t *smem;   // the box (nullptr: nothing staged for this block)
    int byte0;             // source byte offset (within a row) of the box's first byte
    int row0;              // source row of the box's first row
};

// interior sample (all four taps inside the image) from the staged box; false: the taps are not all inside the box
__device__ __forceinline__ bool sample_from_box(const WarpParams &P, const StagedBox &S, float x, float y, uint32_t &out)
{
    if (!(fabsf(x) < 67108864.f) || !(fabsf(y) < 67108864.f)) { out = 0u; return true; }
    const int fx = __float2int_rn(x * 32.f), fy = __float2int_rn(y * 32.f);
    const int sx = fx >> 5, sy = fy >> 5;
    if (!((unsigned)sx < (unsigned)(P.src_w - 1) && (unsigned)sy < (unsigned)(P.src_h - 1))) return false;
    const int ob = 3 * sx - S.byte0, rb = sy - S.row0;
    if (!((unsigned)ob <= (unsigned)(BOX_PITCH - 12) && (unsigned)rb < (unsigned)(BOX_H - 1))) return false;
    const int ax = fx & 31, ay = fy & 31;
    const uint32_t wx = (uint32_t)(32 - ax) | ((uint32_t)ax << 24);
    const int t8 = (ob & 3) * 8;
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(S.smem + rb * BOX_PITCH + (ob & ~3));
    uint32_t hb[2], hg[2], hr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t *w = w0 + r * BOX_W;
        const uint32_t a = w[0], b = w[1], c = w[2];
        const uint32_t lo = __funnelshift_r(a, b, t8), hi = __funnelshift_r(b, c, t8);
        hb[r] = __dp4a(lo, wx, 0u);
        hg[r] = __dp4a(__funnelshift_r(lo, hi, 8), wx, 0u);
        hr[r] = __dp4a(__funnelshift_r(lo, hi, 16), wx, 0u);
    }
    const uint32_t wy0 = 32 - ay, wy1 = ay;
    const uint32_t B = (wy0 * hb[0] + wy1 * hb[1] + 512u) >> 10;
    const uint32_t G = (wy0 * hg[0] + wy1 * hg[1] + 512u) >> 10;
    const uint32_t R = (wy0 * hr[0] + wy1 * hr[1] + 512u) >> 10;
    out = B | (G << 8) | (R << 16);
    return true;
}

// gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15 (cv::cvtColor BGR2GRAY, 8 bit); dark = gray <= 1, i.e.
// 3735 B + 19235 G + 9798 R < 3 * 2^14.  The 16-bit weights are split into bytes so two dp4a do the sum.
__device__ __forceinline__ uint32_t is_dark(uint32_t bgr)
{
    constexpr uint32_t WLO = (3735u & 255u) | ((19235u & 255u) << 8) | ((9798u & 255u) << 16);
    constexpr uint32_t WHI = (3735u >> 8) | ((19235u >> 8) << 8) | ((9798u >> 8) << 16);
    const uint32_t sum = __dp4a(bgr, WLO, 0u) + (__dp4a(bgr, WHI, 0u) << 8);
    return sum < 49152u ? 1u : 0u;
}

__device__ __forceinline__ uint32_t gain_u8(uint32_t v, float a)
{
    int r = __float2int_rn(__fmul_rn((float)v, a));
    return (uint32_t)min(255, max(0, r));
}

__device__ __forceinline__ uint32_t gain_bgr(uint32_t bgr, float a)
{
    return gain_u8(bgr & 255u, a) | (gain_u8((bgr >> 8) & 255u, a) << 8) | (gain_u8((bgr >> 16) & 255u, a) << 16);
}

// mapBackward, OpenCV expression order, no contraction.  One pixel (stereographic, buildMaps) ...
template <int KIND>
__device__ __forceinline__ void project_ray(const WarpParams &P, int u_i, int v_i, float &X, float &Y, float &Z)
{
    const float *m = P.m;
    if (KIND == SPANO_CYLINDRICAL) {
        const float *c = P.col + u_i, *r = P.row + v_i;
        X = __fadd_rn(__fadd_rn(__ldg(c), __ldg(r)), __ldg(c + P.wp));
        Y = __fadd_rn(__fadd_rn(__ldg(c + 2 * P.wp), __ldg(r + P.dst_h)), __ldg(c + 3 * P.wp));
        Z = __fadd_rn(__fadd_rn(__ldg(c + 4 * P.wp), __ldg(r + 2 * P.dst_h)), __ldg(c + 5 * P.wp));
        return;
    }
    float x_, z_, my1, my4, my7;
    if (KIND == SPANO_SPHERICAL) {
        const float *r = P.row + v_i;
        const float sinv = __ldg(r);
        x_ = __fmul_rn(sinv, __ldg(P.col + u_i));
        z_ = __fmul_rn(sinv, __ldg(P.col + P.wp + u_i));
        my1 = __ldg(r + P.dst_h);  my4 = __ldg(r + 2 * P.dst_h);  my7 = __ldg(r + 3 * P.dst_h);
    } else {
        const float u = __fdiv_rn((float)(u_i + P.tl_x), P.scale);
        const float v = __fdiv_rn((float)(v_i + P.tl_y), P.scale);
        const float az = gtrig::atan2f_glibc(v, u);
        const float r = sqrtf(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
        const float pol = __fmul_rn(2.f, gtrig::atanf_glibc(__fdiv_rn(1.f, r)));
        const float t = __fsub_rn(kPiF, pol);
        const float sinv = gtrig::sinf_glibc(t);
        x_ = __fmul_rn(sinv, gtrig::sinf_glibc(az));
        const float y_ = gtrig::cosf_glibc(t);
        z_ = __fmul_rn(sinv, gtrig::cosf_glibc(az));
        my1 = __fmul_rn(m[1], y_);  my4 = __fmul_rn(m[4], y_);  my7 = __fmul_rn(m[7], y_);
    }
    X = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x_), my1), __fmul_rn(m[2], z_));
    Y = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x_), my4), __fmul_rn(m[5], z_));
    Z = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x_), my7), __fmul_rn(m[8], z_));
}

__device__ __forceinline__ void perspective(float X, float Y, float Z, float &x, float &y)
{
    if (Z > 0.f) {
        x = __fdiv_rn(X, Z);
        y = __fdiv_rn(Y, Z);
    } else {
        x = y = -1.f;
    }
}

template <int KIND>
__device__ __forceinline__ void map_backward(const WarpParams &P, int u_i, int v_i, float &x, float &y)
{
    float X, Y, Z;
    project_ray<KIND>(P, u_i, v_i, X, Y, Z);
    perspective(X, Y, Z, x, y);
}

// ... and the 4 consecutive pixels of one warp-kernel thread (x0 is a multiple of 4: float4 table loads; the
// column arrays are padded to a multiple of 4, pixels beyond dst_w are computed on padding and dropped)
template <int KIND>
__device__ __forceinline__ void map_backward4(const WarpParams &P, int x0, int v_i, float (&x)[4], float (&y)[4])
{
    if (KIND == SPANO_STEREOGRAPHIC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) map_backward<KIND>(P, x0 + i, v_i, x[i], y[i]);
        return;
    }
    const float *m = P.m;
    const float *r = P.row + v_i;
    float X[4], Y[4], Z[4];
    if (KIND == SPANO_CYLINDRICAL) {
        const float4 *c = reinterpret_cast<const float4 *>(P.col + x0);
        const int q = P.wp >> 2;
        const float4 a0 = __ldg(c), a2 = __ldg(c + q), a3 = __ldg(c + 2 * q), a5 = __ldg(c + 3 * q), a6 = __ldg(c + 4 * q), a8 = __ldg(c + 5 * q);
        const float r1 = __ldg(r), r4 = __ldg(r + P.dst_h), r7 = __ldg(r + 2 * P.dst_h);
        const float A0[4] = {a0.x, a0.y, a0.z, a0.w}, A2[4] = {a2.x, a2.y, a2.z, a2.w}, A3[4] = {a3.x, a3.y, a3.z, a3.w};
        const float A5[4] = {a5.x, a5.y, a5.z, a5.w}, A6[4] = {a6.x, a6.y, a6.z, a6.w}, A8[4] = {a8.x, a8.y, a8.z, a8.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            X[i] = __fadd_rn(__fadd_rn(A0[i], r1), A2[i]);
            Y[i] = __fadd_rn(__fadd_rn(A3[i], r4), A5[i]);
            Z[i] = __fadd_rn(__fadd_rn(A6[i], r7), A8[i]);
        }
    } else {
        const float4 su = __ldg(reinterpret_cast<const float4 *>(P.col + x0));
        const float4 cu = __ldg(reinterpret_cast<const float4 *>(P.col + P.wp + x0));
        const float sinv = __ldg(r), my1 = __ldg(r + P.dst_h), my4 = __ldg(r + 2 * P.dst_h), my7 = __ldg(r + 3 * P.dst_h);
        const float SU[4] = {su.x, su.y, su.z, su.w}, CU[4] = {cu.x, cu.y, cu.z, cu.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x_ = __fmul_rn(sinv, SU[i]), z_ = __fmul_rn(sinv, CU[i]);
            X[i] = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x_), my1), __fmul_rn(m[2], z_));
            Y[i] = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x_), my4), __fmul_rn(m[5], z_));
            Z[i] = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x_), my7), __fmul_rn(m[8], z_));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) perspective(X[i], Y[i], Z[i], x[i], y[i]);
}

constexpr int WARP_PX_PER_THREAD = 4;
// (block shape: 4, 8 and 12 warps per CTA measured the same, alone and next to a blend CTA of the fused path, where an SM has
// room for ONE such CTA -- 16 K registers and ~1.4 KB of shared memory are free beside blend_ws_kernel)
constexpr int WARP_BLOCK_X = 32, WARP_BLOCK_Y = 8;

// 4 consecutive destination pixels x0..x0+3 of tile row v: coordinates, sampling (from the staged box when there is one),
// dark flags, gain, stores (local tile or the band slices of the tile-sharded path)
template <int KIND>
__device__ __forceinline__ void warp_four_pixels(const WarpParams &P, const uint8_t *s_gain, const StagedBox &S, int x0, int v)
{
    uint32_t px[WARP_PX_PER_THREAD];
    uint32_t dark = 0;
    float mx[WARP_PX_PER_THREAD], my[WARP_PX_PER_THREAD];
    map_backward4<KIND>(P, x0, v, mx, my);
#pragma unroll
    for (int i = 0; i < WARP_PX_PER_THREAD; ++i) {
        const int u = x0 + i;
        uint32_t s = 0, d = 1;
        if (u < P.dst_w) {
            if (!(S.smem && sample_from_box(P, S, mx[i], my[i], s))) s = sample_bilinear(P, mx[i], my[i]);
            d = is_dark(s);
            s = (uint32_t)s_gain[s & 255u] | ((uint32_t)s_gain[(s >> 8) & 255u] << 8) | ((uint32_t)s_gain[(s >> 16) & 255u] << 16);
        }
        px[i] = s;
        dark |= d << (8 * i);
    }
    const bool full = x0 + WARP_PX_PER_THREAD <= P.dst_w;
    auto store_row = [&](uint8_t *drow) {
        if (full && (((uintptr_t)drow) & 3) == 0) {
            // 4 px = 12 B = three 32-bit words; a warp writes 384 contiguous bytes
            uint32_t *q = reinterpret_cast<uint32_t *>(drow);
            q[0] = px[0] | (px[1] << 24);
            q[1] = (px[1] >> 8) | (px[2] << 16);
            q[2] = (px[2] >> 16) | (px[3] << 8);
        } else {
            for (int i = 0; i < WARP_PX_PER_THREAD && x0 + i < P.dst_w; ++i) {
                drow[3 * i] = (uint8_t)px[i];
                drow[3 * i + 1] = (uint8_t)(px[i] >> 8);
                drow[3 * i + 2] = (uint8_t)(px[i] >> 16);
            }
        }
    };
    if (P.sc.n > 0) {
        // tile-sharded multi-GPU path: the row goes to every band slice that reads it (its band plus the blur
        // halo of the neighbours), straight into the owning GPU's memory; the row index is warp-uniform
        for (int d = 0; d < P.sc.n; ++d)
            if (v >= P.sc.row0[d] && v < P.sc.row1[d] && x0 >= P.sc.col0[d] && x0 < P.sc.col1[d])
                store_row(P.sc.base[d] + (size_t)v * P.sc.step[d] + (size_t)x0 * 3);
    } else if (P.dst) {
        store_row(P.dst + (size_t)v * P.dst_step + (size_t)x0 * 3);
    }   // else: flags-only pass (validity masks computed on another rank's behalf), no tile store
    if (P.dark) {
        uint8_t *krow = P.dark + (size_t)v * P.dark_step + x0;
        if (full && (((uintptr_t)krow) & 3) == 0) *reinterpret_cast<uint32_t *>(krow) = dark;
        else
            for (int i = 0; i < WARP_PX_PER_THREAD && x0 + i < P.dst_w; ++i) krow[i] = (uint8_t)(dark >> (8 * i));
    }
}

template <int KIND>
__global__ void __launch_bounds__(WARP_BLOCK_X *WARP_BLOCK_Y) warp_kernel(const WarpParams P)
{
    // 8-bit gain as a 256-entry table (one shared-memory load per channel instead of convert/multiply/round/clamp)
    __shared__ uint8_t s_gain[256];
    {
        const uint32_t t = threadIdx.y * WARP_BLOCK_X + threadIdx.x;
        if (t < 256) s_gain[t] = (uint8_t)(P.apply_gain ? gain_u8(t, P.inv_gain) : t);
    }
    __syncthreads();
    const int x0 = (blockIdx.x * WARP_BLOCK_X + threadIdx.x) * WARP_PX_PER_THREAD;
    const int v = P.row_begin + blockIdx.y * WARP_BLOCK_Y + threadIdx.y;
    if (x0 >= P.dst_w || v >= P.row_end) return;
    const StagedBox none = {nullptr, 0, 0};
    warp_four_pixels<KIND>(P, s_gain, none, x0, v);
}

// The TMA-staged kernel: block = BLK_W x BLK_H destination pixels, 256 threads (16 x 16, 4 px each).
template <int KIND>
__global__ void __launch_bounds__(256) warp_tma_kernel(const WarpParams P, const __grid_constant__ TmaDesc tmap)
{
    __shared__ __align__(128) uint8_t s_box[BOX_H * BOX_PITCH];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint8_t s_gain[256];
    __shared__ int s_meta[3];          // staged?, byte0, row0
    const int tid = threadIdx.x;
    s_gain[tid] = (uint8_t)(P.apply_gain ? gain_u8(tid, P.inv_gain) : tid);
    const int bx0 = blockIdx.x * BLK_W, by0 = P.row_begin + blockIdx.y * BLK_H;
    if (tid < 32) {
        // footprint of the block from 9 probes (corners, edge midpoints, centre) of mapBackward
        float lox = 3.0e38f, loy = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f;
        bool bad = false;
        if (tid < 9) {
            const int pu = min(P.dst_w - 1, bx0 + (tid % 3) * (BLK_W / 2) - ((tid % 3) == 2 ? 1 : 0));
            const int pv = min(P.row_end - 1, by0 + (tid / 3) * (BLK_H / 2) - ((tid / 3) == 2 ? 1 : 0));
            float x, y;
            map_backward<KIND>(P, pu, pv, x, y);
            bad = !(fabsf(x) < 1.0e6f) || !(fabsf(y) < 1.0e6f) || (x == -1.f && y == -1.f);
            lox = hix = x;  loy = hiy = y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));  loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
            hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));  hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        }
        bad = __any_sync(0xffffffffu, bad);
        if (tid == 0) {
            int staged = 0, byte0 = 0, row0 = 0;
            if (!bad) {
                const int px0 = (int)floorf(lox) - 2, px1 = (int)ceilf(hix) + 3;      // pixels px0 .. px1 (taps included)
                row0 = (int)floorf(loy) - 2;
                const int rows = (int)ceilf(hiy) + 3 - row0 + 1;
                const int w0 = ((3 * px0) >> 4) << 2;   // first word: the box must start on a 16-byte boundary of the row (floor, also for negatives)
                byte0 = 4 * w0;
                if (rows <= BOX_H && 3 * (px1 + 1) - byte0 <= BOX_PITCH - 8) staged = 1;
                if (staged) {
                    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar), dst = (uint32_t)__cvta_generic_to_shared(s_box);
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(BOX_H * BOX_PITCH) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(w0), "r"(row0), "r"(bar) : "memory");
                }
            }
            s_meta[0] = staged;  s_meta[1] = byte0;  s_meta[2] = row0;
        }
    }
    __syncthreads();
    StagedBox S = {nullptr, s_meta[1], s_meta[2]};
    if (s_meta[0]) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "WAITW_%=:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
                     "@p bra DONEW_%=;\n"
                     "bra WAITW_%=;\n"
                     "DONEW_%=:\n"
                     "}" ::"r"(bar) : "memory");
        S.smem = s_box;
    }
    const int x0 = bx0 + (tid & 15) * WARP_PX_PER_THREAD;
    const int v = by0 + (tid >> 4);
    if (x0 >= P.dst_w || v >= P.row_end) return;
    warp_four_pixels<KIND>(P, s_gain, S, x0, v);
}

// buildMaps alone (float maps, as cv::detail::RotationWarperBase::buildMaps returns them)
template <int KIND>
__global__ void build_maps_kernel(const WarpParams P, float *xmap, float *ymap)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (u >= P.dst_w || v >= P.dst_h) return;
    float x, y;
    map_backward<KIND>(P, u, v, x, y);
    xmap[(size_t)v * P.dst_w + u] = x;
    ymap[(size_t)v * P.dst_w + u] = y;
}

// cv::remap(INTER_LINEAR, BORDER_CONSTANT) on explicit float maps
__global__ void remap_kernel(const WarpParams P, const float *xmap, const float *ymap)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (u >= P.dst_w || v >= P.dst_h) return;
    const uint32_t s = sample_bilinear(P, xmap[(size_t)v * P.dst_w + u], ymap[(size_t)v * P.dst_w + u]);
    uint8_t *d = P.dst + (size_t)v * P.dst_step + (size_t)u * 3;
    d[0] = (uint8_t)s; d[1] = (uint8_t)(s >> 8); d[2] = (uint8_t)(s >> 16);
}

__global__ void gain_kernel(uint8_t *img, int w3, int h, size_t step, float a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w3 || y >= h) return;
    uint8_t *p = img + (size_t)y * step + x;
    *p = (uint8_t)gain_u8(*p, a);
}

__global__ void dark_flags_kernel(const uint8_t *bgr, int w, int h, size_t step, uint8_t *dark, size_t dark_step)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = bgr + (size_t)y * step + (size_t)x * 3;
    const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    dark[(size_t)y * dark_step + x] = (uint8_t)is_dark(v);
}

} // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda):
// the source image as a 2-D tensor of 32-bit words, box BOX_W x BOX_H, no swizzle, out-of-bounds words read as zero
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool encode_source_tensor_map(TmaDesc *out, const uint8_t *src, int src_w, int src_h, size_t src_step)
{
    static encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<encode_tiled_fn>(p);
    }();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)(((size_t)src_w * 3 + 3) / 4), (cuuint64_t)src_h};
    const cuuint64_t strides[1] = {(cuuint64_t)src_step};
    const cuuint32_t box[2] = {BOX_W, BOX_H}, estr[2] = {1, 1};
    return fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(src), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_remap(spano_ctx *ctx, const uint8_t *src, int src_w, int src_h, size_t src_step, const float *xmap,
                 const float *ymap, int dst_w, int dst_h, uint8_t *dst, size_t dst_step)
{
    WarpParams P = {};
    P.dst_w = dst_w;  P.dst_h = dst_h;
    P.src = src;  P.src_w = src_w;  P.src_h = src_h;  P.src_step = src_step;
    P.dst = dst;  P.dst_step = dst_step;
    P.src_aligned8 = ((((uintptr_t)src) | src_step) & 3) == 0;
    dim3 block(256), grid((dst_w + 255) / 256, dst_h);
    remap_kernel<<<grid, block, 0, ctx->stream>>>(P, xmap, ymap);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_warp(spano_ctx *ctx, const SpanoProjector &proj, const uint8_t *src, int src_w, int src_h, size_t src_step,
                double gain, int tl_x, int tl_y, int dst_w, int dst_h, int row_begin, int row_end, uint8_t *dst,
                size_t dst_step, uint8_t *dark, size_t dark_step, float *xmap, float *ymap, const SpanoScatter *scatter)
{
    if (row_end <= row_begin) return 0;
    int launches = 0;
    float *tables = nullptr;
    const int wp = (dst_w + 3) & ~3;
    const size_t tab_floats = 6 * (size_t)wp + 4 * (size_t)dst_h;
    int rc = spano_reserve(ctx, spano_table_buffer(ctx, spano_ctx::BUF_TABLES, spano_ctx::BUF_TABLES_AUX), tab_floats * sizeof(float), (void **)&tables);
    if (rc) return rc;
    WarpParams P;
    for (int i = 0; i < 9; ++i) P.m[i] = proj.k_rinv[i];
    P.scale = proj.scale;
    P.tl_x = tl_x;  P.tl_y = tl_y;
    P.dst_w = dst_w;  P.dst_h = dst_h;
    P.row_begin = row_begin;  P.row_end = row_end;
    P.src = src;  P.src_w = src_w;  P.src_h = src_h;  P.src_step = src_step;
    P.dst = dst;  P.dst_step = dst_step;
    P.dark = dark;  P.dark_step = dark_step;
    P.inv_gain = (float)(1.0 / gain);
    P.apply_gain = (gain != 1.0);
    P.src_aligned8 = ((((uintptr_t)src) | src_step) & 3) == 0;
    P.col = tables;
    P.row = tables + 6 * (size_t)wp;
    P.wp = wp;
    if (scatter) P.sc = *scatter;

    if (proj.kind != SPANO_STEREOGRAPHIC) {
        const int n = dst_w > dst_h ? dst_w : dst_h;
        const int tb = 256;
        Mat9 M;
        for (int i = 0; i < 9; ++i) M.m[i] = proj.k_rinv[i];
        if (proj.kind == SPANO_SPHERICAL)
            warp_tables_kernel<SPANO_SPHERICAL><<<(n + tb - 1) / tb, tb, 0, ctx->stream>>>(proj.scale, tl_x, tl_y, dst_w, dst_h, wp, M, tables, tables + 6 * (size_t)wp);
        else
            warp_tables_kernel<SPANO_CYLINDRICAL><<<(n + tb - 1) / tb, tb, 0, ctx->stream>>>(proj.scale, tl_x, tl_y, dst_w, dst_h, wp, M, tables, tables + 6 * (size_t)wp);
        ++launches;
    }
    if (xmap) { // maps only
        dim3 mb(256), mg((dst_w + 255) / 256, dst_h);
        switch (proj.kind) {
        case SPANO_SPHERICAL: build_maps_kernel<SPANO_SPHERICAL><<<mg, mb, 0, ctx->stream>>>(P, xmap, ymap); break;
        case SPANO_CYLINDRICAL: build_maps_kernel<SPANO_CYLINDRICAL><<<mg, mb, 0, ctx->stream>>>(P, xmap, ymap); break;
        default: build_maps_kernel<SPANO_STEREOGRAPHIC><<<mg, mb, 0, ctx->stream>>>(P, xmap, ymap); break;
        }
        ++launches;
        SPANO_CUDA(ctx, cudaGetLastError());
        ctx->launches += launches;
        return launches;
    }
    // TMA-staged kernel when the source can be described as a 2-D tensor of 32-bit words (16-byte aligned base and pitch)
    if (ctx->opt_warp_kernel == 1 && src && ((((uintptr_t)src) | src_step) & 15) == 0 && src_step >= (((size_t)src_w * 3 + 3) & ~(size_t)3)) {
        TmaDesc tm;
        if (encode_source_tensor_map(&tm, src, src_w, src_h, src_step)) {
            dim3 tb(256), tg((dst_w + BLK_W - 1) / BLK_W, (row_end - row_begin + BLK_H - 1) / BLK_H);
            switch (proj.kind) {
            case SPANO_SPHERICAL: warp_tma_kernel<SPANO_SPHERICAL><<<tg, tb, 0, ctx->stream>>>(P, tm); break;
            case SPANO_CYLINDRICAL: warp_tma_kernel<SPANO_CYLINDRICAL><<<tg, tb, 0, ctx->stream>>>(P, tm); break;
            default: warp_tma_kernel<SPANO_STEREOGRAPHIC><<<tg, tb, 0, ctx->stream>>>(P, tm); break;
            }
            ++launches;
            SPANO_CUDA(ctx, cudaGetLastError());
            ctx->launches += launches;
            return launches;
        }
    }
    dim3 block(WARP_BLOCK_X, WARP_BLOCK_Y);
    dim3 grid((dst_w + WARP_BLOCK_X * WARP_PX_PER_THREAD - 1) / (WARP_BLOCK_X * WARP_PX_PER_THREAD),
              (row_end - row_begin + WARP_BLOCK_Y - 1) / WARP_BLOCK_Y);
    switch (proj.kind) {
    case SPANO_SPHERICAL: warp_kernel<SPANO_SPHERICAL><<<grid, block, 0, ctx->stream>>>(P); break;
    case SPANO_CYLINDRICAL: warp_kernel<SPANO_CYLINDRICAL><<<grid, block, 0, ctx->stream>>>(P); break;
    default: warp_kernel<SPANO_STEREOGRAPHIC><<<grid, block, 0, ctx->stream>>>(P); break;
    }
    ++launches;
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += launches;
    return launches;
}

int launch_gain(spano_ctx *ctx, uint8_t *img, int w, int h, size_t step, double gain)
{
    if (gain == 1.0 || w <= 0 || h <= 0) return 0;
    dim3 block(256), grid((3 * w + 255) / 256, h);
    gain_kernel<<<grid, block, 0, ctx->stream>>>(img, 3 * w, h, step, (float)(1.0 / gain));
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}

int launch_dark_flags(spano_ctx *ctx, const uint8_t *bgr, int w, int h, size_t step, uint8_t *dark, size_t dark_step)
{
    dim3 block(256), grid((w + 255) / 256, h);
    dark_flags_kernel<<<grid, block, 0, ctx->stream>>>(bgr, w, h, step, dark, dark_step);
    SPANO_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 1;
}
