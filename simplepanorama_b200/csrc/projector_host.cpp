// projector_host.cpp -- host-side projector geometry: camera set-up and result-ROI detection.
//
// Replaces, for the hot path, cv::detail::ProjectorBase::setCameraParams and
// RotationWarperBase::detectResultRoi / detectResultRoiByBorder / SphericalWarper::detectResultRoi
// as reached from proj::*_proj::project (reference src/math/_projection.cpp:51,81,321).
// The ROI decides tile sizes and corners (integers), so it is computed with the host libm in
// exactly OpenCV's float expression order; this file must be compiled with -ffp-contract=off.
// Stereographic needs the forward map of EVERY source pixel (OpenCV's base detectResultRoi):
// that scan is spread over host threads.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

#include "spano_internal.h"

namespace {

constexpr double kPi = 3.1415926535897932384626433832795;

// 3x3 float product the way cv::gemm's small-matrix branch does it: float, left to right.
void mul3(const float *A, const float *B, float *D)
{
    float out[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            out[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
    std::memcpy(D, out, sizeof(out));
}

// cv::invert on 3x3 CV_32F: adjugate and determinant in double, result stored as float.
void inv3(const float *M, float *D)
{
    const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) {
        std::memset(D, 0, 9 * sizeof(float));
        return;
    }
    const double s = 1.0 / det;
    D[0] = (float)((e * i - f * h) * s);
    D[1] = (float)((c * h - b * i) * s);
    D[2] = (float)((b * f - c * e) * s);
    D[3] = (float)((f * g - d * i) * s);
    D[4] = (float)((a * i - c * g) * s);
    D[5] = (float)((c * d - a * f) * s);
    D[6] = (float)((d * h - e * g) * s);
    D[7] = (float)((b * g - a * h) * s);
    D[8] = (float)((a * e - b * d) * s);
}

struct Extent {
    float u0 = FLT_MAX, v0 = FLT_MAX, u1 = -FLT_MAX, v1 = -FLT_MAX;
    void add(float u, float v)
    {
        // std::min/std::max argument order of OpenCV: a NaN never replaces the running value
        u0 = (u < u0) ? u : u0;
        v0 = (v < v0) ? v : v0;
        u1 = (u1 < u) ? u : u1;
        v1 = (v1 < v) ? v : v1;
    }
    void merge(const Extent &o)
    {
        u0 = std::min(u0, o.u0);
        v0 = std::min(v0, o.v0);
        u1 = std::max(u1, o.u1);
        v1 = std::max(v1, o.v1);
    }
};

} // namespace

void spano_host_set_camera(SpanoProjector *p, int kind, float scale, const float *K, const float *R)
{
    p->kind = kind;
    p->scale = scale;
    std::memcpy(p->k, K, sizeof(p->k));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p->rinv[3 * r + c] = R[3 * c + r];
    float kinv[9];
    inv3(K, kinv);
    mul3(R, kinv, p->r_kinv);
    mul3(K, p->rinv, p->k_rinv);
}

void spano_host_map_forward(const SpanoProjector *p, float x, float y, float *u, float *v)
{
    const float *m = p->r_kinv;
    const float X = m[0] * x + m[1] * y + m[2];
    const float Y = m[3] * x + m[4] * y + m[5];
    const float Z = m[6] * x + m[7] * y + m[8];
    switch (p->kind) {
    case SPANO_SPHERICAL: {
        *u = p->scale * atan2f(X, Z);
        const float w = Y / sqrtf(X * X + Y * Y + Z * Z);
        *v = p->scale * ((float)kPi - acosf(w == w ? w : 0));
        break;
    }
    case SPANO_CYLINDRICAL:
        *u = p->scale * atan2f(X, Z);
        *v = p->scale * Y / sqrtf(X * X + Z * Z);
        break;
    default: {
        const float az = atan2f(X, Z);
        const float pol = (float)kPi - acosf(Y / sqrtf(X * X + Y * Y + Z * Z));
        const float r = sinf(pol) / (1 - cosf(pol));
        *u = p->scale * r * cosf(az);
        *v = p->scale * r * sinf(az);
    }
    }
}

void spano_host_roi(const SpanoProjector *p, int src_w, int src_h, int roi[4])
{
    Extent ext;
    if (p->kind == SPANO_STEREOGRAPHIC) {
        unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        if ((size_t)src_w * src_h < (1u << 16)) nt = 1;
        std::vector<Extent> part(nt);
        auto work = [&](unsigned t) {
            Extent e;
            float u, v;
            for (int y = (int)t; y < src_h; y += (int)nt)
                for (int x = 0; x < src_w; ++x) {
                    spano_host_map_forward(p, (float)x, (float)y, &u, &v);
                    e.add(u, v);
                }
            part[t] = e;
        };
        if (nt == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
            for (auto &t : th) t.join();
        }
        for (auto &e : part) ext.merge(e);
    } else {
        float u, v;
        for (int i = 0; i < src_w; ++i) {
            spano_host_map_forward(p, (float)i, 0.f, &u, &v);
            ext.add(u, v);
            spano_host_map_forward(p, (float)i, (float)(src_h - 1), &u, &v);
            ext.add(u, v);
        }
        for (int i = 0; i < src_h; ++i) {
            spano_host_map_forward(p, 0.f, (float)i, &u, &v);
            ext.add(u, v);
            spano_host_map_forward(p, (float)(src_w - 1), (float)i, &u, &v);
            ext.add(u, v);
        }
    }
    int x0 = (int)ext.u0, y0 = (int)ext.v0, x1 = (int)ext.u1, y1 = (int)ext.v1;
    if (p->kind == SPANO_SPHERICAL) {
        // SphericalWarper::detectResultRoi: a pole inside the image stretches the ROI to u = 0
        // and v = pi*scale (south) / 0 (north).
        float fu0 = (float)x0, fv0 = (float)y0, fu1 = (float)x1, fv1 = (float)y1;
        for (int south = 1; south >= 0; --south) {
            const float ax = p->rinv[1];
            const float ay = south ? p->rinv[4] : -p->rinv[4];
            const float az = p->rinv[7];
            if (ay > 0.f) {
                const float px = (p->k[0] * ax + p->k[1] * ay) / az + p->k[2];
                const float py = p->k[4] * ay / az + p->k[5];
                if (px > 0.f && px < src_w && py > 0.f && py < src_h) {
                    const float pv = south ? (float)(kPi * p->scale) : 0.f;
                    fu0 = std::min(fu0, 0.f);
                    fv0 = std::min(fv0, pv);
                    fu1 = std::max(fu1, 0.f);
                    fv1 = std::max(fv1, pv);
                }
            }
        }
        x0 = (int)fu0;
        y0 = (int)fv0;
        x1 = (int)fu1;
        y1 = (int)fv1;
    }
    roi[0] = x0;
    roi[1] = y0;
    roi[2] = x1;
    roi[3] = y1;
}

// cv::getGaussianKernel(n, sigma, CV_32F): exp(-x^2 / 2 sigma^2) in double, normalised, stored float.
void spano_host_gaussian_taps(int n, double sigma, float *taps)
{
    const int half = (n - 1) / 2;
    std::vector<double> e(half + 1);
    const double coef = -0.5 / (sigma * sigma);
    double total = 0.0;
    for (int i = 0; i < half; ++i) {
        const double d = (double)(i - half);
        e[i] = std::exp(coef * d * d);
        total += e[i];
    }
    total = total * 2.0 + 1.0;
    const double norm = 1.0 / total;
    for (int i = 0; i < half; ++i) taps[i] = taps[n - 1 - i] = (float)(e[i] * norm);
    taps[half] = (float)norm;
}
