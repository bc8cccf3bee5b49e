// spano_shim.cpp -- drop-in callee bodies for SimplePanorama's compositing path.
//
// Build this file INSTEAD of the bodies it names; the entry points in
// src/classes/_panorama.cpp (stitch_parameters::set_config / get_preview / return_full / blend)
// stay untouched and keep calling the same C++ signatures:
//
//   proj::spherical_proj::project, cylindrical_proj::project, sten_proj::project   src/math/_projection.cpp:27-84,297-324
//   proj::get_proj_parameters                                                     src/math/_projection.cpp:422-454
//   sten_proj::disk_reproj                                                        src/math/_projection.cpp:193-294
//   blnd::createSurroundingMask                                                   src/math/_blending.cpp:278-324
//   blnd::multi_blend, simple_blend, no_blend                                     src/math/_blending.cpp:83-252
//   dcut::distance_transform, dcut::dist_cut                                      src/math/_distance_cut.cpp:7-73
//   gain::get_overlapp_intensity                                                  src/math/_gain_compensation.cpp:7-75
//   test::equalizeIntensities, test::adjust_intensity                             src/test/_test.cpp:9-122
//
// Each body only converts cv::Mat / Eigen arguments to pointers + sizes and forwards to the C ABI
// (include/spano.h).  Errors come back as status codes and are re-thrown as std::runtime_error, the
// exception type the reference already uses on this path (_blending.cpp:86); the worker thread's
// try/catch (src/ui/_image_viewer.cpp:535-550) keeps working.
//
// Not BUILT in the build container (no OpenCV C++ headers / Eigen there), but type-checked there against the reference's
// real headers with stand-ins for <opencv2/...> and <Eigen/...> (tests/test_shim_compiles.py).
#include <stdexcept>
#include <string>
#include <vector>

#include "_projection.h" // reference headers: proj::, blnd::, util::
#include "_blending.h"
#include "_distance_cut.h"      // dcut::
#include "_gain_compensation.h" // gain::OverlapInfo
#include "_test.h"              // test::adjust_intensity
#include "spano.h"

namespace {

// one context per thread: stitch_panorama/get_preview run on a worker thread, get_panorama on the GTK thread
spano_ctx *ctx()
{
    thread_local struct Holder {
        spano_ctx *c = nullptr;
        Holder() { if (spano_create(&c, 0) != SPANO_OK) throw std::runtime_error("spano: no CUDA device (no CPU fallback)"); }
        ~Holder() { spano_destroy(c); }
    } h;
    return h.c;
}

void check(int rc)
{
    if (rc != SPANO_OK) throw std::runtime_error(std::string("spano: ") + spano_last_error(ctx()));
}

// get_proj_parameters -> projector->project(): when set, the project() bodies below also produce the validity mask
// (createSurroundingMask + 3 erosions) in the same device pass instead of a second upload of the warped tile
thread_local cv::Mat *g_mask_request = nullptr;

struct Camera { float K[9], R[9]; };

// K_adj = [f 0 w-cx; 0 f h-cy; 0 0 1] and R, both cast to float32 (what the reference hands to OpenCV)
Camera adjusted(const Eigen::MatrixXd &R, const Eigen::MatrixXd &K, const cv::Mat &img)
{
    Camera c{};
    const double f = K(0, 0);
    const double k[9] = {f, 0, img.cols - K(0, 2), 0, f, img.rows - K(1, 2), 0, 0, 1};
    for (int i = 0; i < 9; ++i) { c.K[i] = (float)k[i]; c.R[i] = (float)R(i / 3, i % 3); }
    return c;
}

proj::warped warp_with(int kind, float focal, const Eigen::MatrixXd &R, const Eigen::MatrixXd &K, const cv::Mat &img, cv::Mat *mask)
{
    const Camera c = adjusted(R, K, img);
    int x, y, w, h;
    check(spano_warp_roi(ctx(), kind, focal, c.K, c.R, img.cols, img.rows, &x, &y, &w, &h));
    proj::warped out;
    out.imgs.create(h, w, CV_8UC3);
    if (mask) mask->create(h, w, CV_8UC1);
    check(spano_warp(ctx(), kind, focal, c.K, c.R, img.data, img.cols, img.rows, img.step, 1.0, out.imgs.data, out.imgs.step,
                     mask ? mask->data : nullptr, mask ? mask->step : 0));
    out.corners = cv::Point(x, y);
    return out;
}

} // namespace

namespace proj {

// `focal` is the private member set by the constructor / change_focal (unchanged in _projection.h)
warped spherical_proj::project(const Eigen::MatrixXd &R, const Eigen::MatrixXd &K, const cv::Mat &img) const
{ return warp_with(SPANO_SPHERICAL, focal, R, K, img, g_mask_request); }
warped cylindrical_proj::project(const Eigen::MatrixXd &R, const Eigen::MatrixXd &K, const cv::Mat &img) const
{ return warp_with(SPANO_CYLINDRICAL, focal, R, K, img, g_mask_request); }
warped sten_proj::project(const Eigen::MatrixXd &R, const Eigen::MatrixXd &K, const cv::Mat &img) const
{ return warp_with(SPANO_STEREOGRAPHIC, focal, R, K, img, g_mask_request); }

proj_data get_proj_parameters(const std::vector<cv::Mat> &images, std::vector<Eigen::MatrixXd> &R, std::vector<Eigen::MatrixXd> &K,
                              std::vector<double> &con, std::shared_ptr<projection> projector, bool get_masks)
{
    proj_data out;
    for (size_t i = 0; i < images.size(); ++i) {
        if (!(con[i] > 0)) continue;
        cv::Mat m;
        struct Request {   // reset on every exit path, exceptions included
            explicit Request(cv::Mat *p) { g_mask_request = p; }
            ~Request() { g_mask_request = nullptr; }
        };
        warped w;
        {
            Request rq(get_masks ? &m : nullptr);
            w = projector->project(R[i], K[i], images[i]);
        }
        if (get_masks) out.msks.push_back(m);
        out.imgs.push_back(w.imgs);
        out.corners.push_back(w.corners);
    }
    return out;
}

void sten_proj::disk_reproj(proj_data &p, bool quadratic)
{
    const int n = (int)p.imgs.size();
    std::vector<const uint8_t *> src(n);
    std::vector<size_t> sstep(n), ostep(n), mstep(n);
    std::vector<int> x(n), y(n), w(n), h(n), nx(n), ny(n), nw(n), nh(n);
    for (int i = 0; i < n; ++i) {
        src[i] = p.imgs[i].data; sstep[i] = p.imgs[i].step;
        x[i] = p.corners[i].x; y[i] = p.corners[i].y; w[i] = p.imgs[i].cols; h[i] = p.imgs[i].rows;
    }
    check(spano_disk_reproj_size(ctx(), n, x.data(), y.data(), w.data(), h.data(), ansatz.x, ansatz.y, radius, quadratic,
                                 nx.data(), ny.data(), nw.data(), nh.data()));
    std::vector<cv::Mat> tiles(n), masks(n);
    std::vector<uint8_t *> optr(n), mptr(n);
    for (int i = 0; i < n; ++i) {
        tiles[i].create(nh[i], nw[i], CV_8UC3); masks[i].create(nh[i], nw[i], CV_8UC1);
        optr[i] = tiles[i].data; ostep[i] = tiles[i].step; mptr[i] = masks[i].data; mstep[i] = masks[i].step;
    }
    check(spano_disk_reproj(ctx(), n, src.data(), sstep.data(), x.data(), y.data(), w.data(), h.data(), ansatz.x, ansatz.y, radius,
                            quadratic, optr.data(), ostep.data(), mptr.data(), mstep.data()));
    p.msks.resize(n);
    for (int i = 0; i < n; ++i) { p.imgs[i] = tiles[i]; p.msks[i] = masks[i]; p.corners[i] = cv::Point(nx[i], ny[i]); }
}

} // namespace proj

namespace blnd {

cv::Mat createSurroundingMask(const cv::Mat &img, bool invert, uchar /*thresholdValue == 1 at every call site*/)
{
    if (img.empty()) return cv::Mat();
    cv::Mat m(img.rows, img.cols, CV_8UC1);
    check(spano_surrounding_mask(ctx(), img.data, img.cols, img.rows, img.step, 0, m.data, m.step));
    if (!invert) cv::bitwise_not(m, m);
    return m;
}

cv::Mat multi_blend(const std::vector<cv::Mat> &images, const std::vector<cv::Mat> &masks, const std::vector<cv::Mat> &masks_orig,
                    const std::vector<cv::Point> &top_lefts, int bands, double sigma)
{
    const int n = (int)images.size();
    std::vector<const uint8_t *> t(n), c(n), v(n);
    std::vector<size_t> ts(n), cs(n), vs(n);
    std::vector<int> x(n), y(n), w(n), h(n);
    for (int i = 0; i < n; ++i) {
        t[i] = images[i].data; ts[i] = images[i].step; c[i] = masks[i].data; cs[i] = masks[i].step;
        v[i] = masks_orig[i].data; vs[i] = masks_orig[i].step;
        x[i] = top_lefts[i].x; y[i] = top_lefts[i].y; w[i] = images[i].cols; h[i] = images[i].rows;
    }
    int cw, ch, mx, my;
    check(spano_pan_dimension(n, x.data(), y.data(), w.data(), h.data(), &cw, &ch, &mx, &my));
    cv::Mat out(ch, cw, CV_32FC3);
    check(spano_multiblend(ctx(), n, t.data(), ts.data(), c.data(), cs.data(), v.data(), vs.data(), x.data(), y.data(), w.data(),
                           h.data(), bands, sigma, SPANO_OUT_F32, out.data, out.step));
    return out;
}

static cv::Mat simple_or_no(bool simple, const std::vector<cv::Mat> &images, const std::vector<cv::Mat> &masks,
                            const std::vector<cv::Point> &top_lefts)
{
    if (images.empty() || images.size() != masks.size() || images.size() != top_lefts.size())
        throw std::runtime_error("Input consistency!");
    const int n = (int)images.size();
    std::vector<const uint8_t *> t(n), m(n);
    std::vector<size_t> ts(n), ms(n);
    std::vector<int> x(n), y(n), w(n), h(n);
    for (int i = 0; i < n; ++i) {
        t[i] = images[i].data; ts[i] = images[i].step; m[i] = masks[i].data; ms[i] = masks[i].step;
        x[i] = top_lefts[i].x; y[i] = top_lefts[i].y; w[i] = images[i].cols; h[i] = images[i].rows;
    }
    int cw, ch, mx, my;
    check(spano_pan_dimension(n, x.data(), y.data(), w.data(), h.data(), &cw, &ch, &mx, &my));
    cv::Mat out(ch, cw, CV_8UC3);
    check((simple ? spano_simple_blend : spano_no_blend)(ctx(), n, t.data(), ts.data(), m.data(), ms.data(), x.data(), y.data(),
                                                         w.data(), h.data(), out.data, out.step));
    return out;
}

// src/math/_blending.cpp:83-153 (stitch_parameters::blend, SIMPLE_BLEND)
cv::Mat simple_blend(const std::vector<cv::Mat> &images, const std::vector<cv::Mat> &masks, const std::vector<cv::Point> &top_lefts)
{
    return simple_or_no(true, images, masks, top_lefts);
}

// src/math/_blending.cpp:157-182 (stitch_parameters::blend, NO_BLEND)
cv::Mat no_blend(const std::vector<cv::Mat> &images, const std::vector<cv::Mat> &masks, const std::vector<cv::Point> &top_lefts)
{
    return simple_or_no(false, images, masks, top_lefts);
}

} // namespace blnd

namespace test {

// src/test/_test.cpp:9-106 (called from stitch_parameters::set_config when conf.blend_intensity is on)
std::vector<cv::Mat> equalizeIntensities(const std::vector<cv::Mat> &images, const std::vector<cv::Mat> &masks,
                                         const std::vector<cv::Point> &top_lefts, float ratio)
{
    const int n = (int)images.size();
    std::vector<const uint8_t *> t(n), m(n);
    std::vector<size_t> ts(n), ms(n), fs(n);
    std::vector<int> x(n), y(n), w(n), h(n);
    std::vector<cv::Mat> fields(n);
    std::vector<float *> f(n);
    for (int i = 0; i < n; ++i) {
        t[i] = images[i].data; ts[i] = images[i].step; m[i] = masks[i].data; ms[i] = masks[i].step;
        x[i] = top_lefts[i].x; y[i] = top_lefts[i].y; w[i] = images[i].cols; h[i] = images[i].rows;
        int fw = 0, fh = 0;
        check(spano_equalize_intensities_size(w[i], h[i], ratio, &fw, &fh));
        fields[i].create(fh, fw, CV_32FC1);
        f[i] = fields[i].ptr<float>(); fs[i] = fields[i].step;
    }
    check(spano_equalize_intensities(ctx(), n, t.data(), ts.data(), m.data(), ms.data(), x.data(), y.data(), w.data(), h.data(), ratio,
                                     f.data(), fs.data()));
    return fields;
}

// src/test/_test.cpp:110-122 (called from return_full / get_preview when conf.blend_intensity is on)
void adjust_intensity(std::vector<cv::Mat> &images, const std::vector<cv::Mat> &intensities)
{
    for (size_t i = 0; i < images.size(); ++i) {
        cv::Mat &im = images[i];
        const cv::Mat &f = intensities[i];   // CV_32FC1, low resolution
        check(spano_adjust_intensity(ctx(), im.data, im.cols, im.rows, im.step, f.ptr<float>(), f.cols, f.rows, f.step));
    }
}

} // namespace test

namespace dcut {

// src/math/_distance_cut.cpp:57-73
std::vector<cv::Mat> distance_transform(const std::vector<cv::Mat> &masks)
{
    std::vector<cv::Mat> ret;
    for (const cv::Mat &m : masks) {
        cv::Mat d(m.rows, m.cols, CV_32FC1);
        check(spano_distance_transform(ctx(), m.data, m.cols, m.rows, m.step, d.ptr<float>(), d.step));
        ret.push_back(d / 255);
    }
    return ret;
}

// src/math/_distance_cut.cpp:7-51 (called from stitch_parameters::set_config when conf.cut is on)
std::vector<cv::Mat> dist_cut(const std::vector<cv::Mat> &masks, const std::vector<cv::Point> &top_lefts)
{
    const int n = (int)masks.size();
    std::vector<const uint8_t *> m(n);
    std::vector<uint8_t *> o(n);
    std::vector<size_t> ms(n), os(n);
    std::vector<int> x(n), y(n), w(n), h(n);
    std::vector<cv::Mat> cut(n);
    for (int i = 0; i < n; ++i) {
        cut[i].create(masks[i].rows, masks[i].cols, masks[i].type());
        m[i] = masks[i].data; ms[i] = masks[i].step; o[i] = cut[i].data; os[i] = cut[i].step;
        x[i] = top_lefts[i].x; y[i] = top_lefts[i].y; w[i] = masks[i].cols; h[i] = masks[i].rows;
    }
    check(spano_dist_cut(ctx(), n, m.data(), ms.data(), x.data(), y.data(), w.data(), h.data(), o.data(), os.data()));
    return cut;
}

} // namespace dcut

namespace gain {

// src/math/_gain_compensation.cpp:7-75 (struct OverlapInfo {int i, j; double area, I_i, I_j;} from _gain_compensation.h)
std::vector<OverlapInfo> get_overlapp_intensity(const std::vector<cv::Mat> &warped_images, const std::vector<cv::Point> &corners,
                                                const cv::Mat &adj_pass)
{
    const int n = (int)warped_images.size();
    std::vector<const uint8_t *> t(n);
    std::vector<size_t> ts(n);
    std::vector<int> x(n), y(n), w(n), h(n);
    for (int i = 0; i < n; ++i) {
        t[i] = warped_images[i].data; ts[i] = warped_images[i].step;
        x[i] = corners[i].x; y[i] = corners[i].y; w[i] = warped_images[i].cols; h[i] = warped_images[i].rows;
    }
    cv::Mat adj;
    adj_pass.convertTo(adj, CV_64F);               // contiguous n x n doubles
    std::vector<spano_overlap_info> out((size_t)n * (n + 1) / 2);
    int count = 0;
    check(spano_overlap_intensity(ctx(), n, t.data(), ts.data(), x.data(), y.data(), w.data(), h.data(), adj.ptr<double>(), out.data(), &count));
    std::vector<OverlapInfo> results;
    for (int k = 0; k < count; ++k) results.push_back({out[k].i, out[k].j, out[k].area, out[k].I_i, out[k].I_j});
    return results;
}

} // namespace gain
