// Minimal stand-in for <opencv2/opencv.hpp>: just enough of cv:: for the REFERENCE's own headers
// (src/math/_projection.h, _blending.h, _distance_cut.h, _gain_compensation.h, _img_manipulation.h, _homography.h,
// src/system/_util.h, src/test/_test.h) and shim/spano_shim.cpp to be parsed and type-checked without OpenCV.
// Used only by tests/test_shim_compiles.py (g++ -fsyntax-only): it catches signature drift between the shim and the
// reference's declarations.  Nothing here is linked or executed.
#pragma once
#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {

template <typename T> struct Point_ {
    T x{}, y{};
    Point_() = default;
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Size_ {
    T width{}, height{};
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x{}, y{}, width{}, height{};
};
typedef Rect_<int> Rect;
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
template <typename T, int N> struct Vec { T val[N]; };
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
struct Matx33f {
    float val[9];
    static Matx33f eye() { return Matx33f{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
};
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };

struct MatStep {
    size_t v = 0;
    operator size_t() const { return v; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    MatStep step;
    Mat() = default;
    Mat(int r, int c, int type);
    Mat(Size s, int type);
    void create(int r, int c, int type);
    bool empty() const;
    int type() const;
    template <typename T> T *ptr(int row = 0);
    template <typename T> const T *ptr(int row = 0) const;
    template <typename T> T &at(int r, int c);
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const;
};
template <typename T> class Mat_ : public Mat {};
Mat operator/(const Mat &a, double s);
void bitwise_not(const Mat &src, Mat &dst);

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REFLECT = 2 };
void remap(const Mat &src, Mat &dst, const Mat &map1, const Mat &map2, int interpolation, int borderMode = BORDER_CONSTANT,
           const Scalar &borderValue = Scalar());

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A &&...a) { return std::make_shared<T>(std::forward<A>(a)...); }

namespace detail {
class RotationWarper {
public:
    virtual ~RotationWarper() = default;
};
class SphericalWarper : public RotationWarper { public: explicit SphericalWarper(float) {} };
class CylindricalWarper : public RotationWarper { public: explicit CylindricalWarper(float) {} };
class StereographicWarper : public RotationWarper { public: explicit StereographicWarper(float) {} };
} // namespace detail

} // namespace cv
