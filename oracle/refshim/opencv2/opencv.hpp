/*
 * oracle/refshim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A header-level stand-in for <opencv2/opencv.hpp>, just wide enough to compile the reference's own hot-path sources
 * UNMODIFIED, where they lie under /root/reference:
 *     src/DC_lidar_only/img_completion.cpp      (img_completion)
 *     src/DC_lidar_camera/img_completion_lc.cpp (interpolate_with_superpixels)
 *     src/DC_lidar_camera/slic.cpp              (Slic::generate_superpixels and friends)
 * The image has no OpenCV C++ headers or libraries, but it has the very same library behind Python
 * bindings (cv2, opencv-python-headless 4.13.0.92).  So this header implements only the containers (cv::Mat with the
 * handful of members those files use, Point, Scalar, Vec3b, Size) and forwards every imgproc call the reference makes
 * (dilate, morphologyEx, medianBlur, GaussianBlur, bilateralFilter) through function pointers that
 * oracle/ref_oracle.py points at cv2.  Scalar loops, control flow, buffer handling: the reference's compiled code;
 * filter arithmetic: the real OpenCV.  Nothing of this is shipped or linked by the product.
 *
 * Semantics kept from OpenCV because the reference relies on them:
 *   - Mat is a shared handle (copy = alias, clone() = deep copy);
 *   - Mat(rows, cols, type, void*) wraps caller memory without copying, and reads it as `type` whatever the
 *     pointee really is (img_completion.cpp:71-77 wraps an int[5][5] as CV_8UC1 -- see the oracle's header);
 *   - Mat::copyTo(dst, mask) (re)allocates dst ZERO-FILLED when its size / type differ, then copies where mask != 0;
 *   - cv::bilateralFilter asserts src.data != dst.data (the reference's in-place call, img_completion.cpp:174, throws).
 */
#ifndef DCMT_REFSHIM_OPENCV_HPP
#define DCMT_REFSHIM_OPENCV_HPP

#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))

#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception("refshim: assertion failed: " #expr); } while (0)

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& what) : std::runtime_error(what) {}
};

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int x_, int y_) : x(x_), y(y_) {}
    Point(double x_, double y_) : x((int)x_), y((int)y_) {}  // only display_center_grid (not on the path) passes doubles
};

struct Scalar {
    double val[4];
    Scalar() : val{0, 0, 0, 0} {}
    Scalar(double a, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

struct Vec3b {
    uchar val[3];
    Vec3b() : val{0, 0, 0} {}
    Vec3b(uchar a, uchar b, uchar c) : val{a, b, c} {}
    Vec3b(double a, double b, double c) : val{(uchar)a, (uchar)b, (uchar)c} {}  // colour_with_cluster_means only
    uchar& operator[](int i) { return val[i]; }
    const uchar& operator[](int i) const { return val[i]; }
};

inline int elem_size_of_type(int type) {
    const int depth = type & ((1 << CV_CN_SHIFT) - 1), cn = (type >> CV_CN_SHIFT) + 1;
    const int bytes = depth == CV_8U ? 1 : depth == CV_16U ? 2 : depth == CV_32F ? 4 : 0;
    if (!bytes) throw Exception("refshim: Mat depth not supported by the stand-in header");
    return bytes * cn;
}

// cv::MatStep as far as the reference uses it: converts to size_t, and step[0] is the row step (utils.cpp:54)
struct MatStep {
    size_t p;
    MatStep(size_t v = 0) : p(v) {}
    operator size_t() const { return p; }
    size_t operator[](int) const { return p; }
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;  // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(int r, int c, int type, void* user) : rows(r), cols(c), data(static_cast<uchar*>(user)), step((size_t)c * elem_size_of_type(type)), type_(type) {}

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        const size_t es = elem_size_of_type(type), n = (size_t)r * c * es;
        own_.reset(new uchar[n ? n : 1], std::default_delete<uchar[]>());
        rows = r; cols = c; type_ = type; step = (size_t)c * es; data = own_.get();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, (size_t)r * m.step); return m; }
    static Mat ones(int r, int c, int type) {
        if (type != CV_8UC1) throw Exception("refshim: Mat::ones is provided for CV_8UC1 only");
        Mat m(r, c, type); std::memset(m.data, 1, (size_t)r * m.step); return m;
    }
    int type() const { return type_; }
    int depth() const { return type_ & ((1 << CV_CN_SHIFT) - 1); }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
    void copyTo(Mat& dst) const {  // unmasked: (re)allocates like OpenCV, then copies row by row
        dst.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
    }
    size_t elemSize() const { return (size_t)elem_size_of_type(type_); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst, const Mat& mask) const {
        if (mask.type() != CV_8UC1 || mask.rows != rows || mask.cols != cols) throw Exception("refshim: copyTo mask must be CV_8UC1 of the source size");
        if (dst.data == nullptr || dst.rows != rows || dst.cols != cols || dst.type() != type_) {
            dst = Mat::zeros(rows, cols, type_);  // OpenCV: a (re)allocated destination is zero-filled before the masked copy
        }
        const size_t es = elemSize();
        for (int r = 0; r < rows; ++r) {
            const uchar* m = mask.data + (size_t)r * mask.step;
            const uchar* s = data + (size_t)r * step;
            uchar* d = dst.data + (size_t)r * dst.step;
            for (int c = 0; c < cols; ++c)
                if (m[c]) std::memcpy(d + c * es, s + c * es, es);
        }
    }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }

private:
    int type_;
    std::shared_ptr<uchar> own_;
};

enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum NormTypes { NORM_MINMAX = 32 };

// imgproc entry points the reference calls; defined in oracle/refshim/refshim.cpp (forwarded to cv2)
void dilate(const Mat& src, Mat& dst, const Mat& kernel);
void morphologyEx(const Mat& src, Mat& dst, int op, const Mat& kernel);
void medianBlur(const Mat& src, Mat& dst, int ksize);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX);
void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace);
// core: cv::normalize(src, dst, alpha, beta, NORM_MINMAX) (main_sl.cpp:523), forwarded to cv2 like the filters
void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type);
// drawing (slic.cpp:330, main_sl.cpp:515: display only, not on the path)
inline void circle(Mat&, Point, int, const Scalar&, int) {}

}  // namespace cv

#endif
