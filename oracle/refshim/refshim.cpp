/*
 * oracle/refshim/refshim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Glue around the reference's own sources compiled from /root/reference (see opencv2/opencv.hpp in this directory and
 * oracle/Makefile, target _ref/libdcmt_ref.so):
 *   - the five cv:: imgproc functions of the stand-in header, forwarded to callbacks that oracle/ref_oracle.py
 *     registers (cv2.dilate, cv2.morphologyEx, cv2.medianBlur, cv2.GaussianBlur, cv2.bilateralFilter);
 *   - extern "C" entry points that wrap caller buffers in cv::Mat / fill a Slic and call the reference functions
 *     with their own signatures.  The reference prints progress to std::cout; the entry points park std::cout on a
 *     null buffer for the duration of the call so that bench.py's stdout stays one JSON line.
 */
#include <opencv2/opencv.hpp>

#include <cmath>
#include <cstdint>
#include <streambuf>

#include "slic.h"  // /root/reference/src/DC_lidar_camera/slic.h (-I on the compile line)

// the reference functions (defined in the reference's .cpp files compiled next to this one)
void img_completion(const cv::Mat& sparse_r_img, cv::Mat& dense_r_img, const bool& extr, const std::string& blur_type);
void interpolate_with_superpixels(Slic& slic, const cv::Mat& sparse_r_img, cv::Mat& dense_r_img, const std::string& blur_type, int use_superpixel);

extern "C" {
// src / dst: dense row-major rows x cols planes of the given OpenCV type; dst never aliases src (the shim copies)
typedef int (*ref_morph_cb)(int op /*1 dilate, 3 close*/, const void* src, void* dst, int rows, int cols, int type,
                            const unsigned char* kernel, int krows, int kcols);
typedef int (*ref_blur_cb)(int kind /*0 median, 1 gaussian, 2 bilateral, 3 normalize(p0, p1, NORM_MINMAX)*/, const void* src, void* dst, int rows, int cols,
                           int type, int ksize, double p0, double p1);
}

namespace {
ref_morph_cb g_morph = nullptr;
ref_blur_cb g_blur = nullptr;
long long g_calls[5] = {0, 0, 0, 0, 0};  // dilate, morphologyEx, medianBlur, GaussianBlur, bilateralFilter

struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
};
struct QuietCout {
    NullBuf nb;
    std::streambuf* old;
    QuietCout() : old(std::cout.rdbuf(&nb)) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};

cv::Mat dense_copy(const cv::Mat& m) { return m.clone(); }  // clone() packs rows (step = cols * elemSize)

void store(const cv::Mat& packed, cv::Mat& dst, int rows, int cols, int type) {
    dst.create(rows, cols, type);
    for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step, packed.data + (size_t)r * packed.step, (size_t)cols * packed.elemSize());
}

void run_morph(int op, const cv::Mat& src, cv::Mat& dst, const cv::Mat& kernel) {
    if (!g_morph) throw cv::Exception("refshim: callbacks not registered");
    if (kernel.type() != CV_8UC1) throw cv::Exception("refshim: structuring element must be CV_8UC1");
    const cv::Mat s = dense_copy(src), k = dense_copy(kernel);
    cv::Mat out(src.rows, src.cols, src.type());
    if (g_morph(op, s.data, out.data, src.rows, src.cols, src.type(), k.data, k.rows, k.cols)) throw cv::Exception("refshim: cv2 morphology callback failed");
    store(out, dst, src.rows, src.cols, src.type());
}

void run_blur(int kind, const cv::Mat& src, cv::Mat& dst, int ksize, double p0, double p1) {
    if (!g_blur) throw cv::Exception("refshim: callbacks not registered");
    const cv::Mat s = dense_copy(src);
    cv::Mat out(src.rows, src.cols, src.type());
    if (g_blur(kind, s.data, out.data, src.rows, src.cols, src.type(), ksize, p0, p1)) throw cv::Exception("refshim: cv2 blur callback failed");
    store(out, dst, src.rows, src.cols, src.type());
}
}  // namespace

namespace cv {
void dilate(const Mat& src, Mat& dst, const Mat& kernel) { ++g_calls[0]; run_morph(MORPH_DILATE, src, dst, kernel); }
void morphologyEx(const Mat& src, Mat& dst, int op, const Mat& kernel) { ++g_calls[1]; run_morph(op, src, dst, kernel); }
void medianBlur(const Mat& src, Mat& dst, int ksize) { ++g_calls[2]; run_blur(0, src, dst, ksize, 0.0, 0.0); }
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX) {
    ++g_calls[3];
    if (ksize.width != ksize.height) throw Exception("refshim: square Gaussian kernels only");
    run_blur(1, src, dst, ksize.width, sigmaX, 0.0);
}
void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type) {
    if (norm_type != NORM_MINMAX) throw Exception("refshim: NORM_MINMAX only");
    run_blur(3, src, dst, 0, alpha, beta);
}
void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace) {
    ++g_calls[4];
    // OpenCV (modules/imgproc/src/bilateral_filter.dispatch.cpp): CV_Assert(... && src.data != dst.data) -- in-place use throws
    if (src.data == dst.data) throw Exception("bilateralFilter: (-215:Assertion failed) src.data != dst.data");
    run_blur(2, src, dst, d, sigmaColor, sigmaSpace);
}
}  // namespace cv

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API void dcmt_ref_set_callbacks(ref_morph_cb m, ref_blur_cb b) { g_morph = m; g_blur = b; }
REF_API void dcmt_ref_call_counts(long long* out5) { for (int i = 0; i < 5; ++i) out5[i] = g_calls[i]; }

// img_completion (src/DC_lidar_only/img_completion.cpp:17-204).  blur: 0 "none", 1 "gaussian", 2 "bilateral".
// returns 0, or -1 when the reference threw (err receives the message)
REF_API int dcmt_ref_img_completion(const float* sparse, float* dense, int rows, int cols, int blur, char* err, int err_cap) {
    QuietCout q;
    try {
        const cv::Mat in(rows, cols, CV_32FC1, const_cast<float*>(sparse));
        cv::Mat out;
        const bool extr = false;
        const std::string bt = blur == 1 ? "gaussian" : blur == 2 ? "bilateral" : "none";
        img_completion(in, out, extr, bt);
        if (out.rows != rows || out.cols != cols || out.type() != CV_32FC1) throw cv::Exception("refshim: unexpected output Mat");
        for (int r = 0; r < rows; ++r) std::memcpy(dense + (size_t)r * cols, out.data + (size_t)r * out.step, (size_t)cols * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        if (err && err_cap > 0) { std::strncpy(err, e.what(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
        return -1;
    }
}

// interpolate_with_superpixels (src/DC_lidar_camera/img_completion_lc.cpp:34-203).  labels: row-major [row][col] int32,
// copied into the reference's own container Slic::clusters[col][row]; n_centers = slic.centers.size().
REF_API int dcmt_ref_interpolate_with_superpixels(const float* sparse, const int32_t* labels, int n_centers, float* dense, int rows,
                                                  int cols, int blur, int use_superpixel, char* err, int err_cap) {
    QuietCout q;
    try {
        Slic slic;
        slic.clusters.assign((size_t)cols, std::vector<int>((size_t)rows, -1));
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) slic.clusters[c][r] = labels[(size_t)r * cols + c];
        slic.centers.assign((size_t)n_centers, std::vector<double>(5, 0.0));
        const cv::Mat in(rows, cols, CV_32FC1, const_cast<float*>(sparse));
        cv::Mat out;
        const std::string bt = blur == 1 ? "gaussian" : blur == 2 ? "bilateral" : "none";
        interpolate_with_superpixels(slic, in, out, bt, use_superpixel);
        if (out.rows != rows || out.cols != cols || out.type() != CV_32FC1) throw cv::Exception("refshim: unexpected output Mat");
        for (int r = 0; r < rows; ++r) std::memcpy(dense + (size_t)r * cols, out.data + (size_t)r * out.step, (size_t)cols * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        if (err && err_cap > 0) { std::strncpy(err, e.what(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
        return -1;
    }
}

// Slic::generate_superpixels (src/DC_lidar_camera/slic.cpp:101-182) on an 8-bit 3-channel (Lab) image.
// labels_out: row-major int32 rows x cols (transposed from clusters[col][row]); centers_out: up to max_centers x 5 doubles
// (L, a, b, x, y); counts_out: up to max_centers ints.  Returns the number of centres, or -1.
REF_API int dcmt_ref_slic(const unsigned char* lab, int rows, int cols, int step, int nc, int32_t* labels_out, double* centers_out,
                          int* counts_out, int max_centers, char* err, int err_cap) {
    QuietCout q;
    try {
        cv::Mat image(rows, cols, CV_8UC3, const_cast<unsigned char*>(lab));
        Slic slic;
        slic.generate_superpixels(image, step, nc);
        const int n = (int)slic.centers.size();
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) labels_out[(size_t)r * cols + c] = slic.clusters[c][r];
        for (int i = 0; i < n && i < max_centers; ++i) {
            for (int k = 0; k < 5; ++k) centers_out[(size_t)i * 5 + k] = slic.centers[i][k];
            if (counts_out) counts_out[i] = slic.center_counts[i];
        }
        return n;
    } catch (const std::exception& e) {
        if (err && err_cap > 0) { std::strncpy(err, e.what(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
        return -1;
    }
}
}  // extern "C"
