/*
 * oracle/refshim/refshim_front.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Third translation unit of oracle/_ref/libdcmt_ref.so: the callers either side of the hot path,
 * compiled from the reference's own lines:
 *   - the LiDAR projection of withSuperPixels (src/DC_stereo_lidar/main_sl.cpp:474-523): rigid transform, z > 0 filter,
 *     `P * p.homogeneous()`, perspective division, bounds test, (int) truncation, last-writer-wins scatter and
 *     cv::normalize(0, 80, NORM_MINMAX).  The lines are a function body inside a whole-program file, so the build
 *     streams them into a scratch include (ref_project.inc) that is included inside a wrapper declaring the variables
 *     they use, with stand-ins for the two PCL types (a point of four floats, a vector of points) and the Eigen
 *     types (oracle/refshim/Eigen/Dense, which documents the one evaluation order that is Eigen's own choice);
 *   - read_M / write_M (src/DC_lidar_only/utils.cpp:15-58), the raw cv::Mat .bin format.
 */
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#include <cstdint>
#include <fstream>
#include <memory>
#include <streambuf>
#include <vector>

namespace pcl {
struct PointXYZ {  // PCL's PointXYZ is 16 bytes: x, y, z and one float of padding -- which is why the reference can read the
    float x, y, z, pad;  // N x 4 float32 Velodyne payload straight into cloud->points (main_sl.cpp:468-469)
    PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
};
template <class P>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    size_t size() const { return points.size(); }
    const P& at(size_t i) const { return points.at(i); }
    void push_back(const P& p) { points.push_back(p); }
    void resize(size_t n) { points.resize(n); }
};
}  // namespace pcl

#include "ref_utils.inc"  // read_M, write_M

namespace {
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
};
struct QuietCout {
    NullBuf nb;
    std::streambuf* old;
    QuietCout() : old(std::cout.rdbuf(&nb)) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};

// the variables main_sl.cpp:474-523 uses, then the lines themselves
int project_body(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, Eigen::Matrix4f& T, Eigen::Matrix<float, 3, 4>& P, cv::Mat& image,
                 cv::Mat& projected_depths, cv::Mat& normalized_out) {
#include "ref_project.inc"
    normalized_out = normalized_depths;
    return projected;
}
}  // namespace

extern "C" {
#define REF_API __attribute__((visibility("default")))

// points: n x 4 float32 (the .bin payload); T 4x4 and P 3x4 row-major; outputs rows x cols float32
REF_API int dcmt_ref_lidar_project(const float* points, int n, const float* T_rm, const float* P_rm, int rows, int cols,
                                   float* projected_out, float* normalized_out, int* n_projected, char* err, int err_cap) {
    QuietCout q;
    try {
        pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>);
        cloud->resize((size_t)n);
        if (n > 0) std::memcpy(reinterpret_cast<char*>(&cloud->points[0]), points, (size_t)n * 16);  // main_sl.cpp:469
        Eigen::Matrix4f T;
        Eigen::Matrix<float, 3, 4> P;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) T(r, c) = T_rm[r * 4 + c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) P(r, c) = P_rm[r * 4 + c];
        cv::Mat image = cv::Mat::zeros(rows, cols, CV_8UC3);
        cv::Mat projected_depths = cv::Mat::zeros(rows, cols, CV_32F);  // main_sl.cpp:1145
        cv::Mat normalized;
        const int projected = project_body(cloud, T, P, image, projected_depths, normalized);
        if (n_projected) *n_projected = projected;
        for (int r = 0; r < rows; ++r) {
            std::memcpy(projected_out + (size_t)r * cols, projected_depths.data + (size_t)r * projected_depths.step, (size_t)cols * 4);
            std::memcpy(normalized_out + (size_t)r * cols, normalized.data + (size_t)r * normalized.step, (size_t)cols * 4);
        }
        return 0;
    } catch (const std::exception& e) {
        if (err && err_cap > 0) { std::strncpy(err, e.what(), (size_t)err_cap - 1); err[err_cap - 1] = 0; }
        return -1;
    }
}

// write_M (utils.cpp:39-58) of a rows x cols CV_32FC1 matrix wrapping `data`
REF_API int dcmt_ref_write_M(const char* path, const float* data, int rows, int cols) {
    QuietCout q;
    const cv::Mat m(rows, cols, CV_32FC1, const_cast<float*>(data));
    write_M(path, m);
    return 0;
}
// read_M (utils.cpp:15-37): returns rows * cols (0 if the file could not be read) and the header fields; the payload is
// copied into out (capacity out_cap floats) when it is CV_32FC1
REF_API int dcmt_ref_read_M(const char* path, int* rows, int* cols, int* type, float* out, int out_cap) {
    QuietCout q;
    cv::Mat m;
    read_M(path, m);
    if (m.empty()) return 0;
    *rows = m.rows; *cols = m.cols; *type = m.type();
    const int n = m.rows * m.cols;
    if (m.type() == CV_32FC1 && out && n <= out_cap)
        for (int r = 0; r < m.rows; ++r) std::memcpy(out + (size_t)r * m.cols, m.data + (size_t)r * m.step, (size_t)m.cols * 4);
    return n;
}
}  // extern "C"
