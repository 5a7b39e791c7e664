// img_completion.h -- the header the reference includes but never shipped
// (`#include "img_completion.h"`: src/DC_lidar_only/img_completion.cpp:15, main.cpp:1, utils.cpp:3), implemented
// on top of the C ABI of libdcmt.so (include/dcmt.h).
//
// With OpenCV available (the reference's own build), define nothing and the reference signatures are provided on
// cv::Mat, so that main.cpp / main_lc.cpp / main_sl.cpp link against libdcmt.so instead of compiling
// img_completion.cpp / img_completion_lc.cpp:
//
//     void img_completion(const cv::Mat&, cv::Mat&, const bool& extr, const std::string& blur_type);      (img_completion.cpp:17)
//     void interpolate_with_superpixels(Slic&, const cv::Mat&, cv::Mat&, const std::string&, int);        (img_completion_lc.cpp:34)
//     void calculateMeasuementDerivatives / get_initial_disparity / optimize_IG / retrieve_optimized_depth (main_sl.cpp:715-885)
//
// Without OpenCV (this repository's CI) the same functions are available on dcmt::MatView, a plain
// {rows, cols, step, data} view with cv::Mat's CV_32FC1 memory layout, so the shim itself can be compiled and
// tested.  Errors of the C ABI become std::runtime_error (the reference functions return void).
#ifndef DCMT_IMG_COMPLETION_H_
#define DCMT_IMG_COMPLETION_H_

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "dcmt.h"

#if !defined(DCMT_NO_OPENCV) && defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define DCMT_HAVE_OPENCV 1
#endif
#endif

namespace dcmt {

// cv::Mat-compatible view of a single-channel image: element (r, c) lives at data + r * step + c * sizeof(T)
struct MatView {
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes between rows (cv::Mat::step)
    void* data = nullptr;
};

inline void check(int status) {
    if (status != DCMT_OK) throw std::runtime_error(std::string("dcmt: ") + dcmt_status_string(status) + ": " + dcmt_last_error());
}

// blur_type semantics of img_completion.cpp:172-189
inline int blur_code(const std::string& blur_type) {
    return blur_type == "bilateral" ? DCMT_BLUR_BILATERAL : blur_type == "gaussian" ? DCMT_BLUR_GAUSSIAN : DCMT_BLUR_NONE;
}

// img_completion.cpp:17-204.  `dense` must already describe a rows x cols float buffer (cv::Mat overload allocates it,
// like the reference's clone()).  `extr` is ignored, as in the reference (:103).
inline void img_completion(const MatView& sparse, MatView& dense, const bool& /*extr*/, const std::string& blur_type,
                           int path = DCMT_PATH_AUTO) {
    if (dense.rows != sparse.rows || dense.cols != sparse.cols || !dense.data) throw std::invalid_argument("dcmt: dense view not allocated");
    check(dcmt_img_completion_f32_host(static_cast<const float*>(sparse.data), static_cast<float*>(dense.data), sparse.rows,
                                       sparse.cols, sparse.step, 0, 1, blur_code(blur_type), path, nullptr));
    // stride of input and output must agree in the ABI; MatViews with different steps are handled by the cv::Mat overload
}

// main.cpp:75-93 in one call: the uint16 payload of a KITTI depth PNG (CV_16UC1 view, metres * 256) -> dense float32 metres;
// stands for image_r.convertTo(projected_depths, CV_32F, 1.0 / 256.0) (:79) followed by img_completion (:93)
inline void img_completion_u16(const MatView& sparse_u16, MatView& dense, const std::string& blur_type, int path = DCMT_PATH_AUTO) {
    if (dense.rows != sparse_u16.rows || dense.cols != sparse_u16.cols || !dense.data) throw std::invalid_argument("dcmt: dense view not allocated");
    check(dcmt_img_completion_u16_host(static_cast<const uint16_t*>(sparse_u16.data), static_cast<float*>(dense.data), sparse_u16.rows,
                                       sparse_u16.cols, sparse_u16.step, 0, dense.step, 0, 1, blur_code(blur_type), path, nullptr));
}

// The evaluation loops of the three programs; GT and result must share one step.
inline dcmt_eval_result evaluate(const MatView& GT_img, const MatView& r_img, float tolerance, int mode) {
    if (GT_img.rows != r_img.rows || GT_img.cols != r_img.cols || GT_img.step != r_img.step) throw std::invalid_argument("dcmt: evaluate needs same-shaped Mats");
    dcmt_eval_result res;
    check(dcmt_evaluate_f32_host(static_cast<const float*>(GT_img.data), static_cast<const float*>(r_img.data), GT_img.rows, GT_img.cols,
                                 GT_img.step, 0, 1, tolerance, mode, &res));
    return res;
}
// src/DC_lidar_only/main.cpp:16-34: "mse" = mean of (gt - r) over gt > 0 (a signed mean, sic)
inline void evaluate_performance(const MatView& GT_img, const MatView& r_img, float& mse) {
    mse = evaluate(GT_img, r_img, 0.0f, DCMT_EVAL_GT_VALID).mean_err;
}
// src/DC_lidar_camera/main_lc.cpp:85-116: `int tolerance = 0.1` is 0; "mse" = sqrt(sum d^2 / count) (sic), mae
inline void evaluate_performance(const MatView& GT_img, const MatView& r_img, float& mse, float& mae) {
    const dcmt_eval_result r = evaluate(GT_img, r_img, 0.0f, DCMT_EVAL_BOTH_VALID);
    mse = r.rmse;
    mae = r.mae;
}
// src/DC_stereo_lidar/main_sl.cpp:1031-1061: tolerance 2
inline void evaluate_performances(const MatView& GT_img, const MatView& r_img, float& mae, float& rmse) {
    const dcmt_eval_result r = evaluate(GT_img, r_img, 2.0f, DCMT_EVAL_BOTH_VALID);
    mae = r.mae;
    rmse = r.rmse;
}

// main_sl.cpp:478-523: the Velodyne .bin payload (n_points x 4 floats) -> projected_depths and its cv::normalize(0, 80).
// T (4x4) and P (3x4) are row-major here; from Eigen pass `Eigen::Matrix<float, 4, 4, Eigen::RowMajor>(T).data()`.
inline int project_lidar(const float* points, int n_points, const float* T_row_major, const float* P_row_major, MatView& projected_depths,
                         MatView& normalized_depths, float norm_a = 0.0f, float norm_b = 80.0f) {
    const int rows = projected_depths.rows, cols = projected_depths.cols;
    if (normalized_depths.rows != rows || normalized_depths.cols != cols || projected_depths.step != (size_t)cols * 4 ||
        normalized_depths.step != (size_t)cols * 4)
        throw std::invalid_argument("dcmt: project_lidar needs continuous same-shaped Mats");
    int32_t n = 0;
    check(dcmt_lidar_project_f32_host(points, n_points, T_row_major, P_row_major, rows, cols, static_cast<float*>(projected_depths.data),
                                      static_cast<float*>(normalized_depths.data), norm_a, norm_b, &n));
    return n;
}

// img_completion_lc.cpp:34-203.  `clusters_col_major` is Slic::clusters, indexed [col][row] (:83); it is transposed
// into the row-major int32 label map of the ABI here.  `n_centers` is slic.centers.size().
inline void interpolate_with_superpixels(const std::vector<std::vector<int>>& clusters_col_major, size_t n_centers,
                                         const MatView& sparse, MatView& dense, const std::string& /*blur_type*/,
                                         int use_superpixel) {
    if (dense.rows != sparse.rows || dense.cols != sparse.cols || !dense.data) throw std::invalid_argument("dcmt: dense view not allocated");
    std::vector<int32_t> labels;
    if (use_superpixel) {
        labels.resize((size_t)sparse.rows * sparse.cols);
        for (int j = 0; j < sparse.cols; ++j)
            for (int i = 0; i < sparse.rows; ++i) labels[(size_t)i * sparse.cols + j] = clusters_col_major[j][i];
    }
    check(dcmt_interpolate_with_superpixels_f32_host(static_cast<const float*>(sparse.data), use_superpixel ? labels.data() : nullptr,
                                                     (int)n_centers, static_cast<float*>(dense.data), sparse.rows, sparse.cols,
                                                     sparse.step, 0, 1, use_superpixel, nullptr));
}

// main_sl.cpp:1165-1253 in one call: gray images (CV_8UC1 views) + initial dense depth -> refined depth
inline void stereo_refine(const MatView& depth_ig, const MatView& left_gray, const MatView& right_gray, MatView& depth_out,
                          const dcmt_stereo_params* params = nullptr) {
    dcmt_stereo_params p;
    if (params) p = *params;
    else dcmt_stereo_params_default(&p);
    const int rows = depth_ig.rows, cols = depth_ig.cols;
    if (depth_ig.step != (size_t)cols * 4 || depth_out.step != (size_t)cols * 4 || left_gray.step != (size_t)cols ||
        right_gray.step != (size_t)cols)
        throw std::invalid_argument("dcmt: stereo_refine needs continuous Mats");
    check(dcmt_stereo_refine_f32_host(static_cast<const float*>(depth_ig.data), static_cast<const uint8_t*>(left_gray.data),
                                      static_cast<const uint8_t*>(right_gray.data), static_cast<float*>(depth_out.data), nullptr, rows,
                                      cols, 1, &p));
}

}  // namespace dcmt

#ifdef DCMT_HAVE_OPENCV
// ---- the reference's own signatures -------------------------------------------------------------------------

namespace dcmt {
inline MatView view_of(const cv::Mat& m) { return MatView{m.rows, m.cols, (size_t)m.step, m.data}; }
}  // namespace dcmt

// img_completion.cpp:17-20.  A CV_16UC1 Mat (the KITTI PNG as read by cv::imread(..., IMREAD_ANYDEPTH), main.cpp:75) is
// accepted as well and stands for convertTo(CV_32F, 1.0 / 256.0) + img_completion.
inline void img_completion(const cv::Mat& sparse_r_img, cv::Mat& dense_r_img, const bool& extr, const std::string& blur_type) {
    if (sparse_r_img.type() == CV_16UC1) {
        dense_r_img.create(sparse_r_img.rows, sparse_r_img.cols, CV_32FC1);
        dcmt::MatView s16 = dcmt::view_of(sparse_r_img), d16 = dcmt::view_of(dense_r_img);
        dcmt::img_completion_u16(s16, d16, blur_type);
        return;
    }
    CV_Assert(sparse_r_img.type() == CV_32FC1);
    cv::Mat in = sparse_r_img.isContinuous() ? sparse_r_img : sparse_r_img.clone();
    dense_r_img.create(in.rows, in.cols, CV_32FC1);  // the reference's `dense = sparse.clone()` (:27)
    dcmt::MatView s = dcmt::view_of(in), d = dcmt::view_of(dense_r_img);
    dcmt::img_completion(s, d, extr, blur_type);
}

// img_completion_lc.cpp:34-38.  Slic is the reference's class (slic.h:30-71): public `clusters` ([col][row]) and `centers`.
template <class SlicT>
inline void interpolate_with_superpixels(SlicT& slic, const cv::Mat& sparse_r_img, cv::Mat& dense_r_img, const std::string& blur_type,
                                         int use_superpixel) {
    CV_Assert(sparse_r_img.type() == CV_32FC1);
    cv::Mat in = sparse_r_img.isContinuous() ? sparse_r_img : sparse_r_img.clone();
    dense_r_img.create(in.rows, in.cols, CV_32FC1);
    dcmt::MatView s = dcmt::view_of(in), d = dcmt::view_of(dense_r_img);
    dcmt::interpolate_with_superpixels(slic.clusters, slic.centers.size(), s, d, blur_type, use_superpixel);
}

// main_sl.cpp:1165-1253: the sequence entry fill -> calculateMeasuementDerivatives -> get_initial_disparity ->
// optimize_IG -> retrieve_optimized_depth -> GaussianBlur on the caller's Mats (gray CV_8UC1, depth CV_32FC1)
inline void stereo_refine(const cv::Mat& dense_range_img, const cv::Mat& left_gray, const cv::Mat& right_gray, cv::Mat& optimized_depth,
                          const dcmt_stereo_params* params = nullptr) {
    CV_Assert(dense_range_img.type() == CV_32FC1 && left_gray.type() == CV_8UC1 && right_gray.type() == CV_8UC1);
    cv::Mat ig = dense_range_img.isContinuous() ? dense_range_img : dense_range_img.clone();
    cv::Mat l = left_gray.isContinuous() ? left_gray : left_gray.clone(), r = right_gray.isContinuous() ? right_gray : right_gray.clone();
    optimized_depth.create(ig.rows, ig.cols, CV_32FC1);
    dcmt::MatView vi = dcmt::view_of(ig), vl = dcmt::view_of(l), vr = dcmt::view_of(r), vo = dcmt::view_of(optimized_depth);
    dcmt::stereo_refine(vi, vl, vr, vo, params);
}

// the evaluation functions of main.cpp:16 / main_lc.cpp:85 / main_sl.cpp:1031 (CV_32FC1 Mats)
inline void evaluate_performance(const cv::Mat& GT_img, const cv::Mat& r_img, float& mse) {
    cv::Mat g = GT_img.isContinuous() ? GT_img : GT_img.clone(), r = r_img.isContinuous() ? r_img : r_img.clone();
    dcmt::evaluate_performance(dcmt::view_of(g), dcmt::view_of(r), mse);
}
inline void evaluate_performance(const cv::Mat& GT_img, const cv::Mat& r_img, float& mse, float& mae) {
    cv::Mat g = GT_img.isContinuous() ? GT_img : GT_img.clone(), r = r_img.isContinuous() ? r_img : r_img.clone();
    dcmt::evaluate_performance(dcmt::view_of(g), dcmt::view_of(r), mse, mae);
}
inline void evaluate_performances(cv::Mat& GT_img, cv::Mat& r_img, float& mae, float& rmse) {
    cv::Mat g = GT_img.isContinuous() ? GT_img : GT_img.clone(), r = r_img.isContinuous() ? r_img : r_img.clone();
    dcmt::evaluate_performances(dcmt::view_of(g), dcmt::view_of(r), mae, rmse);
}
#endif  // DCMT_HAVE_OPENCV

#endif  // DCMT_IMG_COMPLETION_H_
