// Test program for the cv::Mat half of include/img_completion.h: the calls of the three reference programs, written
// the way the reference writes them, on cv::Mat / Slic / EntryType -- compiled against the stand-in OpenCV and Eigen
// headers of oracle/refshim (containers only; test infrastructure) because the image has no OpenCV C++ headers, and
// linked against libdcmt.so (GPU box) or the CPU emulator build of the same sources (CPU suite).
//
//   lidar  ROWS COLS in.f32 out.f32 BLUR                         src/DC_lidar_only/main.cpp:93
//   lidar16 ROWS COLS in.u16 out.f32 BLUR                        main.cpp:75-93 on the PNG payload
//   guided ROWS COLS in.f32 labels_colmajor.i32 K out.f32 out_sp.f32     src/DC_lidar_camera/main_lc.cpp:219-220
//   stereo ROWS COLS dense.f32 left.u8 right.u8 disp.f32 depth.f32 entries_right.f32   src/DC_stereo_lidar/main_sl.cpp:1162-1246
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#ifdef DCMT_TEST_REFERENCE_SLIC_H
#include "slic.h"  // the reference's own class declaration (src/DC_lidar_camera/slic.h), where the tree is present
#else
// the public members of the reference's Slic (slic.h:30-71) that interpolate_with_superpixels reads
class Slic {
public:
    std::vector<std::vector<int>> clusters;    // [col][row]
    std::vector<std::vector<double>> centers;  // one per superpixel
};
#endif

// main_sl.cpp:23-26
struct EntryType {
    float value;
    Eigen::Vector2f derivative;
};

#include "img_completion.h"
#ifndef DCMT_HAVE_OPENCV
#error "the cv::Mat half of img_completion.h was not enabled"
#endif

template <class T>
static bool read_file(const char* path, T* dst, size_t n) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const bool ok = std::fread(dst, sizeof(T), n, f) == n;
    std::fclose(f);
    return ok;
}
static bool write_mat(const char* path, const cv::Mat& m) {  // rows of cols float32, whatever the step
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    for (int r = 0; r < m.rows; ++r) std::fwrite(m.data + (size_t)r * m.step, sizeof(float), (size_t)m.cols, f);
    std::fclose(f);
    return true;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string mode = argv[1];
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    try {
        if (mode == "lidar" && argc == 7) {
            cv::Mat projected_depths(rows, cols, CV_32FC1);
            if (!read_file(argv[4], reinterpret_cast<float*>(projected_depths.data), (size_t)rows * cols)) return 3;
            // a destination that is NOT a fresh matrix: wider parent, wrong type -- the reference replaces the header (:27)
            cv::Mat dense_r_img(rows + 3, cols + 5, CV_8UC1);
            std::string blur_type = argv[6];
            img_completion(projected_depths, dense_r_img, false, blur_type);  // main.cpp:93
            if (dense_r_img.rows != rows || dense_r_img.cols != cols || dense_r_img.type() != CV_32FC1) return 5;
            float mse = 0;
            evaluate_performance(projected_depths, dense_r_img, mse);  // main.cpp:101 (against its own input here)
            std::printf("%.9g\n", mse);
            return write_mat(argv[5], dense_r_img) ? 0 : 3;
        }
        if (mode == "lidar16" && argc == 7) {
            cv::Mat image_r(rows, cols, CV_16UC1);  // cv::imread(..., IMREAD_ANYDEPTH), main.cpp:75
            if (!read_file(argv[4], reinterpret_cast<uint16_t*>(image_r.data), (size_t)rows * cols)) return 3;
            cv::Mat dense_r_img;
            img_completion(image_r, dense_r_img, false, argv[6]);
            return write_mat(argv[5], dense_r_img) ? 0 : 3;
        }
        if (mode == "guided" && argc == 9) {
            cv::Mat projected_depths(rows, cols, CV_32FC1);
            if (!read_file(argv[4], reinterpret_cast<float*>(projected_depths.data), (size_t)rows * cols)) return 3;
            std::vector<int> lab((size_t)rows * cols);
            if (!read_file(argv[5], lab.data(), lab.size())) return 3;
            Slic slic;
            slic.clusters.assign((size_t)cols, std::vector<int>((size_t)rows));
            for (int c = 0; c < cols; ++c)
                for (int r = 0; r < rows; ++r) slic.clusters[c][r] = lab[(size_t)c * rows + r];
            slic.centers.assign((size_t)std::atoi(argv[6]), std::vector<double>(5, 0.0));
            cv::Mat dense_r_img, dense_r_img_sp;
            std::string blur_type = "gaussian";
            img_completion(projected_depths, dense_r_img, 0, blur_type);                                // main_lc.cpp:219
            interpolate_with_superpixels(slic, projected_depths, dense_r_img_sp, blur_type, 1);          // main_lc.cpp:220
            float mse = 0, mae = 0;
            evaluate_performance(projected_depths, dense_r_img_sp, mse, mae);                           // main_lc.cpp:224
            std::printf("%.9g %.9g\n", mse, mae);
            return write_mat(argv[7], dense_r_img) && write_mat(argv[8], dense_r_img_sp) ? 0 : 3;
        }
        if (mode == "stereo" && argc == 10) {
            cv::Mat dense_range_img(rows, cols, CV_32FC1);
            cv::Mat image_left_gray_entry(rows, cols, CV_8UC1), image_right_gray_entry(rows, cols, CV_8UC1);
            if (!read_file(argv[4], reinterpret_cast<float*>(dense_range_img.data), (size_t)rows * cols)) return 3;
            if (!read_file(argv[5], image_left_gray_entry.data, (size_t)rows * cols)) return 3;
            if (!read_file(argv[6], image_right_gray_entry.data, (size_t)rows * cols)) return 3;
            // main_sl.cpp:1162-1187
            cv::Mat entryMatrix_left(rows, cols, CV_32FC(sizeof(EntryType)));
            cv::Mat entryMatrix_right(rows, cols, CV_32FC(sizeof(EntryType)));
            for (int r = 0; r < rows; r++) {
                for (int c = 0; c < cols; c++) {
                    EntryType& entry_left = entryMatrix_left.at<EntryType>(r, c);
                    entry_left.value = (float)image_left_gray_entry.at<uchar>(r, c);
                    entry_left.derivative.x() = 0.0;
                    entry_left.derivative.y() = 0.0;
                    EntryType& entry_right = entryMatrix_right.at<EntryType>(r, c);
                    entry_right.value = (float)image_right_gray_entry.at<uchar>(r, c);
                    entry_right.derivative.x() = 0.0;
                    entry_right.derivative.y() = 0.0;
                }
            }
            cv::Mat disparity_IG = cv::Mat::zeros(rows, cols, CV_32F);
            calculateMeasuementDerivatives(entryMatrix_left);   // main_sl.cpp:1192
            calculateMeasuementDerivatives(entryMatrix_right);  // :1193
            get_initial_disparity(dense_range_img, disparity_IG);  // :1195
            // one probe of the per-sample accessor (main_sl.cpp:747), the way optimize_IG calls it (:813-816)
            {
                Eigen::Vector2f img_point(rows / 2, cols / 2 - 0.25f);
                float value = 0;
                Eigen::Vector2f derivative;
                const bool ok = calculateObservationDerivatives(entryMatrix_right, img_point, value, derivative);
                std::printf("%d %.9g %.9g %.9g\n", (int)ok, value, derivative[0], derivative[1]);
            }
            optimize_IG(entryMatrix_left, entryMatrix_right, disparity_IG);  // :1240
            cv::Mat optimized_depth = cv::Mat::zeros(rows, cols, CV_32F);
            retrieve_optimized_depth(disparity_IG, optimized_depth);  // :1246
            // the right entry matrix as (value, dx, dy) triples, for the derivative check
            FILE* f = std::fopen(argv[9], "wb");
            if (!f) return 3;
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c < cols; ++c) std::fwrite(&entryMatrix_right.at<EntryType>(r, c), sizeof(EntryType), 1, f);
            std::fclose(f);
            return write_mat(argv[7], disparity_IG) && write_mat(argv[8], optimized_depth) ? 0 : 3;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    return 2;
}
