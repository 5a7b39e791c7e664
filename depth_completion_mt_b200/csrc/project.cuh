// project.cuh -- LiDAR point cloud -> sparse depth image -> cv::normalize.
#pragma once
#include "common.cuh"

namespace dcmt {

// device workspace of one projection call
struct ProjectWork {
    unsigned long long* keys;  // rows * cols: ((point index + 1) << 32) | depth bits, 0 = no point
    unsigned* minmax;          // 8 words: [2] number of points that landed, [3] / [4] ordered bits of the image min / max
};
size_t project_key_count(int rows, int cols);
// T: 4x4 row-major (rows 0..2 used), P: 3x4 row-major.  projected / normalized may be null.
cudaError_t project_run(const float* points, int n_points, const float* T, const float* P, int rows, int cols, float* projected,
                        float* normalized, float norm_a, float norm_b, int32_t* n_projected, const ProjectWork& w, cudaStream_t st);

// The same for n_clouds clouds in four launches: cloud c holds counts_dev[c] points (device array; null: n_points each) at
// points + c * cloud_stride_points * 4 floats, n_points bounds every count; the work arrays hold n_clouds x (keys, 8 min / max
// words); images and n_projected are n_clouds planes / ints.
cudaError_t project_run_batch(const float* points, int n_points, const int32_t* counts_dev, size_t cloud_stride_points, int n_clouds,
                              const float* T, const float* P, int rows, int cols, float* projected, float* normalized, float norm_a,
                              float norm_b, int32_t* n_projected, const ProjectWork& w, cudaStream_t st);

}  // namespace dcmt
