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

}  // namespace dcmt
