// evaluate.cuh -- masked error reductions of the reference's evaluation functions.
#pragma once
#include "common.cuh"

namespace dcmt {

// per-frame result, device or host memory (mirrors dcmt_eval_result in include/dcmt.h)
struct EvalResult {
    double count, sum_err, sum_abs, sum_sq;  // pixels in the mask, sum of (gt - r), of |gt - r|, of (gt - r)^2
    float mean_err, mae, rmse;               // sum_err / count, sum_abs / count, sqrt(sum_sq / count); NaN when count == 0
    int32_t pad;
};

constexpr int kEvalMaxBlocks = 128;  // partial sums per frame
size_t eval_partial_doubles(int n_frames);
// mode 0: mask = gt > tol (DC_lidar_only/main.cpp:16-34); mode 1: mask = gt > tol && r > tol (main_lc.cpp:85-116, main_sl.cpp:1031-1061)
cudaError_t eval_run(const float* gt, const float* r, int rows, int cols, size_t pitch, size_t fstride, int n_frames, float tol, int mode,
                     double* partials, EvalResult* out, cudaStream_t st);

}  // namespace dcmt
