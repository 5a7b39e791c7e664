// stereo.cuh -- host-side interface of stereo.cu (device pointers, contiguous planes).
#pragma once
#include "common.cuh"

namespace dcmt {
cudaError_t stereo_measurement_derivatives(const float* val, float* dx, float* dy, int rows, int cols, int n_frames,
                                           cudaStream_t st);
cudaError_t stereo_get_initial_disparity(const float* depth, float* disp, int rows, int cols, int n_frames, float baseline,
                                         float focal, cudaStream_t st);
cudaError_t stereo_optimize_ig(const float* vl, const float* vr, float* disp, int rows, int cols, int n_frames, int iters,
                               float damp, float clip, cudaStream_t st);
cudaError_t stereo_retrieve_depth(const float* disp, float* depth, int rows, int cols, int n_frames, float baseline,
                                  float focal, float clip, cudaStream_t st);
cudaError_t stereo_refine(const float* depth_ig, const uint8_t* left, const uint8_t* right, float* depth_out,
                          float* disp_out, int rows, int cols, int n_frames, float baseline, float focal, float damp,
                          float err_clip, float depth_clip, int iters, int final_gauss, cudaStream_t st);
}  // namespace dcmt
