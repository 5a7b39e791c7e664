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
// the same on the reference's own containers: EntryType matrices (12-byte elements at `elem_stride` inside rows of
// `row_step` bytes, main_sl.cpp:23-26,:1165) and pitched CV_32FC1 matrices with the in-place semantics of :852 / :871
cudaError_t stereo_entries_derivatives(void* entries, size_t row_step, size_t elem_stride, int rows, int cols, cudaStream_t st);
cudaError_t stereo_entries_optimize_ig(const void* left, size_t left_row_step, const void* right, size_t right_row_step, size_t elem_stride,
                                       float* disp, size_t disp_pitch, int rows, int cols, int iters, float damp, float clip,
                                       cudaStream_t st);
cudaError_t stereo_initial_disparity_mat(const float* depth, size_t depth_pitch, float* disp, size_t disp_pitch, int rows, int cols,
                                         float baseline, float focal, cudaStream_t st);
cudaError_t stereo_retrieve_depth_mat(const float* disp, size_t disp_pitch, float* depth, size_t depth_pitch, int rows, int cols,
                                      float baseline, float focal, float clip, cudaStream_t st);
// cv::cvtColor(COLOR_BGR2GRAY), CV_8UC3 -> CV_8UC1 (main_sl.cpp:1167,1171); pitches / frame strides in bytes
cudaError_t stereo_bgr2gray(const uint8_t* bgr, size_t bgr_pitch, size_t bgr_fstride, uint8_t* gray, size_t gray_pitch, size_t gray_fstride,
                            int rows, int cols, int n_frames, cudaStream_t st);
cudaError_t stereo_refine(const float* depth_ig, const uint8_t* left, const uint8_t* right, float* depth_out,
                          float* disp_out, int rows, int cols, int n_frames, float baseline, float focal, float damp,
                          float err_clip, float depth_clip, int iters, int final_gauss, cudaStream_t st);
}  // namespace dcmt
