// rank_f32.cuh -- host-side interface of rank_f32.cu: the order-preserving dictionary that lets float32 frames run on
// the packed-uint16 kernels of fused_q8.cu.
#pragma once
#include "common.cuh"

namespace dcmt {

constexpr int kRankMaxValid = 32768;        // dictionary entries per frame (one CTA sorts them in 128 KB of shared memory)
constexpr int kRankFirstCode = 27;          // code of the smallest valid inverted value: valid <=> code >= 27, like fused_q8.cu
constexpr uint32_t kRankHundred = 65535u;   // code of the constant 100.0 (empty columns, img_completion.cpp:110)

cudaError_t rank_configure();
// Sorts the inverted valid values of every frame into lut[frame * kRankMaxValid ...] (lut_count[frame] entries) and writes
// the uint16 code plane (rows x code_pitch per frame, code_pitch a multiple of 8).  Frames that do not fit the
// dictionary or hold a NaN get ctr[frame].needs_generic = 1 (ctr must have been zeroed before).
cudaError_t rank_build(const float* in, size_t in_pitch, size_t in_fstride, int rows, int cols, int n_frames, float* lut, int* lut_count,
                       uint16_t* codes, size_t code_pitch, FrameCounters* ctr, cudaStream_t st);

}  // namespace dcmt
