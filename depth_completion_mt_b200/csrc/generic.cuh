// generic.cuh -- host-side interface of the generic float32 pipeline (generic.cu).
#pragma once
#include "common.cuh"

namespace dcmt {

constexpr int kOracleStages = 10;  // stage numbering of oracle/dcmt_oracle.c

// One chunk of frames through the generic pipeline.  All pointers are device pointers.
struct GenericChunk {
    const float* in;
    size_t in_pitch, in_fstride;  // elements
    const int32_t* labels;        // contiguous rows*cols per frame (guided only)
    int n_clusters;
    bool guided;                  // interpolate_with_superpixels front (img_completion_lc.cpp:78-103)
    float* out;
    size_t out_pitch, out_fstride;  // elements
    int rows, cols, n_frames;
    int blur;                     // DCMT_BLUR_*
    // workspace (contiguous rows*cols*n_frames floats each)
    float* w1;
    float* w2;
    FrameCounters* ctr;           // n_frames
    unsigned int* minmax;         // 2 * n_frames (bilateral only)
    float* lut;                   // generic_lut_floats() * n_frames (bilateral only)
    int32_t* stats;               // optional, 4 * n_frames
    float* stages;                // optional debug snapshots (n_frames must be 1)
    uint32_t* stage_mask;         // host, optional
    bool skip_front;              // w1 already holds the A4 output (used by other front ends)
};

size_t generic_lut_floats();
cudaError_t generic_configure();
cudaError_t generic_run_chunk(const GenericChunk& c, cudaStream_t stream);

}  // namespace dcmt
