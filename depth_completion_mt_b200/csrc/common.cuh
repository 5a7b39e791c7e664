// common.cuh -- shared definitions for the dcmt CUDA kernels (sm_100a).
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#ifdef DCMT_EMU
// CPU emulation of the CUDA execution model, tests only (tests/emu/cuda_emu.h); never defined for libdcmt.so
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define DCMT_DYN_SMEM(type, name)                                         \
    extern __shared__ __align__(16) unsigned char dcmt_dyn_smem_raw[];    \
    type* name = reinterpret_cast<type*>(dcmt_dyn_smem_raw)
// 16-byte asynchronous global -> shared copy (LDGSTS): no registers, many copies in flight per thread
#define DCMT_CP_ASYNC_16(smem_ptr, gmem_ptr)                                                                        \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_ptr)), \
                 "l"(gmem_ptr)                                                                                     \
                 : "memory")
#define DCMT_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory")
namespace dcmt { void note_launch(); }  // launch counter behind dcmt_launch_count() (api.cu)
#define DCMT_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (dcmt::note_launch(), kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#endif

namespace dcmt {

// img_completion.cpp:59,96,113,140,154,184,194 compare float against the double literal 0.1:
//   (double)d > 0.1  <=>  d >= 0.1f      (valid pixel)
//   (double)d < 0.1  <=>  d <  0.1f      (hole)
// 0.1f is the float immediately above the real number 0.1, so the two predicates are exact complements.
__device__ __forceinline__ bool is_valid(float d) { return d >= 0.1f; }
__device__ __forceinline__ bool is_hole(float d) { return d < 0.1f; }

constexpr float kMaxDepth = 100.0f;  // img_completion.cpp:23
constexpr float kAbsentMax = -FLT_MAX;  // OpenCV morphologyDefaultBorderValue for dilate
constexpr float kAbsentMin = FLT_MAX;   // ... and for erode

// img_completion.cpp:55-67 / :191-202
__device__ __forceinline__ float invert_valid(float d) { return is_valid(d) ? kMaxDepth - d : d; }

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// per-frame counters kept in the workspace (int32 each)
struct FrameCounters {
    int holes_after_extrapolation;  // counted by the first 31x31 fill before filling (stats[2])
    int holes_after_first_fill;     // left after the first 31x31 fill (stats[1])
    int holes_remaining;            // != 0 only if the fill loop hit its pass bound
    int extra_passes;               // effective passes after the first fill
    int path;                       // 0 generic, 1 fused q8
    int needs_generic;              // fused kernel could not finish this frame
    int pad[2];
};

// plain description of a batch of row-major float frames
struct Batch {
    int rows, cols, n_frames;
    size_t pitch;         // elements between rows
    size_t frame_stride;  // elements between frames
};

}  // namespace dcmt
