// median_f32.cuh -- exact median of a 5x5 float window by forgetful selection (cv::medianBlur ksize 5,
// img_completion.cpp:170): keep 14 samples, drop the minimum and the maximum, add one, ... down to 3.
#pragma once
#include "common.cuh"

namespace dcmt {

__device__ __forceinline__ void cswap(float& a, float& b) {
    const float lo = fminf(a, b);
    b = fmaxf(a, b);
    a = lo;
}

// p points at the top-left sample of the window, `stride` elements between rows
__device__ __forceinline__ float median25(const float* __restrict__ p, int stride) {
    float w[25];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) w[dy * 5 + dx] = p[dy * stride + dx];
    float v[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) v[i] = w[i];
#pragma unroll
    for (int n = 14; n >= 3; --n) {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) cswap(v[i], v[n - 1 - i]);
#pragma unroll
        for (int i = 1; i < (n + 1) / 2; ++i) cswap(v[0], v[i]);
#pragma unroll
        for (int i = n / 2; i < n - 1; ++i) cswap(v[i], v[n - 1]);
        if (n > 3) v[0] = w[14 + (14 - n)];  // drop min (slot 0) and max (slot n-1), add the next sample
    }
    return v[1];
}

}  // namespace dcmt
