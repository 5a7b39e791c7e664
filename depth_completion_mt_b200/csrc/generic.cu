// generic.cu -- generic float32 multi-kernel img_completion pipeline (any rows/cols, any finite input).
//
// This is the correctness anchor and the fallback of the fused q8 strip kernel (fused_q8.cu): one
// kernel per dependency region of the reference function
//   img_completion            /root/reference/src/DC_lidar_only/img_completion.cpp:17-204
//   interpolate_with_superpixels  .../DC_lidar_camera/img_completion_lc.cpp:34-203 (guided front)
//     k_front        :55-100   invert, 2-tap dilate, 5x5 close, 7x7 dilate + hole fill   (2-D tiles)
//     k_colextrap    :103-129  per-column extrapolation                                 (thread/column)
//     k_fill31       :131-144  31x31 dilate + hole fill, hole counters                  (2-D tiles)
//     k_fill31_loop  :146-166  further fill passes, only for frames that still have holes (1 CTA/frame)
//     k_tail         :170-202  5x5 median, Gaussian / none, final inversion             (2-D tiles)
//     k_minmax/k_bilateral_lut/k_tail_bilateral  :172-175 bilateral variant (intended out-of-place call)
// All min/max/select work is exact, so every stage up to the blur is bit-identical to OpenCV's.
#include "generic.cuh"
#include "median_f32.cuh"

#include <cmath>

namespace dcmt {

namespace {

constexpr int kThreads = 256;
constexpr int TH = 32;  // tile core rows
constexpr int TW = 64;  // tile core cols

// ------------------------------------------------------------------------------------------------
// k_front: A1..A4 on a (TH x TW) tile.  Region = core + {up 8, down 9, left 7, right 9}: the exact
// dependency cone of 2-tap (-1..+2 rows, +1..+2 cols), close5 (+-4), dilate7 (+-3).
// ------------------------------------------------------------------------------------------------
constexpr int F_UP = 8, F_DN = 9, F_LF = 7, F_RT = 9;
constexpr int F_RH = TH + F_UP + F_DN;
constexpr int F_RW = TW + F_LF + F_RT;

struct FrontArgs {
    const float* in;
    size_t in_pitch, in_fstride;
    const int32_t* labels;  // guided only, contiguous rows*cols per frame
    int n_clusters;
    float* out;  // contiguous rows*cols per frame
    int rows, cols;
};

template <bool kIsMax>
__device__ __forceinline__ float ext(float a, float b) {
    return kIsMax ? fmaxf(a, b) : fminf(a, b);
}

// One separable pass over the whole region: dst = extremum over [-R, R] along one axis of src at
// in-image positions, `fill` (identity of the NEXT operator) elsewhere.
template <int R, bool kHoriz, bool kIsMax>
__device__ __forceinline__ void region_pass(const float* __restrict__ src, float* __restrict__ dst, int gy0, int gx0,
                                            int rows, int cols, float fill) {
    for (int i = threadIdx.x; i < F_RH * F_RW; i += kThreads) {
        const int ry = i / F_RW, rx = i - ry * F_RW;
        const int gy = gy0 + ry, gx = gx0 + rx;
        float v = fill;
        if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
            v = kIsMax ? kAbsentMax : kAbsentMin;
#pragma unroll
            for (int d = -R; d <= R; ++d) {
                const int yy = kHoriz ? ry : ry + d, xx = kHoriz ? rx + d : rx;
                if (yy >= 0 && yy < F_RH && xx >= 0 && xx < F_RW) v = ext<kIsMax>(v, src[yy * F_RW + xx]);
            }
        }
        dst[i] = v;
    }
}

// Guided closed form for one pixel p with label c:
//   out(p) = min_{q in 5x5(p)} max_{r in 5x5(q)} max(tap(r-1row,+1col), tap(r+2rows,+2cols)),
//   tap(s) = label(s)==c ? D(s) : 0, absent outside the image.
// D / L are the region-local smem planes; (ry, rx) region coordinates of p.
__device__ __forceinline__ float guided_pixel(const float* __restrict__ D, const int* __restrict__ L, int ry, int rx,
                                              int gy, int gx, int rows, int cols, int c) {
    float acc[5][5];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) acc[a][b] = kAbsentMax;
#pragma unroll
    for (int i = 0; i < 9; ++i) {  // r rows p.y-4 .. p.y+4
        const int yr = gy - 4 + i;
        if (yr < 0 || yr >= rows) continue;  // uniform per thread; r outside the image is absent
        float r1[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {  // r cols p.x-4 .. p.x+4
            const int xr = gx - 4 + j;
            float v = kAbsentMax;
            if (xr >= 0 && xr < cols) {
                float t1 = kAbsentMax, t2 = kAbsentMax;
                if (yr - 1 >= 0 && xr + 1 < cols) {
                    const int s = (ry - 4 + i - 1) * F_RW + (rx - 4 + j + 1);
                    t1 = (L[s] == c) ? D[s] : 0.0f;
                }
                if (yr + 2 < rows && xr + 2 < cols) {
                    const int s = (ry - 4 + i + 2) * F_RW + (rx - 4 + j + 2);
                    t2 = (L[s] == c) ? D[s] : 0.0f;
                }
                v = fmaxf(t1, t2);
            }
            r1[j] = v;  // absent for r outside the image
        }
        float h[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) h[b] = fmaxf(fmaxf(fmaxf(r1[b], r1[b + 1]), fmaxf(r1[b + 2], r1[b + 3])), r1[b + 4]);
#pragma unroll
        for (int a = 0; a < 5; ++a)
            if (a <= i && i <= a + 4) {
#pragma unroll
                for (int b = 0; b < 5; ++b) acc[a][b] = fmaxf(acc[a][b], h[b]);
            }
    }
    float er = kAbsentMin;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const int yq = gy - 2 + a;
        if (yq < 0 || yq >= rows) continue;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const int xq = gx - 2 + b;
            if (xq >= 0 && xq < cols) er = fminf(er, acc[a][b]);
        }
    }
    return er;
}

template <bool kGuided>
__global__ void __launch_bounds__(kThreads) k_front(FrontArgs a) {
    DCMT_DYN_SMEM(float, smem);
    float* A = smem;
    float* B = A + F_RH * F_RW;
    float* Cc = B + F_RH * F_RW;
    int* Ls = reinterpret_cast<int*>(Cc + F_RH * F_RW);

    const int frame = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int gy0 = y0 - F_UP, gx0 = x0 - F_LF;
    const int rows = a.rows, cols = a.cols;
    const float* in = a.in + (size_t)frame * a.in_fstride;
    const size_t fpix = (size_t)rows * cols;

    // A1: load + invert (:55-67); out-of-image samples are absent for the 2-tap dilate
    for (int i = threadIdx.x; i < F_RH * F_RW; i += kThreads) {
        const int ry = i / F_RW, rx = i - ry * F_RW;
        const int gy = gy0 + ry, gx = gx0 + rx;
        const bool inimg = gy >= 0 && gy < rows && gx >= 0 && gx < cols;
        float v = kAbsentMax;
        if (inimg) v = invert_valid(__ldg(in + (size_t)gy * a.in_pitch + gx));
        A[i] = v;
        if (kGuided) Ls[i] = inimg ? __ldg(a.labels + (size_t)frame * fpix + (size_t)gy * cols + gx) : -1;
    }
    __syncthreads();

    if (!kGuided) {
        // A2: 2-tap dilate (:71-80): max(src(y-1,x+1), src(y+2,x+2)), absent taps = -FLT_MAX
        for (int i = threadIdx.x; i < F_RH * F_RW; i += kThreads) {
            const int ry = i / F_RW, rx = i - ry * F_RW;
            const int gy = gy0 + ry, gx = gx0 + rx;
            float v = kAbsentMax;  // identity of the following dilate outside the image
            if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
                const float t1 = (ry - 1 >= 0 && rx + 1 < F_RW) ? A[(ry - 1) * F_RW + rx + 1] : kAbsentMax;
                const float t2 = (ry + 2 < F_RH && rx + 2 < F_RW) ? A[(ry + 2) * F_RW + rx + 2] : kAbsentMax;
                v = fmaxf(t1, t2);
            }
            B[i] = v;
        }
        __syncthreads();
        // A3: close5 (:84-85) = dilate5 then erode5, separable
        region_pass<2, true, true>(B, A, gy0, gx0, rows, cols, kAbsentMax);
        __syncthreads();
        region_pass<2, false, true>(A, B, gy0, gx0, rows, cols, kAbsentMin);
        __syncthreads();
        region_pass<2, false, false>(B, A, gy0, gx0, rows, cols, kAbsentMin);
        __syncthreads();
        region_pass<2, true, false>(A, Cc, gy0, gx0, rows, cols, kAbsentMax);
        __syncthreads();
    } else {
        // guided front (img_completion_lc.cpp:78-103), needed on core +-3 (dilate7 reach)
        constexpr int GH = TH + 6, GW = TW + 6;
        for (int i = threadIdx.x; i < F_RH * F_RW; i += kThreads) Cc[i] = kAbsentMax;
        __syncthreads();
        for (int i = threadIdx.x; i < GH * GW; i += kThreads) {
            const int py = i / GW, px = i - py * GW;
            const int ry = F_UP - 3 + py, rx = F_LF - 3 + px;
            const int gy = gy0 + ry, gx = gx0 + rx;
            if (gy < 0 || gy >= rows || gx < 0 || gx >= cols) continue;
            const int c = Ls[ry * F_RW + rx];
            float v;
            if (c < 0 || c >= a.n_clusters) v = A[ry * F_RW + rx];
            else v = guided_pixel(A, Ls, ry, rx, gy, gx, rows, cols, c);
            Cc[ry * F_RW + rx] = v;
        }
        __syncthreads();
    }
    // A4: dilate7 + hole fill (:88-100); Cc holds the closed image D
    region_pass<3, true, true>(Cc, A, gy0, gx0, rows, cols, kAbsentMax);
    __syncthreads();
    float* out = a.out + (size_t)frame * fpix;
    for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
        const int cy = i / TW, cx = i - cy * TW;
        const int gy = y0 + cy, gx = x0 + cx;
        if (gy >= rows || gx >= cols) continue;
        const int ry = cy + F_UP, rx = cx + F_LF;
        float d = Cc[ry * F_RW + rx];
        if (is_hole(d)) {
            float m = kAbsentMax;
#pragma unroll
            for (int dy = -3; dy <= 3; ++dy) m = fmaxf(m, A[(ry + dy) * F_RW + rx]);
            d = m;
        }
        out[(size_t)gy * cols + gx] = d;
    }
}

// ------------------------------------------------------------------------------------------------
// k_colextrap: A5 (:103-129).  One thread per (frame, column); consecutive threads own consecutive
// columns so every row access is coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_colextrap(float* __restrict__ w, int rows, int cols, int n_frames) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_frames * cols) return;
    const int frame = (int)(t / cols), j = (int)(t - (long long)frame * cols);
    float* p = w + (size_t)frame * rows * cols + j;
    int last = 0, first = rows - 1;
    float mv = -1.0f, nv = 100.0f;
    bool found = false;
    for (int i = 0; i < rows; ++i) {
        const float v = p[(size_t)i * cols];
        if (is_valid(v)) {
            if (!found) { first = i; nv = v; found = true; }
            last = i;
            mv = v;
        }
    }
    for (int i = last; i < rows; ++i) p[(size_t)i * cols] = mv;
    for (int i = first; i >= 0; --i) p[(size_t)i * cols] = nv;  // second loop wins (:125-127)
}

// ------------------------------------------------------------------------------------------------
// k_fill31: A6 (:131-144).  dst = hole(src) ? dilate31(src) : src; counts holes before and after.
// ------------------------------------------------------------------------------------------------
constexpr int L_R = 15;
constexpr int L_RH = TH + 2 * L_R, L_RW = TW + 2 * L_R;

__global__ void __launch_bounds__(kThreads) k_fill31(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                                     int cols, FrameCounters* __restrict__ ctr) {
    DCMT_DYN_SMEM(float, smem);
    float* A = smem;            // region L_RH x L_RW
    float* B = A + L_RH * L_RW;  // row-max, L_RH x TW
    const int frame = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const size_t fpix = (size_t)rows * cols;
    const float* s = src + (size_t)frame * fpix;
    float* d = dst + (size_t)frame * fpix;

    int holes = 0;
    for (int i = threadIdx.x; i < L_RH * L_RW; i += kThreads) {
        const int ry = i / L_RW, rx = i - ry * L_RW;
        const int gy = y0 - L_R + ry, gx = x0 - L_R + rx;
        float v = kAbsentMax;
        if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
            v = s[(size_t)gy * cols + gx];
            if (ry >= L_R && ry < L_R + TH && rx >= L_R && rx < L_R + TW && is_hole(v)) ++holes;
        }
        A[i] = v;
    }
    const int tile_holes = __syncthreads_count(holes > 0);
    if (tile_holes == 0) {  // nothing to fill: copy the core through
        for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
            const int cy = i / TW, cx = i - cy * TW;
            const int gy = y0 + cy, gx = x0 + cx;
            if (gy < rows && gx < cols) d[(size_t)gy * cols + gx] = A[(cy + L_R) * L_RW + cx + L_R];
        }
        return;
    }
    for (int i = threadIdx.x; i < L_RH * TW; i += kThreads) {
        const int ry = i / TW, cx = i - ry * TW;
        const float* row = A + ry * L_RW + cx;  // window = region cols cx .. cx+30
        float m = row[0];
#pragma unroll
        for (int k = 1; k <= 2 * L_R; ++k) m = fmaxf(m, row[k]);
        B[i] = m;
    }
    __syncthreads();
    int remaining = 0;
    for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
        const int cy = i / TW, cx = i - cy * TW;
        const int gy = y0 + cy, gx = x0 + cx;
        if (gy >= rows || gx >= cols) continue;
        float v = A[(cy + L_R) * L_RW + cx + L_R];
        if (is_hole(v)) {
            float m = B[cy * TW + cx];
#pragma unroll
            for (int k = 1; k <= 2 * L_R; ++k) m = fmaxf(m, B[(cy + k) * TW + cx]);
            v = m;
            if (is_hole(v)) ++remaining;
        }
        d[(size_t)gy * cols + gx] = v;
    }
    // block reduction of the two counters
    __shared__ int s_h, s_r;
    if (threadIdx.x == 0) { s_h = 0; s_r = 0; }
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        holes += __shfl_down_sync(0xffffffffu, holes, o);
        remaining += __shfl_down_sync(0xffffffffu, remaining, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (holes) atomicAdd(&s_h, holes);
        if (remaining) atomicAdd(&s_r, remaining);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_h) atomicAdd(&ctr[frame].holes_after_extrapolation, s_h);
        if (s_r) atomicAdd(&ctr[frame].holes_after_first_fill, s_r);
    }
}

// ------------------------------------------------------------------------------------------------
// k_fill31_loop: A7 (:146-166) for the rare frame that still has holes after the first fill.
// One CTA per frame iterates whole-frame passes in global memory (L2-resident) until no hole is
// left; frames without holes return immediately.  img holds the frame (updated in place), tmp is a
// scratch plane for the row maxima.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_fill31_loop(float* __restrict__ img, float* __restrict__ tmp, int rows,
                                                      int cols, FrameCounters* __restrict__ ctr, int max_passes) {
    const int frame = blockIdx.x;
    if (ctr[frame].holes_after_first_fill == 0) return;  // uniform: written by the previous kernel
    const size_t fpix = (size_t)rows * cols;
    float* D = img + (size_t)frame * fpix;
    float* T = tmp + (size_t)frame * fpix;
    int passes = 0, any = 1;
    while (any && passes < max_passes) {
        for (size_t i = threadIdx.x; i < fpix; i += blockDim.x) {
            const int y = (int)(i / cols), x = (int)(i - (size_t)y * cols);
            const int lo = max(x - L_R, 0), hi = min(x + L_R, cols - 1);
            const float* row = D + (size_t)y * cols;
            float m = row[lo];
            for (int k = lo + 1; k <= hi; ++k) m = fmaxf(m, row[k]);
            T[i] = m;
        }
        __syncthreads();
        int remaining = 0;
        for (size_t i = threadIdx.x; i < fpix; i += blockDim.x) {
            const float v = D[i];
            if (!is_hole(v)) continue;
            const int y = (int)(i / cols), x = (int)(i - (size_t)y * cols);
            const int lo = max(y - L_R, 0), hi = min(y + L_R, rows - 1);
            float m = T[(size_t)lo * cols + x];
            for (int k = lo + 1; k <= hi; ++k) m = fmaxf(m, T[(size_t)k * cols + x]);
            D[i] = m;  // reads only T and its own D: in-place is safe
            if (is_hole(m)) ++remaining;
        }
        ++passes;
        any = __syncthreads_count(remaining > 0);  // also orders the D writes before the next row pass
    }
    if (threadIdx.x == 0) {
        ctr[frame].extra_passes = passes;
        ctr[frame].holes_remaining = any;
    }
}

// ------------------------------------------------------------------------------------------------
// k_tail: A8..A10 (:170-202): median5 (BORDER_REPLICATE), Gaussian 5x5 (REFLECT_101) where valid,
// final inversion.
// ------------------------------------------------------------------------------------------------
constexpr int T_RH = TH + 8, T_RW = TW + 8;  // input region (halo 4)
constexpr int T_MH = TH + 4, T_MW = TW + 4;  // median region (halo 2)

struct TailArgs {
    const float* src;  // contiguous
    float* out;
    size_t out_pitch, out_fstride;
    float* median_out;  // optional contiguous snapshot / bilateral input (nullptr = none)
    int rows, cols;
    int blur;  // DCMT_BLUR_*
    int write_final;  // 0: only produce median_out (bilateral first phase)
};

__global__ void __launch_bounds__(kThreads) k_tail(TailArgs a) {
    DCMT_DYN_SMEM(float, smem);
    float* R = smem;                 // T_RH x T_RW
    float* M = R + T_RH * T_RW;      // T_MH x T_MW
    float* G = M + T_MH * T_MW;      // T_MH x TW (row-filtered)
    const int frame = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int rows = a.rows, cols = a.cols;
    const size_t fpix = (size_t)rows * cols;
    const float* s = a.src + (size_t)frame * fpix;

    for (int i = threadIdx.x; i < T_RH * T_RW; i += kThreads) {
        const int ry = i / T_RW, rx = i - ry * T_RW;
        const int gy = clampi(y0 - 4 + ry, 0, rows - 1), gx = clampi(x0 - 4 + rx, 0, cols - 1);  // BORDER_REPLICATE
        R[i] = s[(size_t)gy * cols + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T_MH * T_MW; i += kThreads) {
        const int my = i / T_MW, mx = i - my * T_MW;
        const int gy = y0 - 2 + my, gx = x0 - 2 + mx;
        float m = 0.0f;
        if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) m = median25(R + my * T_RW + mx, T_RW);
        M[i] = m;
    }
    __syncthreads();
    if (a.median_out) {
        float* mo = a.median_out + (size_t)frame * fpix;
        for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
            const int cy = i / TW, cx = i - cy * TW;
            const int gy = y0 + cy, gx = x0 + cx;
            if (gy < rows && gx < cols) mo[(size_t)gy * cols + gx] = M[(cy + 2) * T_MW + cx + 2];
        }
    }
    if (!a.write_final) return;
    float* out = a.out + (size_t)frame * a.out_fstride;
    if (a.blur == 1) {
        // GaussianBlur 5x5 sigma 0 (:179): [1,4,6,4,1]/16 separable, REFLECT_101
        const float k0 = 0.375f, k1 = 0.25f, k2 = 0.0625f;
        for (int i = threadIdx.x; i < T_MH * TW; i += kThreads) {
            const int my = i / TW, cx = i - my * TW;
            const int gy = y0 - 2 + my, gx = x0 + cx;
            float g = 0.0f;
            if (gy >= 0 && gy < rows && gx < cols) {
                const float* row = M + my * T_MW;
                const int b = 2 - x0;  // region column of image column 0
                const float c0 = row[cx + 2];
                const float m1 = row[reflect101(gx - 1, cols) + b], p1 = row[reflect101(gx + 1, cols) + b];
                const float m2 = row[reflect101(gx - 2, cols) + b], p2 = row[reflect101(gx + 2, cols) + b];
                g = __fadd_rn(__fadd_rn(__fmul_rn(c0, k0), __fmul_rn(__fadd_rn(m1, p1), k1)),
                              __fmul_rn(__fadd_rn(m2, p2), k2));
            }
            G[i] = g;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
            const int cy = i / TW, cx = i - cy * TW;
            const int gy = y0 + cy, gx = x0 + cx;
            if (gy >= rows || gx >= cols) continue;
            float d = M[(cy + 2) * T_MW + cx + 2];
            if (is_valid(d)) {  // :181-188 masked copy
                const int b = 2 - y0;
                const float c0 = G[(cy + 2) * TW + cx];
                const float m1 = G[(reflect101(gy - 1, rows) + b) * TW + cx], p1 = G[(reflect101(gy + 1, rows) + b) * TW + cx];
                const float m2 = G[(reflect101(gy - 2, rows) + b) * TW + cx], p2 = G[(reflect101(gy + 2, rows) + b) * TW + cx];
                d = __fadd_rn(__fadd_rn(__fmul_rn(c0, k0), __fmul_rn(__fadd_rn(m1, p1), k1)),
                              __fmul_rn(__fadd_rn(m2, p2), k2));
            }
            out[(size_t)gy * a.out_pitch + gx] = invert_valid(d);  // :191-202
        }
    } else {
        for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
            const int cy = i / TW, cx = i - cy * TW;
            const int gy = y0 + cy, gx = x0 + cx;
            if (gy < rows && gx < cols) out[(size_t)gy * a.out_pitch + gx] = invert_valid(M[(cy + 2) * T_MW + cx + 2]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bilateral variant (:172-175; the reference's in-place call asserts inside OpenCV, the evident
// intent is the out-of-place filter): cv::bilateralFilter(d=5, sigmaColor=1.5, sigmaSpace=2.0) on
// CV_32FC1 = per-frame min/max, 4096-bin interpolated exp LUT over [0, max-min], 12 neighbour taps
// with r<=2 plus the centre, BORDER_REFLECT_101.
// ------------------------------------------------------------------------------------------------
constexpr int kLutBins = 1 << 12;
constexpr int kLutSize = kLutBins + 2;

__device__ __forceinline__ unsigned int f2ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void k_minmax_init(unsigned int* mm, int n_frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_frames) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

__global__ void __launch_bounds__(kThreads) k_minmax(const float* __restrict__ src, size_t fpix, unsigned int* __restrict__ mm) {
    const int frame = blockIdx.y;
    const float* s = src + (size_t)frame * fpix;
    unsigned int lo = 0xffffffffu, hi = 0u;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < fpix; i += (size_t)gridDim.x * kThreads) {
        const unsigned int o = f2ord(s[i]);
        lo = min(lo, o);
        hi = max(hi, o);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_down_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_down_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[2 * frame], lo);
        atomicMax(&mm[2 * frame + 1], hi);
    }
}

// one CTA per frame: LUT + scale.  lut layout per frame: [0]=scale_index, [1]=degenerate flag, [2..] table
__global__ void __launch_bounds__(kThreads) k_bilateral_lut(const unsigned int* __restrict__ mm, float* __restrict__ lut) {
    const int frame = blockIdx.x;
    float* L = lut + (size_t)frame * (kLutSize + 2);
    const float mn = ord2f(mm[2 * frame]), mx = ord2f(mm[2 * frame + 1]);
    const bool degenerate = fabs((double)mn - (double)mx) < (double)FLT_EPSILON;
    const float len = (float)((double)mx - (double)mn);
    const float scale_index = degenerate ? 0.0f : (float)kLutBins / len;
    if (threadIdx.x == 0) { L[0] = scale_index; L[1] = degenerate ? 1.0f : 0.0f; }
    if (degenerate) return;
    const double coeff = -0.5 / (1.5 * 1.5);
    // OpenCV stops evaluating exp once a table entry underflowed to 0 (monotone), the rest stays 0:
    // evaluating every entry gives the same table.
    for (int i = threadIdx.x; i < kLutSize; i += kThreads) {
        const double v = (double)i / (double)scale_index;
        L[2 + i] = (float)exp(v * v * coeff);
    }
}

struct BilateralArgs {
    const float* src;  // median output, contiguous
    const float* lut;
    float* out;
    size_t out_pitch, out_fstride;
    int rows, cols;
};

__global__ void __launch_bounds__(kThreads) k_tail_bilateral(BilateralArgs a) {
    const int frame = blockIdx.z;
    const int gx = blockIdx.x * 32 + (threadIdx.x & 31), gy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int rows = a.rows, cols = a.cols;
    if (gy >= rows || gx >= cols) return;
    const size_t fpix = (size_t)rows * cols;
    const float* s = a.src + (size_t)frame * fpix;
    const float* L = a.lut + (size_t)frame * (kLutSize + 2);
    const float scale_index = L[0];
    const float val0 = s[(size_t)gy * cols + gx];
    float res = val0;
    if (L[1] == 0.0f) {
        const double sc = -0.5 / (2.0 * 2.0);
        float sum = val0, wsum = 1.0f;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int r2 = dy * dy + dx * dx;
                if (r2 > 4 || r2 == 0) continue;
                const float sw = (float)exp((double)r2 * sc);  // constant-folded
                const float val = s[(size_t)reflect101(gy + dy, rows) * cols + reflect101(gx + dx, cols)];
                float alpha = __fmul_rn(fabsf(__fsub_rn(val, val0)), scale_index);
                const int idx = (int)floorf(alpha);
                alpha = __fsub_rn(alpha, (float)idx);
                const float e0 = L[2 + idx], e1 = L[2 + idx + 1];
                const float w = __fmul_rn(sw, __fadd_rn(e0, __fmul_rn(alpha, __fsub_rn(e1, e0))));
                sum = __fadd_rn(sum, __fmul_rn(val, w));
                wsum = __fadd_rn(wsum, w);
            }
        res = __fdiv_rn(sum, wsum);
    }
    a.out[(size_t)frame * a.out_fstride + (size_t)gy * a.out_pitch + gx] = invert_valid(res);
}

__global__ void k_zero_counters(FrameCounters* c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = FrameCounters{};
}

__global__ void k_write_stats(const FrameCounters* __restrict__ c, int32_t* __restrict__ stats, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    stats[4 * i + 0] = c[i].extra_passes + 1;  // the reference always runs one more (no-op) pass
    stats[4 * i + 1] = c[i].holes_after_first_fill;
    stats[4 * i + 2] = c[i].holes_after_extrapolation;
    stats[4 * i + 3] = c[i].path;
}

}  // namespace

static size_t generic_front_smem(bool guided) { return (size_t)F_RH * F_RW * sizeof(float) * (guided ? 4 : 3); }
static size_t fill31_smem() { return ((size_t)L_RH * L_RW + (size_t)L_RH * TW) * sizeof(float); }
static size_t tail_smem() { return ((size_t)T_RH * T_RW + (size_t)T_MH * T_MW + (size_t)T_MH * TW) * sizeof(float); }

size_t generic_lut_floats() { return (size_t)kLutSize + 2; }

cudaError_t generic_configure() {
    cudaError_t e;
    // per-device function attributes: cheap, so set on every call (the current device may differ)
    if ((e = cudaFuncSetAttribute(k_front<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)generic_front_smem(false))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_front<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)generic_front_smem(true))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_fill31, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill31_smem())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem())) != cudaSuccess) return e;
    return cudaSuccess;
}

static cudaError_t snapshot(const GenericChunk& c, int stage, const float* src, cudaStream_t st) {
    if (!c.stages) return cudaSuccess;
    if (c.stage_mask) *c.stage_mask |= 1u << stage;
    return cudaMemcpyAsync(c.stages + (size_t)stage * c.rows * c.cols, src, (size_t)c.rows * c.cols * sizeof(float),
                           cudaMemcpyDeviceToDevice, st);
}

#define DCMT_TRY(x)                      \
    do {                                 \
        cudaError_t e__ = (x);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

cudaError_t generic_run_chunk(const GenericChunk& c, cudaStream_t st) {
    const int rows = c.rows, cols = c.cols, nf = c.n_frames;
    if (nf <= 0) return cudaSuccess;
    if (nf > 65535) return cudaErrorInvalidValue;  // grid.z limit; the caller chunks
    const dim3 tiles((cols + TW - 1) / TW, (rows + TH - 1) / TH, nf);
    const size_t fpix = (size_t)rows * cols;

    DCMT_LAUNCH(k_zero_counters, dim3((nf + 127) / 128), dim3(128), 0, st, c.ctr, nf);
    if (!c.skip_front) {
        FrontArgs fa{c.in, c.in_pitch, c.in_fstride, c.labels, c.n_clusters, c.w1, rows, cols};
        if (c.guided) DCMT_LAUNCH(k_front<true>, tiles, dim3(kThreads), generic_front_smem(true), st, fa);
        else DCMT_LAUNCH(k_front<false>, tiles, dim3(kThreads), generic_front_smem(false), st, fa);
    }
    DCMT_TRY(cudaGetLastError());
    DCMT_TRY(snapshot(c, 3, c.w1, st));
    {
        const long long nt = (long long)nf * cols;
        DCMT_LAUNCH(k_colextrap, dim3((unsigned)((nt + 127) / 128)), dim3(128), 0, st, c.w1, rows, cols, nf);
    }
    DCMT_TRY(snapshot(c, 4, c.w1, st));
    DCMT_LAUNCH(k_fill31, tiles, dim3(kThreads), fill31_smem(), st, c.w1, c.w2, rows, cols, c.ctr);
    DCMT_TRY(cudaGetLastError());
    DCMT_TRY(snapshot(c, 5, c.w2, st));
    {
        const int max_passes = (rows > cols ? rows : cols) / L_R + 2;
        DCMT_LAUNCH(k_fill31_loop, dim3(nf), dim3(1024), 0, st, c.w2, c.w1, rows, cols, c.ctr, max_passes);
    }
    DCMT_TRY(snapshot(c, 6, c.w2, st));
    if (c.blur != 2) {
        TailArgs ta{c.w2, c.out, c.out_pitch, c.out_fstride, c.stages ? c.w1 : nullptr, rows, cols, c.blur, 1};
        DCMT_LAUNCH(k_tail, tiles, dim3(kThreads), tail_smem(), st, ta);
        DCMT_TRY(cudaGetLastError());
        DCMT_TRY(snapshot(c, 7, c.w1, st));
    } else {
        TailArgs ta{c.w2, c.out, c.out_pitch, c.out_fstride, c.w1, rows, cols, c.blur, 0};
        DCMT_LAUNCH(k_tail, tiles, dim3(kThreads), tail_smem(), st, ta);
        DCMT_TRY(cudaGetLastError());
        DCMT_TRY(snapshot(c, 7, c.w1, st));
        DCMT_LAUNCH(k_minmax_init, dim3((nf + 127) / 128), dim3(128), 0, st, c.minmax, nf);
        int gx = (int)((fpix + (size_t)kThreads * 8 - 1) / ((size_t)kThreads * 8));
        if (gx > 1024) gx = 1024;
        if (gx < 1) gx = 1;
        DCMT_LAUNCH(k_minmax, dim3(gx, nf), dim3(kThreads), 0, st, c.w1, fpix, c.minmax);
        DCMT_LAUNCH(k_bilateral_lut, dim3(nf), dim3(kThreads), 0, st, c.minmax, c.lut);
        BilateralArgs ba{c.w1, c.lut, c.out, c.out_pitch, c.out_fstride, rows, cols};
        DCMT_LAUNCH(k_tail_bilateral, dim3((cols + 31) / 32, (rows + 7) / 8, nf), dim3(kThreads), 0, st, ba);
        DCMT_TRY(cudaGetLastError());
    }
    if (c.stats) DCMT_LAUNCH(k_write_stats, dim3((nf + 127) / 128), dim3(128), 0, st, c.ctr, c.stats, nf);
    DCMT_TRY(cudaGetLastError());
    if (c.stages && c.out_pitch == (size_t)cols) DCMT_TRY(snapshot(c, 9, c.out, st));
    return cudaSuccess;
}

}  // namespace dcmt
