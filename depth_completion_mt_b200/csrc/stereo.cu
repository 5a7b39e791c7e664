// stereo.cu -- stereo/LiDAR disparity refinement of /root/reference/src/DC_stereo_lidar/main_sl.cpp.
//
//   k_measurement_derivatives   calculateMeasuementDerivatives  :715-745   (a4)
//   k_get_initial_disparity     get_initial_disparity           :846-861   (a5)
//   k_optimize_ig               optimize_IG + calculateObservationDerivatives :747-843 (a6, a7)
//   k_retrieve_depth            retrieve_optimized_depth        :863-885   (a8)
//   k_stereo_refine             the whole sequence :1165-1253 fused per tile: u8 gray -> entries,
//                               initial disparity, 4 damped Gauss-Newton steps, depth, 5x5 Gaussian (a9)
//
// Every pixel is independent (an iteration reads/writes only its own disparity), so the k loop of
// optimize_IG is hoisted into the thread.  All float arithmetic uses explicit round-to-nearest
// intrinsics in source order (no FMA contraction) -- the reference is built for baseline x86-64,
// which has no FMA -- so results are bit-identical to the scalar C++.  Reads the reference performs
// at column index == cols (undefined behaviour there, :763-779) are defined as 0.
#include "stereo.cuh"

namespace dcmt {
namespace {

constexpr int kThreads = 256;
constexpr int TH = 32, TW = 128;  // tile core of the fused kernel (2-pixel halo for the Gaussian: 16 % extra refinement work)

// dx of calculateMeasuementDerivatives at (r, c): 0 on the 1-px border (:719,:724)
template <class Plane>
__device__ __forceinline__ float deriv_x(const Plane& v, int r, int c, int rows, int cols) {
    if (r < 1 || r >= rows - 1 || c < 1 || c >= cols - 1) return 0.0f;
    // `.5 * v_c1 - .5 * v_c0` is double arithmetic, stored to float (:742)
    return __double2float_rn(__dsub_rn(0.5 * (double)v(r, c + 1), 0.5 * (double)v(r, c - 1)));
}

struct PlaneF32 {
    const float* p;
    int cols;
    __device__ __forceinline__ float operator()(int r, int c) const { return __ldg(p + (size_t)r * cols + c); }
};
// dx entries recomputed from the value plane: what calculateMeasuementDerivatives stored at (r, c)
template <class Plane>
struct GradFromValues {
    const Plane& v;
    int rows, cols;
    __device__ __forceinline__ float operator()(int r, int c) const { return deriv_x(v, r, c, rows, cols); }
};

// One optimize_IG pixel: `iters` damped Gauss-Newton steps on the disparity d of pixel (i, j).  `grad(r, c)` is the
// derivative.x() entry of the right image (recomputed from the values, or read from the caller's EntryType matrix).
template <class Plane, class Grad>
__device__ __forceinline__ float refine_pixel(const Plane& right, const Grad& grad, float left_val, float d, int i, int j, int rows,
                                              int cols, int iters, float damp, float clip) {
    for (int k = 0; k < iters; ++k) {
        if (d == 0.0f) break;                       // :817 `disparity != 0` (never changes once 0)
        const float c = __fsub_rn((float)j, d);     // :812 float pixel_right = j - disp
        if (c != c) break;                          // NaN: (int) conversion is INT_MIN on x86 -> rejected
        const int c0 = __double2int_rz((double)c + 0.5);  // :759 (int)(c + 0.5), double add, truncation
        // :763-772  (row tests never fire for 0 <= i < rows); saturated conversions are rejected like
        // the x86 "integer indefinite" value
        if (c0 < 0 || c0 > cols || c0 == 2147483647 || c0 + 1 > cols) continue;
        const int c1 = c0 + 1;
        const float dc = __fsub_rn(c, (float)c0);                       // :787
        const float dc1 = __double2float_rn(__dsub_rn(1.0, (double)dc));  // :789 `1. - dc` in double
        const float p00 = c0 < cols ? right(i, c0) : 0.0f;
        const float p01 = c1 < cols ? right(i, c1) : 0.0f;
        const float g00 = c0 < cols ? grad(i, c0) : 0.0f;
        const float g01 = c1 < cols ? grad(i, c1) : 0.0f;
        // :794-797 with dr == 0, dr1 == 1: the second-row term contributes an exact zero
        const float value = __fadd_rn(__fmul_rn(p00, dc1), __fmul_rn(p01, dc));
        const float gx = __fadd_rn(__fmul_rn(g00, dc1), __fmul_rn(g01, dc));
        float err = __fsub_rn(value, left_val);     // :819
        if (err > clip) err = clip;                 // :821-826
        if (err < -clip) err = -clip;
        const float jcr = -gx;                      // :831-833  J = -1
        const float H = __fadd_rn(__fmul_rn(jcr, jcr), damp);
        const float b = __fmul_rn(jcr, err);
        const float dd = __fdiv_rn(-b, H);          // :837
        d = __fadd_rn(d, dd);                       // :838
    }
    return d;
}

// refine_pixel for gray uint8 planes (the fused kernel): the same arithmetic with the four neighbouring bytes of the
// right image row loaded once per iteration.  `.5 * v[c+1] - .5 * v[c-1]` (:742, double in the source) is exact in
// float for byte values, so the derivative needs no double here.
__device__ __forceinline__ float refine_pixel_u8(const uint8_t* __restrict__ rrow, bool row_inner, float left_val, float d, int j, int cols,
                                                 int iters, float damp, float clip) {
    for (int k = 0; k < iters; ++k) {
        if (d == 0.0f) break;
        const float c = __fsub_rn((float)j, d);
        if (c != c) break;
        const int c0 = __double2int_rz((double)c + 0.5);
        if (c0 < 0 || c0 > cols || c0 == 2147483647 || c0 + 1 > cols) continue;
        const float dc = __fsub_rn(c, (float)c0);
        const float dc1 = __double2float_rn(__dsub_rn(1.0, (double)dc));
        float vm1, v0, v1, v2;  // right(i, c0 - 1 .. c0 + 2), 0 where the column does not exist
        if (c0 >= 1 && c0 + 2 < cols) {
            vm1 = (float)__ldg(rrow + c0 - 1); v0 = (float)__ldg(rrow + c0); v1 = (float)__ldg(rrow + c0 + 1); v2 = (float)__ldg(rrow + c0 + 2);
        } else {
            vm1 = c0 >= 1 ? (float)__ldg(rrow + c0 - 1) : 0.0f;
            v0 = c0 < cols ? (float)__ldg(rrow + c0) : 0.0f;
            v1 = c0 + 1 < cols ? (float)__ldg(rrow + c0 + 1) : 0.0f;
            v2 = c0 + 2 < cols ? (float)__ldg(rrow + c0 + 2) : 0.0f;
        }
        // derivative entries are 0 on the 1-pixel image border (:719,:724) and where the column does not exist
        const float g00 = (row_inner && c0 >= 1 && c0 < cols - 1) ? __fsub_rn(__fmul_rn(0.5f, v1), __fmul_rn(0.5f, vm1)) : 0.0f;
        const float g01 = (row_inner && c0 + 1 < cols - 1) ? __fsub_rn(__fmul_rn(0.5f, v2), __fmul_rn(0.5f, v0)) : 0.0f;
        const float value = __fadd_rn(__fmul_rn(v0, dc1), __fmul_rn(v1, dc));
        const float gx = __fadd_rn(__fmul_rn(g00, dc1), __fmul_rn(g01, dc));
        float err = __fsub_rn(value, left_val);
        if (err > clip) err = clip;
        if (err < -clip) err = -clip;
        const float jcr = -gx;
        const float H = __fadd_rn(__fmul_rn(jcr, jcr), damp);
        const float b = __fmul_rn(jcr, err);
        d = __fadd_rn(d, __fdiv_rn(-b, H));
    }
    return d;
}

__device__ __forceinline__ float initial_disparity(float depth, float bf) {
    return depth > 0.0f ? __fdiv_rn(bf, depth) : 0.0f;  // :852-856
}
__device__ __forceinline__ float depth_from_disparity(float disp, float bf, float clip) {
    if (!(disp > 0.0f)) return 0.0f;  // :871, output pre-zeroed (:1244)
    float depth = __fdiv_rn(bf, disp);
    if (depth > clip) depth = clip;  // :875-879
    return depth;
}

__global__ void __launch_bounds__(kThreads) k_measurement_derivatives(const float* __restrict__ val, float* __restrict__ dx,
                                                                      float* __restrict__ dy, int rows, int cols,
                                                                      long long total) {
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const size_t fpix = (size_t)rows * cols;
    const size_t f = t / fpix, o = t - f * fpix;
    const int r = (int)(o / cols), c = (int)(o - (size_t)r * cols);
    const float* v = val + f * fpix;
    float gx = 0.0f, gy = 0.0f;
    if (r >= 1 && r < rows - 1 && c >= 1 && c < cols - 1) {
        gx = __double2float_rn(__dsub_rn(0.5 * (double)v[o + 1], 0.5 * (double)v[o - 1]));
        gy = __double2float_rn(__dsub_rn(0.5 * (double)v[o + cols], 0.5 * (double)v[o - cols]));
    }
    dx[t] = gx;
    if (dy) dy[t] = gy;
}

__global__ void __launch_bounds__(kThreads) k_get_initial_disparity(const float* __restrict__ depth, float* __restrict__ disp,
                                                                    long long total, float bf) {
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t < total) disp[t] = initial_disparity(depth[t], bf);
}

__global__ void __launch_bounds__(kThreads) k_retrieve_depth(const float* __restrict__ disp, float* __restrict__ depth,
                                                             long long total, float bf, float clip) {
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t < total) depth[t] = depth_from_disparity(disp[t], bf, clip);
}

__global__ void __launch_bounds__(kThreads) k_optimize_ig(const float* __restrict__ vl, const float* __restrict__ vr,
                                                          float* __restrict__ disp, int rows, int cols, long long total,
                                                          int iters, float damp, float clip) {
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const size_t fpix = (size_t)rows * cols;
    const size_t f = t / fpix, o = t - f * fpix;
    const int i = (int)(o / cols), j = (int)(o - (size_t)i * cols);
    const PlaneF32 right{vr + f * fpix, cols};
    disp[t] = refine_pixel(right, GradFromValues<PlaneF32>{right, rows, cols}, vl[t], disp[t], i, j, rows, cols, iters, damp, clip);
}

// ---- the reference's own containers (a3): EntryType {float value; Eigen::Vector2f derivative;} (main_sl.cpp:23-26) stored
// in a cv::Mat of type CV_32FC(sizeof(EntryType)) (:1165,:1169).  at<EntryType>(r, c) addresses
// data + r * step + c * sizeof(EntryType): 12-byte elements packed at the start of rows that are four times too wide.
struct EntriesView {
    char* base;
    size_t row_step, elem_stride;  // bytes
    __device__ __forceinline__ float* at(int r, int c) const { return reinterpret_cast<float*>(base + (size_t)r * row_step + (size_t)c * elem_stride); }
};
struct EntryValues {
    EntriesView e;
    __device__ __forceinline__ float operator()(int r, int c) const { return e.at(r, c)[0]; }
};
struct EntryDx {
    EntriesView e;
    __device__ __forceinline__ float operator()(int r, int c) const { return e.at(r, c)[1]; }
};

// calculateMeasuementDerivatives on an EntryType matrix, in place (:715-745): interior entries get (dx, dy), the
// one-pixel border keeps whatever the caller stored there
__global__ void __launch_bounds__(kThreads) k_entries_derivatives(EntriesView e, int rows, int cols) {
    const int c = blockIdx.x * kThreads + threadIdx.x, r = blockIdx.y;
    if (r < 1 || r >= rows - 1 || c < 1 || c >= cols - 1) return;
    float* o = e.at(r, c);
    o[1] = __double2float_rn(__dsub_rn(0.5 * (double)e.at(r, c + 1)[0], 0.5 * (double)e.at(r, c - 1)[0]));
    o[2] = __double2float_rn(__dsub_rn(0.5 * (double)e.at(r + 1, c)[0], 0.5 * (double)e.at(r - 1, c)[0]));
}

// optimize_IG on EntryType matrices (:804-843): value of the left entry, value and STORED derivative.x() of the right one
__global__ void __launch_bounds__(kThreads) k_entries_optimize_ig(EntriesView left, EntriesView right, float* __restrict__ disp,
                                                                  size_t disp_pitch, int rows, int cols, int iters, float damp,
                                                                  float clip) {
    const int j = blockIdx.x * kThreads + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    float* d = disp + (size_t)i * disp_pitch + j;
    *d = refine_pixel(EntryValues{right}, EntryDx{right}, left.at(i, j)[0], *d, i, j, rows, cols, iters, damp, clip);
}

// get_initial_disparity / retrieve_optimized_depth with the reference's in-place semantics: pixels that fail the test
// (depth > 0, disparity > 0) keep what the output matrix held (:852, :871)
__global__ void __launch_bounds__(kThreads) k_initial_disparity_mat(const float* __restrict__ depth, size_t depth_pitch,
                                                                    float* __restrict__ disp, size_t disp_pitch, int cols, float bf) {
    const int j = blockIdx.x * kThreads + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const float z = depth[(size_t)i * depth_pitch + j];
    if (z > 0.0f) disp[(size_t)i * disp_pitch + j] = __fdiv_rn(bf, z);
}
__global__ void __launch_bounds__(kThreads) k_retrieve_depth_mat(const float* __restrict__ disp, size_t disp_pitch,
                                                                 float* __restrict__ depth, size_t depth_pitch, int cols, float bf,
                                                                 float clip) {
    const int j = blockIdx.x * kThreads + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const float d = disp[(size_t)i * disp_pitch + j];
    if (d > 0.0f) depth[(size_t)i * depth_pitch + j] = depth_from_disparity(d, bf, clip);
}

struct RefineArgs {
    const float* depth_ig;
    const uint8_t* left;
    const uint8_t* right;
    float* depth_out;
    float* disp_out;  // optional
    int rows, cols;
    float bf, damp, err_clip, depth_clip;
    int iters, final_gauss;
};

// Fused: each CTA refines a (TH+4) x (TW+4) patch (2-px halo for the Gaussian) and blurs it on chip.
__global__ void __launch_bounds__(kThreads) k_stereo_refine(RefineArgs a) {
    constexpr int PH = TH + 4, PW = TW + 4;
    __shared__ float P[PH * PW];   // refined depth, halo 2
    __shared__ float G[PH * TW];   // row-filtered
    const int frame = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int rows = a.rows, cols = a.cols;
    const size_t fpix = (size_t)rows * cols;
    const float* dig = a.depth_ig + (size_t)frame * fpix;
    const uint8_t* left = a.left + (size_t)frame * fpix;
    const uint8_t* right = a.right + (size_t)frame * fpix;
    const int halo = a.final_gauss ? 2 : 0;

    for (int idx = threadIdx.x; idx < PH * PW; idx += kThreads) {
        const int py = idx / PW, px = idx - py * PW;
        const int gy = y0 - 2 + py, gx = x0 - 2 + px;
        float depth = 0.0f;
        const bool core = py >= 2 && py < 2 + TH && px >= 2 && px < 2 + TW;
        const bool wanted = py >= 2 - halo && py < 2 + TH + halo && px >= 2 - halo && px < 2 + TW + halo;
        if (wanted && gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
            const int o = gy * cols + gx;  // frames are < 2^30 pixels
            float d = initial_disparity(__ldg(dig + o), a.bf);
            d = refine_pixel_u8(right + gy * cols, gy >= 1 && gy < rows - 1, (float)__ldg(left + o), d, gx, cols, a.iters, a.damp, a.err_clip);
            depth = depth_from_disparity(d, a.bf, a.depth_clip);
            if (core && a.disp_out) a.disp_out[(size_t)frame * fpix + (size_t)gy * cols + gx] = d;
        }
        P[idx] = depth;
    }
    __syncthreads();
    float* out = a.depth_out + (size_t)frame * fpix;
    if (!a.final_gauss) {
        for (int idx = threadIdx.x; idx < TH * TW; idx += kThreads) {
            const int cy = idx / TW, cx = idx - cy * TW;
            const int gy = y0 + cy, gx = x0 + cx;
            if (gy < rows && gx < cols) out[(size_t)gy * cols + gx] = P[(cy + 2) * PW + cx + 2];
        }
        return;
    }
    // cv::GaussianBlur(5x5, sigma 0) (:1253): [1,4,6,4,1]/16 separable, BORDER_REFLECT_101
    const float k0 = 0.375f, k1 = 0.25f, k2 = 0.0625f;
    for (int idx = threadIdx.x; idx < PH * TW; idx += kThreads) {
        const int py = idx / TW, cx = idx - py * TW;
        const int gy = y0 - 2 + py, gx = x0 + cx;
        float g = 0.0f;
        if (gy >= 0 && gy < rows && gx < cols) {
            const float* row = P + py * PW;
            const int b = 2 - x0;
            const float c0 = row[cx + 2];
            float m1, p1, m2, p2;
            if (gx >= 2 && gx + 2 < cols) {
                m1 = row[cx + 1]; p1 = row[cx + 3]; m2 = row[cx]; p2 = row[cx + 4];
            } else {
                m1 = row[reflect101(gx - 1, cols) + b]; p1 = row[reflect101(gx + 1, cols) + b];
                m2 = row[reflect101(gx - 2, cols) + b]; p2 = row[reflect101(gx + 2, cols) + b];
            }
            g = __fadd_rn(__fadd_rn(__fmul_rn(c0, k0), __fmul_rn(__fadd_rn(m1, p1), k1)), __fmul_rn(__fadd_rn(m2, p2), k2));
        }
        G[idx] = g;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TH * TW; idx += kThreads) {
        const int cy = idx / TW, cx = idx - cy * TW;
        const int gy = y0 + cy, gx = x0 + cx;
        if (gy >= rows || gx >= cols) continue;
        const int b = 2 - y0;
        const float c0 = G[(cy + 2) * TW + cx];
        float m1, p1, m2, p2;
        if (gy >= 2 && gy + 2 < rows) {
            m1 = G[(cy + 1) * TW + cx]; p1 = G[(cy + 3) * TW + cx]; m2 = G[cy * TW + cx]; p2 = G[(cy + 4) * TW + cx];
        } else {
            m1 = G[(reflect101(gy - 1, rows) + b) * TW + cx]; p1 = G[(reflect101(gy + 1, rows) + b) * TW + cx];
            m2 = G[(reflect101(gy - 2, rows) + b) * TW + cx]; p2 = G[(reflect101(gy + 2, rows) + b) * TW + cx];
        }
        out[(size_t)gy * cols + gx] =
            __fadd_rn(__fadd_rn(__fmul_rn(c0, k0), __fmul_rn(__fadd_rn(m1, p1), k1)), __fmul_rn(__fadd_rn(m2, p2), k2));
    }
}

// cv::cvtColor(COLOR_BGR2GRAY) on CV_8UC3 (main_sl.cpp:1167,1171): OpenCV's 8-bit path is fixed point with 15-bit
// coefficients, gray = (3735 B + 19235 G + 9798 R + 16384) >> 15 (bit-equal to cv2 4.13 for every (B, G, R), tests/test_stereo_gray.py).
// One thread per 4 pixels: 12 input bytes (three aligned 32-bit loads where the row allows), one 32-bit store.
__global__ void __launch_bounds__(kThreads) k_bgr2gray(const uint8_t* __restrict__ bgr, size_t bgr_pitch, size_t bgr_fstride,
                                                       uint8_t* __restrict__ gray, size_t gray_pitch, size_t gray_fstride, int cols) {
    const int q = blockIdx.x * kThreads + threadIdx.x, r = blockIdx.y;
    if (q * 4 >= cols) return;
    const uint8_t* src = bgr + (size_t)blockIdx.z * bgr_fstride + (size_t)r * bgr_pitch + (size_t)q * 12;
    uint8_t* dst = gray + (size_t)blockIdx.z * gray_fstride + (size_t)r * gray_pitch + (size_t)q * 4;
    auto g = [](unsigned b, unsigned gg, unsigned rr) { return (3735u * b + 19235u * gg + 9798u * rr + 16384u) >> 15; };
    if (q * 4 + 4 <= cols && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(src)), w1 = __ldg(reinterpret_cast<const uint32_t*>(src) + 1),
                       w2 = __ldg(reinterpret_cast<const uint32_t*>(src) + 2);
        const unsigned g0 = g(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        const unsigned g1 = g(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        const unsigned g2 = g((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        const unsigned g3 = g((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        *reinterpret_cast<uint32_t*>(dst) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (int j = 0; j < 4 && q * 4 + j < cols; ++j) dst[j] = (uint8_t)g(src[3 * j], src[3 * j + 1], src[3 * j + 2]);
    }
}

inline unsigned blocks_for(long long total) { return (unsigned)((total + kThreads - 1) / kThreads); }

}  // namespace

cudaError_t stereo_measurement_derivatives(const float* val, float* dx, float* dy, int rows, int cols, int n_frames,
                                           cudaStream_t st) {
    const long long total = (long long)rows * cols * n_frames;
    if (total == 0) return cudaSuccess;
    DCMT_LAUNCH(k_measurement_derivatives, dim3(blocks_for(total)), dim3(kThreads), 0, st, val, dx, dy, rows, cols, total);
    return cudaGetLastError();
}

cudaError_t stereo_get_initial_disparity(const float* depth, float* disp, int rows, int cols, int n_frames, float baseline,
                                         float focal, cudaStream_t st) {
    const long long total = (long long)rows * cols * n_frames;
    if (total == 0) return cudaSuccess;
    volatile float bf = baseline * focal;  // float product as in :852 `(baseline*focal)/depth`
    DCMT_LAUNCH(k_get_initial_disparity, dim3(blocks_for(total)), dim3(kThreads), 0, st, depth, disp, total, (float)bf);
    return cudaGetLastError();
}

cudaError_t stereo_optimize_ig(const float* vl, const float* vr, float* disp, int rows, int cols, int n_frames, int iters,
                               float damp, float clip, cudaStream_t st) {
    const long long total = (long long)rows * cols * n_frames;
    if (total == 0) return cudaSuccess;
    DCMT_LAUNCH(k_optimize_ig, dim3(blocks_for(total)), dim3(kThreads), 0, st, vl, vr, disp, rows, cols, total, iters, damp,
                clip);
    return cudaGetLastError();
}

cudaError_t stereo_retrieve_depth(const float* disp, float* depth, int rows, int cols, int n_frames, float baseline,
                                  float focal, float clip, cudaStream_t st) {
    const long long total = (long long)rows * cols * n_frames;
    if (total == 0) return cudaSuccess;
    volatile float bf = baseline * focal;
    DCMT_LAUNCH(k_retrieve_depth, dim3(blocks_for(total)), dim3(kThreads), 0, st, disp, depth, total, (float)bf, clip);
    return cudaGetLastError();
}

cudaError_t stereo_bgr2gray(const uint8_t* bgr, size_t bgr_pitch, size_t bgr_fstride, uint8_t* gray, size_t gray_pitch, size_t gray_fstride,
                            int rows, int cols, int n_frames, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    DCMT_LAUNCH(k_bgr2gray, dim3(((cols + 3) / 4 + kThreads - 1) / kThreads, rows, n_frames), dim3(kThreads), 0, st, bgr, bgr_pitch, bgr_fstride,
                gray, gray_pitch, gray_fstride, cols);
    return cudaGetLastError();
}

cudaError_t stereo_entries_derivatives(void* entries, size_t row_step, size_t elem_stride, int rows, int cols, cudaStream_t st) {
    if (rows < 3 || cols < 3) return cudaSuccess;  // no interior
    DCMT_LAUNCH(k_entries_derivatives, dim3((cols + kThreads - 1) / kThreads, rows), dim3(kThreads), 0, st,
                EntriesView{static_cast<char*>(entries), row_step, elem_stride}, rows, cols);
    return cudaGetLastError();
}

cudaError_t stereo_entries_optimize_ig(const void* left, size_t left_row_step, const void* right, size_t right_row_step, size_t elem_stride,
                                       float* disp, size_t disp_pitch, int rows, int cols, int iters, float damp, float clip,
                                       cudaStream_t st) {
    DCMT_LAUNCH(k_entries_optimize_ig, dim3((cols + kThreads - 1) / kThreads, rows), dim3(kThreads), 0, st,
                EntriesView{static_cast<char*>(const_cast<void*>(left)), left_row_step, elem_stride},
                EntriesView{static_cast<char*>(const_cast<void*>(right)), right_row_step, elem_stride}, disp, disp_pitch, rows, cols, iters, damp,
                clip);
    return cudaGetLastError();
}

cudaError_t stereo_initial_disparity_mat(const float* depth, size_t depth_pitch, float* disp, size_t disp_pitch, int rows, int cols,
                                         float baseline, float focal, cudaStream_t st) {
    volatile float bf = baseline * focal;
    DCMT_LAUNCH(k_initial_disparity_mat, dim3((cols + kThreads - 1) / kThreads, rows), dim3(kThreads), 0, st, depth, depth_pitch, disp,
                disp_pitch, cols, (float)bf);
    return cudaGetLastError();
}

cudaError_t stereo_retrieve_depth_mat(const float* disp, size_t disp_pitch, float* depth, size_t depth_pitch, int rows, int cols,
                                      float baseline, float focal, float clip, cudaStream_t st) {
    volatile float bf = baseline * focal;
    DCMT_LAUNCH(k_retrieve_depth_mat, dim3((cols + kThreads - 1) / kThreads, rows), dim3(kThreads), 0, st, disp, disp_pitch, depth,
                depth_pitch, cols, (float)bf, clip);
    return cudaGetLastError();
}

cudaError_t stereo_refine(const float* depth_ig, const uint8_t* left, const uint8_t* right, float* depth_out,
                          float* disp_out, int rows, int cols, int n_frames, float baseline, float focal, float damp,
                          float err_clip, float depth_clip, int iters, int final_gauss, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    volatile float bf = baseline * focal;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const size_t off = (size_t)f0 * rows * cols;
        RefineArgs a{depth_ig + off, left + off, right + off, depth_out + off, disp_out ? disp_out + off : nullptr,
                     rows, cols, (float)bf, damp, err_clip, depth_clip, iters, final_gauss};
        DCMT_LAUNCH(k_stereo_refine, dim3((cols + TW - 1) / TW, (rows + TH - 1) / TH, nf), dim3(kThreads), 0, st, a);
    }
    return cudaGetLastError();
}

}  // namespace dcmt
