// project.cu -- the LiDAR projection in front of the completion call (src/DC_stereo_lidar/main_sl.cpp:478-523, the same
// loop in vedi_pc :340-385): Velodyne points (x, y, z, intensity) -> T (velodyne to camera) -> keep z > 0 -> P (3 x 4
// projection) -> perspective division -> bounds test on the float coordinates -> (int) truncation -> scatter of the
// depth, LAST point in file order wins (:515) -> cv::normalize(NORM_MINMAX, 0, 80) (:521).
//
// All arithmetic is float32 in the order of the source: the transform is written out term by term (:485-487); the
// Eigen product P * p.homogeneous() (:500) is evaluated the way Eigen >= 3.3 evaluates it (Geometry/Homogeneous.h:
// dst = P.leftCols<3>() * p, then dst += P.col(3); the three-term sum of the small fixed-size product is the balanced
// tree a0 + (a1 + a2) of redux_novec_unroller) -- Eigen 3.2 summed left to right; the reference pins no version.
// Explicit __fmul_rn / __fadd_rn / __fdiv_rn keep the compiler from contracting anything into FMAs.  "Last writer wins" is made deterministic with a 64-bit
// atomicMax on (point index + 1) << 32 | depth bits per pixel.  cv::normalize is restated from OpenCV 4.x
// (scale = float((b - a) / (max - min)), shift = float(a) - float(min * scale), dst = src * scale + shift).
#include "project.cuh"

namespace dcmt {
namespace {

struct Mats {
    float T[12];  // rows 0..2 of the 4x4
    float P[12];
};

// Every kernel below serves a BATCH of clouds: blockIdx.y is the cloud; its keys, min / max words, points and images
// sit at cloud * (their per-cloud size).
__global__ void k_project_clear(unsigned long long* __restrict__ keys, size_t n, unsigned* __restrict__ minmax) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    keys += (size_t)blockIdx.y * n;
    minmax += (size_t)blockIdx.y * 8;
    if (i < n) keys[i] = 0ull;
    if (i == 0) { minmax[2] = 0u; minmax[3] = 0xffffffffu; minmax[4] = 0u; }
}

__device__ __forceinline__ float row_dot(const float* m, float x, float y, float z) {  // m0 x + m1 y + m2 z + m3, left to right
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
}

__device__ __forceinline__ float row_dot_eigen(const float* m, float x, float y, float z) {  // (m0 x + (m1 y + m2 z)) + m3
    return __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fadd_rn(__fmul_rn(m[1], y), __fmul_rn(m[2], z))), m[3]);
}

__global__ void __launch_bounds__(256) k_project_scatter(const float4* __restrict__ pts, int n, const int32_t* __restrict__ counts,
                                                         size_t cloud_stride, Mats m, int rows, int cols,
                                                         unsigned long long* __restrict__ keys, unsigned* __restrict__ minmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (counts) n = counts[blockIdx.y];
    pts += (size_t)blockIdx.y * cloud_stride;
    keys += (size_t)blockIdx.y * rows * cols;
    minmax += (size_t)blockIdx.y * 8;
    int landed = 0;
    if (i < n) {
        const float4 p = __ldg(pts + i);
        const float tx = row_dot(m.T, p.x, p.y, p.z), ty = row_dot(m.T + 4, p.x, p.y, p.z), tz = row_dot(m.T + 8, p.x, p.y, p.z);
        if (tz > 0.0f) {  // :488
            const float X = row_dot_eigen(m.P, tx, ty, tz), Y = row_dot_eigen(m.P + 4, tx, ty, tz), Z = row_dot_eigen(m.P + 8, tx, ty, tz);
            const float u = __fdiv_rn(X, Z), v = __fdiv_rn(Y, Z);  // :501-502
            if (u >= 0.0f && u < (float)cols && v >= 0.0f && v < (float)rows) {  // :505-506 (NaN fails every comparison)
                const int iu = (int)u, iv = (int)v;  // :510-511
                atomicMax(keys + (size_t)iv * cols + iu, ((unsigned long long)(unsigned)(i + 1) << 32) | (unsigned long long)__float_as_uint(Z));
                landed = 1;
            }
        }
    }
    const int cnt = __syncthreads_count(landed);
    if (threadIdx.x == 0 && cnt) atomicAdd(minmax + 2, (unsigned)cnt);
}

// keys -> depth image; min / max of the image for cv::normalize.  Depths here are > 0 or the image's initial 0, except
// for a projected depth Z <= 0 (possible only with an exotic P): such frames take the signed path of the ordering trick.
__device__ __forceinline__ unsigned order_bits(float f) {  // monotone map float -> unsigned
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// four pixels per thread, one reduction per 1024-pixel block, and only a block whose values can change the image minimum /
// maximum touches the shared words: a reservation per warp made 13 000 requests per cloud to ONE address in L2 -- 17 us of
// a 20 us cloud, atomics or plain reads alike
constexpr int kGatherPx = 4;
__global__ void __launch_bounds__(256) k_project_gather(const unsigned long long* __restrict__ keys, size_t n, float* __restrict__ projected,
                                                        unsigned* __restrict__ minmax) {
    __shared__ unsigned s_lo[8], s_hi[8];
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kGatherPx;
    keys += (size_t)blockIdx.y * n;
    minmax += (size_t)blockIdx.y * 8;
    if (projected) projected += (size_t)blockIdx.y * n;
    unsigned lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int j = 0; j < kGatherPx; ++j) {
        const size_t i = i0 + j;
        if (i < n) {
            const unsigned long long k = keys[i];
            const float d = k ? __uint_as_float((unsigned)(k & 0xffffffffull)) : 0.0f;
            if (projected) projected[i] = d;
            const unsigned o = order_bits(d);
            lo = min(lo, o);
            hi = max(hi, o);
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        // monotone values: a stale read only costs an atomic
        if (lo != 0xffffffffu && lo < *reinterpret_cast<volatile unsigned*>(minmax + 3)) atomicMin(minmax + 3, lo);
        if (hi > *reinterpret_cast<volatile unsigned*>(minmax + 4)) atomicMax(minmax + 4, hi);
    }
}

__global__ void __launch_bounds__(256) k_project_normalize(const unsigned long long* __restrict__ keys, size_t n, const unsigned* __restrict__ minmax,
                                                           float a, float b, float* __restrict__ normalized, int32_t* __restrict__ n_projected) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    keys += (size_t)blockIdx.y * n;
    minmax += (size_t)blockIdx.y * 8;
    if (i == 0 && n_projected) n_projected[blockIdx.y] = (int32_t)minmax[2];
    if (i >= n || !normalized) return;
    normalized += (size_t)blockIdx.y * n;
    // cv::normalize, NORM_MINMAX, dtype CV_32F (OpenCV 4.x modules/core/src/norm.cpp)
    const double smin = (double)unorder_bits(minmax[3]), smax = (double)unorder_bits(minmax[4]);
    const double dmin = a < b ? (double)a : (double)b, dmax = a < b ? (double)b : (double)a;
    double scale = (dmax - dmin) * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    scale = (double)(float)scale;
    const float shift = __fsub_rn((float)dmin, (float)(smin * scale));
    const unsigned long long k = keys[i];
    const float d = k ? __uint_as_float((unsigned)(k & 0xffffffffull)) : 0.0f;
    normalized[i] = __fadd_rn(__fmul_rn(d, (float)scale), shift);
}

}  // namespace

size_t project_key_count(int rows, int cols) { return (size_t)rows * cols; }

cudaError_t project_run_batch(const float* points, int n_points, const int32_t* counts_dev, size_t cloud_stride_points, int n_clouds,
                              const float* T, const float* P, int rows, int cols, float* projected, float* normalized, float norm_a,
                              float norm_b, int32_t* n_projected, const ProjectWork& w, cudaStream_t st) {
    if (n_clouds == 0) return cudaSuccess;
    Mats m;
    for (int k = 0; k < 12; ++k) { m.T[k] = T[k]; m.P[k] = P[k]; }
    const size_t n = (size_t)rows * cols;
    const unsigned nb = (unsigned)((n + 255) / 256);
    // four launches for the whole batch: clear, scatter (64-bit atomicMax, last point in file order wins), gather + min / max,
    // normalize.  (A single cloud used to cost the same four launches: 58 us each, launch bound.)
    DCMT_LAUNCH(k_project_clear, dim3(nb, n_clouds), dim3(256), 0, st, w.keys, n, w.minmax);
    if (n_points > 0)
        DCMT_LAUNCH(k_project_scatter, dim3((n_points + 255) / 256, n_clouds), dim3(256), 0, st, reinterpret_cast<const float4*>(points),
                    n_points, counts_dev, cloud_stride_points, m, rows, cols, w.keys, w.minmax);
    DCMT_LAUNCH(k_project_gather, dim3((unsigned)((n + 256 * kGatherPx - 1) / (256 * kGatherPx)), n_clouds), dim3(256), 0, st, w.keys, n, projected,
                w.minmax);
    DCMT_LAUNCH(k_project_normalize, dim3(nb, n_clouds), dim3(256), 0, st, w.keys, n, w.minmax, norm_a, norm_b, normalized, n_projected);
    return cudaGetLastError();
}

cudaError_t project_run(const float* points, int n_points, const float* T, const float* P, int rows, int cols, float* projected,
                        float* normalized, float norm_a, float norm_b, int32_t* n_projected, const ProjectWork& w, cudaStream_t st) {
    return project_run_batch(points, n_points, nullptr, 0, 1, T, P, rows, cols, projected, normalized, norm_a, norm_b, n_projected, w, st);
}

}  // namespace dcmt
