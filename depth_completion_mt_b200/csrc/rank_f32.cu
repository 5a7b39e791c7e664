// rank_f32.cu -- float32 frames on the fused kernels: the order-preserving dictionary in front of fused_q8.cu.
//
// Every stage of img_completion between the inversion (img_completion.cpp:55-67) and the blur (:172-189) only SELECTS
// among the values it is given: dilations and erosions take maxima / minima, the fills copy, the column extrapolation
// copies, the median picks the 13th of 25 -- the one constant that enters is 100.0 for empty columns (:110).  So a frame
// of arbitrary finite floats can run through the packed-uint16 kernels of fused_q8.cu unchanged if its valid pixels
// are first replaced by their RANKS among the frame's inverted values (any strictly monotone code gives the same
// selections), and the codes are turned back into the floats they stand for before the Gaussian:
//
//   k_rank_sort    one CTA per frame: the inverted values 100 - v of the valid pixels (v >= 0.1f and 100 - v >= 0.1f)
//                  are gathered in shared memory and sorted (bitonic network); the sorted list is the frame's
//                  dictionary LUT[0 .. count)
//   k_rank_encode  every pixel -> code: 1 for a hole (the encoding of fused_q8.cu), 27 + lower_bound(LUT, 100 - v) for a
//                  valid pixel (equal values share a code); the code plane is the uint16 input of k_q8_front
//   k_q8_tail<true> decodes with the LUT (code 65535 = the constant 100.0), blurs in float32 and inverts back.
//
// With blur "none" the output is bit-identical to the reference for any finite input (the same float subtractions on the
// same selected values); with the Gaussian it is within the tolerance stated for float input (1e-4).  Frames with more
// than kRankMaxValid valid pixels (7.6 % of a KITTI frame) or a NaN are left to the generic pipeline (flagged).
#include "rank_f32.cuh"

namespace dcmt {
namespace {

constexpr int kSortThreads = 1024;

// hole / valid classification and the inverted value of one input pixel (img_completion.cpp:55-67 followed by the
// `< 0.1` tests of every later stage): valid <=> v >= 0.1f and 100 - v >= 0.1f
__device__ __forceinline__ bool rank_valid(float v, float& inv) {
    inv = __fsub_rn(kMaxDepth, v);
    return v >= 0.1f && inv >= 0.1f;
}

__global__ void __launch_bounds__(kSortThreads, 1) k_rank_sort(const float* __restrict__ in, size_t in_pitch, size_t in_fstride, int rows,
                                                               int cols, float* __restrict__ lut, int* __restrict__ lut_count,
                                                               FrameCounters* __restrict__ ctr) {
    DCMT_DYN_SMEM(uint32_t, keys);  // kRankMaxValid keys
    __shared__ int s_warp_sum[kSortThreads / 32];
    __shared__ int s_total, s_bad;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = in + (size_t)f * in_fstride;
    const int n = rows * cols;
    if (tid == 0) s_bad = 0;
    // pass 1: valid pixels per thread (pixels tid, tid + T, ...), then an exclusive scan over the CTA
    int mine = 0, bad = 0;
    for (int i = tid; i < n; i += kSortThreads) {
        const int r = i / cols, c = i - r * cols;
        const float v = __ldg(src + (size_t)r * in_pitch + c);
        float inv;
        mine += rank_valid(v, inv) ? 1 : 0;
        bad |= (v != v) ? 1 : 0;  // NaN: outside the domain parity is defined on -> generic pipeline
    }
    int incl = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    if (bad) s_bad = 1;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp_sum[lane];
        int winc = w;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, s);
            if (lane >= s) winc += t;
        }
        s_warp_sum[lane] = winc - w;  // exclusive
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    const int total = s_total;
    if (total > kRankMaxValid || s_bad) {  // the dictionary does not fit / NaN: this frame goes to the generic pipeline
        if (tid == 0) {
            ctr[f].needs_generic = 1;
            lut_count[f] = 0;
        }
        return;
    }
    // pass 2: the keys (bit patterns of positive floats order like the floats), at thread-private offsets
    int at = s_warp_sum[warp] + incl - mine;
    for (int i = tid; i < n; i += kSortThreads) {
        const int r = i / cols, c = i - r * cols;
        const float v = __ldg(src + (size_t)r * in_pitch + c);
        float inv;
        if (rank_valid(v, inv)) keys[at++] = __float_as_uint(inv);
    }
    int P = 1;
    while (P < total) P <<= 1;
    if (P < 2) P = 2;
    for (int i = total + tid; i < P; i += kSortThreads) keys[i] = 0xffffffffu;
    __syncthreads();
    // bitonic sort, ascending
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += kSortThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // the lower index of pair t at distance j
                const int p = i | j;
                const uint32_t a = keys[i], b = keys[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { keys[i] = b; keys[p] = a; }
            }
            __syncthreads();
        }
    }
    float* o = lut + (size_t)f * kRankMaxValid;
    for (int i = tid; i < total; i += kSortThreads) o[i] = __uint_as_float(keys[i]);
    if (tid == 0) lut_count[f] = total;
}

// one thread per 8 pixels of a row (one 16-byte store of codes)
__global__ void __launch_bounds__(256) k_rank_encode(const float* __restrict__ in, size_t in_pitch, size_t in_fstride, int rows, int cols,
                                                     const float* __restrict__ lut, const int* __restrict__ lut_count,
                                                     uint16_t* __restrict__ codes, size_t code_pitch) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, f = blockIdx.z;
    if (q * 8 >= cols) return;
    const int count = lut_count[f];
    const float* d = lut + (size_t)f * kRankMaxValid;
    const float* src = in + (size_t)f * in_fstride + (size_t)r * in_pitch + q * 8;
    uint32_t e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        e[j] = 0u;  // beyond the last column: the padding of the code plane, never read as image
        if (q * 8 + j < cols) {
            float inv;
            if (rank_valid(__ldg(src + j), inv)) {
                int lo = 0, hi = count;  // first index with d[index] >= inv
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(d + mid) < inv) lo = mid + 1;
                    else hi = mid;
                }
                e[j] = (uint32_t)(kRankFirstCode + lo);
            } else {
                e[j] = 1u;
            }
        }
    }
    uint16_t* o = codes + ((size_t)f * rows + r) * code_pitch + q * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
}

}  // namespace

cudaError_t rank_configure() {
    return cudaFuncSetAttribute(k_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankMaxValid * (int)sizeof(uint32_t));
}

cudaError_t rank_build(const float* in, size_t in_pitch, size_t in_fstride, int rows, int cols, int n_frames, float* lut, int* lut_count,
                       uint16_t* codes, size_t code_pitch, FrameCounters* ctr, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    DCMT_LAUNCH(k_rank_sort, dim3(n_frames), dim3(kSortThreads), kRankMaxValid * sizeof(uint32_t), st, in, in_pitch, in_fstride, rows, cols,
                lut, lut_count, ctr);
    DCMT_LAUNCH(k_rank_encode, dim3(((cols + 7) / 8 + 255) / 256, rows, n_frames), dim3(256), 0, st, in, in_pitch, in_fstride, rows, cols,
                lut, lut_count, codes, code_pitch);
    return cudaGetLastError();
}

}  // namespace dcmt
