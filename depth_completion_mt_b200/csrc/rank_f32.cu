// rank_f32.cu -- float32 frames on the fused kernels: the order-preserving dictionary in front of fused_q8.cu.
//
// Every stage of img_completion between the inversion (img_completion.cpp:55-67) and the blur (:172-189) only SELECTS
// among the values it is given: dilations and erosions take maxima / minima, the fills copy, the column extrapolation
// copies, the median picks the 13th of 25 -- the one constant that enters is 100.0 for empty columns (:110).  So a frame
// of arbitrary finite floats can run through the packed-uint16 kernels of fused_q8.cu unchanged if its valid pixels
// are first replaced by their RANKS among the frame's inverted values (any strictly monotone code gives the same
// selections), and the codes are turned back into the floats they stand for before the Gaussian:
//
//   k_rank_compact the inverted values 100 - v of the valid pixels (v >= 0.1f and 100 - v >= 0.1f) are gathered per frame
//   k_rank_sort    one CTA per frame sorts them in shared memory (bitonic network); the sorted list is the frame's
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

// Gather the keys of a frame's valid pixels (bit patterns of the inverted values: positive floats order like their bits)
// into lut[frame][0 .. count), in arbitrary order, with every SM: one thread per 8 pixels of a row, one reservation per
// warp.  count keeps counting past the capacity (the sort kernel flags such frames); a NaN flags the frame at once.
__global__ void __launch_bounds__(256) k_rank_compact(const float* __restrict__ in, size_t in_pitch, size_t in_fstride, int rows, int cols,
                                                      int vec_ok, float* __restrict__ lut, int* __restrict__ lut_count,
                                                      FrameCounters* __restrict__ ctr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, f = blockIdx.z, lane = threadIdx.x & 31;
    float iv[8];
    uint32_t m = 0u;  // bit j: pixel j of the octet is valid
    bool nan = false;
    if (q * 8 < cols) {
        const float* src = in + (size_t)f * in_fstride + (size_t)r * in_pitch + q * 8;
        float v[8];
        if (vec_ok && q * 8 + 8 <= cols) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = q * 8 + j < cols ? __ldg(src + j) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            nan |= v[j] != v[j];
            if (rank_valid(v[j], iv[j])) m |= 1u << j;
        }
    }
    const int n = __popc(m);
    if (__any_sync(0xffffffffu, nan) && lane == 0) ctr[f].needs_generic = 1;  // outside the domain parity is defined on
    // offsets: inclusive scan inside the warp, warp totals through shared memory, ONE atomic reservation per block (the
    // blocks of a frame run side by side: a reservation per warp serialised 1 700 atomics per frame on one address)
    __shared__ int s_warp[8], s_base;
    const int warp = threadIdx.x >> 5;
    int incl = n;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(lut_count + f, tot) : 0;
    }
    __syncthreads();
    int base = s_base + s_warp[warp] + incl - n;
    float* o = lut + (size_t)f * kRankMaxValid;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if ((m >> j) & 1u) {
            if (base < kRankMaxValid) o[base] = iv[j];  // the float itself: positive floats order like their bit patterns
            ++base;
        }
}

// One CTA per frame: the gathered keys are sorted in shared memory (bitonic network) and written back -- the frame's
// dictionary.  Frames whose keys do not fit (or that were flagged for a NaN) are left to the generic pipeline.
__global__ void __launch_bounds__(kSortThreads, 1) k_rank_sort(float* __restrict__ lut, const int* __restrict__ lut_count,
                                                               FrameCounters* __restrict__ ctr) {
    DCMT_DYN_SMEM(uint32_t, keys);  // kRankMaxValid keys
    const int f = blockIdx.x, tid = threadIdx.x;
    const int total = lut_count[f];
    if (total > kRankMaxValid) {
        if (tid == 0) ctr[f].needs_generic = 1;
        return;
    }
    if (ctr[f].needs_generic) return;
    float* o = lut + (size_t)f * kRankMaxValid;
    int P = 2;
    while (P < total) P <<= 1;
    for (int i = tid; i < P; i += kSortThreads) keys[i] = i < total ? __float_as_uint(o[i]) : 0xffffffffu;
    __syncthreads();
    // bitonic sort, ascending, four keys per access: a stage at distance j >= 4 exchanges QUADS (two 16-byte loads, four
    // min / max pairs, two 16-byte stores: the direction is uniform over a quad because k >= 8 there); the stages at
    // distances 2 and 1 of every k run on one quad in registers.  91 + 15 passes over shared memory instead of 120 scalar ones.
    if (P < 4) {  // at most two keys (P is 2): one compare-exchange
        if (tid == 0 && keys[0] > keys[1]) { const uint32_t t = keys[0]; keys[0] = keys[1]; keys[1] = t; }
        __syncthreads();
    } else {
        uint4* kq = reinterpret_cast<uint4*>(keys);
        const int nq = P >> 2;
        auto cx = [](uint32_t& a, uint32_t& b, bool up) {
            const uint32_t lo = min(a, b), hi = max(a, b);
            a = up ? lo : hi;
            b = up ? hi : lo;
        };
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j >= 4; j >>= 1) {
                const int jq = j >> 2;  // distance in quads
                for (int t = tid; t < (nq >> 1); t += kSortThreads) {
                    const int iq = ((t & ~(jq - 1)) << 1) | (t & (jq - 1));  // the lower quad of pair t
                    uint4 a = kq[iq], b = kq[iq | jq];
                    const bool up = ((iq << 2) & k) == 0;
                    cx(a.x, b.x, up); cx(a.y, b.y, up); cx(a.z, b.z, up); cx(a.w, b.w, up);
                    kq[iq] = a;
                    kq[iq | jq] = b;
                }
                __syncthreads();
            }
            for (int q = tid; q < nq; q += kSortThreads) {
                uint4 v = kq[q];
                const int i = q << 2;
                if (k == 2) {  // pairs (0,1) ascending, (2,3) descending
                    cx(v.x, v.y, true);
                    cx(v.z, v.w, false);
                } else {
                    const bool up = (i & k) == 0;  // k == 4: (i & 4); larger k: uniform over the quad as well
                    cx(v.x, v.z, up); cx(v.y, v.w, up);
                    cx(v.x, v.y, up); cx(v.z, v.w, up);
                }
                kq[q] = v;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < total; i += kSortThreads) o[i] = __uint_as_float(keys[i]);
}

// One thread per 8 pixels of a row (one 16-byte store of codes).  The valid pixels of a warp's 256 pixels (about 13 at 5 %)
// are queued in shared memory and searched one per lane: the searches are chains of dependent loads from the dictionary,
// so what counts is how many run side by side -- not every lane walking its own eight pixels in turn.  The kernel keeps
// its shared memory small (8 KB per block) on purpose: the 86 KB dictionary of the frame the SM is working on then stays
// in L1 (a 32 KB-per-block variant that pushed it out to L2 was 50 % slower).
constexpr int kEncOct = 1;  // octets per thread
__global__ void __launch_bounds__(256) k_rank_encode(const float* __restrict__ in, size_t in_pitch, size_t in_fstride, int rows, int cols,
                                                     int vec_ok, const float* __restrict__ lut, const int* __restrict__ lut_count,
                                                     uint16_t* __restrict__ codes, size_t code_pitch) {
    __shared__ uint16_t s_slot[8][256];                 // queue of the warp: which of its 256 pixels are valid
    __shared__ __align__(16) uint16_t s_code[8][256];   // codes of the warp's pixels: [lane * 8 + j]
    const int q = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, f = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = min(lut_count[f], kRankMaxValid);  // (frames past the capacity are flagged and redone: any code will do)
    const float* d = lut + (size_t)f * kRankMaxValid;
    const float* row = in + (size_t)f * in_fstride + (size_t)r * in_pitch;
    uint16_t* sc = s_code[warp];
    uint16_t* ss = s_slot[warp];
    uint32_t m = 0u, inside = 0u;
    if (q * 8 < cols) {
        const float* src = row + q * 8;
        float v[8];
        if (vec_ok && q * 8 + 8 <= cols) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            inside = 0xffu;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = q * 8 + j < cols ? __ldg(src + j) : 0.0f;
                inside |= q * 8 + j < cols ? 1u << j : 0u;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float inv;
            if (rank_valid(v[j], inv)) m |= 1u << j;
        }
        m &= inside;
    }
    // holes: 1, beyond the last column: 0 (padding of the code plane, never read as image); valid ones are overwritten below
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[lane * 8 + j] = (uint16_t)((inside >> j) & 1u);
    const int n = __popc(m);
    int incl = n;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int at = incl - n;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if ((m >> j) & 1u) ss[at++] = (uint16_t)(lane * 8 + j);
    __syncwarp();
    const int px0 = (q - lane) * 8;  // first pixel of the warp in the row
    for (int k = lane; k < total; k += 32) {
        const int slot = ss[k];
        const float inv = __fsub_rn(kMaxDepth, __ldg(row + px0 + slot));  // the pixel again (L1): cheaper than queueing the value
        int lo = 0, hi = count;  // first index with d[index] >= inv
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(d + mid) < inv) lo = mid + 1;
            else hi = mid;
        }
        sc[slot] = (uint16_t)(kRankFirstCode + lo);
    }
    __syncwarp();
    if (q * 8 < cols) {
        uint16_t* o = codes + ((size_t)f * rows + r) * code_pitch + (size_t)q * 8;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(sc + lane * 8);
    }
}

}  // namespace

cudaError_t rank_configure() {
    return cudaFuncSetAttribute(k_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kRankMaxValid * (int)sizeof(uint32_t));
}

cudaError_t rank_build(const float* in, size_t in_pitch, size_t in_fstride, int rows, int cols, int n_frames, float* lut, int* lut_count,
                       uint16_t* codes, size_t code_pitch, FrameCounters* ctr, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(lut_count, 0, (size_t)n_frames * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int vec_ok = in_pitch % 4 == 0 && in_fstride % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    DCMT_LAUNCH(k_rank_compact, dim3(((cols + 7) / 8 + 255) / 256, rows, n_frames), dim3(256), 0, st, in, in_pitch, in_fstride, rows, cols, vec_ok,
                lut, lut_count, ctr);
    DCMT_LAUNCH(k_rank_sort, dim3(n_frames), dim3(kSortThreads), kRankMaxValid * sizeof(uint32_t), st, lut, lut_count, ctr);
    DCMT_LAUNCH(k_rank_encode, dim3(((cols + 7) / 8 + 256 * kEncOct - 1) / (256 * kEncOct), rows, n_frames), dim3(256), 0, st, in, in_pitch, in_fstride, rows, cols,
                vec_ok, lut, lut_count, codes, code_pitch);
    return cudaGetLastError();
}

}  // namespace dcmt
