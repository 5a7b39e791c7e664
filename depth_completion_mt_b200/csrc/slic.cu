// slic.cu -- Slic::generate_superpixels (src/DC_lidar_camera/slic.cpp:19-182) on the device.
//
// The reference loops over the centres and, for each, over the pixels of its 2 step x 2 step window, keeping per pixel
// the smallest distance seen (strict <, so the lowest centre index wins ties, :131-134).  Here every pixel looks up the
// centres whose window covers it (centres are binned on a step-sized grid each iteration; the covering ones lie in the
// 3 x 3 bins around the pixel) and takes the minimum over (distance, index) -- the same result, all pixels in parallel.
//   * window test exactly as written: k from (int)(cx - step) while k < cx + step (:123-124), double arithmetic;
//   * distance exactly as compute_dist (:61-69) -- doubles, pow(x, 2) = x * x, sqrt, no FMA contraction -- wherever it decides:
//     cheaper monotone stand-ins (without the square roots in double; in float in the band kernel) pick the winner only when
//     their margin exceeds their own error bound, everything closer is re-evaluated in the reference's operation sequence;
//   * pixels no window covers keep their previous label (the reference only resets `distances`, :115-119);
//   * centre update (:141-171) sums integers (pixel values, coordinates): exact in any order, done with 64-bit
//     integer atomics, then divided in double; an empty cluster becomes 0 / 0 = NaN and never wins a pixel again.
// So the labels are bit-identical to a scalar build of the reference, not merely "mostly equal".
// create_connectivity (:186-260) writes only its local `new_clusters`, never `clusters`: nothing to do.
#include "slic.cuh"

#include <cstdlib>

namespace dcmt {
namespace {

__device__ __forceinline__ double sq(double x) { return __dmul_rn(x, x); }

// slic.cpp:72-99, one thread per centre
__global__ void k_slic_init(const uint8_t* __restrict__ lab, int rows, int cols, int step, int ny, int n, double* __restrict__ centers) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    lab += (size_t)blockIdx.y * rows * cols * 3;  // frame
    centers += (size_t)blockIdx.y * n * 5;
    // centres are created column by column: i = step, 2 step, ... (outer), j = step, 2 step, ... (inner)  (:33-34)
    const int ci = c / ny, cj = c - ci * ny;
    const int cx = step + ci * step, cy = step + cj * step;
    double min_grad = (double)FLT_MAX;
    int lx = cx, ly = cy;
    for (int i = cx - 1; i < cx + 2; i++)
        for (int j = cy - 1; j < cy + 2; j++) {
            const double i1 = lab[((size_t)(j + 1) * cols + i) * 3], i2 = lab[((size_t)j * cols + i + 1) * 3], i3 = lab[((size_t)j * cols + i) * 3];
            const double g = __dadd_rn(fabs(__dsub_rn(i1, i3)), fabs(__dsub_rn(i2, i3)));
            if (g < min_grad) { min_grad = g; lx = i; ly = j; }
        }
    const uint8_t* p = lab + ((size_t)ly * cols + lx) * 3;
    double* o = centers + (size_t)c * 5;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = lx; o[4] = ly;
}

__global__ void k_slic_fill_labels(int32_t* __restrict__ labels, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) labels[i] = -1;
}

__device__ __forceinline__ int bin_of(double v, int step, int nb) {
    const int b = (int)floor(v / (double)step);
    return b < 0 ? 0 : (b >= nb ? nb - 1 : b);
}

constexpr int kScanThreads = 1024;

// Everything between two assignments in ONE launch of one block (the centre and bin counts are a few thousand at most):
// centre update from the sums of the previous iteration (:166-172), reset of the sums, binning of the centres on the
// step-sized grid (count, exclusive scan, fill).  Five tiny dependent launches per iteration were 40 % of the run time.
__global__ void __launch_bounds__(kScanThreads) k_slic_prepare(double* __restrict__ centers, unsigned long long* __restrict__ sums, int n,
                                                              int step, int bx, int by, int* __restrict__ count, int* __restrict__ fill,
                                                              int* __restrict__ items, double* __restrict__ sorted, int do_update) {
    __shared__ int part[kScanThreads];
    const int nbins = bx * by;
    {  // one block per frame
        const size_t f = blockIdx.x;
        centers += f * n * 5;
        sums += f * n * 6;
        count += f * (nbins + 1);
        fill += f * nbins;
        items += f * n;
        sorted += f * n * 5;
    }
    for (int c = threadIdx.x; c < n; c += kScanThreads) {
        if (do_update) {
            const double cnt = (double)sums[(size_t)c * 6 + 5];
            for (int k = 0; k < 5; ++k) centers[(size_t)c * 5 + k] = __ddiv_rn((double)sums[(size_t)c * 6 + k], cnt);
        }
        for (int k = 0; k < 6; ++k) sums[(size_t)c * 6 + k] = 0ull;
    }
    // bin counts / offsets / fill cursors live in shared memory when the grid of bins fits (it does up to 5120 bins);
    // only the finished offsets, the items and the bin-ordered centres go to global memory
    constexpr int kSmemBins = 5120;
    __shared__ int s_count[kSmemBins + 1];
    __shared__ int s_fill[kSmemBins];
    const bool in_smem = nbins <= kSmemBins;
    int* cnt = in_smem ? s_count : count;
    int* fil = in_smem ? s_fill : fill;
    for (int b = threadIdx.x; b <= nbins; b += kScanThreads) cnt[b] = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += kScanThreads) {
        const double x = centers[(size_t)c * 5 + 3], y = centers[(size_t)c * 5 + 4];
        if (x != x || y != y) continue;  // NaN centre (empty cluster): covers nothing
        atomicAdd(cnt + bin_of(y, step, by) * bx + bin_of(x, step, bx), 1);
    }
    __syncthreads();
    const int chunk = (nbins + kScanThreads - 1) / kScanThreads;
    const int b0 = threadIdx.x * chunk, b1 = min(b0 + chunk, nbins);
    int total = 0;
    for (int b = b0; b < b1; ++b) total += cnt[b];
    part[threadIdx.x] = total;
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int acc = part[threadIdx.x] - total;
    for (int b = b0; b < b1; ++b) {
        const int c = cnt[b];
        cnt[b] = acc;
        fil[b] = 0;
        acc += c;
    }
    if (threadIdx.x == kScanThreads - 1) cnt[nbins] = part[kScanThreads - 1];
    __syncthreads();
    if (in_smem)
        for (int b = threadIdx.x; b <= nbins; b += kScanThreads) count[b] = cnt[b];
    for (int c = threadIdx.x; c < n; c += kScanThreads) {
        const double x = centers[(size_t)c * 5 + 3], y = centers[(size_t)c * 5 + 4];
        if (x != x || y != y) continue;
        const int b = bin_of(y, step, by) * bx + bin_of(x, step, bx);
        const int pos = cnt[b] + atomicAdd(fil + b, 1);
        items[pos] = c;
        for (int k = 0; k < 5; ++k) sorted[(size_t)pos * 5 + k] = centers[(size_t)c * 5 + k];  // bin order: the assignment reads them in one hop
    }
}

// assignment (:121-139) + accumulation of the new centres (:148-163), one thread per pixel
__global__ void __launch_bounds__(256) k_slic_assign(const uint8_t* __restrict__ lab, int rows, int cols, int step, int nc, int bx, int by,
                                                     const double* __restrict__ centers, const int* __restrict__ offset,
                                                     const int* __restrict__ items, const double* __restrict__ sorted,
                                                     int32_t* __restrict__ labels, unsigned long long* __restrict__ sums, int n_centers,
                                                     double inv_nc2, double inv_ns2) {
    // a block is a 16 x 16 pixel tile: the centres that can cover any of its pixels sit in the bins around it; they are
    // staged in shared memory once (the per-pixel loops below would otherwise chase offset -> item -> centre through
    // global memory: the kernel was bound by that latency, 18 % issue-active)
    constexpr int kCap = 96;
    __shared__ double s_cent[kCap][5];
    // the window of a centre (:123-124: `for (int k = cx - step; k < cx + step; k++)`, truncation towards zero, then a double
    // comparison) as integers: first column / row lo = (int)(c - step), and -- for integer k -- k < c + step <=> k < ceil(c + step);
    // stored as {lo_x, ceil_x - lo_x, lo_y, ceil_y - lo_y} so that a pixel tests (unsigned)(x - lo) < width: one 16-byte load,
    // two subtractions, two compares per candidate instead of four double compares
    __shared__ int4 s_win[kCap];
    __shared__ int s_idx[kCap];
    __shared__ unsigned s_sum[kCap][6];  // per-tile sums of L, a, b, x, y, count for the staged centres
    __shared__ int s_n;
    {  // frame
        const size_t f = blockIdx.z;
        lab += f * rows * cols * 3;
        labels += f * rows * cols;
        centers += f * n_centers * 5;
        sums += f * n_centers * 6;
        offset += f * (bx * by + 1);
        items += f * n_centers;
        sorted += f * n_centers * 5;
    }
    const int tid = threadIdx.y * 16 + threadIdx.x;
    const int x = blockIdx.x * 16 + threadIdx.x, y = blockIdx.y * 16 + threadIdx.y;
    if (tid == 0) s_n = 0;
    for (int i = tid; i < kCap * 6; i += 256) (&s_sum[0][0])[i] = 0u;
    __syncthreads();
    {
        const int bx0 = max((int)(blockIdx.x * 16) / step - 1, 0), bx1 = min((int)(blockIdx.x * 16 + 15) / step + 1, bx - 1);
        const int by0 = max((int)(blockIdx.y * 16) / step - 1, 0), by1 = min((int)(blockIdx.y * 16 + 15) / step + 1, by - 1);
        const int nbx = bx1 - bx0 + 1, nb = nbx * (by1 - by0 + 1);
        // 16 threads per bin (at most 16 bins around a 16 x 16 tile for step >= 16; more bins loop): thread j of a bin
        // copies its j-th centre, so the global loads of all candidates are in flight together
        for (int i = tid >> 4; i < nb; i += 16) {
            const int b = (by0 + i / nbx) * bx + bx0 + i % nbx;
            const int o0 = offset[b], o1 = offset[b + 1];
            for (int k = o0 + (tid & 15); k < o1; k += 16) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < kCap) {
                    s_idx[pos] = items[k];
                    double ce[5];
                    for (int q = 0; q < 5; ++q) ce[q] = sorted[(size_t)k * 5 + q];
                    for (int q = 0; q < 5; ++q) s_cent[pos][q] = ce[q];
                    const int lox = (int)__dsub_rn(ce[3], (double)step), loy = (int)__dsub_rn(ce[4], (double)step);
                    const int hix = (int)ceil(__dadd_rn(ce[3], (double)step)), hiy = (int)ceil(__dadd_rn(ce[4], (double)step));
                    s_win[pos] = make_int4(lox, hix - lox, loy, hiy - loy);
                }
            }
        }
    }
    __syncthreads();
    const int n_cand = s_n;
    const bool staged = n_cand <= kCap;  // otherwise (pathological clustering of centres) fall back to the bins in global memory
    const bool live = x < cols && y < rows;  // lanes beyond the image stay for the warp collectives
    const uint8_t* p = lab + ((size_t)(live ? y : 0) * cols + (live ? x : 0)) * 3;
    const double L = p[0], A = p[1], B = p[2];
    const int pbx = x / step < bx ? x / step : bx - 1, pby = y / step < by ? y / step : by - 1;
    // Two stages.  (1) A cheap, monotone stand-in for compute_dist without square roots or divisions,
    //       D' = Sc / nc^2 + Ss / ns^2  (the reference's distance is sqrt of exactly that, up to ~1e-15 relative rounding),
    //     finds the winner whenever it beats the runner-up by more than 1e-12 relative: no rounding of the reference's own
    //     chain (three sqrt, two divisions, three squarings, each within 2^-53) can reverse such a margin.
    //     (2) Otherwise -- exact ties on symmetric pixels, near ties -- the candidates are re-evaluated with the
    //     reference's exact operation sequence and its tie rule (strict <, lowest centre index).  Same labels, ~4x less work.
    const double xd = (double)x, yd = (double)y;
    double q1 = 1.0e300, q2 = 1.0e300;  // smallest and second smallest D'
    int best_c = -1, best_i = -1;     // winning centre and its slot in the staged list
    const int gy_lo = max(pby - 1, 0), gy_hi = min(pby + 1, by - 1), gx_lo = max(pbx - 1, 0), gx_hi = min(pbx + 1, bx - 1);
    auto covers = [&](double cx, double cy) {
        // for (int k = cx - step; k < cx + step; k++) (:123): truncation towards zero, then a double comparison
        return !(x < (int)__dsub_rn(cx, (double)step) || !((double)x < __dadd_rn(cx, (double)step)) ||
                 y < (int)__dsub_rn(cy, (double)step) || !((double)y < __dadd_rn(cy, (double)step)));
    };
    auto cheap = [&](const double* ce, int c, int slot) {
        const double cx = ce[3], cy = ce[4];
        if (slot >= 0) {
            const int4 w = s_win[slot];
            if ((unsigned)(x - w.x) >= (unsigned)w.y || (unsigned)(y - w.z) >= (unsigned)w.w) return;
        } else if (!covers(cx, cy)) {
            return;
        }
        const double d0 = ce[0] - L, d1 = ce[1] - A, d2 = ce[2] - B, dx = cx - xd, dy = cy - yd;
        const double q = (d0 * d0 + d1 * d1 + d2 * d2) * inv_nc2 + (dx * dx + dy * dy) * inv_ns2;
        if (q < q1 || (q == q1 && c < best_c)) { q2 = q1; q1 = q; best_c = c; best_i = slot; }
        else if (q < q2) q2 = q;
    };
    if (staged) {
        for (int i = 0; i < n_cand; ++i) cheap(s_cent[i], s_idx[i], i);
    } else {
        for (int gy = gy_lo; gy <= gy_hi; ++gy)
            for (int gx = gx_lo; gx <= gx_hi; ++gx) {
                const int b = gy * bx + gx;
                for (int k = offset[b]; k < offset[b + 1]; ++k) cheap(centers + (size_t)items[k] * 5, items[k], -1);
            }
    }
    if (best_c >= 0 && !(q2 > q1 * (1.0 + 1.0e-12))) {
        // too close to call: the reference's own arithmetic decides
        double best = (double)FLT_MAX;  // :117
        best_c = -1;
        best_i = -1;
        auto exact = [&](const double* ce, int c, int slot) {
            const double cx = ce[3], cy = ce[4];
            if (slot >= 0) {
                const int4 w = s_win[slot];
                if ((unsigned)(x - w.x) >= (unsigned)w.y || (unsigned)(y - w.z) >= (unsigned)w.w) return;
            } else if (!covers(cx, cy)) {
                return;
            }
            const double dc = __dsqrt_rn(__dadd_rn(__dadd_rn(sq(__dsub_rn(ce[0], L)), sq(__dsub_rn(ce[1], A))), sq(__dsub_rn(ce[2], B))));
            const double ds = __dsqrt_rn(__dadd_rn(sq(__dsub_rn(cx, (double)x)), sq(__dsub_rn(cy, (double)y))));
            const double d = __dsqrt_rn(__dadd_rn(sq(__ddiv_rn(dc, (double)nc)), sq(__ddiv_rn(ds, (double)step))));  // ns = step (:105)
            if (d < best || (d == best && best_c >= 0 && c < best_c)) { best = d; best_c = c; best_i = slot; }
        };
        if (staged) {
            for (int i = 0; i < n_cand; ++i) exact(s_cent[i], s_idx[i], i);
        } else {
            for (int gy = gy_lo; gy <= gy_hi; ++gy)
                for (int gx = gx_lo; gx <= gx_hi; ++gx) {
                    const int b = gy * bx + gx;
                    for (int k = offset[b]; k < offset[b + 1]; ++k) exact(centers + (size_t)items[k] * 5, items[k], -1);
                }
        }
    }
    int label = -1;
    if (live) {
        label = labels[(size_t)y * cols + x];
        if (best_c >= 0) {
            label = best_c;
            labels[(size_t)y * cols + x] = label;
        }
    }
    // centre sums (integers: exact in any order).  Pixels that were (re)assigned to a staged centre add into the tile's
    // shared-memory sums, flushed with one global atomic per centre and field at the end -- same-address global atomics
    // were the bottleneck (357 serialised updates per address and iteration).  Pixels keeping a stale label add directly.
    // a warp is 2 x 16 neighbouring pixels with one to three distinct winners: reduce per winner with the hardware warp
    // reduction and let one lane add to shared memory (32 lanes hitting one shared-memory address serialise)
    {
        const int key = (live && label != -1 && best_c >= 0 && best_i >= 0) ? best_i : -1;
        unsigned todo = __ballot_sync(0xffffffffu, key >= 0);
        while (todo) {
            const int leader = __ffs((int)todo) - 1;
            const int k0 = __shfl_sync(0xffffffffu, key, leader);
            const bool mine = key == k0;
            const unsigned v0 = __reduce_add_sync(0xffffffffu, mine ? (unsigned)p[0] : 0u), v1 = __reduce_add_sync(0xffffffffu, mine ? (unsigned)p[1] : 0u),
                           v2 = __reduce_add_sync(0xffffffffu, mine ? (unsigned)p[2] : 0u), v3 = __reduce_add_sync(0xffffffffu, mine ? (unsigned)x : 0u),
                           v4 = __reduce_add_sync(0xffffffffu, mine ? (unsigned)y : 0u), v5 = __reduce_add_sync(0xffffffffu, mine ? 1u : 0u);
            if ((tid & 31) == leader) {
                atomicAdd(&s_sum[k0][0], v0);
                atomicAdd(&s_sum[k0][1], v1);
                atomicAdd(&s_sum[k0][2], v2);
                atomicAdd(&s_sum[k0][3], v3);
                atomicAdd(&s_sum[k0][4], v4);
                atomicAdd(&s_sum[k0][5], v5);
            }
            todo &= ~__ballot_sync(0xffffffffu, mine);
        }
        if (live && label != -1 && key < 0) {  // a stale label (no window covers the pixel) or the unstaged fallback
            unsigned long long* sg = sums + (size_t)label * 6;
            atomicAdd(sg + 0, (unsigned long long)p[0]);
            atomicAdd(sg + 1, (unsigned long long)p[1]);
            atomicAdd(sg + 2, (unsigned long long)p[2]);
            atomicAdd(sg + 3, (unsigned long long)x);
            atomicAdd(sg + 4, (unsigned long long)y);
            atomicAdd(sg + 5, 1ull);
        }
    }
    __syncthreads();
    if (staged)
        for (int i = tid; i < n_cand * 6; i += 256) {
            const int slot = i / 6, k = i - slot * 6;
            const unsigned v = s_sum[slot][k];
            if (v) atomicAdd(sums + (size_t)s_idx[slot] * 6 + k, (unsigned long long)v);
        }
}

// ---- assignment over BANDS of rows with the whole frame's centres in shared memory ---------------------------------------
// k_slic_assign stages the candidate centres of every 16 x 16 tile from global memory (offset -> item -> centre, dependent
// loads) between two barriers, and every pixel walks its own 3 x 3 bins in double precision: ~900 instructions per pixel and
// iteration, 40 % issue-active.  Here a CTA copies ALL centres of its frame -- in bin order, with their integer windows, a
// float record and the bin offsets -- into shared memory once (K x 72 bytes: 90 KB at KITTI size, two CTAs per SM) and streams a band of rows
// through them.  A warp owns a strip 32 pixels wide and walks down its rows:
//   * the centres whose window meets the strip's columns at all are found once per row of bins (one candidate per lane, kept
//     in a register); per image row one compare + ballot gives the ~7 that also cover the row, and the warp walks THAT list
//     together: the centre is the same for all lanes (broadcast loads), two centres per trip;
//   * the walk runs in float.  Its result is taken only where the winner beats the runner-up by more than the error bound of
//     the float evaluation (e_abs, kRel below); the other pixels -- exact ties on symmetric pixels, near ties -- are decided
//     by the double-precision two-stage procedure of k_slic_assign (slic_resolve), i.e. by the reference's own arithmetic --
//     after the walk, from a short list in shared memory (a call inside the walk cost it its registers: spills, stalled loads);
//   * going down a strip a lane keeps its winner for about `step` rows: the sums of the new centres are run-length
//     accumulated in registers and added to the CTA's shared-memory sums when the winner changes.
// Same candidates (the window test is the exact integer one), same winner, same sums: labels and centres stay bit-identical.
//
// Error bound of the float stage.  With eps = 2^-24, colours in [0, 255], |c - p| < step + 1 for the coordinates of a covering
// centre and M = max(rows, cols): a colour difference is off by at most 2^-15 (conversion of the centre + rounding of the
// subtraction), its square by 2^-15 * 511; a coordinate difference by 2^-23 M, its square by 2^-22 M (step + 2).  So
//   |q_float - q| <= E + 1e-6 q,   E = 0.0468 / nc^2 + 2^-21 M (step + 2) / step^2
// (six roundings at most on the way of any term through the sum of positive terms: 3.6e-7 relative), and
//   q1_float + 2.5 E < q2_float (1 - 4e-6)   =>   q1 < q2 by far more than the 1e-12 the double stage itself asks for.
#ifndef DCMT_SLIC_BAND_THREADS
#define DCMT_SLIC_BAND_THREADS 384
#endif
constexpr int kBandThreads = DCMT_SLIC_BAND_THREADS;
#ifndef DCMT_SLIC_BAND_CTAS
#define DCMT_SLIC_BAND_CTAS 2  // CTAs per SM the register allocation aims at
#endif
#ifndef DCMT_SLIC_CENT_SMEM
#define DCMT_SLIC_CENT_SMEM 0  // 1: the centres in double precision in shared memory as well (slic_resolve reads them; + 40 bytes per centre)
#endif
#ifndef __CUDACC__
#define __noinline__ __attribute__((noinline))
#endif

#ifndef DCMT_SLIC_PREFETCH_ROWS
#define DCMT_SLIC_PREFETCH_ROWS 4
#endif
constexpr int kBandPrefetchRows = DCMT_SLIC_PREFETCH_ROWS;  // 0: no L2 prefetch
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if !defined(DCMT_EMU)
    if (kBandPrefetchRows > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
constexpr int kBandQueue = 1024;  // deferred pixels a CTA lists (more: it walks its band again)
struct BandSmem {  // layout of the dynamic shared memory for K centres and NB bins
    size_t off_cent, off_win, off_rec, off_sum, off_bin, off_queue, total;
};
static inline BandSmem band_smem(int K, int NB) {
    BandSmem b;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = (o + bytes + 15) & ~size_t(15); return at; };  // 16-byte aligned parts
    b.off_cent = take(DCMT_SLIC_CENT_SMEM ? (size_t)K * 5 * sizeof(double) : 0);
    b.off_win = take((size_t)K * sizeof(int4));
    b.off_rec = take((size_t)K * 2 * sizeof(float4));
    b.off_sum = take((size_t)K * 6 * sizeof(unsigned));
    b.off_bin = take((size_t)(NB + 2) * sizeof(int));  // + the CTA's count of deferred pixels
    b.off_queue = take((size_t)kBandQueue * sizeof(int));
    b.total = o;
    return b;
}

struct BandView {  // what slic_resolve needs of the CTA's shared memory
    const double* cent;  // [k][5], bin order
    const int4* win;     // {lo_x, width_x, lo_y, width_y}
    const float4* rec;   // [k][2]: {L, a, b, x} and {y, lo_x, width_x, original centre index (the last three as integer bits)}
    const int* bin;      // exclusive offsets, nbins + 1
    int bx, by, nc, step;
    double inv_nc2, inv_ns2;
};

// The winner of one pixel by the reference's arithmetic: stage 1, the square-root- and division-free stand-in for compute_dist
// in double (see k_slic_assign); stage 2, compute_dist itself (:61-69, strict <, lowest centre index) when stage 1 is within
// 1e-12.  Returns the winner's slot in the bin-ordered list (-1: no window covers the pixel).  Rare (pixels the float walk
// could not call): kept out of line.
__device__ __noinline__ int slic_resolve(const BandView& v, int x, int y, unsigned pix) {
    const double L = pix & 255u, A = (pix >> 8) & 255u, B = pix >> 16;
    const int pbx = min(x / v.step, v.bx - 1), pby = min(y / v.step, v.by - 1);
    const int gy_lo = max(pby - 1, 0), gy_hi = min(pby + 1, v.by - 1), gx_lo = max(pbx - 1, 0), gx_hi = min(pbx + 1, v.bx - 1);
    const double xd = (double)x, yd = (double)y;
    double q1 = 1.0e300, q2 = 1.0e300;
    int best_c = -1, best_k = -1;
    for (int gy = gy_lo; gy <= gy_hi; ++gy) {
        const int k0 = v.bin[gy * v.bx + gx_lo], k1 = v.bin[gy * v.bx + gx_hi + 1];  // bins of a row are contiguous
        for (int k = k0; k < k1; ++k) {
            const int4 w = v.win[k];
            if ((unsigned)(x - w.x) >= (unsigned)w.y || (unsigned)(y - w.z) >= (unsigned)w.w) continue;
            const double* ce = v.cent + (size_t)k * 5;
            const double d0 = ce[0] - L, d1 = ce[1] - A, d2 = ce[2] - B, dx = ce[3] - xd, dy = ce[4] - yd;
            const double q = (d0 * d0 + d1 * d1 + d2 * d2) * v.inv_nc2 + (dx * dx + dy * dy) * v.inv_ns2;
            const int c = __float_as_int(v.rec[2 * k + 1].w);
            if (q < q1 || (q == q1 && c < best_c)) { q2 = q1; q1 = q; best_c = c; best_k = k; }
            else if (q < q2) q2 = q;
        }
    }
    if (best_c >= 0 && !(q2 > q1 * (1.0 + 1.0e-12))) {
        double best = (double)FLT_MAX;  // :117
        best_c = -1;
        best_k = -1;
        for (int gy = gy_lo; gy <= gy_hi; ++gy) {
            const int k0 = v.bin[gy * v.bx + gx_lo], k1 = v.bin[gy * v.bx + gx_hi + 1];
            for (int k = k0; k < k1; ++k) {
                const int4 w = v.win[k];
                if ((unsigned)(x - w.x) >= (unsigned)w.y || (unsigned)(y - w.z) >= (unsigned)w.w) continue;
                const double* ce = v.cent + (size_t)k * 5;
                const double dc = __dsqrt_rn(__dadd_rn(__dadd_rn(sq(__dsub_rn(ce[0], L)), sq(__dsub_rn(ce[1], A))), sq(__dsub_rn(ce[2], B))));
                const double ds = __dsqrt_rn(__dadd_rn(sq(__dsub_rn(ce[3], xd)), sq(__dsub_rn(ce[4], yd))));
                const double d = __dsqrt_rn(__dadd_rn(sq(__ddiv_rn(dc, (double)v.nc)), sq(__ddiv_rn(ds, (double)v.step))));  // ns = step (:105)
                const int c = __float_as_int(v.rec[2 * k + 1].w);
                if (d < best || (d == best && best_c >= 0 && c < best_c)) { best = d; best_c = c; best_k = k; }
            }
        }
    }
    return best_k;
}

__global__ void __launch_bounds__(kBandThreads, DCMT_SLIC_BAND_CTAS) k_slic_assign_band(const uint8_t* __restrict__ lab, int rows, int cols, int step, int nc,
                                                                    int bx, int by, uint32_t step_magic, int band_rows,
                                                                    const int* __restrict__ offset, const int* __restrict__ items,
                                                                    const double* __restrict__ sorted, int32_t* __restrict__ labels,
                                                                    unsigned long long* __restrict__ sums, int n_centers, double inv_nc2,
                                                                    double inv_ns2, float e_abs, int cand_max, BandSmem lay) {
    DCMT_DYN_SMEM(unsigned char, smem);
    double* s_cent = reinterpret_cast<double*>(smem + lay.off_cent);
    int4* s_win = reinterpret_cast<int4*>(smem + lay.off_win);
    float4* s_rec = reinterpret_cast<float4*>(smem + lay.off_rec);
    unsigned* s_sum = reinterpret_cast<unsigned*>(smem + lay.off_sum);  // [k][6]
    int* s_bin = reinterpret_cast<int*>(smem + lay.off_bin);
    int* s_deferred = s_bin + (bx * by + 1);  // pixels of this CTA left to slic_resolve ...
    int* s_queue = reinterpret_cast<int*>(smem + lay.off_queue);  // ... and the offsets of the first kBandQueue of them
    constexpr int kDeferred = 0x40000000;   // marker in `labels`: kDeferred | (old label + 1); labels themselves stay far below
    const int nbins = bx * by;
    {  // frame
        const size_t f = blockIdx.y;
        lab += f * rows * cols * 3;
        labels += f * rows * cols;
        sums += f * n_centers * 6;
        offset += f * (nbins + 1);
        items += f * n_centers;
        sorted += f * n_centers * 5;
    }
    const int tid = threadIdx.x;
    if (tid == 0) *s_deferred = 0;
    for (int b = tid; b <= nbins; b += kBandThreads) s_bin[b] = offset[b];
    __syncthreads();
    const int n_binned = s_bin[nbins];  // centres that are not NaN
    for (int i = tid; i < n_binned * 6; i += kBandThreads) s_sum[i] = 0u;
    auto add_stale = [&](int label, unsigned px, int x, int y) {  // a pixel that keeps its label counts for that centre: global memory directly
        if (label == -1) return;
        unsigned long long* sg = sums + (size_t)label * 6;
        atomicAdd(sg + 0, (unsigned long long)(px & 255u));
        atomicAdd(sg + 1, (unsigned long long)((px >> 8) & 255u));
        atomicAdd(sg + 2, (unsigned long long)(px >> 16));
        atomicAdd(sg + 3, (unsigned long long)x);
        atomicAdd(sg + 4, (unsigned long long)y);
        atomicAdd(sg + 5, 1ull);
    };
    for (int k = tid; k < n_binned; k += kBandThreads) {
        double ce[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) ce[q] = sorted[(size_t)k * 5 + q];
        if (DCMT_SLIC_CENT_SMEM) {
#pragma unroll
            for (int q = 0; q < 5; ++q) s_cent[(size_t)k * 5 + q] = ce[q];
        }
        // for (int k = cx - step; k < cx + step; k++) (:123): truncation towards zero, then a double comparison; for integer k,
        // k < c + step <=> k < ceil(c + step)
        const int lox = (int)__dsub_rn(ce[3], (double)step), loy = (int)__dsub_rn(ce[4], (double)step);
        const int hix = (int)ceil(__dadd_rn(ce[3], (double)step)), hiy = (int)ceil(__dadd_rn(ce[4], (double)step));
        s_win[k] = make_int4(lox, hix - lox, loy, hiy - loy);
        s_rec[2 * k] = make_float4((float)ce[0], (float)ce[1], (float)ce[2], (float)ce[3]);
        s_rec[2 * k + 1] = make_float4((float)ce[4], __int_as_float(lox), __int_as_float(hix - lox), __int_as_float(items[k]));
    }
    __syncthreads();
    const int y_begin = blockIdx.x * band_rows, n_rows = min(rows, y_begin + band_rows) - y_begin;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kBandThreads / 32;
    // one item = 32 consecutive pixels of one row; the items of the band in strip-major order (down the first strip of 32 columns,
    // then down the next), a contiguous run of them per warp
    const int n_strips = (cols + 31) >> 5, n_items = n_rows * n_strips, per_warp = (n_items + kWarps - 1) / kWarps;
    int item = warp * per_warp;
    const int item_end = min(item + per_warp, n_items);
    int strip = item / max(n_rows, 1), row = item - strip * n_rows;
    const float inc = (float)inv_nc2, ins = (float)inv_ns2;
    const float kInf = __int_as_float(0x7f800000), kRel = 1.0f - 4.0e-6f;
    // offset of the lane's pixel in the frame (the host keeps rows x cols x 3 below 2^31 for this kernel): + cols per row of the strip
    auto px_offset = [&](int r, int s) { return (y_begin + r) * cols + s * 32 + lane; };
    int off = px_offset(row, strip);
    unsigned pix = 0u;  // L | a << 8 | b << 16 of the lane's pixel, 0 beyond the row
    if (item < item_end && strip * 32 + lane < cols) {
        const uint8_t* p = lab + (size_t)off * 3;
        pix = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
    }
    // the lane's candidate: a centre of the bins around (strip, row of bins) whose window meets the strip's columns (its slot; -1: none;
    // -2: more than 32 centres in those bins -- pathological clustering -- every pixel asks slic_resolve), the rows of its window as
    // lo << 12 | height, and the item at which the warp leaves the row of bins or the strip and looks again
    int cand = -1, cand_rows = 0, rebuild_at = item;
    // run of consecutive rows of the strip with the same winner: its sums wait in registers, two 16-bit fields per register (a run is
    // cut at 255 rows: 255 x 255 < 2^16).  The rows of a run are consecutive and its column is the lane's, so the coordinate sums
    // follow from the count: n x, and n y_after - n (n + 1) / 2 for the n rows that end just before row y_after.
    int run_key = -1;
    unsigned run_La = 0u, run_bn = 0u;  // L | a << 16,  b | n << 16
    auto flush_run = [&](int xv, int y_after) {
        if (run_key >= 0) {
            unsigned* sg = s_sum + (size_t)run_key * 6;
            const unsigned n = run_bn >> 16;
            atomicAdd(sg + 0, run_La & 0xffffu); atomicAdd(sg + 1, run_La >> 16); atomicAdd(sg + 2, run_bn & 0xffffu);
            atomicAdd(sg + 3, n * (unsigned)xv); atomicAdd(sg + 4, n * (unsigned)y_after - n * (n + 1u) / 2u); atomicAdd(sg + 5, n);
        }
        run_La = run_bn = 0u;
        run_key = -1;
    };
    for (; item < item_end; ++item) {
        const int y = y_begin + row, x0 = strip * 32, x = x0 + lane;
        const bool live = x < cols;
        const int cur_off = off;
        if (item == rebuild_at) {  // warp-uniform: every `step` rows
            if (row == 0) flush_run(x - 32, y_begin + n_rows);  // the runs of the strip before end with it
            const int pby = min(fast_div(y, step_magic), by - 1);
            const int rows_in_cell = pby < by - 1 ? (pby + 1) * step - y : n_rows;
            rebuild_at = item + min(rows_in_cell, n_rows - row);
            const int gy_lo = max(pby - 1, 0), gy_hi = min(pby + 1, by - 1);
            const int gx_lo = max(min(fast_div(x0, step_magic), bx - 1) - 1, 0), gx_hi = min(min(fast_div(x0 + 31, step_magic), bx - 1) + 1, bx - 1);
            int k = -1, before = 0;
            for (int gy = gy_lo; gy <= gy_hi; ++gy) {
                const int k0 = s_bin[gy * bx + gx_lo], k1 = s_bin[gy * bx + gx_hi + 1];  // bins of a row are contiguous
                if (lane >= before && lane < before + (k1 - k0)) k = k0 + (lane - before);
                before += k1 - k0;
            }
            cand = -1;
            if (k >= 0) {
                const int4 w = s_win[k];
                if (w.x < x0 + 32 && w.x + w.y > x0) { cand = k; cand_rows = (w.z << 12) | w.w; }
            }
            if (before > cand_max) cand = -2;  // 32, the lanes of a warp (less in tests of this path)
        }
        // the next item's pixel is on its way while this one is worked on: three byte loads now, put together at the bottom of the
        // loop (warps issue in order: an instruction that needs the bytes would wait here for them)
        if (++row == n_rows) { row = 0; ++strip; off = px_offset(0, strip); } else off += cols;
        unsigned next_L = 0u, next_a = 0u, next_b = 0u;
        if (item + 1 < item_end && strip * 32 + lane < cols) {
            const uint8_t* p = lab + (size_t)off * 3;
            next_L = p[0]; next_a = p[1]; next_b = p[2];
            // ... and the rows after it on their way into L2 (one item of work did not cover the latency of these loads from DRAM)
            if (row + kBandPrefetchRows < n_rows) prefetch_l2(p + (size_t)kBandPrefetchRows * cols * 3);
        }
        const float xf = (float)x, yf = (float)y;
        float q1 = kInf, q2 = kInf;  // smallest and second smallest float distance over the covering centres
        int best = -1;               // slot of the smallest
        unsigned todo = __ballot_sync(0xffffffffu, cand >= 0 && (unsigned)(y - (cand_rows >> 12)) < (unsigned)(cand_rows & 4095));
        while (todo) {  // two centres per trip (the second may be missing), the same two for every lane
            const int j0 = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const bool two = todo != 0u;
            const int j1 = two ? __ffs((int)todo) - 1 : j0;
            todo &= todo - 1;
            const int ka = __shfl_sync(0xffffffffu, cand, j0), kb = __shfl_sync(0xffffffffu, cand, j1);
            const float4 a0 = s_rec[2 * ka], a1 = s_rec[2 * ka + 1], b0 = s_rec[2 * kb], b1 = s_rec[2 * kb + 1];
            const float L = (float)(pix & 255u), A = (float)((pix >> 8) & 255u), B = (float)(pix >> 16);  // here: three registers less across the loop
            float qa, qb;
            {
                const float d0 = a0.x - L, d1 = a0.y - A, d2 = a0.z - B, dx = a0.w - xf, dy = a1.x - yf;
                qa = (d0 * d0 + d1 * d1 + d2 * d2) * inc + (dx * dx + dy * dy) * ins;
                if ((unsigned)(x - __float_as_int(a1.y)) >= (unsigned)__float_as_int(a1.z)) qa = kInf;  // the window misses this column
            }
            {
                const float d0 = b0.x - L, d1 = b0.y - A, d2 = b0.z - B, dx = b0.w - xf, dy = b1.x - yf;
                qb = (d0 * d0 + d1 * d1 + d2 * d2) * inc + (dx * dx + dy * dy) * ins;
                if (!two || (unsigned)(x - __float_as_int(b1.y)) >= (unsigned)__float_as_int(b1.z)) qb = kInf;
            }
            q2 = fminf(q2, fmaxf(q1, qa));
            if (qa < q1) best = ka;
            q1 = fminf(q1, qa);
            q2 = fminf(q2, fmaxf(q1, qb));
            if (qb < q1) best = kb;
            q1 = fminf(q1, qb);
        }
        int key = -1;
        bool deferred = false;  // too close for floats (or no candidate list): decided after the walk, by slic_resolve
        if (live) {
            if (cand == -2) deferred = true;
            else if (best >= 0) {
                if (q1 + e_abs < q2 * kRel) key = best;
                else deferred = true;
            }
        }
        if (key >= 0) labels[cur_off] = __float_as_int(s_rec[2 * key + 1].w);
        if (deferred) {  // rare: leave a marker that keeps the old label (it stands if no window covers the pixel after all)
            labels[cur_off] = kDeferred | (labels[cur_off] + 1);
            const int slot = atomicAdd(s_deferred, 1);
            if (slot < kBandQueue) s_queue[slot] = cur_off;
        }
        // centre sums (integers: exact in any order)
        if (key != run_key || run_bn >= (255u << 16)) {
            flush_run(x, y);
            run_key = key;
        }
        if (key >= 0) {
            run_La += __byte_perm(pix, 0u, 0x4140);  // L | a << 16
            run_bn += (pix >> 16) + 0x10000u;         // b | 1 << 16
        } else if (live && !deferred) {  // no window covers the pixel: it keeps its label and counts for that centre
            add_stale(labels[cur_off], pix, x, y);
        }
        pix = next_L | (next_a << 8) | (next_b << 16);
    }
    if (item_end > warp * per_warp) {  // the warp had items: its last one tells where its open runs end
        const int last = item_end - 1, last_strip = last / n_rows;
        flush_run(last_strip * 32 + lane, y_begin + (last - last_strip * n_rows) + 1);
    }
    __syncthreads();
    const int n_deferred = *s_deferred;
    if (n_deferred > 0) {
        // the marked pixels by the reference's arithmetic: from the queue, one per thread; when it overflowed (flat images: every pixel
        // ties), in a second walk over the same items (every lane meets the pixels it marked itself)
        BandView view;
        view.cent = DCMT_SLIC_CENT_SMEM ? s_cent : sorted; view.win = s_win; view.rec = s_rec; view.bin = s_bin;
        view.bx = bx; view.by = by; view.nc = nc; view.step = step; view.inv_nc2 = inv_nc2; view.inv_ns2 = inv_ns2;
        auto decide = [&](int o, int x, int y) {
            const int marker = labels[o];
            if (marker < kDeferred) return;
            const uint8_t* p = lab + (size_t)o * 3;
            const unsigned px = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
            const int key = slic_resolve(view, x, y, px);
            if (key >= 0) {
                labels[o] = __float_as_int(s_rec[2 * key + 1].w);
                unsigned* sg = s_sum + (size_t)key * 6;
                atomicAdd(sg + 0, px & 255u); atomicAdd(sg + 1, (px >> 8) & 255u); atomicAdd(sg + 2, px >> 16);
                atomicAdd(sg + 3, (unsigned)x); atomicAdd(sg + 4, (unsigned)y); atomicAdd(sg + 5, 1u);
            } else {
                const int old_label = (marker & (kDeferred - 1)) - 1;
                labels[o] = old_label;
                add_stale(old_label, px, x, y);
            }
        };
        if (n_deferred <= kBandQueue) {
            for (int i = tid; i < n_deferred; i += kBandThreads) {
                const int o = s_queue[i], y = o / cols;
                decide(o, o - y * cols, y);
            }
        } else {
            item = warp * per_warp;
            strip = item / max(n_rows, 1);
            row = item - strip * n_rows;
            for (; item < item_end; ++item) {
                const int x = strip * 32 + lane, y = y_begin + row, o = px_offset(row, strip);
                if (++row == n_rows) { row = 0; ++strip; }
                if (x < cols) decide(o, x, y);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n_binned * 6; i += kBandThreads) {
        const unsigned v = s_sum[i];
        if (v) atomicAdd(sums + (size_t)__float_as_int(s_rec[2 * (i / 6) + 1].w) * 6 + (i % 6), (unsigned long long)v);
    }
}

__global__ void k_slic_update(const unsigned long long* __restrict__ sums, int n, double* __restrict__ centers) {  // :166-172
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    sums += (size_t)blockIdx.y * n * 6;  // frame
    centers += (size_t)blockIdx.y * n * 5;
    const double cnt = (double)sums[(size_t)c * 6 + 5];
    for (int k = 0; k < 5; ++k) centers[(size_t)c * 5 + k] = __ddiv_rn((double)sums[(size_t)c * 6 + k], cnt);
}

}  // namespace

static void slic_grid(int rows, int cols, int step, int* nx, int* ny) {
    int a = 0, b = 0;
    for (int i = step; i < cols - step / 2; i += step) ++a;
    for (int j = step; j < rows - step / 2; j += step) ++b;
    *nx = a;
    *ny = b;
}

int slic_center_count(int rows, int cols, int step) {
    if (step < 1) return 0;
    int nx, ny;
    slic_grid(rows, cols, step, &nx, &ny);
    return nx * ny;
}

size_t slic_bins(int rows, int cols, int step, int* bins_x, int* bins_y) {
    *bins_x = (cols + step - 1) / step;
    *bins_y = (rows + step - 1) / step;
    return (size_t)*bins_x * *bins_y;
}

cudaError_t slic_run(const uint8_t* lab, int rows, int cols, int n_frames, int step, int nc, int iterations, int32_t* labels,
                     int n_centers, const SlicWork& w, cudaStream_t st) {
    int nx, ny;
    slic_grid(rows, cols, step, &nx, &ny);
    const size_t npx = (size_t)rows * cols * n_frames;
    const unsigned cb = (unsigned)((n_centers + 127) / 128);
    DCMT_LAUNCH(k_slic_fill_labels, dim3((unsigned)((npx + 255) / 256)), dim3(256), 0, st, labels, npx);
    if (n_centers == 0 || n_frames == 0) return cudaGetLastError();
    DCMT_LAUNCH(k_slic_init, dim3(cb, n_frames), dim3(128), 0, st, lab, rows, cols, step, ny, n_centers, w.centers);
    // assignment kernel: bands of rows with the frame's centres in shared memory when they fit (K x 72 bytes: two CTAs per SM at KITTI size),
    // else 16 x 16 tiles.  Measured on the B200 (step 18, 1 254 centres), frames/s for batches of 1 / 2 / 8 / 64 / 256 frames:
    // bands 2.1 k / 4.1 k / 9.2 k / 12.7 k / 14.1 k, tiles 2.5 k / 3.4 k / 4.9 k / 3.0 k / 3.0 k -- a single frame keeps the tiles
    // (a band CTA copies all centres first).  DCMT_SLIC_BAND_MIN_FRAMES: batch size from which the band kernel is used.
    const BandSmem lay = band_smem(n_centers, w.bins_x * w.bins_y);
    static const int band_min = [] { const char* e = getenv("DCMT_SLIC_BAND_MIN_FRAMES"); return e ? atoi(e) : 2; }();
    // DCMT_SLIC_CAND_MAX (tests): fewer than 32 candidates per warp before the walk gives up and every pixel asks slic_resolve
    static const int cand_max = [] { const char* e = getenv("DCMT_SLIC_CAND_MAX"); const int v = e ? atoi(e) : 32; return v < 0 ? 0 : (v > 32 ? 32 : v); }();
    constexpr size_t kBandSmemMax = 200 * 1024;
    const bool bands_fit = lay.total <= kBandSmemMax && n_frames >= band_min && (long long)rows * cols * 3 < (1ll << 31) && step <= 2000;
    int n_bands = 1, band_rows = rows;
    float e_abs = 0.f;
    if (bands_fit) {
        // one or two CTAs per SM at a time: the number of bands per frame (of at least 4 rows: every CTA copies the whole frame's centres) that
        // fills the last wave of CTAs best, slightly preferring fewer bands
        const int max_bands = rows / 4 < 1 ? 1 : (rows / 4 > 64 ? 64 : rows / 4);
        const long per_sm = (233472 / (lay.total + 1024) >= 2 && 2 * kBandThreads <= 2048) ? 2 : 1, wave = 148 * per_sm;  // resident CTAs
        double best_score = -1.0;
        for (int nb = 1; nb <= max_bands; ++nb) {
            const double ctas = (double)nb * n_frames, waves = (double)((nb * (long)n_frames + wave - 1) / wave);
            const double score = ctas / (waves * wave) - 0.002 * nb;
            if (score > best_score) { best_score = score; n_bands = nb; }
        }
        band_rows = (rows + n_bands - 1) / n_bands;
        // a CTA sums coordinates in 32 bits: band pixels x largest coordinate must stay below 2^32
        while (band_rows > 1 && (unsigned long long)band_rows * cols * (rows > cols ? rows : cols) >= (1ull << 32)) band_rows = (band_rows + 1) / 2;
        n_bands = (rows + band_rows - 1) / band_rows;
        cudaError_t e = cudaFuncSetAttribute(k_slic_assign_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBandSmemMax);
        if (e != cudaSuccess) return e;
        // 2.5 E of the float stage (see k_slic_assign_band), rounded up
        const double M = rows > cols ? rows : cols;
        const double E = 0.0468 / ((double)nc * (double)nc) + M * ((double)step + 2.0) / ((double)step * (double)step) / 2097152.0;
        e_abs = (float)(2.5 * E * 1.001);
    }
    const uint32_t step_magic = (uint32_t)((1ull << 32) / (unsigned)step) + 1u;
    for (int it = 0; it < iterations; ++it) {
        DCMT_LAUNCH(k_slic_prepare, dim3(n_frames), dim3(kScanThreads), 0, st, w.centers, w.sums, n_centers, step, w.bins_x, w.bins_y,
                    w.bin_count, w.bin_fill, w.bin_items, w.sorted, it > 0 ? 1 : 0);
        if (bands_fit)
            DCMT_LAUNCH(k_slic_assign_band, dim3(n_bands, n_frames), dim3(kBandThreads), lay.total, st, lab, rows, cols, step, nc, w.bins_x,
                        w.bins_y, step_magic, band_rows, w.bin_count, w.bin_items, w.sorted, labels, w.sums, n_centers,
                        1.0 / ((double)nc * (double)nc), 1.0 / ((double)step * (double)step), e_abs, cand_max, lay);
        else
            DCMT_LAUNCH(k_slic_assign, dim3((cols + 15) / 16, (rows + 15) / 16, n_frames), dim3(16, 16), 0, st, lab, rows, cols, step, nc,
                        w.bins_x, w.bins_y, w.centers, w.bin_count, w.bin_items, w.sorted, labels, w.sums, n_centers,
                        1.0 / ((double)nc * (double)nc), 1.0 / ((double)step * (double)step));
    }
    if (iterations > 0) DCMT_LAUNCH(k_slic_update, dim3(cb, n_frames), dim3(128), 0, st, w.sums, n_centers, w.centers);
    return cudaGetLastError();
}

}  // namespace dcmt
