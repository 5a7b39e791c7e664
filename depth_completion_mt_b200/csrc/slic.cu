// slic.cu -- Slic::generate_superpixels (src/DC_lidar_camera/slic.cpp:19-182) on the device.
//
// The reference loops over the centres and, for each, over the pixels of its 2 step x 2 step window, keeping per pixel
// the smallest distance seen (strict <, so the lowest centre index wins ties, :131-134).  Here every pixel looks up the
// centres whose window covers it (centres are binned on a step-sized grid each iteration; the covering ones lie in the
// 3 x 3 bins around the pixel) and takes the minimum over (distance, index) -- the same result, all pixels in parallel.
//   * window test exactly as written: k from (int)(cx - step) while k < cx + step (:123-124), double arithmetic;
//   * distance exactly as compute_dist (:61-69): doubles, pow(x, 2) = x * x, sqrt, no FMA contraction;
//   * pixels no window covers keep their previous label (the reference only resets `distances`, :115-119);
//   * centre update (:141-171) sums integers (pixel values, coordinates): exact in any order, done with 64-bit
//     integer atomics, then divided in double; an empty cluster becomes 0 / 0 = NaN and never wins a pixel again.
// So the labels are bit-identical to a scalar build of the reference, not merely "mostly equal".
// create_connectivity (:186-260) writes only its local `new_clusters`, never `clusters`: nothing to do.
#include "slic.cuh"

namespace dcmt {
namespace {

__device__ __forceinline__ double sq(double x) { return __dmul_rn(x, x); }

// slic.cpp:72-99, one thread per centre
__global__ void k_slic_init(const uint8_t* __restrict__ lab, int rows, int cols, int step, int ny, int n, double* __restrict__ centers) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
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

__global__ void k_slic_bin_count(const double* __restrict__ centers, int n, int step, int bx, int by, int* __restrict__ count,
                                 unsigned long long* __restrict__ sums) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int k = 0; k < 6; ++k) sums[(size_t)c * 6 + k] = 0ull;
    const double x = centers[(size_t)c * 5 + 3], y = centers[(size_t)c * 5 + 4];
    if (x != x || y != y) return;  // NaN centre (empty cluster): covers nothing
    atomicAdd(count + bin_of(y, step, by) * bx + bin_of(x, step, bx), 1);
}

// exclusive scan of the bin counts by one block: thread t owns `chunk` consecutive bins, the per-thread totals are
// scanned in shared memory (Hillis-Steele), then every thread writes the offsets of its bins
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) k_slic_bin_scan(int* __restrict__ count, int* __restrict__ fill, int nbins) {
    __shared__ int part[kScanThreads];
    const int chunk = (nbins + kScanThreads - 1) / kScanThreads;
    const int b0 = threadIdx.x * chunk, b1 = min(b0 + chunk, nbins);
    int total = 0;
    for (int b = b0; b < b1; ++b) total += count[b];
    part[threadIdx.x] = total;
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int acc = part[threadIdx.x] - total;  // exclusive prefix of this thread's chunk
    for (int b = b0; b < b1; ++b) {
        const int c = count[b];
        count[b] = acc;
        fill[b] = 0;
        acc += c;
    }
    if (threadIdx.x == kScanThreads - 1) count[nbins] = part[kScanThreads - 1];
}

__global__ void k_slic_bin_fill(const double* __restrict__ centers, int n, int step, int bx, int by, const int* __restrict__ offset,
                                int* __restrict__ fill, int* __restrict__ items) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double x = centers[(size_t)c * 5 + 3], y = centers[(size_t)c * 5 + 4];
    if (x != x || y != y) return;
    const int b = bin_of(y, step, by) * bx + bin_of(x, step, bx);
    items[offset[b] + atomicAdd(fill + b, 1)] = c;
}

__global__ void k_slic_clear_counts(int* __restrict__ count, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) count[i] = 0;
}

// assignment (:121-139) + accumulation of the new centres (:148-163), one thread per pixel
__global__ void __launch_bounds__(256) k_slic_assign(const uint8_t* __restrict__ lab, int rows, int cols, int step, int nc, int bx, int by,
                                                     const double* __restrict__ centers, const int* __restrict__ offset,
                                                     const int* __restrict__ items, int32_t* __restrict__ labels,
                                                     unsigned long long* __restrict__ sums) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* p = lab + ((size_t)y * cols + x) * 3;
    const double L = p[0], A = p[1], B = p[2];
    const int pbx = x / step < bx ? x / step : bx - 1, pby = y / step < by ? y / step : by - 1;
    double best = (double)FLT_MAX;  // :117
    int best_c = -1;
    for (int gy = max(pby - 1, 0); gy <= min(pby + 1, by - 1); ++gy)
        for (int gx = max(pbx - 1, 0); gx <= min(pbx + 1, bx - 1); ++gx) {
            const int b = gy * bx + gx;
            for (int k = offset[b]; k < offset[b + 1]; ++k) {
                const int c = items[k];
                const double* ce = centers + (size_t)c * 5;
                const double cx = ce[3], cy = ce[4];
                // for (int k = cx - step; k < cx + step; k++) (:123): truncation towards zero, then a double comparison
                if (x < (int)__dsub_rn(cx, (double)step) || !((double)x < __dadd_rn(cx, (double)step))) continue;
                if (y < (int)__dsub_rn(cy, (double)step) || !((double)y < __dadd_rn(cy, (double)step))) continue;
                const double dc = __dsqrt_rn(__dadd_rn(__dadd_rn(sq(__dsub_rn(ce[0], L)), sq(__dsub_rn(ce[1], A))), sq(__dsub_rn(ce[2], B))));
                const double ds = __dsqrt_rn(__dadd_rn(sq(__dsub_rn(cx, (double)x)), sq(__dsub_rn(cy, (double)y))));
                const double d = __dsqrt_rn(__dadd_rn(sq(__ddiv_rn(dc, (double)nc)), sq(__ddiv_rn(ds, (double)step))));  // ns = step (:105)
                if (d < best || (d == best && best_c >= 0 && c < best_c)) { best = d; best_c = c; }
            }
        }
    int label = labels[(size_t)y * cols + x];
    if (best_c >= 0) {
        label = best_c;
        labels[(size_t)y * cols + x] = label;
    }
    if (label != -1) {
        unsigned long long* s = sums + (size_t)label * 6;
        atomicAdd(s + 0, (unsigned long long)p[0]);
        atomicAdd(s + 1, (unsigned long long)p[1]);
        atomicAdd(s + 2, (unsigned long long)p[2]);
        atomicAdd(s + 3, (unsigned long long)x);
        atomicAdd(s + 4, (unsigned long long)y);
        atomicAdd(s + 5, 1ull);
    }
}

__global__ void k_slic_update(const unsigned long long* __restrict__ sums, int n, double* __restrict__ centers) {  // :166-172
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
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

cudaError_t slic_run(const uint8_t* lab, int rows, int cols, int step, int nc, int iterations, int32_t* labels, int n_centers,
                     const SlicWork& w, cudaStream_t st) {
    int nx, ny;
    slic_grid(rows, cols, step, &nx, &ny);
    const size_t npx = (size_t)rows * cols;
    const int nbins = w.bins_x * w.bins_y;
    const unsigned cb = (unsigned)((n_centers + 127) / 128);
    DCMT_LAUNCH(k_slic_fill_labels, dim3((unsigned)((npx + 255) / 256)), dim3(256), 0, st, labels, npx);
    if (n_centers == 0) return cudaGetLastError();
    DCMT_LAUNCH(k_slic_init, dim3(cb), dim3(128), 0, st, lab, rows, cols, step, ny, n_centers, w.centers);
    for (int it = 0; it < iterations; ++it) {
        DCMT_LAUNCH(k_slic_clear_counts, dim3((nbins + 1 + 255) / 256), dim3(256), 0, st, w.bin_count, nbins + 1);
        DCMT_LAUNCH(k_slic_bin_count, dim3(cb), dim3(128), 0, st, w.centers, n_centers, step, w.bins_x, w.bins_y, w.bin_count, w.sums);
        DCMT_LAUNCH(k_slic_bin_scan, dim3(1), dim3(kScanThreads), 0, st, w.bin_count, w.bin_fill, nbins);
        DCMT_LAUNCH(k_slic_bin_fill, dim3(cb), dim3(128), 0, st, w.centers, n_centers, step, w.bins_x, w.bins_y, w.bin_count, w.bin_fill, w.bin_items);
        DCMT_LAUNCH(k_slic_assign, dim3((cols + 255) / 256, rows), dim3(256), 0, st, lab, rows, cols, step, nc, w.bins_x, w.bins_y, w.centers,
                    w.bin_count, w.bin_items, labels, w.sums);
        DCMT_LAUNCH(k_slic_update, dim3(cb), dim3(128), 0, st, w.sums, n_centers, w.centers);
    }
    return cudaGetLastError();
}

}  // namespace dcmt
