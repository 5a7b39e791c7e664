// slic.cuh -- SLIC superpixels, the producer of the label map of interpolate_with_superpixels.
#pragma once
#include "common.cuh"

namespace dcmt {

// number of cluster centres Slic::init_data creates for this image / step (slic.cpp:33-34)
int slic_center_count(int rows, int cols, int step);
struct SlicWork {               // every array once per frame of the batch
    double* centers;             // K x 5: L, a, b, x, y
    unsigned long long* sums;    // K x 6: sums of L, a, b, x, y and the pixel count
    int* bin_count;              // nbins + 1 (counts, then exclusive offsets)
    int* bin_fill;               // nbins
    int* bin_items;              // K
    double* sorted;              // K x 5: the centres in bin order
    int bins_x, bins_y;
};
size_t slic_bins(int rows, int cols, int step, int* bins_x, int* bins_y);
// lab: n_frames x rows x cols x 3 uint8 (cv::Mat CV_8UC3 after COLOR_BGR2Lab), labels: n_frames x rows x cols int32
// row-major (-1 = unassigned); frames are independent and run side by side (grid.z)
cudaError_t slic_run(const uint8_t* lab, int rows, int cols, int n_frames, int step, int nc, int iterations, int32_t* labels,
                     int n_centers, const SlicWork& w, cudaStream_t st);

}  // namespace dcmt
