// Test program for include/img_completion.h without OpenCV: reads a raw float32 frame, calls the shim exactly like
// src/DC_lidar_only/main.cpp:93 calls img_completion, writes the dense frame.  Linked against the CPU emulator
// build of the library by tests/test_cpp_shim.py (test infrastructure), or against libdcmt.so on a GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>

#define DCMT_NO_OPENCV 1
#include "img_completion.h"

int main(int argc, char** argv) {
    if (argc != 6) return 2;
    const int rows = std::atoi(argv[1]), cols = std::atoi(argv[2]);
    std::vector<float> in((size_t)rows * cols), out((size_t)rows * cols);
    FILE* f = std::fopen(argv[3], "rb");
    if (!f || std::fread(in.data(), sizeof(float), in.size(), f) != in.size()) return 3;
    std::fclose(f);
    dcmt::MatView sparse{rows, cols, (size_t)cols * sizeof(float), in.data()};
    dcmt::MatView dense{rows, cols, (size_t)cols * sizeof(float), out.data()};
    try {
        dcmt::img_completion(sparse, dense, false, argv[5]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    // the caller's evaluation against its own input as "ground truth" (main.cpp:101, main_lc.cpp:224, main_sl.cpp:1232)
    try {
        float mse = 0, rmse = 0, mae = 0, mae2 = 0, rmse2 = 0;
        dcmt::evaluate_performance(sparse, dense, mse);
        dcmt::evaluate_performance(sparse, dense, rmse, mae);
        dcmt::evaluate_performances(sparse, dense, mae2, rmse2);
        std::printf("%.9g %.9g %.9g %.9g %.9g\n", mse, rmse, mae, mae2, rmse2);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    f = std::fopen(argv[4], "wb");
    std::fwrite(out.data(), sizeof(float), out.size(), f);
    std::fclose(f);
    return 0;
}
