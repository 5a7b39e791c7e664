// pipe_bench.cu -- instruction-throughput microbenchmark used to size the fused kernels (sm_100a).
// Prints lane-ops per clock per SM for the packed min/max, permute and shared-memory instructions the
// completion kernels are built from.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned seed, long long* cycles) {
    unsigned a[UNROLL], b = seed ^ threadIdx.x, c = seed * 3 + 7;
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = seed + i * 77 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (MODE == 0) a[i] = __vmaxu2(a[i], b);                       // VIMNMX.U16x2
            if (MODE == 1) a[i] = __vimax3_u16x2(a[i], b, c);              // VIMNMX3.U16x2
            if (MODE == 2) { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&b);
                             x = __hmax2(x, y); a[i] = *reinterpret_cast<unsigned*>(&x); }  // HMNMX2
            if (MODE == 3) { if (i & 1) a[i] = __vmaxu2(a[i], b);
                             else { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&b);
                                    x = __hmax2(x, y); a[i] = *reinterpret_cast<unsigned*>(&x); } }  // mixed
            if (MODE == 4) a[i] = __byte_perm(a[i], b, 0x5432);            // PRMT
            if (MODE == 5) a[i] = a[i] * 3 + b;                            // IMAD
            if (MODE == 6) a[i] = (a[i] & b) ^ c;                          // LOP3
            if (MODE == 7) { if (i & 1) a[i] = __vmaxu2(a[i], b); else a[i] = a[i] * 3 + b; }  // VIMNMX + IMAD
            if (MODE == 8) a[i] = __vadd2(a[i], b);                        // VIADD.16x2
            if (MODE == 9) a[i] = sm[(a[i] + threadIdx.x) & 4095];         // LDS.32 dependent-address
            if (MODE == 10) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);  // SHFL
            if (MODE == 11) a[i] = fmaxf(__uint_as_float(a[i]), __uint_as_float(b)) > 0 ? a[i] : b;  // FMNMX-ish
            if (MODE == 12) { if (i & 1) a[i] = __vmaxu2(a[i], b); else a[i] = __byte_perm(a[i], b, 0x5432); }  // VIMNMX+PRMT
        }
        b += 0x00010001u;
    }
    long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// streaming LDS bandwidth: every thread reads W-byte vectors, conflict free
template <int W>
__global__ void __launch_bounds__(1024) k_lds(unsigned* out, long long* cycles) {
    __shared__ __align__(16) unsigned sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < 2048; ++it) {
        const int base = ((it * 1024 + threadIdx.x) * (W / 4)) & 8191 & ~(W / 4 - 1);
        if (W == 4) acc ^= sm[base];
        if (W == 8) { uint2 v = *reinterpret_cast<uint2*>(&sm[base]); acc ^= v.x ^ v.y; }
        if (W == 16) { uint4 v = *reinterpret_cast<uint4*>(&sm[base]); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, unsigned* d_out, long long* d_cyc) {
    k<MODE><<<148, 1024>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, 1024>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    double ops = (double)ITERS * UNROLL * 1024;
    printf("%-28s %8.1f lane-ops/clk/SM   (%.0f cycles)\n", name, ops / avg, avg);
}

template <int W>
void run_lds(const char* name, unsigned* d_out, long long* d_cyc) {
    k_lds<W><<<148, 1024>>>(d_out, d_cyc);
    cudaDeviceSynchronize();
    k_lds<W><<<148, 1024>>>(d_out, d_cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-28s %8.1f bytes/clk/SM\n", name, 2048.0 * 1024 * W / avg);
}

int main() {
    unsigned* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4);
    cudaMalloc(&d_cyc, 148 * 8);
    run<0>("VIMNMX.U16x2", d_out, d_cyc);
    run<1>("VIMNMX3.U16x2", d_out, d_cyc);
    run<2>("HMNMX2", d_out, d_cyc);
    run<3>("VIMNMX + HMNMX2 (1:1)", d_out, d_cyc);
    run<4>("PRMT", d_out, d_cyc);
    run<5>("IMAD", d_out, d_cyc);
    run<6>("LOP3", d_out, d_cyc);
    run<7>("VIMNMX + IMAD (1:1)", d_out, d_cyc);
    run<8>("VIADD.16x2", d_out, d_cyc);
    run<9>("LDS.32 (dependent)", d_out, d_cyc);
    run<10>("SHFL", d_out, d_cyc);
    run<11>("FMNMX + SEL", d_out, d_cyc);
    run<12>("VIMNMX + PRMT (1:1)", d_out, d_cyc);
    run_lds<4>("LDS.32 stream", d_out, d_cyc);
    run_lds<8>("LDS.64 stream", d_out, d_cyc);
    run_lds<16>("LDS.128 stream", d_out, d_cyc);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
