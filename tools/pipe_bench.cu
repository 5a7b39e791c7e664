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
            if (MODE == 13) a[i] = __vmaxu2(a[i], a[(i + 3) & 7]);          // VIMNMX, two distinct varying registers
            if (MODE == 15) a[i] = __dp2a_lo(a[i], 0x0401u, b);             // IDP.2A
            if (MODE == 16) { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&a[(i + 3) & 7]);
                              x = __hmax2(x, y); a[i] = *reinterpret_cast<unsigned*>(&x); }  // HMNMX2 two varying regs
            if (MODE == 17) { if (i & 1) a[i] = __vmaxu2(a[i], a[(i + 3) & 7]);
                              else { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&a[(i + 3) & 7]);
                                     x = __hmin2(x, y); a[i] = *reinterpret_cast<unsigned*>(&x); } }  // VIMNMX / HMNMX2 alternating
            if (MODE == 18) a[i] = max((int)a[i], (int)a[(i + 3) & 7]);     // 32-bit IMNMX two varying regs
            if (MODE == 19) a[i] = __float_as_uint(fmaxf(__uint_as_float(a[i]), __uint_as_float(a[(i + 3) & 7])));  // FMNMX
            if (MODE == 20) { if (i & 1) a[i] = __vmaxu2(a[i], a[(i + 3) & 7]); else a[i] = a[i] * 5 + a[(i + 3) & 7]; }  // VIMNMX / IMAD varying
        }
        if (MODE == 14) {  // compare-exchange network (odd-even transposition) on 8 registers: min + max of the same pair
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int i = 0; i + 1 < UNROLL; i += 2) { unsigned lo = __vminu2(a[i], a[i + 1]), hi = __vmaxu2(a[i], a[i + 1]); a[i] = lo; a[i + 1] = hi; }
#pragma unroll
                for (int i = 1; i + 1 < UNROLL; i += 2) { unsigned lo = __vminu2(a[i], a[i + 1]), hi = __vmaxu2(a[i], a[i + 1]); a[i] = lo; a[i + 1] = hi; }
            }
            a[0] ^= b; a[7] += b;
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

__global__ void k_hcheck(const unsigned* a, const unsigned* b, unsigned* mx, unsigned* mn) {
    const int i = threadIdx.x;
    __half2 x = *reinterpret_cast<const __half2*>(&a[i]), y = *reinterpret_cast<const __half2*>(&b[i]);
    __half2 hi = __hmax2(x, y), lo = __hmin2(x, y);
    mx[i] = *reinterpret_cast<unsigned*>(&hi);
    mn[i] = *reinterpret_cast<unsigned*>(&lo);
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
    run<13>("VIMNMX two varying regs", d_out, d_cyc);
    run<14>("compare-exchange x56/iter (8/iter counted)", d_out, d_cyc);
    run<15>("IDP.2A", d_out, d_cyc);
    run<16>("HMNMX2 two varying regs", d_out, d_cyc);
    run<17>("VIMNMX / HMNMX2 alternating", d_out, d_cyc);
    run<18>("IMNMX 32-bit two varying", d_out, d_cyc);
    run<19>("FMNMX two varying", d_out, d_cyc);
    run<20>("VIMNMX / IMAD alternating", d_out, d_cyc);
    {   // does HMNMX2 order small (denormal-pattern) and large u16 codes like integers?
        unsigned h_a[8] = {0x00010000u, 0x00010002u, 0x03ff0400u, 0x64010001u, 0x00000001u, 0x7bff6401u, 0x00016401u, 0x00020001u};
        unsigned h_b[8] = {0x00000001u, 0x00020001u, 0x04000001u, 0x00016401u, 0x00010000u, 0x64017bffu, 0x64010001u, 0x00010002u};
        unsigned *da, *db, *dmx, *dmn, hmx[8], hmn[8];
        cudaMalloc(&da, 32); cudaMalloc(&db, 32); cudaMalloc(&dmx, 32); cudaMalloc(&dmn, 32);
        cudaMemcpy(da, h_a, 32, cudaMemcpyHostToDevice); cudaMemcpy(db, h_b, 32, cudaMemcpyHostToDevice);
        k_hcheck<<<1, 8>>>(da, db, dmx, dmn);
        cudaMemcpy(hmx, dmx, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hmn, dmn, 32, cudaMemcpyDeviceToHost);
        int ok = 1;
        for (int i = 0; i < 8; ++i) {
            unsigned alo = h_a[i] & 0xffff, ahi = h_a[i] >> 16, blo = h_b[i] & 0xffff, bhi = h_b[i] >> 16;
            unsigned emx = (alo > blo ? alo : blo) | ((ahi > bhi ? ahi : bhi) << 16), emn = (alo < blo ? alo : blo) | ((ahi < bhi ? ahi : bhi) << 16);
            if (emx != hmx[i] || emn != hmn[i]) { ok = 0; printf("HMNMX2 mismatch %08x %08x -> max %08x (want %08x) min %08x (want %08x)\n", h_a[i], h_b[i], hmx[i], emx, hmn[i], emn); }
        }
        printf("HMNMX2 orders u16 codes below 0x7c00 like integers (incl. denormal patterns): %s\n", ok ? "yes" : "NO");
    }
    run_lds<4>("LDS.32 stream", d_out, d_cyc);
    run_lds<8>("LDS.64 stream", d_out, d_cyc);
    run_lds<16>("LDS.128 stream", d_out, d_cyc);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
