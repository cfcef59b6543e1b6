// Microbenchmark: what does a red.global.add.v4.f32 warp instruction cost as a function of WHAT it touches?
// The MSDA backward issues one such instruction per (point, corner) for the four units of a warp tile: 4 lane groups x 8
// lanes x 16 B = four different 128-byte rows.  Variants (same instruction count unless noted):
//   0  four random rows per instruction (the backward's shape)
//   1  same, two of the four lane groups predicated off (what pair aggregation produces)
//   2  same, three of the four groups off
//   3  the four groups target four ADJACENT rows (one contiguous 512-byte span; a head-major layout would give pairs)
//   4  two adjacent pairs (2 x 256 B): x-neighbour corners contiguous
//   5  all 32 lanes in ONE row with scalar red.f32 (128 B per instruction, 4x the instructions for the same bytes)
//   6  red.v2.f32, 16 lanes per row (256 B per instruction)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_shapes red_shapes.cu && ./red_shapes
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *gbuf, unsigned row_shift, int iters) {
    const int lane = threadIdx.x & 31, j = lane & 7, g = lane >> 3;
    const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int it = 0; it < iters; ++it) {
        const unsigned base = (warp_global * 977u + (unsigned)it) * 2654435761u;
        if (MODE <= 2) {
            const unsigned row = ((base + (unsigned)g * 0x9E3779B9u) * 2246822519u) >> row_shift;
            float *dst = gbuf + (size_t)row * 32 + j * 4;
            const bool on = MODE == 0 || (MODE == 1 && (g & 1) == 0) || (MODE == 2 && g == 0);
            if (on) asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(dst), "f"(1.0f) : "memory");
        } else if (MODE == 3) {
            const unsigned row = ((base >> row_shift) & ~3u) + g;
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(gbuf + (size_t)row * 32 + j * 4), "f"(1.0f) : "memory");
        } else if (MODE == 4) {
            const unsigned row = ((((base + (unsigned)(g >> 1) * 0x9E3779B9u) * 2246822519u) >> row_shift) & ~1u) + (g & 1);
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(gbuf + (size_t)row * 32 + j * 4), "f"(1.0f) : "memory");
        } else if (MODE == 5) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned row = ((base + (unsigned)r * 0x9E3779B9u) * 2246822519u) >> row_shift;
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(gbuf + (size_t)row * 32 + lane), "f"(1.0f) : "memory");
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const unsigned row = ((base + (unsigned)(2 * r + (lane >> 4)) * 0x9E3779B9u) * 2246822519u) >> row_shift;
                asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1,%1};" ::"l"(gbuf + (size_t)row * 32 + (lane & 15) * 2), "f"(1.0f) : "memory");
            }
        }
    }
}

template <int MODE> void run(const char *name, float *gbuf, unsigned row_shift, double rows_per_iter) {
    const int iters = 4000, grid = 148, threads = 512;
    k<MODE><<<grid, threads>>>(gbuf, row_shift, 50);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(gbuf, row_shift, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 16, rows = warps * iters * rows_per_iter;
    printf("%-62s %7.3f ms  %6.2f TB/s of row adds  %5.1f clk per warp iteration per SM  %s\n", name, ms,
           rows * 128 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (16.0 * iters), cudaGetErrorString(err));
}

int main() {
    const unsigned rows = 1u << 18;   // 32 MiB, L2 resident
    float *gbuf;
    cudaMalloc(&gbuf, (size_t)rows * 128);
    cudaMemset(gbuf, 0, (size_t)rows * 128);
    const unsigned shift = 32 - 18;
    run<0>("0 four random rows per red.v4 instruction", gbuf, shift, 4);
    run<1>("1 two of the four lane groups predicated off", gbuf, shift, 2);
    run<2>("2 three of the four lane groups predicated off", gbuf, shift, 1);
    run<3>("3 four adjacent rows (512 B contiguous)", gbuf, shift, 4);
    run<4>("4 two adjacent pairs (2 x 256 B contiguous)", gbuf, shift, 4);
    run<5>("5 scalar red.f32, one row per instruction, 4 instructions", gbuf, shift, 4);
    run<6>("6 red.v2.f32, two rows per instruction, 2 instructions", gbuf, shift, 4);
    return 0;
}
