// Microbenchmark: can the TMA engine (cp.reduce.async.bulk ... add.f32, 128-byte rows staged in shared memory) add
// grad_img rows into L2 in parallel with the LSU path (REDG.E.ADD.F32x4)?  The MSDA backward is bound by the
// LSU->XBAR request port (1 sector/cycle/SM); if bulk reductions travel a different path the two could be combined.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce tma_reduce.cu && ./tma_reduce
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// MODE 0: LSU red.v4 only; 1: TMA bulk reduce only; 2: alternate (even iterations LSU, odd iterations TMA)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *gbuf, int rows, int iters, long long *cycles) {
    // per warp: ring of 4 stages x 4 rows x 128 B
    __shared__ __align__(128) float stage[16][4][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7, g = lane >> 3;
    unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        seed = mix32(seed + it * 7919u);
        const int row = (int)(((unsigned long long)seed * (unsigned)rows) >> 32);
        float *dst = gbuf + (size_t)row * 32;
        const float v = 1.0f + j;
        const bool use_tma = MODE == 1 || (MODE == 2 && (it & 1));
        if (!use_tma) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j * 4), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
        } else {
            const int s = (MODE == 2 ? (it >> 1) : it) & 3;
            if (s == 0 && it >= 4) {
                // the ring wraps: wait until the bulk reductions that read these slots have finished reading
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            float4 *slot = reinterpret_cast<float4 *>(&stage[warp][s][g][j * 4]);
            *slot = make_float4(v, v, v, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (j == 0) {
                const unsigned saddr = (unsigned)__cvta_generic_to_shared(&stage[warp][s][g][0]);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(dst), "r"(saddr) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int rows, float *gbuf, long long *cyc) {
    const int iters = 4000, grid = 148, threads = 512;
    k<MODE><<<grid, threads>>>(gbuf, rows, 16, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(gbuf, rows, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double bytes = 148.0 * 16 * iters * 512;
    printf("%-34s rows=%7d  %7.1f cycles per warp row-quad per SM  %6.1f B/cycle/SM  %7.2f TB/s (%.3f ms, %s)\n", name, rows,
           avg / (16.0 * iters), 512.0 * 16 * iters / avg, bytes / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(err));
}

int main() {
    const int max_rows = 174080;  // the benchmark grad_img: 4 x 5440 x 8 rows of 128 B = 22 MB
    float *gbuf; long long *cyc;
    cudaMalloc(&gbuf, (size_t)max_rows * 128);
    cudaMemset(gbuf, 0, (size_t)max_rows * 128);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int rows : {2048, 174080}) {
        run<0>("LSU red.global.add.v4.f32", rows, gbuf, cyc);
        run<1>("TMA cp.reduce.async.bulk add.f32", rows, gbuf, cyc);
        run<2>("alternating LSU / TMA", rows, gbuf, cyc);
    }
    return 0;
}
