// Microbenchmark: throughput of shared-memory accumulation primitives on sm_100a, in the access shape the MSDA
// backward would use (8 lanes x 16 B per 128-byte row, 4 random rows per warp instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// mode 0: fp32 atomicAdd (CAS loop) x4 per lane; 1: 64-bit CAS on float pairs x2; 2: int32 ATOMS.ADD x4;
// mode 3: plain LDS.128 + STS.128 (racy, upper bound); 4: global REDG.128 to an L2-resident buffer
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *gbuf, int rows, int iters, long long *cycles) {
    extern __shared__ __align__(16) float s[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, j = lane & 7, g = lane >> 3;
    unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        seed = mix32(seed + it * 7919u);
        const int row = (int)(((unsigned long long)seed * (unsigned)rows) >> 32);
        // rotate the element order by the group id so the 4 groups of a warp hit different banks
        float *p = s + row * 32 + ((j * 4 + g * 8) & 31);
        const float v = 1.0f + j;
        if (MODE == 0) {
            atomicAdd(p + 0, v); atomicAdd(p + 1, v); atomicAdd(p + 2, v); atomicAdd(p + 3, v);
        } else if (MODE == 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                unsigned long long *q = reinterpret_cast<unsigned long long *>(p + 2 * h);
                unsigned long long old = *q, assumed;
                do {
                    assumed = old;
                    float2 f = *reinterpret_cast<float2 *>(&assumed);
                    f.x += v; f.y += v;
                    old = atomicCAS(q, assumed, *reinterpret_cast<unsigned long long *>(&f));
                } while (old != assumed);
            }
        } else if (MODE == 2) {
            int *q = reinterpret_cast<int *>(p);
            const int iv = (int)(v * 1024.f);
            atomicAdd(q + 0, iv); atomicAdd(q + 1, iv); atomicAdd(q + 2, iv); atomicAdd(q + 3, iv);
        } else if (MODE == 3) {
            float4 *q = reinterpret_cast<float4 *>(p);
            float4 o = *q; o.x += v; o.y += v; o.z += v; o.w += v; *q = o;
        } else {
            float *q = gbuf + (size_t)(row + (blockIdx.x & 7) * rows) * 32 + j * 4;
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (s[threadIdx.x] == 12345.678f) gbuf[0] = 1.f;
}

template <int MODE> void run(const char *name, int rows, float *gbuf, long long *cyc) {
    const int iters = 2000, grid = 148, threads = 512;
    const size_t smem = (size_t)rows * 128;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<grid, threads, smem>>>(gbuf, rows, 10, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads, smem>>>(gbuf, rows, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // per SM: 16 warps x iters warp-level row-quads (4 rows of 128 B each)
    printf("%-28s rows=%5d  %8.1f cycles per warp-quad-op per SM (%.3f ms, %s)  => %.1f B/cycle/SM\n", name, rows,
           avg / (16.0 * iters), ms, cudaGetErrorString(err), 512.0 * 16 * iters / avg);
}

int main() {
    float *gbuf; long long *cyc;
    cudaMalloc(&gbuf, (size_t)8 * 1344 * 128 + 1024);
    cudaMemset(gbuf, 0, (size_t)8 * 1344 * 128 + 1024);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int rows : {64, 320, 1344}) {
        run<0>("fp32 atomicAdd (CAS loop) x4", rows, gbuf, cyc);
        run<1>("64-bit CAS on float2 x2", rows, gbuf, cyc);
        run<2>("int32 ATOMS.ADD x4", rows, gbuf, cyc);
        run<3>("LDS.128+STS.128 (racy)", rows, gbuf, cyc);
        run<4>("global REDG.128 (L2)", rows, gbuf, cyc);
    }
    return 0;
}
