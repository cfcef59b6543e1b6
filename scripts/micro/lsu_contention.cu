// Microbenchmark behind the owner-warp backward (csrc/msda_bwd_owner.cu): how long does ONE warp's dependent
// read-modify-write chain take while the other 15 warps of the SM saturate the LSU with red.global.add.v4.f32 (the state
// the MSDA backward is in), for accumulators that live in
//   (a) shared memory  : LDS.128 -> FFMA -> STS.128            (goes through the same LSU / MIO queue as the reds)
//   (b) tensor memory  : tcgen05.ld.32x32b -> FFMA -> tcgen05.st (TMEM has its own datapath)
//   (c) registers      : FFMA only (lower bound)
// First measurement of the owner-warp backward: 8 700 clk per warp tile instead of the ~500 estimated from the unloaded
// LDS latency (29 clk) -- the suspicion is that every LDS of the owner queues behind the workers' reds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_contention lsu_contention.cu && ./lsu_contention
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int THREADS = 512, WARPS = 16;
constexpr int ACC_ROWS = 320;

// MODE: 0 = shared memory chain, 1 = TMEM chain (x1: one row per access), 2 = register chain, 3 = TMEM x4 (4 adjacent
// columns per access), 4 = shared memory, 4 independent rows in flight per step
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(float *gbuf, int rows, int red_iters, int chain_iters, int producers_on,
                                                long long *owner_cycles, long long *total_cycles, float *sink) {
    extern __shared__ __align__(16) unsigned char raw[];
    float *acc = reinterpret_cast<float *>(raw);   // [ACC_ROWS][32]
    __shared__ unsigned s_tmem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7;
    for (int i = threadIdx.x; i < ACC_ROWS * 32; i += THREADS) acc[i] = 0.0f;
    if (MODE == 1 || MODE == 3 || MODE == 7) {
        if (warp == WARPS - 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    __syncthreads();
    if (MODE == 1 || MODE == 3 || MODE == 7) asm volatile("tcgen05.fence::after_thread_sync;");
    const long long t0 = clock64();
    if (warp < WARPS - 1) {
        if (producers_on) {
            unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
            for (int it = 0; it < red_iters; ++it) {
                seed = mix32(seed + it * 7919u);
                const int row = (int)(((unsigned long long)seed * (unsigned)rows) >> 32);
                float *dst = gbuf + (size_t)row * 32;
                const float v = 1.0f + j;
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j * 4), "f"(v), "f"(v),
                             "f"(v), "f"(v) : "memory");
            }
        }
    } else {
        // the owner: chain_iters dependent read-modify-write steps on pseudo-random rows
        unsigned seed = blockIdx.x * 977u + 13u;
        float wsum = 0.0f;
        const long long c0 = clock64();
        if (MODE == 0) {
            const int c = lane >> 3;
            for (int it = 0; it < chain_iters; ++it) {
                seed = mix32(seed + it);
                const unsigned row = (seed % (ACC_ROWS / 4)) * 4 + c;      // four distinct rows per step
                float4 *p = reinterpret_cast<float4 *>(acc + row * 32 + j * 4);
                float4 v = *p;
                v.x = fmaf(0.5f, 1.0f, v.x); v.y = fmaf(0.5f, 2.0f, v.y); v.z = fmaf(0.5f, 3.0f, v.z); v.w = fmaf(0.5f, 4.0f, v.w);
                *p = v;
                __syncwarp();
            }
        } else if (MODE == 4) {
            const int c = lane >> 3;
            for (int it = 0; it < chain_iters; it += 4) {
                float4 *p[4];
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    seed = mix32(seed + it + u);
                    const unsigned row = (u * (ACC_ROWS / 16) + seed % (ACC_ROWS / 16)) * 4 + c;   // disjoint row ranges
                    p[u] = reinterpret_cast<float4 *>(acc + row * 32 + j * 4);
                    v[u] = *p[u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u].x += 0.5f; v[u].y += 1.0f; v[u].z += 1.5f; v[u].w += 2.0f;
                    *p[u] = v[u];
                }
                __syncwarp();
            }
        } else if (MODE == 1) {
            const unsigned base = s_tmem;     // lane field = this warp's quarter (warp 15 -> lanes 96..127)
            const unsigned lane_base = base + ((unsigned)((warp & 3) * 32) << 16);
            // zero the columns first
            for (int col = 0; col < 512; ++col)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + col), "r"(0u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            for (int it = 0; it < chain_iters; ++it) {
                seed = mix32(seed + it);
                const unsigned col = seed % 512u;
                unsigned r;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(lane_base + col) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float f = fmaf(0.5f, 1.0f + lane, __uint_as_float(r));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + col), "r"(__float_as_uint(f)) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            unsigned r;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(lane_base + 7u) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            wsum += __uint_as_float(r);
        } else if (MODE == 3) {
            const unsigned base = s_tmem;
            const unsigned lane_base = base + ((unsigned)((warp & 3) * 32) << 16);
            for (int col = 0; col < 512; ++col)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + col), "r"(0u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            for (int it = 0; it < chain_iters; ++it) {
                seed = mix32(seed + it);
                const unsigned col = (seed % 127u) * 4u;
                unsigned r0, r1, r2, r3;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(lane_base + col) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float g = 1.0f + lane;
                r0 = __float_as_uint(fmaf(0.5f, g, __uint_as_float(r0)));
                r1 = __float_as_uint(fmaf(0.25f, g, __uint_as_float(r1)));
                r2 = __float_as_uint(fmaf(0.125f, g, __uint_as_float(r2)));
                r3 = __float_as_uint(fmaf(0.0625f, g, __uint_as_float(r3)));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                             ::"r"(lane_base + col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            unsigned r;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(lane_base + 8u) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            wsum += __uint_as_float(r);
        } else if (MODE == 5 || MODE == 6) {
            // broadcast chain: every step moves one 32-bit word from a (changing) source lane to all lanes
            unsigned v = seed + lane;
            for (int it = 0; it < chain_iters; ++it) {
                const int src = (it * 7 + (v & 3)) & 31;
                unsigned b;
                if (MODE == 5) b = __reduce_or_sync(0xffffffffu, lane == src ? v : 0u);
                else b = __shfl_sync(0xffffffffu, v, src);
                v = v * 1664525u + b;
            }
            wsum = __uint_as_float(v & 0x3fffffffu);
        } else if (MODE == 7) {
            // the owner's inner loop as planned: 5 words broadcast with REDUX (column + 4 weights), two x2 TMEM
            // read-modify-writes (columns c, c+1 and c+16, c+17), lane = channel
            const unsigned base = s_tmem;
            const unsigned lane_base = base + ((unsigned)((warp & 3) * 32) << 16);
            for (int col = 0; col < 512; ++col)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + col), "r"(0u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            unsigned rec[5];
#pragma unroll
            for (int e = 0; e < 5; ++e) rec[e] = mix32(seed + lane * 5 + e);
            const float g = 1.0f + lane;
            for (int it = 0; it < chain_iters; ++it) {
                const int src = it & 31;
                unsigned b[5];
#pragma unroll
                for (int e = 0; e < 5; ++e) b[e] = __reduce_or_sync(0xffffffffu, lane == src ? rec[e] : 0u);
                const unsigned col = b[0] % 400u;
                unsigned r0, r1, r2, r3;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(lane_base + col) : "memory");
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r2), "=r"(r3) : "r"(lane_base + col + 16u) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                r0 = __float_as_uint(fmaf(__uint_as_float((b[1] & 0x007fffffu) | 0x3e000000u), g, __uint_as_float(r0)));
                r1 = __float_as_uint(fmaf(__uint_as_float((b[2] & 0x007fffffu) | 0x3e000000u), g, __uint_as_float(r1)));
                r2 = __float_as_uint(fmaf(__uint_as_float((b[3] & 0x007fffffu) | 0x3e000000u), g, __uint_as_float(r2)));
                r3 = __float_as_uint(fmaf(__uint_as_float((b[4] & 0x007fffffu) | 0x3e000000u), g, __uint_as_float(r3)));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(lane_base + col), "r"(r0), "r"(r1) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(lane_base + col + 16u), "r"(r2), "r"(r3) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            unsigned r;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(lane_base + 9u) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            wsum += __uint_as_float(r);
        } else {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int it = 0; it < chain_iters; ++it) {
                seed = mix32(seed + it);
                const float w = __uint_as_float((seed & 0x007fffffu) | 0x3f000000u);
                a0 = fmaf(w, 1.0f, a0); a1 = fmaf(w, 2.0f, a1); a2 = fmaf(w, 3.0f, a2); a3 = fmaf(w, 4.0f, a3);
            }
            wsum = a0 + a1 + a2 + a3;
        }
        const long long c1 = clock64();
        if (lane == 0) owner_cycles[blockIdx.x] = c1 - c0;
        if (wsum == 123.456f) sink[0] = wsum;
    }
    __syncthreads();
    if (MODE == 1 || MODE == 3 || MODE == 7) {
        if (warp == WARPS - 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512));
    }
    if (threadIdx.x == 0) {
        total_cycles[blockIdx.x] = clock64() - t0;
        if (acc[5] == 123.456f) sink[1] = acc[5];
    }
}

template <int MODE>
static void run(const char *name, int producers_on, float *gbuf, int rows, long long *cyc, long long *tot, float *sink) {
    const int ctas = 148, red_iters = 4000, chain_iters = 2000;
    auto kern = k<MODE>;
    const int smem = ACC_ROWS * 128;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<ctas, THREADS, smem>>>(gbuf, rows, 100, 100, producers_on, cyc, tot, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<ctas, THREADS, smem>>>(gbuf, rows, red_iters, chain_iters, producers_on, cyc, tot, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148], t[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(t, tot, sizeof(t), cudaMemcpyDeviceToHost);
    double mean = 0, tmean = 0;
    for (int i = 0; i < ctas; ++i) { mean += (double)h[i] / ctas; tmean += (double)t[i] / ctas; }
    const double row_adds = (double)ctas * (WARPS - 1) * 4.0 * red_iters;
    printf("%-44s producers %s: owner %.1f clk per step; kernel %.3f ms (%.0f clk / CTA), reds %.2f TB/s  %s\n", name,
           producers_on ? "on " : "off", mean / chain_iters, ms, tmean,
           producers_on ? row_adds * 128 / ms * 1e-9 : 0.0, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const int rows = 174080;
    float *gbuf, *sink; long long *cyc, *tot;
    cudaMalloc(&gbuf, (size_t)rows * 128);
    cudaMemset(gbuf, 0, (size_t)rows * 128);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&tot, 148 * sizeof(long long));
    cudaMalloc(&sink, 16);
    for (int on = 0; on < 2; ++on) {
        run<0>("shared memory, 4 rows per step (LDS.128)", on, gbuf, rows, cyc, tot, sink);
        run<4>("shared memory, 4 steps in flight", on, gbuf, rows, cyc, tot, sink);
        run<1>("tensor memory, 1 row per step (32x32b.x1)", on, gbuf, rows, cyc, tot, sink);
        run<3>("tensor memory, 4 rows per step (32x32b.x4)", on, gbuf, rows, cyc, tot, sink);
        run<2>("registers (FFMA chain)", on, gbuf, rows, cyc, tot, sink);
        run<5>("broadcast chain, REDUX.OR", on, gbuf, rows, cyc, tot, sink);
        run<6>("broadcast chain, SHFL.IDX", on, gbuf, rows, cyc, tot, sink);
        run<7>("5 x REDUX + 2 x TMEM x2 read-modify-write", on, gbuf, rows, cyc, tot, sink);
    }
    return 0;
}
