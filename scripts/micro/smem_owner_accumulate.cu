// Microbenchmark behind DESIGN.md section 10 item 1: can a few "owner" warps per CTA absorb the row adds of the coarse
// pyramid levels in shared memory (single writer: plain LDS + FFMA + STS, no atomics) while the other warps keep
// issuing red.global.add.v4.f32 for the fine levels -- and does the chip-wide row-add rate go up accordingly?
//
// 148 CTAs x 512 threads (16 warps).  Every warp iteration stands for four 128-byte row adds (one per 8-lane group), as
// in the MSDA backward.  MODE 0: all 16 warps issue red.v4 into a 22 MB buffer (today's backward).  MODE 1/2: a fraction
// `coarse` of the row adds (0.25 = one pyramid level of four, 0.5 = two) is diverted: the producing group writes an
// 8-byte record {weight, row | unit << 16} into its warp's ring in shared memory, and OWNERS owner warps (1 or 2) drain
// the rings: per record one LDS.64 (record), one LDS.32 per lane of the unit's grad_out row (parked in shared memory),
// one LDS.32 + FFMA + STS.32 per lane on the accumulator row (32 lanes x 4 B = one 128-byte row).  Rings are
// single-producer / single-consumer with head / tail counters in shared memory (volatile, __syncwarp + fence).
// The benchmark reports row adds per second for the whole chip; bytes = rows x 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_owner_accumulate smem_owner_accumulate.cu && ./smem_owner_accumulate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

constexpr int THREADS = 512, WARPS = THREADS / 32;
constexpr int RING = 128;          // records per producer ring
constexpr int ACC_ROWS = 320;      // 8x8 + 16x16 level of one (b,h) slice
constexpr int UNITS = 128;         // grad_out rows parked in shared memory

struct Shared {
    float acc[ACC_ROWS][32];                 // 40 KB accumulator, owned by the owner warps
    float go[UNITS][32];                     // 16 KB parked grad_out rows
    volatile unsigned long long ring[WARPS][RING];    // 16 KB of records
    volatile unsigned head[WARPS];           // written by the producer (lane 0)
    volatile unsigned tail[WARPS];           // written by the consumer
    volatile unsigned done;                  // number of producer warps that have finished
};

template <int OWNERS, int BATCH>
__global__ void __launch_bounds__(THREADS, 1) k(float *gbuf, int rows, int iters, unsigned coarse_per_256,
                                                long long *cycles, unsigned long long *absorbed) {
    extern __shared__ __align__(16) unsigned char raw[];
    Shared &s = *reinterpret_cast<Shared *>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7;
    constexpr int OW = OWNERS > 0 ? OWNERS : 1;
    for (int i = threadIdx.x; i < ACC_ROWS * 32; i += THREADS) (&s.acc[0][0])[i] = 0.0f;
    for (int i = threadIdx.x; i < UNITS * 32; i += THREADS) (&s.go[0][0])[i] = 1.0f + (i & 31);
    if (threadIdx.x < WARPS) { s.head[threadIdx.x] = 0; s.tail[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s.done = 0;
    __syncthreads();
    const int producers = WARPS - OWNERS;
    const long long t0 = clock64();

    if (warp < producers) {
        unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
        unsigned head = 0;                                    // this warp's ring head (records pushed so far)
        for (int it = 0; it < iters; ++it) {
            seed = mix32(seed + it * 7919u);
            const bool coarse = OWNERS > 0 && (mix32(seed ^ 0x9e3779b9u) & 255u) < coarse_per_256;
            // how many of the four groups divert their row add this iteration
            const unsigned vote = __ballot_sync(0xffffffffu, coarse && j == 0);
            const int n_push = __popc(vote);
            if (n_push) {
                // wait for room in the ring (bounded by the consumer's progress)
                if (lane == 0) {
                    int guard = 0;   // bounded: a benchmark must never hang the box
                    while (head + n_push - s.tail[warp] > RING && ++guard < (1 << 22)) __nanosleep(64);
                }
                __syncwarp();
                if (coarse && j == 0) {
                    const int slot = __popc(vote & ((1u << lane) - 1u));
                    const unsigned row = mix32(seed * 31u) % ACC_ROWS, unit = mix32(seed * 17u) % UNITS;
                    const unsigned long long rec = ((unsigned long long)__float_as_uint(0.5f) << 32) | row | (unit << 16);
                    s.ring[warp][(head + slot) % RING] = rec;
                }
                head += n_push;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) s.head[warp] = head;
            }
            if (!coarse) {
                const int row = (int)(((unsigned long long)seed * (unsigned)rows) >> 32);
                float *dst = gbuf + (size_t)row * 32;
                const float v = 1.0f + j;
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j * 4), "f"(v), "f"(v),
                             "f"(v), "f"(v) : "memory");
            }
        }
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<unsigned *>(&s.done), 1u);
    } else {
        // owner warp `o` drains the rings of producers o, o + OWNERS, ... into ITS rows of the accumulator (rows with
        // row % OWNERS == o; a record's row is folded into that class -- in a real kernel the producers would pick the
        // ring by row class so that every row has exactly one writer)
        const int o = warp - producers;
        unsigned long long mine = 0;
        unsigned tails[WARPS];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) tails[w] = 0;
        for (long long spins = 0; spins < (1ll << 26); ++spins) {   // bounded, see above
            bool any = false;
            for (int w = o; w < producers; w += OW) {
                const unsigned h = s.head[w];
                unsigned t = tails[w];
                if (t == h) continue;
                any = true;
                // BATCH records at a time: their loads are independent, so the LDS -> FFMA -> STS latency is paid once per
                // batch; two records of a batch that hit the same row would lose an update, so such a batch takes the
                // one-at-a-time path (1.9 % of the batches for 4 random rows out of 320)
                for (; h - t >= (unsigned)BATCH; t += BATCH) {
                    unsigned r[BATCH], unit[BATCH];
                    float wgt[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; ++b) {
                        const unsigned long long rec = s.ring[w][(t + b) % RING];
                        wgt[b] = __uint_as_float((unsigned)(rec >> 32));
                        const unsigned row = ((unsigned)rec & 0xFFFFu) / OW * OW + o;   // this owner's row class
                        unit[b] = ((unsigned)rec >> 16) & 0xFFFFu;
                        r[b] = row < ACC_ROWS ? row : o;
                    }
                    bool clash = false;
#pragma unroll
                    for (int b = 1; b < BATCH; ++b)
#pragma unroll
                        for (int c = 0; c < b; ++c) clash |= r[b] == r[c];
                    if (!clash) {
                        float a[BATCH], gq[BATCH];
#pragma unroll
                        for (int b = 0; b < BATCH; ++b) { a[b] = s.acc[r[b]][lane]; gq[b] = s.go[unit[b]][lane]; }
#pragma unroll
                        for (int b = 0; b < BATCH; ++b) s.acc[r[b]][lane] = fmaf(wgt[b], gq[b], a[b]);
                    } else {
#pragma unroll
                        for (int b = 0; b < BATCH; ++b) s.acc[r[b]][lane] = fmaf(wgt[b], s.go[unit[b]][lane], s.acc[r[b]][lane]);
                    }
                    mine += BATCH;
                }
                for (; t != h; ++t) {
                    const unsigned long long rec = s.ring[w][t % RING];
                    const float wgt = __uint_as_float((unsigned)(rec >> 32));
                    const unsigned row = ((unsigned)rec & 0xFFFFu) / OW * OW + o;   // this owner's row class
                    const unsigned unit = ((unsigned)rec >> 16) & 0xFFFFu;
                    const unsigned r = row < ACC_ROWS ? row : o;
                    s.acc[r][lane] = fmaf(wgt, s.go[unit][lane], s.acc[r][lane]);
                    ++mine;
                }
                tails[w] = t;
                __syncwarp();
                if (lane == 0) s.tail[w] = t;
            }
            if (!any && s.done == (unsigned)producers) {
                bool empty = true;
                for (int w = o; w < producers; w += OW) empty &= (tails[w] == s.head[w]);
                if (empty) break;
            }
        }
        if (lane == 0) atomicAdd(absorbed, mine);
        // flush: one red per accumulator row of this owner
        for (int r = o; r < ACC_ROWS; r += OW)
            asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(gbuf + (size_t)r * 32 + lane), "f"(s.acc[r][lane]) : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------------
// Variant 2: row classes.  A record goes to the ring of its row class (row mod 4) of its producer warp; group g of an
// owner warp (8 lanes x 16 bytes = one 128-byte row) is the single writer of class g and takes one record per step, so
// an owner warp retires up to four records per step with a quarter of the instructions per record.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RINGC = 32;   // records per (producer warp, class) ring

struct SharedC {
    float acc[ACC_ROWS][32];
    float go[UNITS][32];
    volatile unsigned long long ring[WARPS][4][RINGC];   // 16 KB
    volatile unsigned head[WARPS][4];
    volatile unsigned tail[WARPS][4];
    volatile unsigned done;
};

template <int OWNERS>
__global__ void __launch_bounds__(THREADS, 1) kc(float *gbuf, int rows, int iters, unsigned coarse_per_256,
                                                 unsigned long long *absorbed) {
    extern __shared__ __align__(16) unsigned char raw[];
    SharedC &s = *reinterpret_cast<SharedC *>(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7, g = lane >> 3;
    constexpr int OW = OWNERS > 0 ? OWNERS : 1;
    for (int i = threadIdx.x; i < ACC_ROWS * 32; i += THREADS) (&s.acc[0][0])[i] = 0.0f;
    for (int i = threadIdx.x; i < UNITS * 32; i += THREADS) (&s.go[0][0])[i] = 1.0f + (i & 31);
    if (threadIdx.x < WARPS * 4) { (&s.head[0][0])[threadIdx.x] = 0; (&s.tail[0][0])[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s.done = 0;
    __syncthreads();
    const int producers = WARPS - OWNERS;
    if (warp < producers) {
        unsigned seed = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
        unsigned heads[4] = {0, 0, 0, 0};
        for (int it = 0; it < iters; ++it) {
            seed = mix32(seed + it * 7919u);
            const bool coarse = (mix32(seed ^ 0x9e3779b9u) & 255u) < coarse_per_256;
            const unsigned row = mix32(seed * 31u) % ACC_ROWS, unit = mix32(seed * 17u) % UNITS;
            const int cls = row & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const unsigned vote = __ballot_sync(0xffffffffu, coarse && j == 0 && cls == c);
                const int n_push = __popc(vote);
                if (n_push == 0) continue;                      // warp-uniform
                if (lane == 0) {
                    int guard = 0;
                    while (heads[c] + n_push - s.tail[warp][c] > RINGC && ++guard < (1 << 22)) __nanosleep(32);
                }
                __syncwarp();
                if (coarse && j == 0 && cls == c) {
                    const int slot = __popc(vote & ((1u << lane) - 1u));
                    s.ring[warp][c][(heads[c] + slot) % RINGC] =
                        ((unsigned long long)__float_as_uint(0.5f) << 32) | row | (unit << 16);
                }
                heads[c] += n_push;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) s.head[warp][c] = heads[c];
            }
            if (!coarse) {
                const int grow = (int)(((unsigned long long)seed * (unsigned)rows) >> 32);
                float *dst = gbuf + (size_t)grow * 32;
                const float v = 1.0f + j;
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j * 4), "f"(v), "f"(v),
                             "f"(v), "f"(v) : "memory");
            }
        }
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<unsigned *>(&s.done), 1u);
    } else {
        const int o = warp - producers;
        unsigned long long mine = 0;
        for (long long spins = 0; spins < (1ll << 26); ++spins) {
            bool any = false;
            for (int w = o; w < producers; w += OW) {
                // group g drains ring (w, g): every group at its own pace, one record per step
                const unsigned h = s.head[w][g];
                unsigned t = s.tail[w][g];
                const unsigned todo = h - t;
                unsigned steps = todo;
                for (int m = 8; m < 32; m <<= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, m));
                if (steps == 0) continue;
                any = true;
                for (unsigned k = 0; k < steps; ++k) {
                    if (k < todo) {
                        const unsigned long long rec = s.ring[w][g][(t + k) % RINGC];
                        const float wgt = __uint_as_float((unsigned)(rec >> 32));
                        const unsigned row = (unsigned)rec & 0xFFFFu, unit = ((unsigned)rec >> 16) & 0xFFFFu;
                        float4 a = *reinterpret_cast<float4 *>(&s.acc[row][j * 4]);
                        const float4 q = *reinterpret_cast<const float4 *>(&s.go[unit][j * 4]);
                        a.x = fmaf(wgt, q.x, a.x); a.y = fmaf(wgt, q.y, a.y); a.z = fmaf(wgt, q.z, a.z); a.w = fmaf(wgt, q.w, a.w);
                        *reinterpret_cast<float4 *>(&s.acc[row][j * 4]) = a;
                        if (j == 0) ++mine;
                    }
                }
                __syncwarp();
                if (j == 0 && todo) s.tail[w][g] = h;
            }
            if (!any && s.done == (unsigned)producers) {
                bool empty = true;
                for (int w = o; w < producers; w += OW)
                    for (int c = 0; c < 4; ++c) empty &= (s.tail[w][c] == s.head[w][c]);
                if (empty) break;
            }
        }
        for (int m = 8; m < 32; m <<= 1) mine += __shfl_xor_sync(0xffffffffu, mine, m);
        if (lane == 0) atomicAdd(absorbed, mine);
        for (int r = o; r < ACC_ROWS; r += OW)
            asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(gbuf + (size_t)r * 32 + lane), "f"(s.acc[r][lane]) : "memory");
    }
}

template <int OWNERS>
static void run_classes(const char *name, float coarse, int rows, float *gbuf, unsigned long long *absorbed) {
    const int iters = 4000, ctas = 148;
    auto kern = kc<OWNERS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SharedC));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned per256 = (unsigned)(coarse * 256.0f + 0.5f);
    kern<<<ctas, THREADS, sizeof(SharedC)>>>(gbuf, rows, 200, per256, absorbed);
    cudaMemset(absorbed, 0, sizeof(unsigned long long));
    cudaEventRecord(e0);
    kern<<<ctas, THREADS, sizeof(SharedC)>>>(gbuf, rows, iters, per256, absorbed);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long ab = 0;
    cudaMemcpy(&ab, absorbed, sizeof(ab), cudaMemcpyDeviceToHost);
    const double row_adds = (double)ctas * (WARPS - OWNERS) * 4.0 * iters;
    printf("%-44s coarse %.2f  owners %d (row classes): %.3f ms, %.2f G row adds/s = %.2f TB/s of 128-byte rows (%.1f %% absorbed)  %s\n",
           name, coarse, OWNERS, ms, row_adds / ms * 1e-6, row_adds * 128 / ms * 1e-9, 100.0 * ab / row_adds,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

template <int OWNERS, int BATCH = 1>
static void run(const char *name, float coarse, int rows, float *gbuf, long long *cyc, unsigned long long *absorbed) {
    const int iters = 4000, ctas = 148;
    auto kern = k<OWNERS, BATCH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
    cudaMemset(absorbed, 0, sizeof(unsigned long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned per256 = (unsigned)(coarse * 256.0f + 0.5f);
    kern<<<ctas, THREADS, sizeof(Shared)>>>(gbuf, rows, 200, per256, cyc, absorbed);   // warm-up
    cudaMemset(absorbed, 0, sizeof(unsigned long long));
    cudaEventRecord(e0);
    kern<<<ctas, THREADS, sizeof(Shared)>>>(gbuf, rows, iters, per256, cyc, absorbed);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long ab = 0;
    cudaMemcpy(&ab, absorbed, sizeof(ab), cudaMemcpyDeviceToHost);
    const double row_adds = (double)ctas * (WARPS - OWNERS) * 4.0 * iters;   // every producer iteration = 4 row adds
    printf("%-44s coarse %.2f  owners %d batch %d: %.3f ms, %.2f G row adds/s = %.2f TB/s of 128-byte rows (%.1f %% absorbed in shared memory)  %s\n",
           name, coarse, OWNERS, BATCH, ms, row_adds / ms * 1e-6, row_adds * 128 / ms * 1e-9, 100.0 * ab / row_adds,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const int rows = 174080;   // the benchmark grad_img: 4 x 5440 x 8 rows of 128 B = 22 MB
    float *gbuf; long long *cyc; unsigned long long *absorbed;
    cudaMalloc(&gbuf, (size_t)rows * 128);
    cudaMemset(gbuf, 0, (size_t)rows * 128);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&absorbed, sizeof(unsigned long long));
    run<0>("all warps red.global.add.v4.f32", 0.0f, rows, gbuf, cyc, absorbed);
    run<1>("15 producers + 1 owner warp", 0.25f, rows, gbuf, cyc, absorbed);
    run<2>("14 producers + 2 owner warps", 0.25f, rows, gbuf, cyc, absorbed);
    run<2>("14 producers + 2 owner warps", 0.50f, rows, gbuf, cyc, absorbed);
    run<4>("12 producers + 4 owner warps", 0.50f, rows, gbuf, cyc, absorbed);
    run<1, 4>("15 producers + 1 owner warp", 0.25f, rows, gbuf, cyc, absorbed);
    run<2, 4>("14 producers + 2 owner warps", 0.25f, rows, gbuf, cyc, absorbed);
    run<1, 8>("15 producers + 1 owner warp", 0.25f, rows, gbuf, cyc, absorbed);
    run<2, 8>("14 producers + 2 owner warps", 0.50f, rows, gbuf, cyc, absorbed);
    run<4, 8>("12 producers + 4 owner warps", 0.50f, rows, gbuf, cyc, absorbed);
    run_classes<1>("15 producers + 1 owner warp", 0.25f, rows, gbuf, absorbed);
    run_classes<2>("14 producers + 2 owner warps", 0.25f, rows, gbuf, absorbed);
    run_classes<2>("14 producers + 2 owner warps", 0.50f, rows, gbuf, absorbed);
    run_classes<4>("12 producers + 4 owner warps", 0.50f, rows, gbuf, absorbed);
    return 0;
}
