// msda_peer.cu -- query-sharded MSDA over NVLink peer memory: the pixel-shard all-gather in front of the forward and the
// grad_img reduce-scatter behind the backward as ONE kernel each, reading the peers' buffers directly (P2P loads over
// NVLink / NVSwitch) instead of going through NCCL.
//
// Setting (the reference has no distributed code): the ranks of a group share images.  Every rank holds a
// pixel shard of `img` and a query shard of the sampling points; forward needs the full pyramid on every rank, backward
// produces a full-size PARTIAL grad_img on every rank whose sum has to end up pixel-sharded again
// (/root/reference/src/msda_triton/kernels.py:549-553 is the only coupling between output rows).
//
// All buffers the peers touch live in symmetric memory (the host side allocates them with torch symmetric memory and
// hands over the peers' device pointers).  Cross-rank ordering uses four flag rows per rank,
//      flags[slot][src]   slot in {AG_READY, AG_DONE, RS_READY, RS_DONE},   written ONLY by rank `src`,
// that carry monotonically increasing epochs, so nothing is ever reset.  The epochs are call counters kept ON THE DEVICE
// (counters[4] = all-gathers issued, counters[5] = reduce-scatters issued; every kernel reads them at its start and its
// last CTA bumps its own at the end), so a launch has no call-dependent argument and a whole training step -- all-gather,
// forward, backward, reduce-scatter -- can be captured into ONE CUDA graph and replayed:
//   all-gather e (PUSH): signal AG_READY = e ("I am in all-gather e: the kernels that read my pyramid in step e-1 have
//                   completed -- stream order -- so you may write into it"); wait RS_DONE >= r (peers finished reading my
//                   partial grad_img of the previous backward: the coming backward overwrites it); copy my shard into
//                   my own pyramid; wait for everybody's AG_READY, then WRITE my shard into rows
//                   [rank*chunk, (rank+1)*chunk) of every peer's pyramid (posted stores over NVLink; pulling the same
//                   bytes with loads ran at 270-330 GB/s per GPU); signal AG_DONE = e ("my writes are complete"); wait for
//                   everybody's AG_DONE: my pyramid is complete.
//   reduce-scatter: signal RS_READY = r (stream order: my backward has completed) -> wait for everybody's RS_READY
//                   -> out[b, i] = sum over ranks (fixed order 0..world-1: deterministic given the partials) of
//                   partial_rank[b, my_rank * chunk + i] -> signal RS_DONE = r
// The kernels are persistent-sized (one CTA per SM): CTAs spin on the flags, so the whole grid has to be resident.
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_b200.h"

namespace {

constexpr int kThreads = 512;
enum { AG_READY = 0, AG_DONE = 1, RS_READY = 2, RS_DONE = 3 };
constexpr int kMaxWorld = 16;

struct PeerArgs {
    uint4 *pyramids[kMaxWorld];           // every rank's gathered pyramid [B, world * chunk, H, D]
    const float4 *partials[kMaxWorld];    // every rank's partial grad_img [B, world * chunk, H, D] fp32
    uint32_t *flags[kMaxWorld];           // every rank's flag block [4][world]
    uint32_t *counters;                   // local: [0..2] CTA arrival counters (zero between kernels), [4] / [5] epochs
    int world, rank;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Weak (ordinary) load without L1 allocation: the line belongs to another GPU and is read once.  Ordering against the
// producer comes from the acquire on its flag (thread 0) + __syncthreads(); system-scope RELAXED loads, the first
// version, ran the pulls at ~360 GB/s.
__device__ __forceinline__ uint4 ld_remote(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// thread 0 of the CTA: wait until every rank's entry of `slot` in MY flag block has reached `epoch`
__device__ void wait_all(const PeerArgs &a, int slot, uint32_t epoch) {
    if (threadIdx.x == 0) {
        const uint32_t *mine = a.flags[a.rank] + slot * a.world;
        for (int s = 0; s < a.world; ++s)
            while ((int32_t)(ld_acquire_sys(mine + s) - epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
}
// every CTA calls this after its part of the preceding phase; the last one to arrive publishes `epoch` in `slot` of
// every rank's flag block (entry [slot][my rank]) and, if asked to, records the epoch as this rank's call count
__device__ void arrive_and_signal(const PeerArgs &a, int counter, int slot, uint32_t epoch, int bump = -1) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(a.counters + counter, 1u) == gridDim.x - 1) {
            a.counters[counter] = 0u;
            if (bump >= 0) a.counters[bump] = epoch;    // every CTA of this launch has read the old value long ago
            __threadfence_system();
            for (int p = 0; p < a.world; ++p) st_release_sys(a.flags[p] + slot * a.world + a.rank, epoch);
        }
    }
}
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_remote(uint4 *p, const uint4 v) {   // posted store into a peer's memory
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// n16: uint4 per image and rank (chunk * H * D * elem_size / 16)
__global__ void __launch_bounds__(kThreads) peer_all_gather_kernel(const PeerArgs a, const uint4 *__restrict__ shard,
                                                                   const long long B, const long long n16) {
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x, stride = (long long)gridDim.x * kThreads;
    const uint32_t epoch_ag = ld_volatile(a.counters + 4) + 1u;      // this is all-gather number ...
    const uint32_t epoch_rs_done = ld_volatile(a.counters + 5);      // ... after that many reduce-scatters
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // stream order: whatever read my pyramid in the previous step is done
        __threadfence_system();
        for (int p = 0; p < a.world; ++p) st_release_sys(a.flags[p] + AG_READY * a.world + a.rank, epoch_ag);
    }
    wait_all(a, RS_DONE, epoch_rs_done);
    const long long per_rank = B * n16;
    uint4 *__restrict__ own = a.pyramids[a.rank];
    for (long long i = tid; i < per_rank; i += stride) {        // my own part of my own pyramid
        const long long b = i / n16, k = i - b * n16;
        own[(b * a.world + a.rank) * n16 + k] = __ldg(shard + i);
    }
    // my shard into every peer's pyramid: one flat index space over (peer, element) so that all world-1 links carry
    // traffic at once and a thread has several independent stores in flight (a shard is only ~5 elements per thread;
    // one phase per peer cost 7 flag waits and 7 short bursts: 0.152 ms at 8 GPUs); the peer order is rotated by the rank
    wait_all(a, AG_READY, epoch_ag);
    const long long remote = per_rank * (a.world - 1);
    for (long long i = tid; i < remote; i += stride) {
        const int d = (int)(i / per_rank);
        const long long e = i - (long long)d * per_rank;
        const int p = (a.rank + 1 + d) % a.world;
        const long long b = e / n16, k = e - b * n16;
        st_remote(a.pyramids[p] + (b * a.world + a.rank) * n16 + k, __ldg(shard + e));
    }
    arrive_and_signal(a, 1, AG_DONE, epoch_ag, 4);    // fence + "my writes have landed"
    wait_all(a, AG_DONE, epoch_ag);                   // everybody's writes into MY pyramid have landed
}

// n4: float4 per image and rank
__global__ void __launch_bounds__(kThreads) peer_reduce_scatter_kernel(const PeerArgs a, float4 *__restrict__ out,
                                                                       const long long B, const long long n4) {
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x, stride = (long long)gridDim.x * kThreads;
    const uint32_t epoch_rs = ld_volatile(a.counters + 5) + 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {    // stream order: the backward that filled my partial has completed
        __threadfence_system();
        for (int p = 0; p < a.world; ++p) st_release_sys(a.flags[p] + RS_READY * a.world + a.rank, epoch_rs);
    }
    wait_all(a, RS_READY, epoch_rs);
    for (long long i = tid; i < B * n4; i += stride) {
        const long long b = i / n4, k = i - b * n4;
        const long long off = (b * a.world + a.rank) * n4 + k;   // my pixel chunk inside every rank's partial image
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 v[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < a.world) v[r] = ld_remote(reinterpret_cast<const uint4 *>(a.partials[r] + off));
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < a.world) {
                acc.x += __uint_as_float(v[r].x);
                acc.y += __uint_as_float(v[r].y);
                acc.z += __uint_as_float(v[r].z);
                acc.w += __uint_as_float(v[r].w);
            }
        out[i] = acc;
    }
    arrive_and_signal(a, 2, RS_DONE, epoch_rs, 5);
}

int fill(PeerArgs &k, const msda_peer_ctx *ctx) {
    if (!ctx || ctx->world < 1 || ctx->world > kMaxWorld || ctx->rank < 0 || ctx->rank >= ctx->world) return -1;
    if (!ctx->peer_pyramids || !ctx->peer_partials || !ctx->peer_flags || !ctx->counters) return -1;
    k.world = ctx->world;
    k.rank = ctx->rank;
    k.counters = ctx->counters;
    for (int r = 0; r < ctx->world; ++r) {
        k.pyramids[r] = static_cast<uint4 *>(const_cast<void *>(ctx->peer_pyramids[r]));
        k.partials[r] = static_cast<const float4 *>(ctx->peer_partials[r]);
        k.flags[r] = ctx->peer_flags[r];
        if (!k.pyramids[r] || !k.partials[r] || !k.flags[r]) return -1;
    }
    return 0;
}

int sm_count_of_current_device() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

}  // namespace

extern "C" {

int msda_peer_all_gather(const void *shard, const msda_peer_ctx *ctx, int64_t B, int64_t shard_bytes_per_image, void *stream) {
    PeerArgs k;
    if (fill(k, ctx) != 0 || !shard || B < 0 || shard_bytes_per_image < 0 || shard_bytes_per_image % 16 != 0)
        return MSDA_ERR_BAD_SHAPE;
    if ((uintptr_t)shard & 15u) return MSDA_ERR_BAD_SHAPE;
    const int sms = sm_count_of_current_device();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    peer_all_gather_kernel<<<sms, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        k, static_cast<const uint4 *>(shard), (long long)B, (long long)(shard_bytes_per_image / 16));
    return (int)cudaGetLastError();
}

int msda_peer_reduce_scatter(void *grad_shard, const msda_peer_ctx *ctx, int64_t B, int64_t shard_floats_per_image,
                             void *stream) {
    PeerArgs k;
    if (fill(k, ctx) != 0 || !grad_shard || B < 0 || shard_floats_per_image < 0 || shard_floats_per_image % 4 != 0)
        return MSDA_ERR_BAD_SHAPE;
    if ((uintptr_t)grad_shard & 15u) return MSDA_ERR_BAD_SHAPE;
    const int sms = sm_count_of_current_device();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    peer_reduce_scatter_kernel<<<sms, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        k, static_cast<float4 *>(grad_shard), (long long)B, (long long)(shard_floats_per_image / 4));
    return (int)cudaGetLastError();
}

}  // extern "C"
