// msda_pace.cu -- arrival counters for the wave pacing of the persistent tuned kernels (wave_pace() in msda_tiled.cuh).
//
// The C ABI gives the forward no workspace, so the counters live in a small ring of device words owned by the library:
// a launch that runs more than one wave takes the next slot and zeroes it on its stream right before the kernel.
// Launches in flight at the same time use different slots (unless more than kPaceSlots multi-wave launches overlap, in
// which case a CTA may pass a wave early or run into the bounded wait -- pacing is a performance hint only).
// CUDA-graph capture records the memset and the slot address, so a replayed graph reuses its slot.
#include <atomic>
#include <cstdlib>

#include "msda_launch.h"
#include "msda_tuning.h"

namespace msda {

constexpr unsigned kPaceSlots = 256;
constexpr int kMaxDevices = 64;

__device__ unsigned g_pace_ring[kPaceSlots];

static bool pacing_enabled() { return tuning().wave_pacing != 0; }   // measurement knob: 0 disables

bool pacing_forced() { return tuning().wave_pacing == 2; }   // test knob: 2 paces every multi-wave launch

cudaError_t acquire_pace_counter(cudaStream_t st, unsigned **slot) {
    static std::atomic<unsigned> ticket{0};
    static std::atomic<unsigned *> base[kMaxDevices];
    *slot = nullptr;
    if (!pacing_enabled()) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaSuccess;   // no pacing rather than an error
    unsigned *ring = base[dev].load(std::memory_order_acquire);
    if (ring == nullptr) {
        void *p = nullptr;
        e = cudaGetSymbolAddress(&p, g_pace_ring);
        if (e != cudaSuccess) {        // pacing is optional: run unpaced rather than fail the launch
            (void)cudaGetLastError();
            return cudaSuccess;
        }
        ring = static_cast<unsigned *>(p);
        base[dev].store(ring, std::memory_order_release);
    }
    unsigned *mine = ring + ticket.fetch_add(1, std::memory_order_relaxed) % kPaceSlots;
    e = cudaMemsetAsync(mine, 0, sizeof(unsigned), st);
    if (e == cudaSuccess) *slot = mine;
    return e;
}

}  // namespace msda
