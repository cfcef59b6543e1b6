// msda_bwd_split.cu -- tuned backward for units with more than 16 sampling points (SPLIT instantiations of
// msda_bwd_tiled.cuh: sub-units of 8 or 16 point slots), in its own translation unit.
#include "msda_bwd_tiled.cuh"

namespace msda {

// More than 16 sampling points per unit (5-level pyramids, K = 8): sub-units of SLOTS points, see decode_tile().
template <typename T, int LANES, int SLOTS, bool PADDED>
static cudaError_t launch_split_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    const int subs = (a.LK + SLOTS - 1) / SLOTS;
    return launch_tiled_t<T, LANES, SLOTS, false, 4, PADDED, true>(a, sm_count, st, subs);
}

template <typename T> static cudaError_t launch_split(const KernelArgs &a, int sm_count, cudaStream_t st) {
    if (a.D == 32) {
        if (a.LK % 16 == 0) return launch_split_t<T, 8, 16, false>(a, sm_count, st);
        if (a.LK % 8 == 0) return launch_split_t<T, 8, 8, false>(a, sm_count, st);
        // ragged: the slot count that wastes fewer dead slots (20 points: 3 x 8 rather than 2 x 16; 0.91 vs 1.10 ms)
        const int dead16 = (a.LK + 15) / 16 * 16 - a.LK, dead8 = (a.LK + 7) / 8 * 8 - a.LK;
        bool use8 = dead8 <= dead16;   // tie: 8 slots measured faster (28 points: 0.89 vs 0.94 ms)
        if (tuning().split_slots > 0) use8 = tuning().split_slots == 8;   // tuning knob
        if (use8) return launch_split_t<T, 8, 8, true>(a, sm_count, st);
        return launch_split_t<T, 8, 16, true>(a, sm_count, st);
    }
    if (a.D == 64 && a.LK % 16 == 0) return launch_split_t<T, 16, 16, false>(a, sm_count, st);
    return cudaErrorNotSupported;
}

cudaError_t launch_backward_tiled_split(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK <= 16) return cudaErrorNotSupported;
    if (dtype == 0) return launch_split<float>(a, sm_count, st);
    if (dtype == 1) return launch_split<__half>(a, sm_count, st);
    if (dtype == 2) return launch_split<__nv_bfloat16>(a, sm_count, st);
    return cudaErrorNotSupported;
}

}  // namespace msda
