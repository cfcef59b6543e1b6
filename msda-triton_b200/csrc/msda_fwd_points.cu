// msda_fwd_points.cu -- tuned forward for point counts other than 16 per unit (exact 8 / 32 slots, padded 1..31), in its own
// translation unit (see msda_fwd_tiled.cuh).
#include "msda_fwd_tiled.cuh"

namespace msda {

cudaError_t launch_forward_tiled_points(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.L > 8 || a.LK > 32 || a.LK == 16) return cudaErrorNotSupported;
    if (a.D != 32) return cudaErrorNotSupported;
    // fp32, 9..15 points (L=3, K=4): 256-bit lanes as for 16 points (0.153 -> 0.145 ms); with more than 16 slots the
    // wide layout (>= 6 points per lane) spills and loses (20 points: 0.229 -> 0.254 ms), 8 slots stay as they are
    if (dtype == 0 && a.LK > 8 && a.LK < 16 && reinterpret_cast<uintptr_t>(a.img) % 32 == 0) {
        if (tuning().fwd_variant < 0) return launch_tiled_cfg<float, 4, 16, 512, 2, false, true, 32>(a, sm_count, st);
    }
    if (a.LK == 8) {
        if (dtype == 0) return launch_tiled_t<float, 8, 8>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_t<__half, 4, 8>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 4, 8>(a, sm_count, st);
    } else if (a.LK == 32) {
        if (dtype == 0) return launch_tiled_t<float, 8, 32>(a, sm_count, st);
    } else if (a.LK < 8) {
        if (dtype == 0) return launch_tiled_t<float, 8, 8, true>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_t<__half, 4, 8, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 4, 8, true>(a, sm_count, st);
    } else if (a.LK < 16) {
        if (dtype == 0) return launch_tiled_t<float, 8, 16, true>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_t<__half, 4, 16, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 4, 16, true>(a, sm_count, st);
    } else if (a.LK < 24) {
        if (dtype == 0) return launch_tiled_t<float, 8, 24, true>(a, sm_count, st);
    } else if (a.LK < 32) {
        if (dtype == 0) return launch_tiled_t<float, 8, 32, true>(a, sm_count, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace msda
