// msda_bwd_module.cu -- backward of the fused module core (FUSED instantiations of msda_bwd_tiled.cuh), in its own
// translation unit.
#include "msda_bwd_tiled.cuh"

namespace msda {

// Backward of the fused module core; same eligibility as launch_module_forward_tiled.
cudaError_t launch_module_backward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK != 16 || a.L > 8 || (a.ref_dim != 2 && a.ref_dim != 4)) return cudaErrorNotSupported;
    if (a.D == 32) {
        if (dtype == 0) return launch_tiled_t<float, 8, 16, true>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_t<__half, 4, 16, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 4, 16, true>(a, sm_count, st);
    } else if (a.D == 64) {
        if (dtype == 0) return launch_tiled_t<float, 16, 16, true>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_t<__half, 8, 16, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 8, 16, true>(a, sm_count, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace msda
