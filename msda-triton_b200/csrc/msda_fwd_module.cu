// msda_fwd_module.cu -- forward of the fused module core (FUSED instantiations of the tuned forward kernel,
// msda_fwd_tiled.cuh), in its own translation unit.
#include "msda_fwd_tiled.cuh"

namespace msda {

// Fused module core: (fp32 | fp16 | bf16) x D in {32, 64} x L*K=16 -- hidden 256 or 512 with 8 heads.
cudaError_t launch_module_forward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK != 16 || a.L > 8 || (a.ref_dim != 2 && a.ref_dim != 4)) return cudaErrorNotSupported;
    if (a.D == 32) {
        if (dtype == 0) {
            if (tuning().fwd_variant < 0 && reinterpret_cast<uintptr_t>(a.img) % 32 == 0)
                return launch_tiled_cfg<float, 4, 16, 512, 2, true, false, 32>(a, sm_count, st);
            return launch_tiled_cfg<float, 8, 16, 1024, 2, true>(a, sm_count, st);
        }
        if (dtype == 1) return launch_tiled_cfg<__half, 4, 16, 512, 4, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_cfg<__nv_bfloat16, 4, 16, 512, 4, true>(a, sm_count, st);
    } else if (a.D == 64) {   // hidden 512 / 8 heads: the reference README's module example
        if (dtype == 0) return launch_tiled_cfg<float, 16, 16, 1024, 2, true>(a, sm_count, st);
        if (dtype == 1) return launch_tiled_cfg<__half, 8, 16, 1024, 2, true>(a, sm_count, st);
        if (dtype == 2) return launch_tiled_cfg<__nv_bfloat16, 8, 16, 1024, 2, true>(a, sm_count, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace msda
