// msda_bwd_dense.cu -- instantiations of the tuned backward with dense-level owner warps (DENSE, see msda_bwd_tiled.cuh):
// an opt-in experiment (MSDA_B200_BWD_DENSE=n), in its own translation unit so that it compiles next to the shipped paths.
#include "msda_bwd_tiled.cuh"

namespace msda {

// DENSE instantiations: 12 warps x 168 registers (the register file is split over the SM's four schedulers, so 13-16 warps
// cap a thread at 128 registers -- not enough for 64 accumulators plus three units of loads in flight without spilling into
// the LSU; 12 warps run the plain main loop as fast as 16, profiles/r2_tmem_backward.md), of which `nown` are owners.
template <int NOWN, int DPF> static cudaError_t launch_dense_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, NOWN, DPF, 384>(a, sm_count, st);
}
cudaError_t launch_backward_dense(const KernelArgs &a, int sm_count, cudaStream_t st, int nown, int prefetch) {
    // owners come in pairs (one per half of the 32 channels); 4 owners: two pairs, each taking every other unit
    if (nown >= 4) {
        if (prefetch >= 3) return launch_dense_t<4, 3>(a, sm_count, st);
        return launch_dense_t<4, 2>(a, sm_count, st);
    }
    if (prefetch >= 3) return launch_dense_t<2, 3>(a, sm_count, st);
    return launch_dense_t<2, 2>(a, sm_count, st);
}

}  // namespace msda
