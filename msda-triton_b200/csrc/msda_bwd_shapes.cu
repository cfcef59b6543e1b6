// msda_bwd_shapes.cu -- launch-shape experiments of the tuned fp32 backward (MSDA_B200_BWD_SHAPE=2..8, see
// launch_backward_tiled in msda_bwd_tiled.cu and profiles/r2_dense_backward.md section 2): deeper gather batches, the tap
// exchange issued one batch ahead, no-allocate gathers of the finest levels.  None of them is a default.
#include "msda_bwd_tiled.cuh"

namespace msda {

cudaError_t launch_backward_shape_variant(const KernelArgs &a, int shape, int sm_count, cudaStream_t st) {
    if (shape == 2)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 4>(a, sm_count, st);
    if (shape == 3)   // + tap exchange one batch ahead
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 2, true>(a, sm_count, st);
    if (shape == 4)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 512, 2, true>(a, sm_count, st);
    if (shape == 5)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 4, true>(a, sm_count, st);
    if (shape == 6)   // 12 x 168, first 4 / 8 / 12 point slots gathered with no-allocate loads
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 2, false, 4>(a, sm_count, st);
    if (shape == 7)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 2, false, 8>(a, sm_count, st);
    if (shape == 8)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 2, false, 12>(a, sm_count, st);
    return cudaErrorNotSupported;
}

}  // namespace msda
