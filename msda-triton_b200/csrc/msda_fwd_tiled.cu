// msda_fwd_tiled.cu -- tuned forward kernel (see msda_tiled.cuh for the schedule).
//
// Per warp iteration: G = 32/LANES units.  Each lane resolves PPL = LK/LANES points, then the group walks the LK
// points in batches of NB: 5 shuffles + 4 independent 128-bit gathers per point, 4*NB gathers in flight per lane on
// top of whatever the compiler hoists from the next batch.
#include "msda_fwd_tiled.cuh"

namespace msda {

// Eligibility: D == 32 (or 64 for L*K == 16) and up to 32 sampling points per unit.  L*K in {8, 16, 32} run exact
// instantiations; any other L*K <= 32 runs the next larger slot count with the spare slots dead (RT-DETR / Mask2Former
// style L=3,K=4 -> 12 points, 5-level pyramids -> 20 points, the reference's own K=3 test fixture -> 12 points).
cudaError_t launch_forward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.L > 8 || a.LK > 32) return cudaErrorNotSupported;
    if (a.LK == 16) {
        if (dtype == 0) {
            // fp32 rows are gathered with 256-bit loads (LDG.E.256, sm_100): 4 lanes x 32 bytes per 128-byte row, so a
            // warp iteration covers 8 units -- half the load, shuffle and address instructions per unit of the
            // 8 lanes x 128-bit layout.  Measured (bench border / zeros / DETR encoder): 0.141 / 0.141 / 0.166 ms
            // versus 0.148 / 0.154 / 0.172 ms (no gain for 256-byte fp32 rows or 128-byte 16-bit rows, which stay on
            // 128-bit lanes).  512 threads x 2-point batches is the best launch shape of those tried (640 x 2, 640 x 1,
            // 384 x 4: 0.145-0.148 / 0.146-0.159 / 0.168-0.209 ms).  MSDA_B200_FWD_VARIANT=0|1 select the 128-bit layouts.
            // (256-bit loads need a 32-byte aligned pyramid; the ABI only demands 16)
            const bool wide = tuning().fwd_variant < 0 && reinterpret_cast<uintptr_t>(a.img) % 32 == 0;
            if (a.D == 32) {
                if (wide) return launch_tiled_cfg<float, 4, 16, 512, 2, false, false, 32>(a, sm_count, st);
                return launch_tiled_t<float, 8, 16>(a, sm_count, st);
            }
            if (a.D == 64) return launch_tiled_t<float, 16, 16>(a, sm_count, st);   // 256-byte rows: no gain from wide lanes
        } else if (dtype == 1) {
            if (a.D == 32) return launch_tiled_t<__half, 4, 16>(a, sm_count, st);
            if (a.D == 64) return launch_tiled_t<__half, 8, 16>(a, sm_count, st);
        } else if (dtype == 2) {
            if (a.D == 32) return launch_tiled_t<__nv_bfloat16, 4, 16>(a, sm_count, st);
            if (a.D == 64) return launch_tiled_t<__nv_bfloat16, 8, 16>(a, sm_count, st);
        }
        return cudaErrorNotSupported;
    }
    return launch_forward_tiled_points(a, dtype, sm_count, st);   // other point counts: msda_fwd_points.cu
}

}  // namespace msda
