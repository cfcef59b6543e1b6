// msda_bwd_tiled.cu -- tuned backward kernel: same persistent (b,h)-major schedule and lane layout as the tuned
// forward (msda_tiled.cuh).
//
// Per sampling point every lane forms, over its VEC channels, the four corner dot products <go, v_c>; from them the
// three per-point partials (grad weight, d/dx, d/dy).  The 3*LK partials of a unit live in registers until the
// end of the unit and are then reduced across the LANES lanes with a TRANSPOSING butterfly (each step halves the
// values a lane keeps), which costs 3*LK*(1 - 1/LANES) shuffles instead of 3*LK*log2(LANES) and leaves lane j
// holding exactly the PPL points it loaded -- so the grad_points / grad_weights stores are the same coalesced
// vector stores as the loads.  grad_img goes out as one REDG.E.ADD.F32x4 per lane per valid corner.
#include "msda_bwd_tiled.cuh"

namespace msda {

// MSDA_B200_BWD_SPLIT=1 selects the experimental split backward: K1 = this file's kernel without grad_img, K2 =
// msda_bwd_scatter.cu (grad_img alone, binned in shared memory).  Measured 0.19 + 0.29 ms on the bench shape versus
// 0.47 ms fused (profiles/r1_ncu_summary.md section 4), so it is opt-in.
constexpr int kMaxSplitPoints = 128;

static bool split_backward_enabled() { return tuning().bwd_split != 0; }

// Tensor-memory backward (msda_bwd_tmem.cu): OPT-IN (MSDA_B200_BWD_TMEM=1).  It removes 25-50 % of the row adds, but the
// per-record tensor-memory read-modify-write costs ~280 clk inside this kernel (50 clk in isolation), which puts
// ~10k clk of serial work on every warp tile and loses against the row adds it saves (bench shape: 0.93 ms against
// 0.47 ms; profiles/r2_tmem_backward.md has the measurements).
static bool tmem_backward_wanted(const KernelArgs &a, int sm_count) {
    (void)sm_count;
    if (!(a.flags & kNeedImg)) return false;
    return tuning().bwd_tmem > 0;
}

// Owner warps of the DENSE instantiation (0 = plain kernel).  The level shapes live on the device, so the host models
// the pyramid from Npix as L levels shrinking 4x each (exact for the benchmark pyramid): the owners are only worth their
// warps when the coarsest level has at most kDenseCells cells.  A wrong guess costs performance only -- the kernel looks at
// the real shapes and turns the owners back into workers when no level qualifies.
static int dense_owner_warps(const KernelArgs &a, int sm_count) {
    if (!(a.flags & kNeedImg) || a.L != 4 || a.K != 4 || a.D != 32) return 0;
    if (reinterpret_cast<uintptr_t>(a.aw) % 16 != 0) return 0;   // the owners read a level's 4 weights as one 16-byte load
    const int want = tuning().bwd_dense;                          // MSDA_B200_BWD_DENSE=n: n owner warps; default off (measured slower)
    (void)sm_count;
    if (want <= 0) return 0;
    return want > 8 ? 8 : want;
}

// Deterministic mode, exact row adds (msda_bwd_detq.cu prepared a.q_slmax / a.q_amax): fp32, D == 32, L*K == 16.
bool quant_backward_supported(const KernelArgs &a, int dtype) {
    return dtype == 0 && a.D == 32 && a.LK == 16 && a.L <= 8 && tiled_offsets_fit(a, sizeof(float));
}
cudaError_t launch_backward_tiled_quant(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (!quant_backward_supported(a, dtype)) return cudaErrorNotSupported;
    // same launch-shape rule as the atomic mode (see launch_backward_tiled): 12 warps x 168 registers for large problems
    const long long tiles = (long long)a.B * a.H * ((a.Q + 3) / 4);
    const int shape = tuning().bwd_shape >= 0 ? tuning().bwd_shape : (tiles >= 16LL * 16 * sm_count ? 1 : 0);
    if (shape == 1)
        return launch_tiled_t<float, 8, 16, false, 4, false, false, true, false, 0, 2, 384, 2>(a, sm_count, st);
    return launch_tiled_t<float, 8, 16, false, 4, false, false, true>(a, sm_count, st);
}

cudaError_t launch_backward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.L > 8 || a.LK > kMaxSplitPoints) return cudaErrorNotSupported;
    if (a.LK > 16)   // 3*L*K partials live in registers, 16 points at a time: larger units run as sub-units
        return launch_backward_tiled_split(a, dtype, sm_count, st);   // msda_bwd_split.cu
    if (a.LK != 16) {
        if (a.D != 32) return cudaErrorNotSupported;
        if (a.LK == 8) {
            if (dtype == 0) return launch_tiled_t<float, 8, 8>(a, sm_count, st);
            if (dtype == 1) return launch_tiled_t<__half, 8, 8, false, 4>(a, sm_count, st);
            if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 8, 8, false, 4>(a, sm_count, st);
        } else if (a.LK < 8) {
            if (dtype == 0) return launch_tiled_t<float, 8, 8, false, 4, true>(a, sm_count, st);
            if (dtype == 1) return launch_tiled_t<__half, 8, 8, false, 4, true>(a, sm_count, st);
            if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 8, 8, false, 4, true>(a, sm_count, st);
        } else {
            if (dtype == 0) return launch_tiled_t<float, 8, 16, false, 4, true>(a, sm_count, st);
            if (dtype == 1) return launch_tiled_t<__half, 8, 16, false, 4, true>(a, sm_count, st);
            if (dtype == 2) return launch_tiled_t<__nv_bfloat16, 8, 16, false, 4, true>(a, sm_count, st);
        }
        return cudaErrorNotSupported;
    }
    // grad_img ONLY (points / weights do not require grad): no pyramid gathers are needed at all, and the scatter-only
    // kernel with in-CTA binning is the faster way to produce it (bench shape: 0.29 ms versus 0.45 ms) -- provided there
    // are enough 384-query super-tiles to fill the machine.
    // The scatter kernel has no L2-sized waves, so it only takes batches whose pyramid + grad_img stay L2-resident
    // as a whole (DETR encoder B=2: 91 MB, 0.39 ms versus 0.51 ms; B=64: 21.1 ms, worse than the full tuned backward).
    const bool l2_resident = (unsigned long long)a.B * a.Npix * a.H * a.D * (2 * sizeof(float)) <= (96ull << 20);
    if (a.flags == kNeedImg && dtype == 0 && a.D == 32 && l2_resident &&
        (long long)a.B * a.H * ((a.Q + 383) / 384) >= 2LL * sm_count) {
        const cudaError_t e = launch_backward_scatter(a, dtype, sm_count, st);
        if (e != cudaErrorNotSupported) return e;
    }
    if ((a.flags & kNeedImg) && dtype == 0 && a.D == 32 && split_backward_enabled()) {
        KernelArgs k1 = a;
        k1.flags &= ~kNeedImg;
        if (k1.flags) {
            const cudaError_t e1 = launch_tiled_t<float, 8, 16>(k1, sm_count, st);
            if (e1 != cudaSuccess) return e1;
        }
        const cudaError_t e2 = launch_backward_scatter(a, dtype, sm_count, st);
        if (e2 != cudaErrorNotSupported) return e2;
        KernelArgs k2 = a;
        k2.flags = kNeedImg;
        return launch_tiled_t<float, 8, 16>(k2, sm_count, st);
    }
    if (dtype == 0) {
        if (a.D == 32) {
            if (tmem_backward_wanted(a, sm_count)) {
                const cudaError_t e = launch_backward_tmem(a, dtype, sm_count, st);
                if (e != cudaErrorNotSupported) return e;
            }
            // opt-in experiment (measured slower, see AGG above): pair aggregation of neighbouring queries' row adds
            if (tuning().bwd_agg > 0 && (a.flags & kNeedImg))
                return launch_tiled_t<float, 8, 16, false, 4, false, false, false, true>(a, sm_count, st);
            const int nown = dense_owner_warps(a, sm_count);
            if (nown > 0) return launch_backward_dense(a, sm_count, st, nown, tuning().dense_prefetch);   // msda_bwd_dense.cu
            // Launch shape (MSDA_B200_BWD_SHAPE, default -1 = by problem size): 12 warps x 168 registers run the main loop
            // without the spills of 16 warps x 128 and keep the row-add port as busy -- 1.5-3.5 % faster on every shape with
            // many warp tiles per warp (bench 0.495 -> 0.487 ms, DETR encoder 0.546 -> 0.527 ms, B=64 encoder 16.7 -> 16.4 ms;
            // scripts/time_bwd_shapes.py) -- but a decoder's 900 queries against a 22k-pixel pyramid (6 tiles per warp,
            // multi-wave) want the 16 warps: 0.149 vs 0.196 ms.  Variants 2-5 are recorded experiments (NB = 4: spills;
            // tap exchange one batch ahead: no gain, the LSU is throughput-bound, not latency-bound).
            int shape = tuning().bwd_shape;
            if (shape < 0) {
                const long long tiles = (long long)a.B * a.H * ((a.Q + 3) / 4);
                shape = tiles >= 16LL * 16 * sm_count ? 1 : 0;
                // + the finest level (first K point slots) gathered without allocating in L1, as in the forward: its
                // 512 KB per (b,h) slice cannot stay there and only evicts the coarse levels (another 1-2 %)
                // (measured: -1 % time, but the no-allocate loads also lose their lines in L2 -- DRAM reads 165 -> 239 MB per
                // launch, 1.34x the algorithmic bytes -- so it stays a knob, MSDA_B200_BWD_SHAPE=6)
            }
            if (shape == 1)
                return launch_tiled_t<float, 8, 16, false, 4, false, false, false, false, 0, 2, 384, 2>(a, sm_count, st);
            if (shape >= 2) return launch_backward_shape_variant(a, shape, sm_count, st);   // msda_bwd_shapes.cu
            return launch_tiled_t<float, 8, 16>(a, sm_count, st);
        }
        if (a.D == 64) return launch_tiled_t<float, 16, 16>(a, sm_count, st);
    } else if (dtype == 1) {
        // (16-bit storage keeps the 16 warps: at 12 x 168 the bf16 backward measured -0.6 % on the bench shape, +0.7 % on the
        // DETR encoder, +1 % at B = 16 -- MSDA_TIME_DTYPE=bf16 scripts/time_bwd_shapes.py)
        if (a.D == 32) return launch_tiled_t<__half, 8, 16, false, 4>(a, sm_count, st);
        if (a.D == 64) return launch_tiled_t<__half, 16, 16, false, 4>(a, sm_count, st);
    } else if (dtype == 2) {
        if (a.D == 32) return launch_tiled_t<__nv_bfloat16, 8, 16, false, 4>(a, sm_count, st);
        if (a.D == 64) return launch_tiled_t<__nv_bfloat16, 16, 16, false, 4>(a, sm_count, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace msda
