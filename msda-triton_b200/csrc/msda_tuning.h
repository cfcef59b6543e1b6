// msda_tuning.h -- measurement / test knobs of libmsda_b200.so, read from the environment ONCE (first use) instead of on
// every launch.  msda_reload_tuning() (C ABI) re-reads them; tests that flip a knob in-process call it afterwards.
#pragma once

namespace msda {

struct Tuning {
    int force_generic = 0;     // MSDA_B200_FORCE_GENERIC=1    : every problem takes the generic kernels
    int slices_per_wave = 0;   // MSDA_B200_SLICES_PER_WAVE=n  : (b,h) slices per L2 wave (0 = sized to L2)
    int pace_slack = 1;        // MSDA_B200_PACE_SLACK=n       : waves a CTA may run ahead of the slowest CTA
    int wave_pacing = 1;       // MSDA_B200_WAVE_PACING=0|1|2  : off | auto | pace every multi-wave launch
    int fwd_variant = -1;      // MSDA_B200_FWD_VARIANT=0|1    : 128-bit forward layouts (default: 256-bit lanes)
    int bwd_split = 0;         // MSDA_B200_BWD_SPLIT=1        : split backward (tuned kernel + scatter kernel)
    int split_slots = 0;       // MSDA_B200_SPLIT_SLOTS=8|16   : slot count of ragged sub-unit backward launches
    int bwd_tmem = -1;         // MSDA_B200_BWD_TMEM=0|1       : coarse levels accumulated in tensor memory
                               //                                (msda_bwd_tmem.cu; default: chosen per problem)
    int tmem_warps = 15;       // MSDA_B200_TMEM_WARPS=12|15   : warps per CTA of the tensor-memory backward
    int tmem_levels = 2;       // MSDA_B200_TMEM_LEVELS=0|1|2  : at most this many of the coarsest levels go to tensor memory
    int l1_keep_kb = 100;      // MSDA_B200_L1_KEEP_KB=n       : pyramid KB per (b,h) slice the forward keeps in L1; finer levels
                               //                                are gathered with no-allocate loads (-1: never)
    int bwd_agg = -1;          // MSDA_B200_BWD_AGG=0|1        : pair aggregation of neighbouring queries' row adds (experiment, default off)
    int bwd_dense = -1;        // MSDA_B200_BWD_DENSE=n        : owner warps of the dense-level backward (experiment, default off)
    int dense_prefetch = 2;    // MSDA_B200_DENSE_PF=2|3       : units an owner warp keeps in flight (3 spills)
    int bwd_shape = -1;        // MSDA_B200_BWD_SHAPE=0..5     : fp32 backward launch shape: 16 warps x 128 regs | 12 x 168 | experiments
                               //                                (default: 12 x 168 for problems with many warp tiles per warp)
    int carveout = -1;         // MSDA_B200_CARVEOUT=0..100    : preferred shared-memory carve-out of the tuned kernels (experiment)
    int det_variant = -1;      // MSDA_B200_DET_VARIANT=0|1    : deterministic grad_img: 0 = radix sort, 1 = slice binning
};

const Tuning &tuning();
void reload_tuning();

}  // namespace msda
