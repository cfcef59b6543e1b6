// msda_fwd_tiled.cuh -- the tuned forward kernel template and its launcher templates, shared by the translation units that
// instantiate it (msda_fwd_tiled.cu: 16 point slots; msda_fwd_points.cu: other point counts; msda_fwd_module.cu: the fused
// module core) so that the instantiations compile in parallel.
//
// Tuned forward kernel (see msda_tiled.cuh for the schedule).
//
// Per warp iteration: G = 32/LANES units.  Each lane resolves PPL = LK/LANES points, then the group walks the LK
// points in batches of NB: 5 shuffles + 4 independent 128-bit gathers per point, 4*NB gathers in flight per lane on
// top of whatever the compiler hoists from the next batch.
#pragma once
#include <cstdint>
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"

namespace msda {

// pyramid bytes of the slices gathered concurrently (one wave) -- well inside the 126 MB L2
constexpr size_t kFwdL2Budget = 24u << 20;

// FUSED = the module core (frontend.py:253-289): operands are the raw query projection [.., L, K, 3] and the reference
// points; softmax and the sampling-point arithmetic happen in registers, sampling_points / attention_weights are never
// materialised.
// PADDED = the unit has a.LK <= LK points; the LK - a.LK trailing slots are skipped (warp-uniformly) by every loop.
// NA     = the first NA point slots (the finest levels) are gathered with no-allocate loads (streamed_points()).
template <typename T, int LANES, int LK, bool BORDER, int NB, int THREADS, bool FUSED, bool PADDED, int VECB = 16,
          int NA = 0>
__global__ void __launch_bounds__(THREADS, 1)
    msda_fwd_tiled_kernel(const KernelArgs a, const WaveSchedule ws) {
    using Cfg = TiledCfg<T, LANES, LK, VECB>;
    using Raw = typename RawSlice<VECB>::type;
    constexpr int VEC = Cfg::VEC, G = Cfg::G, PPL = Cfg::PPL;
    static_assert(LANES % NB == 0, "batch must divide the group");

    __shared__ Level s_lv[8];   // tuned kernels take L <= 8
    __shared__ unsigned s_pace[kPaceRing];
    if (threadIdx.x < kPaceRing) s_pace[threadIdx.x] = 0;
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;

    const T *__restrict__ img = static_cast<const T *>(a.img);
    T *__restrict__ out = static_cast<T *>(a.out);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = THREADS >> 5;
    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);

    const int tiles_per_bh = ws.tiles_per_bh;
    for (int wave = 0; wave < ws.waves; ++wave) {
    int t_begin, t_end;
    wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);

    int tile = t_begin + warp;
    if (tile < t_end) {

    // software pipeline: sampling points / weights of the next warp tile are in flight while this one is processed
    TileUnit tu = decode_tile(tile, tiles_per_bh, g, G, a);
    LaneOperands<T, PPL, FUSED> op;
    load_operands<T, LANES, LK, FUSED, PADDED>(a, tu, j, op);

    for (; tile < t_end; tile += nwarps) {
        const int tile_n = tile + nwarps;
        const bool has_next = tile_n < t_end;
        const TileUnit tu_n = decode_tile(has_next ? tile_n : tile, tiles_per_bh, g, G, a);
        LaneOperands<T, PPL, FUSED> op_n;
        load_operands<T, LANES, LK, FUSED, PADDED>(a, tu_n, j, op_n);
        if constexpr (FUSED) derive_operands<T, LANES, LK>(a, s_lv, j, op);

        const unsigned char *__restrict__ lane_base =
            reinterpret_cast<const unsigned char *>(img + tu.bh_off + j * VEC);

        TileTap tap[PPL];
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp)
            tap[pp] = resolve_tap<BORDER>(op.xy[2 * pp], op.xy[2 * pp + 1], s_lv[slot_level(j * PPL + pp, a)], align, row_bytes);

        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;

#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) {
#pragma unroll
            for (int jj0 = 0; jj0 < LANES; jj0 += NB) {
                Raw raw[NB][4];
                float fx[NB], fy[NB], fw[NB];
                unsigned msk[NB];
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    const int src = jj0 + n;
                    if (PADDED && src * PPL + pp >= a.LK) continue;   // dead slot (warp-uniform)
                    const unsigned off = __shfl_sync(0xffffffffu, tap[pp].off, src, LANES);
                    const unsigned pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                    fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                    fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                    fw[n] = __shfl_sync(0xffffffffu, op.wa[pp], src, LANES);
                    msk[n] = (pack >> kPackMaskShift) & 0xFu;
                    unsigned o[4];
                    corner_offsets(off, pack, row_bytes, o);
                    // corner rows are clamped into the level, so all four gathers are always in range; zeros padding
                    // (kernels.py:227-231: out-of-range corners read as 0) is applied when the values are consumed,
                    // which keeps the 4*NB loads independent and in flight together
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (src * PPL + pp < NA)   // (constant after unrolling) a level that cannot live in L1: do not allocate
                            raw[n][c] = gather_slice_na<VECB>(lane_base, o[c]);
                        else
                            raw[n][c] = gather_slice<VECB>(lane_base, o[c]);
                    }
                }
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    if (PADDED && (jj0 + n) * PPL + pp >= a.LK) continue;
                    const float wy1 = fw[n] * fy[n], wy0 = fw[n] - wy1;  // w*dy, w*(1-dy)
                    float w[4];
                    w[1] = wy0 * fx[n];
                    w[0] = wy0 - w[1];
                    w[3] = wy1 * fx[n];
                    w[2] = wy1 - w[3];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v[VEC];
                        widen_row<T, VEC>(raw[n][c], v);
                        if (BORDER || ((msk[n] >> c) & 1u)) {
#pragma unroll
                            for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w[c], v[e], acc[e]);
                        }
                    }
                }
            }
        }
        if (tu.live) store_vec_stream<T, VEC>(out + (size_t)tu.u * a.D + j * VEC, acc);

        tu = tu_n;
        op = op_n;
    }
    }
    wave_pace_warp(ws, wave, s_pace, lane, nwarps);
    }  // waves
}

template <typename T, int LANES, int LK, int THREADS, int NB, bool FUSED = false, bool PADDED = false, int VECB = 16,
          int NA = -1>
static cudaError_t launch_tiled_cfg(const KernelArgs &a, int sm_count, cudaStream_t st) {
    if constexpr (NA < 0) {
        // 16 exact slots, K == 4: instantiations that stream the first 0 / 4 / 8 / 12 points (whole levels)
        if constexpr (LK == 16 && !PADDED) {
            if (a.K == 4 && a.L == 4) {
                switch (streamed_points(a, (size_t)a.D * sizeof(T))) {
                    case 4: return launch_tiled_cfg<T, LANES, LK, THREADS, NB, FUSED, PADDED, VECB, 4>(a, sm_count, st);
                    case 8: return launch_tiled_cfg<T, LANES, LK, THREADS, NB, FUSED, PADDED, VECB, 8>(a, sm_count, st);
                    case 12: return launch_tiled_cfg<T, LANES, LK, THREADS, NB, FUSED, PADDED, VECB, 12>(a, sm_count, st);
                    default: break;
                }
            }
        }
        return launch_tiled_cfg<T, LANES, LK, THREADS, NB, FUSED, PADDED, VECB, 0>(a, sm_count, st);
    } else {
    constexpr int NAK = NA;
    constexpr int G = TiledCfg<T, LANES, LK, VECB>::G;
    if (!tiled_offsets_fit(a, sizeof(T))) return cudaErrorNotSupported;
    const int tiles_per_bh = (a.Q + G - 1) / G;
    const int total_tiles = a.B * a.H * tiles_per_bh;
    const int warps = THREADS / 32;
    const int want = (total_tiles + warps - 1) / warps;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    WaveSchedule ws = make_wave_schedule(a, tiles_per_bh, sizeof(T), kFwdL2Budget, 4LL * warps * grid);
    // many waves of substantial size: keep the persistent CTAs on the same wave (wave_pace).  Waves with only a few
    // tiles per warp (decoder: 900 queries against a 22k-pixel pyramid) cannot drift far and would only pay the
    // per-wave handshake (measured on that shape: module step 0.93 -> 1.14 ms when paced).
    const bool big_waves = (long long)ws.slices_per_wave * tiles_per_bh >= 4LL * warps * grid;
    if (ws.waves > 1 && grid == sm_count && (big_waves || pacing_forced())) {
        const cudaError_t e = acquire_pace_counter(st, &ws.pace);
        if (e != cudaSuccess) return e;
    }
    if (tuning().carveout >= 0) {   // experiment knob: shared-memory carve-out (percent) = what is left for L1
        cudaFuncSetAttribute(msda_fwd_tiled_kernel<T, LANES, LK, true, NB, THREADS, FUSED, PADDED, VECB, NAK>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, tuning().carveout);
        cudaFuncSetAttribute(msda_fwd_tiled_kernel<T, LANES, LK, false, NB, THREADS, FUSED, PADDED, VECB, NAK>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, tuning().carveout);
    }
    if (a.border)
        msda_fwd_tiled_kernel<T, LANES, LK, true, NB, THREADS, FUSED, PADDED, VECB, NAK><<<grid, THREADS, 0, st>>>(a, ws);
    else
        msda_fwd_tiled_kernel<T, LANES, LK, false, NB, THREADS, FUSED, PADDED, VECB, NAK><<<grid, THREADS, 0, st>>>(a, ws);
    return cudaGetLastError();
    }
}

// Launch shape.  Measured on B200 (cold L2, fp32 D=32): 1024 threads x 2-point gather batches (64 registers) versus
// 512 threads x 4-point batches (128 registers): bench shape border 0.147 vs 0.142 ms, zeros 0.153 vs 0.163 ms,
// DETR encoder 0.173 vs 0.209 ms -- more resident warps hide the L2 latency of the levels that do not fit L1.
template <typename T, int LANES, int LK, bool PADDED = false>
static cudaError_t launch_tiled_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    if constexpr (!PADDED) {
        if (tuning().fwd_variant == 1) return launch_tiled_cfg<T, LANES, LK, 512, 4>(a, sm_count, st);   // tuning knob
    }
    // lanes that own >= 3 points keep more state: stay at 128 registers there
    if constexpr (TiledCfg<T, LANES, LK>::PPL >= 3)
        return launch_tiled_cfg<T, LANES, LK, 512, 4, false, PADDED>(a, sm_count, st);
    else
        return launch_tiled_cfg<T, LANES, LK, 1024, 2, false, PADDED>(a, sm_count, st);
}

}  // namespace msda
