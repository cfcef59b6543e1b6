// msda_bwd_tiled.cuh -- the tuned backward kernel template and its launcher template, shared by the translation units that
// instantiate it: msda_bwd_tiled.cu (the shipped paths), msda_bwd_dense.cu (dense-level owner warps, experiment) and
// msda_bwd_shapes.cu (launch-shape experiments).  Three files so that the ~60 instantiations compile in parallel.
//
// forward (msda_tiled.cuh).
//
// Per sampling point every lane forms, over its VEC channels, the four corner dot products <go, v_c>; from them the
// three per-point partials (grad weight, d/dx, d/dy).  The 3*LK partials of a unit live in registers until the
// end of the unit and are then reduced across the LANES lanes with a TRANSPOSING butterfly (each step halves the
// values a lane keeps), which costs 3*LK*(1 - 1/LANES) shuffles instead of 3*LK*log2(LANES) and leaves lane j
// holding exactly the PPL points it loaded -- so the grad_points / grad_weights stores are the same coalesced
// vector stores as the loads.  grad_img goes out as one REDG.E.ADD.F32x4 per lane per valid corner.
#pragma once
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"

namespace msda {

constexpr int kNeedImg = 1, kNeedPts = 2, kNeedAw = 4;
constexpr size_t kBwdL2Budget = 48u << 20;  // img + grad_img bytes of one wave of (b,h) slices

// One lane's VEC channels of a grad_img row as 16-byte reductions.  VEC == 4 (fp32 storage): channels 4j..4j+3 are
// one red.v4 and the LANES lanes of a group cover the row contiguously.  VEC == 8 (16-bit storage, fp32 accumulation
// image): a lane owns channels 8j..8j+7, i.e. 32 bytes; issued naively the group's two instructions would each touch
// HALF of every 32-byte sector (measured: 2x the L2 atomic sector operations, backward 1.0 ms instead of 0.5 ms).  The
// accumulation image is private scratch, so its channel order is permuted instead: channel 8j + 4h + e is stored at
// position 4*LANES*h + 4j + e, which makes instruction h of all lanes one contiguous 16*LANES-byte run.
// `dst` already points at position 4j of the row.  launch_round_grad_img() undoes the permutation.
template <int VEC, int LANES> __device__ __forceinline__ void red_add_row(float *dst, const float (&gv)[VEC]) {
    red_add_v4(dst, gv[0], gv[1], gv[2], gv[3]);
    if constexpr (VEC == 8) red_add_v4(dst + 4 * LANES, gv[4], gv[5], gv[6], gv[7]);
    static_assert(VEC == 4 || VEC == 8, "tuned kernels use 128-bit lanes");
}

// ---------------------------------------------------------------------------------------------------------------------
// DENSE (opt-in experiment, MSDA_B200_BWD_DENSE=2|4): register-resident accumulation of ONE coarse pyramid level by
// specialised "owner" warps.
//
// The backward is bound by the SM's row-add injection into L2 (5.8 clk per 128-byte row).  Earlier attempts to merge row
// adds on chip all had the worker warps hand their records to an accumulator through shared or tensor memory and lost to
// the hand-off (LSU queueing behind the reds, tensor-memory round trips).  Here nothing is handed over: the last NOWN
// warps of the CTA do not take warp tiles at all; they walk the SAME units of the CTA's range a second time, read only
// the 112 bytes a unit needs for one level and half of the channels (warp-uniform 16-byte loads, DPF units in flight),
// and accumulate the level DENSELY in registers: lane = cell (two cells per lane: levels of at most 64 cells, 8x8 on the
// benchmark pyramid), 16 channels per owner, weight(cell) = sum over the level's 4 points of attention weight x
// tent(x - cell_x) x tent(y - cell_y) -- the bilinear corner weights written per cell, so that no register is indexed
// dynamically and nothing branches (dense_owner_wave below).  The owners flush their rows when the CTA's range leaves a
// (b,h) slice, and the workers skip the row adds of that level: a quarter of the backward's `red` sectors never leave
// the SM.  Eligibility (host): fp32, D == 32, L == 4, K == 4, grad_img requested; the level is picked on the device
// (largest level with h*w <= 64); without one the owners only keep the wave barriers company.  12 warps x 168 registers.
// MEASURED (B200, bench shape, profiles/r2_dense_backward.md): correct (tests/test_dense_backward_gpu.py) and SLOWER than
// the plain kernel (0.47 ms): 2.05 / 1.07 ms with 2 / 4 owners -- time ~ 1 / owners at ~1.9k clk per unit, i.e. one trip
// through the LSU per unit: under the reds every load takes that long, a unit in flight costs 28 registers, and 64
// accumulators' worth of register file leaves room for two (three spill: 1.29 ms).  The first version of this experiment
// (lane = channel, register = cell, the add dispatched through a 64-way brx.idx; commit 9e100fd) needed only 13
// registers per unit in flight and hid the loads, but paid ~3.5k clk per unit for its 16 taken branches: 1.75 / 1.10 / 1.01
// ms with 2 / 4 / 6 owners.  And owner warps are not free: with the owners idle the kernel slows down ~ 1 / workers below
// 12 warps.  Kept as a recorded experiment; default off.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDenseCells = 64;
constexpr unsigned kPackDenseBit = 1u << 30;   // TileTap::pack: the point's level is accumulated by the owner warps

// What an owner warp holds of one unit while its loads are in flight: its half of the grad_out row (16 channels), the
// dense level's 4 sampling points (x, y) and their attention weights -- all loaded with warp-uniform addresses.
struct DenseSlot {
    float4 g[4];
    float4 p0, p1;
    float4 w;
};

// Position of an owner warp in the flattened units i = 4 tile + query slot of the CTA's tile range: (b,h) slice `bh` starts
// at flat index `start` and is `slice_len` long (its last positions may be padding queries, q >= Q).  Advanced without
// divisions; the fetch cursor runs DPF units ahead of the consume cursor.
struct DenseCursor {
    int i, start, bh;
    __device__ __forceinline__ void settle(const int slice_len) {
        while (i >= start + slice_len) {
            start += slice_len;
            ++bh;
            asm volatile("");   // keeps the loop a loop (the compiler otherwise rewrites it into a division per call)
        }
    }
};

// One wave of an owner warp.  LANE = CELL: lane l owns cells l and l + 32 of the level and accumulates, for its 16
// channels (owner o: channels 16 (o & 1) ..), acc[cell][channel] += weight(cell) * grad_out[channel], where weight(cell) is
// the sum over the level's 4 points of  attention weight x tent(x - cell_x) x tent(y - cell_y)  -- the bilinear corner
// weights written per CELL instead of per corner (tent(d) = max(0, 1 - |d|); border padding: on the clamped coordinate,
// which folds clamped twin corners exactly as the row adds would).  No register is indexed dynamically and nothing
// branches: ~120 instructions of weights + 32 FFMA per unit.  Owners with the same (o & 1) split the units between them.
template <bool BORDER, int DPF>
__device__ __forceinline__ void dense_owner_wave(const KernelArgs &a, const Level lv, const int level, const int t_begin,
                                                 const int t_end, const int tiles_per_bh, const int o, const int nown,
                                                 const int lane, const bool align, const float *__restrict__ gout,
                                                 float *__restrict__ gimg) {
    constexpr int G = 4;   // queries per warp tile of the 8-lane layout
    const float *__restrict__ pts = static_cast<const float *>(a.pts);
    const float *__restrict__ aw = static_cast<const float *>(a.aw);
    const int end = t_end * G;
    const int cells = lv.h * lv.w;
    const int slice_len = tiles_per_bh * G;
    const size_t row_stride = (size_t)a.H * a.D;
    const int ch0 = (o & 1) * 16;                 // this owner's channels
    const int step = nown >> 1, phase = o >> 1;   // ... and its share of the units

    // this lane's two cells as (x, y); a cell beyond the level sits far away from every coordinate (weight 0)
    const int cA = lane, cB = lane + 32;
    const float fxA = cA < cells ? (float)(cA % lv.w) : 1.0e30f, fyA = cA < cells ? (float)(cA / lv.w) : 1.0e30f;
    const float fxB = cB < cells ? (float)(cB % lv.w) : 1.0e30f, fyB = cB < cells ? (float)(cB / lv.w) : 1.0e30f;
    const float wf = (float)lv.w, hf = (float)lv.h, wm1 = (float)(lv.w - 1), hm1 = (float)(lv.h - 1);

    float accA[16], accB[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) accA[c] = accB[c] = 0.0f;
    int acc_bh = -1;

    DenseCursor cur, pre;   // consume / fetch
    cur.i = t_begin * G + phase;
    cur.bh = t_begin / tiles_per_bh;
    cur.start = cur.bh * slice_len;
    cur.settle(slice_len);
    pre = cur;

    // loads of the unit under the fetch cursor (nothing when it is past the end or on a padding query), then one step
    auto fetch = [&]() {
        DenseSlot s;
        const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        s.g[0] = s.g[1] = s.g[2] = s.g[3] = s.p0 = s.p1 = s.w = zero;
        const int q = pre.i - pre.start;
        if (pre.i < end && q < a.Q) {   // warp-uniform
            const int b = pre.bh / a.H, h = pre.bh - b * a.H;
            const size_t u = ((size_t)b * a.Q + q) * a.H + h;
            const float4 *gp = reinterpret_cast<const float4 *>(gout + u * 32 + ch0);
#pragma unroll
            for (int k = 0; k < 4; ++k) s.g[k] = __ldg(gp + k);
            const float4 *pp = reinterpret_cast<const float4 *>(pts + (u * 16 + level * 4) * 2);
            s.p0 = __ldg(pp);
            s.p1 = __ldg(pp + 1);
            s.w = __ldg(reinterpret_cast<const float4 *>(aw + u * 16 + level * 4));
        }
        pre.i += step;
        pre.settle(slice_len);
        return s;
    };

    // DPF units in flight; the slot a unit leaves is refilled AFTER the unit has been accumulated
    DenseSlot ring[DPF];
#pragma unroll
    for (int k = 0; k < DPF; ++k) ring[k] = fetch();

    // one unit: flush on a slice change, accumulate the unit held in `slot`, refill `slot`; true when the range is done.
    // The ring is walked with static indices (the loop below is unrolled over it): rotating the slots through ring[0]
    // would keep two copies of a slot alive across the moves.
    auto unit = [&](DenseSlot &slot) -> bool {
        cur.settle(slice_len);
        const bool done = cur.i >= end;
        const bool valid = !done && cur.i - cur.start < a.Q;   // else: padding query, nothing to add
        const int bh = done ? -2 : cur.bh;
        if ((done || valid) && bh != acc_bh) {
            // the range leaves slice acc_bh (or ends): the lane's two rows, 16 channels each, as four row adds per cell
            if (acc_bh >= 0) {
                const int b = acc_bh / a.H, h = acc_bh - b * a.H;
                float *__restrict__ base = gimg + ((size_t)b * a.Npix * a.H + h) * a.D + (size_t)lv.off * row_stride + ch0;
                if (cA < cells) {
                    float *dst = base + (size_t)cA * row_stride;
#pragma unroll
                    for (int k = 0; k < 4; ++k) red_add_v4(dst + 4 * k, accA[4 * k], accA[4 * k + 1], accA[4 * k + 2], accA[4 * k + 3]);
                }
                if (cB < cells) {
                    float *dst = base + (size_t)cB * row_stride;
#pragma unroll
                    for (int k = 0; k < 4; ++k) red_add_v4(dst + 4 * k, accB[4 * k], accB[4 * k + 1], accB[4 * k + 2], accB[4 * k + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) accA[c] = accB[c] = 0.0f;
            acc_bh = bh;
        }
        if (done) return true;
        const DenseSlot &s = slot;
        if (valid) {
            // the four points one after the other (not unrolled: interleaving them costs ~40 registers of temporaries, which
            // the prefetched units need); they rotate through (px0, py0, fw0)
            float px0 = s.p0.x, py0 = s.p0.y, px1 = s.p0.z, py1 = s.p0.w, px2 = s.p1.x, py2 = s.p1.y, px3 = s.p1.z, py3 = s.p1.w;
            float fw0 = s.w.x, fw1 = s.w.y, fw2 = s.w.z, fw3 = s.w.w;
            float wA = 0.0f, wB = 0.0f;
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                // the same un-normalisation, in the same operation order, as locate()
                float x = align ? mul_rn(px0, wm1) : sub_rn(mul_rn(px0, wf), 0.5f);
                float y = align ? mul_rn(py0, hm1) : sub_rn(mul_rn(py0, hf), 0.5f);
                float wp = fw0;
                if constexpr (BORDER) {
                    // a non-finite coordinate gives NaN corner weights on the clamped cell (x - floor(x) is NaN): carry
                    // that over as a NaN attention weight -- (x - x) is 0 for finite x, NaN otherwise
                    wp = wp + __fsub_rn(x, x) + __fsub_rn(y, y);
                    x = fminf(fmaxf(x, 0.0f), wm1);
                    y = fminf(fmaxf(y, 0.0f), hm1);
                }
                const float tA = fmaxf(0.0f, 1.0f - fabsf(x - fxA)) * fmaxf(0.0f, 1.0f - fabsf(y - fyA));
                const float tB = fmaxf(0.0f, 1.0f - fabsf(x - fxB)) * fmaxf(0.0f, 1.0f - fabsf(y - fyB));
                // only the cells the point touches take its weight (a NaN / Inf weight must not reach the others)
                if (tA > 0.0f) wA = fmaf(wp, tA, wA);
                if (tB > 0.0f) wB = fmaf(wp, tB, wB);
                px0 = px1, py0 = py1, fw0 = fw1;
                px1 = px2, py1 = py2, fw1 = fw2;
                px2 = px3, py2 = py3, fw2 = fw3;
            }
            const float gch[16] = {s.g[0].x, s.g[0].y, s.g[0].z, s.g[0].w, s.g[1].x, s.g[1].y, s.g[1].z, s.g[1].w,
                                   s.g[2].x, s.g[2].y, s.g[2].z, s.g[2].w, s.g[3].x, s.g[3].y, s.g[3].z, s.g[3].w};
            // untouched cells (weight 0) are skipped: 0 x Inf of a non-finite grad_out must not poison them
            if (wA != 0.0f) {
#pragma unroll
                for (int c = 0; c < 16; ++c) accA[c] = fmaf(wA, gch[c], accA[c]);
            }
            if (wB != 0.0f) {
#pragma unroll
                for (int c = 0; c < 16; ++c) accB[c] = fmaf(wB, gch[c], accB[c]);
            }
        }
        slot = fetch();
        cur.i += step;
        return false;
    };
    for (bool done = false; !done;) {
#pragma unroll
        for (int k = 0; k < DPF; ++k) {
            if (!done) done = unit(ring[k]);
        }
    }
}

// FUSED = backward of the module core: operands are the raw projection + reference points (see msda_tiled.cuh);
// the epilogue turns (grad weight, grad point) into grad of the projection triples (softmax backward, 1/shape or
// ref_wh/2K scaling) and accumulates grad of the reference points with a handful of scalar atomics per unit.
// VEC = channels per lane: 16 bytes per lane by default; the non-fused 16-bit-storage backward runs 8 lanes x 4
// channels (8-byte gathers) so that its fp32 row adds have the same full-sector shape as the fp32 kernel's.
// PADDED: see the forward kernel -- a.LK <= LK real points, per-point loads / stores; dead slots are gathered (point
// (0,0), weight 0) but add nothing.
// SPLIT: units with more than LK points run as `subs` sub-units of LK slots each (decode_tile in msda_tiled.cuh).
// QUANT: deterministic mode (msda_bwd_detq.cu).  Every value added to grad_img is first rounded to a multiple of a
// power-of-two quantum q chosen per (b, h, level) such that NO partial sum of a row can exceed 2^24 q: all the fp32 adds
// of the row are then exact, hence associative, and the relaxed atomics give the same bits in any order.
// AGG: warp-level aggregation of the row adds of NEIGHBOURING QUERIES.  A warp tile holds four consecutive queries of one
// (b,h); when two of them (lane groups g and g^1) put a sampling point into the same pixel cell with the same corner
// validity -- encoder self-attention: query q sits on pixel q, so adjacent queries hit the same cell on every level
// coarser than their own -- the even group adds BOTH contributions with one row add and the odd group adds nothing.
// The groups do not exchange the 16 products of a point (as many LSU wavefronts as the adds saved) but the three
// numbers they derive from: (attention weight, dx, dy) of the partner's point, plus the partner's grad_out slice once
// per tile.  Finding the pairs costs 4 shuffles + 2 votes per TILE (every lane compares the taps of its own two points
// with the partner group's); tiles without a pair run the plain path.
// MEASURED (B200, DETR encoder B=2): it loses.  Uniform random points (no pairs): 0.543 -> 0.582 ms, the price of the
// check and of the extra live registers; freshly-initialised-DETR points (identical offsets for all queries: ~53 % of the
// point slots pair, 26 % fewer row adds): 0.498 -> 0.568 ms.  A row add costs the SM 5.8 clk per 128-byte row whatever
// the shape of the instruction that carries it (scripts/micro/red_shapes.cu: lane groups predicated off, adjacent rows,
// v2 / scalar forms all give 6.2-6.4 TB/s), so a paired point saves 4 rows = 23 clk -- and pays three full-warp shuffles
// through the same saturated LSU plus the partner's weight arithmetic, which is no cheaper.
// OPT-IN: MSDA_B200_BWD_AGG=1 (tests keep it correct).
template <typename T, int LANES, int LK, bool BORDER, int NB, int THREADS, bool FUSED, int VEC, bool PADDED,
          bool SPLIT, bool QUANT = false, bool AGG = false, int NOWN = 0, int DPF = 2, bool PIPE = false, int NA = 0>
__global__ void __launch_bounds__(THREADS, 1)
    msda_bwd_tiled_kernel(const KernelArgs a, const WaveSchedule ws, const int subs_arg) {
    constexpr bool DENSE = NOWN > 0;   // the last NOWN warps of the CTA are owner warps (see the DENSE notes above)
    static_assert(!AGG || (LANES == 8 && VEC == 4 && !PADDED && !SPLIT), "pair aggregation: four 8-lane groups per warp");
    static_assert(!DENSE || (std::is_same<T, float>::value && LANES == 8 && LK == 16 && VEC == 4 && !FUSED && !PADDED &&
                             !SPLIT && !QUANT && !AGG),
                  "dense owner warps: the plain fp32 D = 32, L = K = 4 instantiation only");
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int G = Cfg::G, PPL = Cfg::PPL;
    using Raw = typename RawSlice<VEC * (int)sizeof(T)>::type;
    static_assert(LANES % NB == 0, "batch must divide the group");

    __shared__ Level s_lv[8];   // tuned kernels take L <= 8
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = THREADS >> 5;
    const int j = lane % LANES, g = lane / LANES;
    // DENSE: the level the owner warps accumulate = the largest level of at most kDenseCells cells (-1: none, the workers
    // add everything).  The workers see it as Level::pad != 0 of the level a point belongs to (no register, no extra load).
    __shared__ int s_dense_level;
    if constexpr (DENSE) {
        __syncthreads();   // every thread has read the verdict build_level_table() left in s_lv[0].pad
        if (threadIdx.x == 0) {
            int best = 0, pick = -1;
            for (int l = 0; l < a.L; ++l) {
                const int cells = s_lv[l].h * s_lv[l].w;
                if (cells <= kDenseCells && cells > best) {
                    best = cells;
                    pick = l;
                }
            }
            for (int l = 0; l < a.L; ++l) s_lv[l].pad = (l == pick) ? 1 : 0;
            s_dense_level = pick;
        }
        __syncthreads();
    }
    constexpr int nown = NOWN, nworkers = nwarps - NOWN;

    const bool align = a.align != 0;
    const bool need_img = (a.flags & kNeedImg) != 0, need_pts = (a.flags & kNeedPts) != 0,
               need_aw = (a.flags & kNeedAw) != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);
    constexpr unsigned kAccScale = sizeof(float) / sizeof(T);  // accumulation row bytes / storage row bytes

    const int tiles_per_bh = ws.tiles_per_bh;
    const int subs = SPLIT ? subs_arg : 1;
    static_assert(!(SPLIT && FUSED), "the fused module core is instantiated for L*K == 16 only");
    if constexpr (DENSE) {
        if (warp >= nworkers) {   // owner warp: its own walk over the waves, nothing shared with the workers
            // dense_level < 0: no level small enough (the host guessed wrong) -- the workers add everything, the owners only
            // keep the wave barriers company
            const int dense_level = s_dense_level;
            const Level dlv = s_lv[dense_level < 0 ? 0 : dense_level];
            for (int wave = 0; wave < ws.waves; ++wave) {
                int t_begin, t_end;
                wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);
                if (dense_level >= 0)
                    dense_owner_wave<BORDER, DPF>(a, dlv, dense_level, t_begin, t_end, tiles_per_bh, warp - nworkers, nown, lane,
                                             align, gout, gimg);
                wave_pace_cta(ws, wave);
            }
            return;
        }
    }
    for (int wave = 0; wave < ws.waves; ++wave) {
    int t_begin, t_end;
    wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);

    int tile = t_begin + warp;
    if (tile < t_end) {

    TileUnit tu = decode_tile(tile, tiles_per_bh, g, G, a, subs, LK);
    LaneOperands<T, PPL, FUSED> op;
    float go[VEC];
    load_operands<T, LANES, LK, FUSED, PADDED>(a, tu, j, op);
    load_vec_stream<T, VEC>(gout + (size_t)tu.u * a.D + j * VEC, go);

    for (; tile < t_end; tile += nworkers) {
        const int tile_n = tile + nworkers;
        const bool has_next = tile_n < t_end;
        const TileUnit tu_n = decode_tile(has_next ? tile_n : tile, tiles_per_bh, g, G, a, subs, LK);
        LaneOperands<T, PPL, FUSED> op_n;
        float go_n[VEC];
        load_operands<T, LANES, LK, FUSED, PADDED>(a, tu_n, j, op_n);
        load_vec_stream<T, VEC>(gout + (size_t)tu_n.u * a.D + j * VEC, go_n);
        if constexpr (FUSED) derive_operands<T, LANES, LK>(a, s_lv, j, op);

        const unsigned char *__restrict__ lane_base =
            reinterpret_cast<const unsigned char *>(img + tu.bh_off + j * VEC);
        // fp32 accumulation row of this (b,h): for 16-bit storage (VEC == 8) the row is kept in the PERMUTED channel
        // order of accum_position() so that each red.v4 instruction of a lane group covers whole 32-byte sectors
        unsigned char *__restrict__ gimg_base = reinterpret_cast<unsigned char *>(gimg + tu.bh_off + j * 4);
        // padding queries of the last tile shadow a real query: they gather like it but add nothing to grad_img
        const bool live = tu.live;
        const int p0 = SPLIT ? tu.p0 : 0;   // first point of this tile's sub-unit

        TileTap tap[PPL];
        float sx[PPL], sy[PPL];
        float quantum[QUANT ? PPL : 1];   // QUANT: quantum of the level of this lane's points (0 = leave unquantised)
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) {
            const int lvl = slot_level(p0 + j * PPL + pp, a);
            const Level lv = s_lv[lvl];
            tap[pp] = resolve_tap<BORDER>(op.xy[2 * pp], op.xy[2 * pp + 1], lv, align, row_bytes);
            sx[pp] = align ? (float)(lv.w - 1) : (float)lv.w;
            sy[pp] = align ? (float)(lv.h - 1) : (float)lv.h;
            if constexpr (QUANT) quantum[pp] = row_quantum(a, (tile / tiles_per_bh) * a.L + lvl);
            // DENSE: the owner warps add this level's rows; the flag travels with the tap (a mask of point slots tested
            // per corner was hoisted out of the tile loop by the compiler as 16 registers, and spilled)
            if constexpr (DENSE) tap[pp].pack |= lv.pad ? kPackDenseBit : 0u;
        }

        // AGG: bit (8 g + jj) of pair_mask[pp] = point jj*PPL+pp of group g's unit shares its cell with the partner
        // group's (g ^ 1) same point; partner_go = the partner unit's grad_out slice of this lane's channels
        unsigned pair_mask[AGG ? PPL : 1];
        float partner_go[AGG ? VEC : 1];
        if constexpr (AGG) {
            unsigned any_pair = 0u;
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                // padding queries (and launches without grad_img) never pair: bit 31 of the pack is otherwise unused
                const unsigned key = (live && need_img) ? tap[pp].pack : (0x80000000u | (unsigned)lane);
                const unsigned p_off = __shfl_xor_sync(0xffffffffu, tap[pp].off, LANES);
                const unsigned p_key = __shfl_xor_sync(0xffffffffu, key, LANES);
                pair_mask[pp] = __ballot_sync(0xffffffffu, p_off == tap[pp].off && p_key == key);
                any_pair |= pair_mask[pp];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) partner_go[e] = 0.0f;
            if (any_pair) {   // warp-uniform
#pragma unroll
                for (int e = 0; e < VEC; ++e) partner_go[e] = __shfl_xor_sync(0xffffffffu, go[e], LANES);
            }
        }

        // part[(jj*PPL + pp)*3 + {0,1,2}] : point jj*PPL+pp  ->  {grad weight, d/dx, d/dy} partial over my channels
        float part[3 * LK];

        // The unit's points are processed in BATCHES of NB points (batch bi: point slot pp of the source lanes jj0 ..
        // jj0 + NB - 1).  A batch is two dependent trips through the LSU -- the tap exchange (shuffles), then the gathers
        // the exchanged offsets address -- and under the row adds each trip takes ~1.2k clk, so a warp tile of 8 batches was
        // 16 trips = the ~20k clk tile period measured (time ~ 1 / warps below 12 warps).  PIPE: the exchange of batch
        // bi + 1 is issued BEFORE batch bi is consumed: one trip per batch on the critical path (costs NB x 10 registers).
        struct Exchanged {
            float fx[NB], fy[NB], fw[NB];
            float fq[QUANT ? NB : 1];
            unsigned o[NB][4];
            unsigned msk[NB];
            unsigned dense_flag[DENSE ? NB : 1];
            float pw[AGG ? NB : 1], pdx[AGG ? NB : 1], pdy[AGG ? NB : 1];   // AGG: the partner's point
            bool add_partner[AGG ? NB : 1], leave_to_partner[AGG ? NB : 1];
        };
        constexpr int kBatches = PPL * (LANES / NB);
        auto exchange = [&](const int pp, const int jj0) {
            Exchanged x;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int src = jj0 + n;
                if constexpr (AGG) {
                    x.add_partner[n] = x.leave_to_partner[n] = false;
                    x.pw[n] = x.pdx[n] = x.pdy[n] = 0.0f;
                    const unsigned groups = (pair_mask[pp] >> src) & 0x01010101u;   // bit 8g: group g pairs on this point
                    if (groups) {   // warp-uniform
                        const int partner_lane = ((g ^ 1) * LANES) + src;
                        x.pw[n] = __shfl_sync(0xffffffffu, op.wa[pp], partner_lane);
                        x.pdx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, partner_lane);
                        x.pdy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, partner_lane);
                        const bool paired = (groups >> (8 * g)) & 1u;
                        x.add_partner[n] = paired && !(g & 1);
                        x.leave_to_partner[n] = paired && (g & 1);
                    }
                }
                // dead slots of a padded instantiation carry point (0,0) with weight 0: they are gathered like any
                // other (one in-range row, L1 hits) so that the NB x 4 loads stay one branch-free batch, and only
                // their row adds are skipped
                const unsigned off = __shfl_sync(0xffffffffu, tap[pp].off, src, LANES);
                const unsigned pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                x.fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                x.fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                x.fw[n] = __shfl_sync(0xffffffffu, op.wa[pp], src, LANES);
                if constexpr (QUANT) x.fq[n] = __shfl_sync(0xffffffffu, quantum[pp], src, LANES);
                corner_offsets(off, pack, row_bytes, x.o[n]);
                x.msk[n] = BORDER ? 0xFu : ((pack >> kPackMaskShift) & 0xFu);
                if constexpr (DENSE) x.dense_flag[n] = pack;
            }
            return x;
        };

        Exchanged cur = exchange(0, 0);
#pragma unroll
        for (int bi = 0; bi < kBatches; ++bi) {
            {
                const int pp = bi / (LANES / NB), jj0 = (bi % (LANES / NB)) * NB;
                Raw raw[NB][4];
                // always in range (clamped rows); zeros padding is applied to the dot products below
                // (no-allocate gathers of the levels that cannot stay in L1, the forward's -15 %, measured neutral to
                // slightly negative here: the backward is bound by its row adds, which never allocate in L1)
#pragma unroll
                for (int n = 0; n < NB; ++n) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        // NA: the first NA point slots (the finest levels) are gathered without allocating in L1
                        if ((jj0 + n) * PPL + pp < NA)
                            raw[n][c] = gather_slice_na<VEC * (int)sizeof(T)>(lane_base, cur.o[n][c]);
                        else
                            raw[n][c] = gather_slice<VEC * (int)sizeof(T)>(lane_base, cur.o[n][c]);
                    }
                }
                Exchanged nxt = cur;
                if constexpr (PIPE) {
                    if (bi + 1 < kBatches) nxt = exchange((bi + 1) / (LANES / NB), ((bi + 1) % (LANES / NB)) * NB);
                }
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    const bool alive = !PADDED || p0 + (jj0 + n) * PPL + pp < a.LK;
                    const float dx = cur.fx[n], dy = cur.fy[n];
                    float bw[4];  // bilinear weights of corners 00, 01, 10, 11
                    bw[1] = (1.0f - dy) * dx;
                    bw[0] = (1.0f - dy) - bw[1];
                    bw[3] = dy * dx;
                    bw[2] = dy - bw[3];
                    float d[4];
                    float pbw[AGG ? 4 : 1];   // AGG: the partner's corner weights (attention x bilinear)
                    if constexpr (AGG) {
                        const float b1 = (1.0f - cur.pdy[n]) * cur.pdx[n], b3 = cur.pdy[n] * cur.pdx[n];
                        pbw[0] = cur.pw[n] * ((1.0f - cur.pdy[n]) - b1);
                        pbw[1] = cur.pw[n] * b1;
                        pbw[2] = cur.pw[n] * (cur.pdy[n] - b3);
                        pbw[3] = cur.pw[n] * b3;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v[VEC];
                        widen_row<T, VEC>(raw[n][c], v);
                        float acc = 0.0f;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc = fmaf(go[e], v[e], acc);
                        d[c] = (BORDER || ((cur.msk[n] >> c) & 1u)) ? acc : 0.0f;
                        if (need_img) {
                            const float s = cur.fw[n] * bw[c];
                            float gv[VEC];
                            if constexpr (AGG) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) gv[e] = go[e] * s;
                                if (cur.add_partner[n]) {
#pragma unroll
                                    for (int e = 0; e < VEC; ++e) gv[e] = fmaf(partner_go[e], pbw[c], gv[e]);
                                }
                            } else if constexpr (!QUANT) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) gv[e] = go[e] * s;
                            } else {
                                // |go * s| < 2^24 q, so |go * s| + 1.5 * 2^23 q lies in [1.5, 3.5) * 2^23 q, where fp32 has a spacing
                                // of q (or 2q):
                                // the FMA rounds the exact product to a multiple of q, the subtraction is exact, the sign
                                // goes back on with one logic op.  fq == 0 (nothing to quantise against): plain product.
                                const float magic = cur.fq[n] * 12582912.0f;   // 1.5 * 2^23
                                const unsigned s_sign = __float_as_uint(s) & 0x80000000u;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    const float mag = fmaf(fabsf(go[e]), fabsf(s), magic) - magic;
                                    gv[e] = __uint_as_float(__float_as_uint(mag) |
                                                            ((__float_as_uint(go[e]) & 0x80000000u) ^ s_sign));
                                }
                            }
                            float *dst = reinterpret_cast<float *>(gimg_base + (size_t)cur.o[n][c] * kAccScale);
                            bool add = live && alive && (BORDER || ((cur.msk[n] >> c) & 1u));
                            if constexpr (AGG) add = add && !cur.leave_to_partner[n];
                            if constexpr (DENSE) add = add && !(cur.dense_flag[n] & kPackDenseBit);
                            if (add) red_add_row<VEC, LANES>(dst, gv);
                        }
                    }
                    const int pidx = (jj0 + n) * PPL + pp;
                    part[3 * pidx + 0] = bw[0] * d[0] + bw[1] * d[1] + bw[2] * d[2] + bw[3] * d[3];
                    part[3 * pidx + 1] = (1.0f - dy) * (d[1] - d[0]) + dy * (d[3] - d[2]);
                    part[3 * pidx + 2] = (1.0f - dx) * (d[2] - d[0]) + dx * (d[3] - d[1]);
                }
                if constexpr (!PIPE) {
                    if (bi + 1 < kBatches) nxt = exchange((bi + 1) / (LANES / NB), ((bi + 1) % (LANES / NB)) * NB);
                }
                cur = nxt;
            }
        }

        // ---- reduce over the LANES lanes; lane j ends with points [j*PPL, (j+1)*PPL) in part[0 .. 3*PPL) ----
        transpose_reduce<3 * LK, LANES / 2>(part, j);

        if constexpr (!FUSED) {
            if (tu.live) {
                T *__restrict__ gaw_u = gaw + (size_t)tu.u * a.LK;
                T *__restrict__ gpts_u = gpts + (size_t)tu.u * a.LK * 2;
                if constexpr (!PADDED) {
                    if (need_aw) {
                        float gw[PPL];
#pragma unroll
                        for (int pp = 0; pp < PPL; ++pp) gw[pp] = part[3 * pp + 0];
                        store_vec_stream<T, PPL>(gaw_u + p0 + j * PPL, gw);
                    }
                    if (need_pts) {
                        float gp[2 * PPL];
#pragma unroll
                        for (int pp = 0; pp < PPL; ++pp) {
                            gp[2 * pp + 0] = part[3 * pp + 1] * (op.wa[pp] * sx[pp]);
                            gp[2 * pp + 1] = part[3 * pp + 2] * (op.wa[pp] * sy[pp]);
                        }
                        store_vec_stream<T, 2 * PPL>(gpts_u + (p0 + j * PPL) * 2, gp);
                    }
                } else {
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) {
                        const int p = p0 + j * PPL + pp;
                        if (p < a.LK) {
                            if (need_aw) {
                                const float gw1[1] = {part[3 * pp + 0]};
                                store_vec_stream<T, 1>(gaw_u + p, gw1);
                            }
                            if (need_pts) {
                                const float gp2[2] = {part[3 * pp + 1] * (op.wa[pp] * sx[pp]),
                                                      part[3 * pp + 2] * (op.wa[pp] * sy[pp])};
                                store_vec_stream<T, 2>(gpts_u + 2 * p, gp2);
                            }
                        }
                    }
                }
            }
        } else {
            // ---- module-core epilogue (frontend.py:253-284 differentiated) ----
            //   logit:  g = w * (gw - sum_p w_p gw_p)                                  (softmax backward)
            //   offset: 2-d ref: g = gpoint / (h | w of the level);  4-d ref: g = gpoint * ref_wh / (2K)
            //   ref:    2-d: sum_p gpoint;  4-d: additionally sum_p gpoint * offset / (2K) for (w, h)
            float gx[PPL], gy[PPL], dot = 0.0f;
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                gx[pp] = part[3 * pp + 1] * (op.wa[pp] * sx[pp]);
                gy[pp] = part[3 * pp + 2] * (op.wa[pp] * sy[pp]);
                dot = fmaf(op.wa[pp], part[3 * pp + 0], dot);
            }
            float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
            const float inv_2k = 1.0f / (float)(2 * a.K);
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                r0 += gx[pp];
                r1 += gy[pp];
                r2 = fmaf(gx[pp], op.raw[3 * pp + 0] * inv_2k, r2);
                r3 = fmaf(gy[pp], op.raw[3 * pp + 1] * inv_2k, r3);
            }
#pragma unroll
            for (int s = LANES / 2; s > 0; s >>= 1) {
                dot += __shfl_xor_sync(0xffffffffu, dot, s);
                r0 += __shfl_xor_sync(0xffffffffu, r0, s);
                r1 += __shfl_xor_sync(0xffffffffu, r1, s);
                r2 += __shfl_xor_sync(0xffffffffu, r2, s);
                r3 += __shfl_xor_sync(0xffffffffu, r3, s);
            }
            if (tu.live) {
                if (need_pts || need_aw) {
                    float gtriple[3 * PPL];
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) {
                        const Level lv = s_lv[(j * PPL + pp) / a.K];
                        if (a.ref_dim == 2) {
                            gtriple[3 * pp + 0] = gx[pp] / (float)lv.h;
                            gtriple[3 * pp + 1] = gy[pp] / (float)lv.w;
                        } else {
                            gtriple[3 * pp + 0] = gx[pp] * (op.ref[2] * inv_2k);
                            gtriple[3 * pp + 1] = gy[pp] * (op.ref[3] * inv_2k);
                        }
                        gtriple[3 * pp + 2] = op.wa[pp] * (part[3 * pp + 0] - dot);
                    }
                    constexpr int E8 = FusedChunk<T, PPL>::kElems;
                    T *__restrict__ gproj = static_cast<T *>(a.gproj) + ((size_t)tu.u * LK + j * PPL) * 3;
#pragma unroll
                    for (int c = 0; c < 3 * PPL / E8; ++c) {
                        float tmp[E8];
#pragma unroll
                        for (int e = 0; e < E8; ++e) tmp[e] = gtriple[c * E8 + e];
                        store_vec_stream<T, E8>(gproj + c * E8, tmp);
                    }
                }
                if ((a.flags & kNeedRef) && j == 0) {
                    float *gr = a.gref + (size_t)(tu.u / a.H) * a.ref_dim;
                    red_add_v1(gr + 0, r0);
                    red_add_v1(gr + 1, r1);
                    if (a.ref_dim == 4) {
                        red_add_v1(gr + 2, r2);
                        red_add_v1(gr + 3, r3);
                    }
                }
            }
        }

        tu = tu_n;
        op = op_n;
#pragma unroll
        for (int i = 0; i < VEC; ++i) go[i] = go_n[i];
    }
    }
    wave_pace_cta(ws, wave);
    }  // waves
}


template <typename T, int LANES, int LK, bool FUSED = false, int VEC = 16 / (int)sizeof(T), bool PADDED = false,
          bool SPLIT = false, bool QUANT = false, bool AGG = false, int NOWN = 0, int DPF = 2, int THREADS = 512,
          int NB = 2, bool PIPE = false, int NA = 0>
static cudaError_t launch_tiled_t(const KernelArgs &a, int sm_count, cudaStream_t st, int subs = 1) {
    constexpr int G = TiledCfg<T, LANES, LK>::G;
    if (!tiled_offsets_fit(a, sizeof(T), subs)) return cudaErrorNotSupported;
    const int tiles_per_bh = subs * ((a.Q + G - 1) / G);
    const int total_tiles = a.B * a.H * tiles_per_bh;
    const int warps = THREADS / 32 - NOWN;   // warps that take warp tiles
    const int want = (total_tiles + warps - 1) / warps;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    // one wave keeps its pyramid slices AND (when grad_img is produced) the fp32 grad_img slices in L2
    const size_t per_slice_factor = (a.flags & kNeedImg) ? sizeof(T) + sizeof(float) : sizeof(T);
    WaveSchedule ws = make_wave_schedule(a, tiles_per_bh, per_slice_factor, kBwdL2Budget);
    // many waves of substantial size: keep the persistent CTAs on the same wave (wave_pace).  Waves with only a few
    // tiles per warp (decoder: 900 queries against a 22k-pixel pyramid) cannot drift far and would only pay the
    // per-wave handshake (measured on that shape: module step 0.93 -> 1.14 ms when paced).
    const bool big_waves = (long long)ws.slices_per_wave * tiles_per_bh >= 4LL * warps * grid;
    if (ws.waves > 1 && grid == sm_count && (big_waves || pacing_forced())) {
        const cudaError_t e = acquire_pace_counter(st, &ws.pace);
        if (e != cudaSuccess) return e;
    }
    if (tuning().carveout >= 0) {   // experiment knob: shared-memory carve-out (percent) = what is left for L1
        cudaFuncSetAttribute(
            msda_bwd_tiled_kernel<T, LANES, LK, true, NB, THREADS, FUSED, VEC, PADDED, SPLIT, QUANT, AGG, NOWN, DPF, PIPE, NA>,
            cudaFuncAttributePreferredSharedMemoryCarveout, tuning().carveout);
        cudaFuncSetAttribute(
            msda_bwd_tiled_kernel<T, LANES, LK, false, NB, THREADS, FUSED, VEC, PADDED, SPLIT, QUANT, AGG, NOWN, DPF, PIPE, NA>,
            cudaFuncAttributePreferredSharedMemoryCarveout, tuning().carveout);
    }
    if (a.border)
        msda_bwd_tiled_kernel<T, LANES, LK, true, NB, THREADS, FUSED, VEC, PADDED, SPLIT, QUANT, AGG, NOWN, DPF, PIPE, NA>
            <<<grid, THREADS, 0, st>>>(a, ws, subs);
    else
        msda_bwd_tiled_kernel<T, LANES, LK, false, NB, THREADS, FUSED, VEC, PADDED, SPLIT, QUANT, AGG, NOWN, DPF, PIPE, NA>
            <<<grid, THREADS, 0, st>>>(a, ws, subs);
    return cudaGetLastError();
}


}  // namespace msda
