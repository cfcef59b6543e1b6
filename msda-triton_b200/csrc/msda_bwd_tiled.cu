// msda_bwd_tiled.cu -- tuned backward kernel: same persistent (b,h)-major schedule and lane layout as the tuned
// forward (msda_tiled.cuh).
//
// Per sampling point every lane forms, over its VEC channels, the four corner dot products <go, v_c>; from them the
// three per-point partials (grad weight, d/dx, d/dy).  The 3*LK partials of a unit live in registers until the
// end of the unit and are then reduced across the LANES lanes with a TRANSPOSING butterfly (each step halves the
// values a lane keeps), which costs 3*LK*(1 - 1/LANES) shuffles instead of 3*LK*log2(LANES) and leaves lane j
// holding exactly the PPL points it loaded -- so the grad_points / grad_weights stores are the same coalesced
// vector stores as the loads.  grad_img goes out as one REDG.E.ADD.F32x4 per lane per valid corner.
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"

namespace msda {

constexpr int kNeedImg = 1, kNeedPts = 2, kNeedAw = 4;
constexpr size_t kBwdL2Budget = 48u << 20;  // img + grad_img bytes of one wave of (b,h) slices

template <int N, int STEP> __device__ __forceinline__ void transpose_reduce(float (&part)[N], const int j) {
    // lanes with bit STEP clear keep the lower half, their partners (j ^ STEP) the upper half
    constexpr int HALF = N / 2;
    const bool upper = (j & STEP) != 0;
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
        const float keep = upper ? part[k + HALF] : part[k];
        const float send = upper ? part[k] : part[k + HALF];
        part[k] = keep + __shfl_xor_sync(0xffffffffu, send, STEP);
    }
    if constexpr (STEP > 1) {
        float(&lower)[HALF] = reinterpret_cast<float(&)[HALF]>(part);
        transpose_reduce<HALF, STEP / 2>(lower, j);
    }
}

template <typename T, int LANES, int LK, bool BORDER, int NB, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    msda_bwd_tiled_kernel(const KernelArgs a, const WaveSchedule ws) {
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int VEC = Cfg::VEC, G = Cfg::G, PPL = Cfg::PPL;
    static_assert(LANES % NB == 0, "batch must divide the group");

    __shared__ Level s_lv[LK];
    build_level_table(s_lv, a.shapes, a.L);

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = THREADS >> 5;
    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const bool need_img = (a.flags & kNeedImg) != 0, need_pts = (a.flags & kNeedPts) != 0,
               need_aw = (a.flags & kNeedAw) != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);
    constexpr unsigned kAccScale = sizeof(float) * VEC / 16;  // accumulation row bytes / storage row bytes

    const int tiles_per_bh = ws.tiles_per_bh;
    for (int wave = 0; wave < ws.waves; ++wave) {
    int t_begin, t_end;
    wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);

    int tile = t_begin + warp;
    if (tile >= t_end) continue;

    TileUnit tu = decode_tile(tile, tiles_per_bh, g, G, a);
    float xy[2 * PPL], wa[PPL], go[VEC];
    load_vec_stream<T, 2 * PPL>(pts + ((size_t)tu.u * LK + j * PPL) * 2, xy);
    load_vec_stream<T, PPL>(aw + (size_t)tu.u * LK + j * PPL, wa);
    load_vec_stream<T, VEC>(gout + (size_t)tu.u * a.D + j * VEC, go);

    for (; tile < t_end; tile += nwarps) {
        const int tile_n = tile + nwarps;
        const bool has_next = tile_n < t_end;
        const TileUnit tu_n = decode_tile(has_next ? tile_n : tile, tiles_per_bh, g, G, a);
        float xy_n[2 * PPL], wa_n[PPL], go_n[VEC];
        load_vec_stream<T, 2 * PPL>(pts + ((size_t)tu_n.u * LK + j * PPL) * 2, xy_n);
        load_vec_stream<T, PPL>(aw + (size_t)tu_n.u * LK + j * PPL, wa_n);
        load_vec_stream<T, VEC>(gout + (size_t)tu_n.u * a.D + j * VEC, go_n);

        const unsigned char *__restrict__ lane_base =
            reinterpret_cast<const unsigned char *>(img + tu.bh_off + j * VEC);
        unsigned char *__restrict__ gimg_base = reinterpret_cast<unsigned char *>(gimg + tu.bh_off + j * VEC);
        // padding queries of the last tile shadow a real query: their image contributions are scaled to zero
        const float live_scale = tu.live ? 1.0f : 0.0f;

        TileTap tap[PPL];
        float sx[PPL], sy[PPL];
#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) {
            const Level lv = s_lv[(j * PPL + pp) / a.K];
            tap[pp] = resolve_tap<BORDER>(xy[2 * pp], xy[2 * pp + 1], lv, align, row_bytes);
            sx[pp] = align ? (float)(lv.w - 1) : (float)lv.w;
            sy[pp] = align ? (float)(lv.h - 1) : (float)lv.h;
        }

        // part[(jj*PPL + pp)*3 + {0,1,2}] : point jj*PPL+pp  ->  {grad weight, d/dx, d/dy} partial over my channels
        float part[3 * LK];

#pragma unroll
        for (int pp = 0; pp < PPL; ++pp) {
#pragma unroll
            for (int jj0 = 0; jj0 < LANES; jj0 += NB) {
                uint4 raw[NB][4];
                float fx[NB], fy[NB], fw[NB];
                unsigned o[NB][4];
                unsigned msk[NB];
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    const int src = jj0 + n;
                    const unsigned off = __shfl_sync(0xffffffffu, tap[pp].off, src, LANES);
                    const unsigned pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                    fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                    fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                    fw[n] = __shfl_sync(0xffffffffu, wa[pp], src, LANES) * live_scale;
                    corner_offsets(off, pack, row_bytes, o[n]);
                    msk[n] = BORDER ? 0xFu : ((pack >> kPackMaskShift) & 0xFu);
                    // always in range (clamped rows); zeros padding is applied to the dot products below
#pragma unroll
                    for (int c = 0; c < 4; ++c) raw[n][c] = gather_row(lane_base, o[n][c]);
                }
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    const float dx = fx[n], dy = fy[n];
                    float bw[4];  // bilinear weights of corners 00, 01, 10, 11
                    bw[1] = (1.0f - dy) * dx;
                    bw[0] = (1.0f - dy) - bw[1];
                    bw[3] = dy * dx;
                    bw[2] = dy - bw[3];
                    float d[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v[VEC];
                        widen_row<T, VEC>(raw[n][c], v);
                        float acc = 0.0f;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc = fmaf(go[e], v[e], acc);
                        d[c] = (BORDER || ((msk[n] >> c) & 1u)) ? acc : 0.0f;
                        if (need_img) {
                            const float s = fw[n] * bw[c];
                            float gv[VEC];
#pragma unroll
                            for (int e = 0; e < VEC; ++e) gv[e] = go[e] * s;
                            float *dst = reinterpret_cast<float *>(gimg_base + (size_t)o[n][c] * kAccScale);
                            if (BORDER || ((msk[n] >> c) & 1u)) red_add_vec<VEC>(dst, gv);
                        }
                    }
                    const int pidx = (jj0 + n) * PPL + pp;
                    part[3 * pidx + 0] = bw[0] * d[0] + bw[1] * d[1] + bw[2] * d[2] + bw[3] * d[3];
                    part[3 * pidx + 1] = (1.0f - dy) * (d[1] - d[0]) + dy * (d[3] - d[2]);
                    part[3 * pidx + 2] = (1.0f - dx) * (d[2] - d[0]) + dx * (d[3] - d[1]);
                }
            }
        }

        // ---- reduce over the LANES lanes; lane j ends with points [j*PPL, (j+1)*PPL) in part[0 .. 3*PPL) ----
        transpose_reduce<3 * LK, LANES / 2>(part, j);

        if (tu.live) {
            if (need_aw) {
                float gw[PPL];
#pragma unroll
                for (int pp = 0; pp < PPL; ++pp) gw[pp] = part[3 * pp + 0];
                store_vec_stream<T, PPL>(gaw + (size_t)tu.u * LK + j * PPL, gw);
            }
            if (need_pts) {
                float gp[2 * PPL];
#pragma unroll
                for (int pp = 0; pp < PPL; ++pp) {
                    gp[2 * pp + 0] = part[3 * pp + 1] * (wa[pp] * sx[pp]);
                    gp[2 * pp + 1] = part[3 * pp + 2] * (wa[pp] * sy[pp]);
                }
                store_vec_stream<T, 2 * PPL>(gpts + ((size_t)tu.u * LK + j * PPL) * 2, gp);
            }
        }

        tu = tu_n;
#pragma unroll
        for (int i = 0; i < 2 * PPL; ++i) xy[i] = xy_n[i];
#pragma unroll
        for (int i = 0; i < PPL; ++i) wa[i] = wa_n[i];
#pragma unroll
        for (int i = 0; i < VEC; ++i) go[i] = go_n[i];
    }
    }  // waves
}


// ---------------------------------------------------------------------------------------------------------------
// Binned variant: in-CTA segmented reduction of grad_img for the coarse levels.
//
// ncu on the plain kernel above: the backward is bound by the SM->L2 request port (1 sector per cycle per SM) that
// carries the `red` sectors -- 64 row-adds of 128 B per unit, 2.62 GB for the benchmark shape.  Most of them go to
// the few rows of the coarse levels (the 8x8 level of one (b,h) slice has 64 rows and receives 25 % of all adds).
// Shared-memory fp32 atomics are no help (CAS loops, slower than global `red` on sm_100a: scripts/micro/), so the
// coarse levels are reduced WITHOUT atomics on floats:
//   * the CTA walks its contiguous (b,h)-major range in super-tiles of TQ queries;
//   * phase A (per unit, as before): gathers, the three per-point partials, direct `red` for the fine levels; for
//     every corner that lands in a "binned" level the lane that owns the point pushes a record onto a per-row
//     linked list in shared memory (one native 32-bit ATOMS.EXCH on the row's head + the weight and next index in
//     natural order, so the query index is implied by the record index); the unit's grad_out row is parked in smem;
//   * phase C: one lane group per destination row walks that row's list, accumulates weight * grad_out[q] in
//     registers and issues ONE `red.v4` per lane for the whole super-tile.
// Binned levels = the longest suffix of the pyramid (coarsest first) with at most MAXROWS rows, decided on device
// from img_shapes.  For the benchmark pyramid that is levels 1-3 (1344 rows): 12 of 16 points.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int LANES, int LK, bool BORDER, int NB, int THREADS, int ROUNDS, int MAXROWS>
__global__ void __launch_bounds__(THREADS, 1)
    msda_bwd_binned_kernel(const KernelArgs a, const int stiles_per_bh, const int total_stiles) {
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int VEC = Cfg::VEC, G = Cfg::G, PPL = Cfg::PPL;
    constexpr int NW = THREADS / 32, TQ = NW * G * ROUNDS, NGROUPS = THREADS / LANES;
    constexpr int DCH = LANES * VEC;                       // channels per row (== D)
    constexpr unsigned END = 0xFFFFu;
    static_assert(TQ * LK * 4 < 0xFFFF, "record index must fit in 16 bits");
    static_assert(LANES % NB == 0, "batch must divide the group");

    extern __shared__ __align__(16) unsigned char s_dyn[];
    float *s_w = reinterpret_cast<float *>(s_dyn);                                 // [4][TQ][LK] record weights
    float *s_go = s_w + 4 * TQ * LK;                                               // [TQ][DCH]   grad_out rows
    unsigned *s_head = reinterpret_cast<unsigned *>(s_go + TQ * DCH);              // [MAXROWS]   list heads
    unsigned short *s_next = reinterpret_cast<unsigned short *>(s_head + MAXROWS); // [4][TQ][LK] next record
    __shared__ Level s_lv[LK];
    __shared__ int s_bin[3];  // first binned point, first binned pixel row, number of binned rows

    build_level_table(s_lv, a.shapes, a.L);
    if (threadIdx.x == 0) {
        int rows = 0, l0 = a.L;
        for (int l = a.L - 1; l >= 0; --l) {
            const int n = s_lv[l].h * s_lv[l].w;
            if (rows + n > MAXROWS) break;
            rows += n;
            l0 = l;
        }
        s_bin[0] = l0 * a.K;
        s_bin[1] = l0 < a.L ? s_lv[l0].off : a.Npix;
        s_bin[2] = rows;
    }
    for (int i = threadIdx.x; i < MAXROWS; i += THREADS) s_head[i] = END;
    __syncthreads();
    const int p0 = s_bin[0], base_row = s_bin[1], nrows = s_bin[2];

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const bool need_pts = (a.flags & kNeedPts) != 0, need_aw = (a.flags & kNeedAw) != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);
    constexpr unsigned kAccScale = sizeof(float) * VEC / 16;

    const int st_begin = (int)((long long)total_stiles * blockIdx.x / gridDim.x);
    const int st_end = (int)((long long)total_stiles * (blockIdx.x + 1) / gridDim.x);
    if (st_begin >= st_end) return;

    // flattened (super-tile, round) index: one warp tile = G consecutive queries
    auto decode = [&](int st, int r, int &q_local) -> TileUnit {
        const int bh = st / stiles_per_bh;
        const int qs = (st - bh * stiles_per_bh) * TQ;
        q_local = (r * NW + warp) * G + g;
        const int b = bh / a.H, h = bh - b * a.H;
        const int q_raw = qs + q_local;
        TileUnit t;
        t.live = q_raw < a.Q;
        const int q = t.live ? q_raw : a.Q - 1;
        t.u = ((long long)b * a.Q + q) * a.H + h;
        t.bh_off = ((size_t)b * a.Npix * a.H + h) * a.D;
        return t;
    };

    int q_local;
    TileUnit tu = decode(st_begin, 0, q_local);
    float xy[2 * PPL], wa[PPL], go[VEC];
    load_vec_stream<T, 2 * PPL>(pts + ((size_t)tu.u * LK + j * PPL) * 2, xy);
    load_vec_stream<T, PPL>(aw + (size_t)tu.u * LK + j * PPL, wa);
    load_vec_stream<T, VEC>(gout + (size_t)tu.u * a.D + j * VEC, go);

    for (int st = st_begin; st < st_end; ++st) {
        size_t st_bh_off = 0;
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            // ---- prefetch the next warp tile ----
            int st_n = st, r_n = r + 1;
            if (r_n == ROUNDS) { r_n = 0; st_n = st + 1; }
            if (st_n >= st_end) { st_n = st; r_n = r; }
            int q_local_n;
            const TileUnit tu_n = decode(st_n, r_n, q_local_n);
            float xy_n[2 * PPL], wa_n[PPL], go_n[VEC];
            load_vec_stream<T, 2 * PPL>(pts + ((size_t)tu_n.u * LK + j * PPL) * 2, xy_n);
            load_vec_stream<T, PPL>(aw + (size_t)tu_n.u * LK + j * PPL, wa_n);
            load_vec_stream<T, VEC>(gout + (size_t)tu_n.u * a.D + j * VEC, go_n);

            st_bh_off = tu.bh_off;
            const unsigned char *__restrict__ lane_base =
                reinterpret_cast<const unsigned char *>(img + tu.bh_off + j * VEC);
            unsigned char *__restrict__ gimg_base = reinterpret_cast<unsigned char *>(gimg + tu.bh_off + j * VEC);
            const float live_scale = tu.live ? 1.0f : 0.0f;

            // park this unit's grad_out row for phase C
            *reinterpret_cast<Pack<float, VEC> *>(s_go + q_local * DCH + j * VEC) =
                *reinterpret_cast<const Pack<float, VEC> *>(go);

            TileTap tap[PPL];
            float sx[PPL], sy[PPL];
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                const int p = j * PPL + pp;
                const Level lv = s_lv[p / a.K];
                const Tap<float> t = locate<float>(xy[2 * pp], xy[2 * pp + 1], lv, BORDER, align);
                const unsigned step_rows = (unsigned)t.pack & (unsigned)kPackDyMask;
                tap[pp].off = (unsigned)t.row00 * row_bytes;
                tap[pp].pack = ((step_rows * row_bytes) >> 4) | ((unsigned)t.pack & ~(unsigned)kPackDyMask);
                tap[pp].dx = t.dx;
                tap[pp].dy = t.dy;
                sx[pp] = align ? (float)(lv.w - 1) : (float)lv.w;
                sy[pp] = align ? (float)(lv.h - 1) : (float)lv.h;
                // ---- records for the binned levels (owner lane only) ----
                if (p >= p0 && tu.live) {
                    const int step_x = (t.pack >> kPackDxBit) & 1;
                    const unsigned mask = (unsigned)(t.pack >> kPackMaskShift) & 0xFu;
                    const int r00 = t.row00 - base_row;
                    const int rows4[4] = {r00, r00 + step_x, r00 + (int)step_rows, r00 + (int)step_rows + step_x};
                    const float wy1 = wa[pp] * t.dy, wy0 = wa[pp] - wy1;
                    float w4[4];
                    w4[1] = wy0 * t.dx;
                    w4[0] = wy0 - w4[1];
                    w4[3] = wy1 * t.dx;
                    w4[2] = wy1 - w4[3];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (BORDER || ((mask >> c) & 1u)) {
                            const unsigned idx = (unsigned)((c * TQ + q_local) * LK + p);
                            const unsigned prev = atomicExch(&s_head[rows4[c]], idx);
                            s_next[idx] = (unsigned short)prev;
                            s_w[idx] = w4[c];
                        }
                    }
                }
            }

            float part[3 * LK];
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
#pragma unroll
                for (int jj0 = 0; jj0 < LANES; jj0 += NB) {
                    uint4 raw[NB][4];
                    float fx[NB], fy[NB], fw[NB];
                    unsigned o[NB][4];
                    unsigned msk[NB];
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const int src = jj0 + n;
                        const unsigned off = __shfl_sync(0xffffffffu, tap[pp].off, src, LANES);
                        const unsigned pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                        fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                        fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                        fw[n] = __shfl_sync(0xffffffffu, wa[pp], src, LANES) * live_scale;
                        corner_offsets(off, pack, row_bytes, o[n]);
                        msk[n] = BORDER ? 0xFu : ((pack >> kPackMaskShift) & 0xFu);
#pragma unroll
                        for (int c = 0; c < 4; ++c) raw[n][c] = gather_row(lane_base, o[n][c]);
                    }
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        const float dx = fx[n], dy = fy[n];
                        float bw[4];
                        bw[1] = (1.0f - dy) * dx;
                        bw[0] = (1.0f - dy) - bw[1];
                        bw[3] = dy * dx;
                        bw[2] = dy - bw[3];
                        const int pidx = (jj0 + n) * PPL + pp;
                        const bool direct = pidx < p0;     // warp-uniform: fine levels go straight to L2
                        float d[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float v[VEC];
                            widen_row<T, VEC>(raw[n][c], v);
                            float acc = 0.0f;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) acc = fmaf(go[e], v[e], acc);
                            d[c] = (BORDER || ((msk[n] >> c) & 1u)) ? acc : 0.0f;
                            if (direct) {
                                const float s = fw[n] * bw[c];
                                float gv[VEC];
#pragma unroll
                                for (int e = 0; e < VEC; ++e) gv[e] = go[e] * s;
                                float *dst = reinterpret_cast<float *>(gimg_base + (size_t)o[n][c] * kAccScale);
                                if (BORDER || ((msk[n] >> c) & 1u)) red_add_vec<VEC>(dst, gv);
                            }
                        }
                        part[3 * pidx + 0] = bw[0] * d[0] + bw[1] * d[1] + bw[2] * d[2] + bw[3] * d[3];
                        part[3 * pidx + 1] = (1.0f - dy) * (d[1] - d[0]) + dy * (d[3] - d[2]);
                        part[3 * pidx + 2] = (1.0f - dx) * (d[2] - d[0]) + dx * (d[3] - d[1]);
                    }
                }
            }

            transpose_reduce<3 * LK, LANES / 2>(part, j);

            if (tu.live) {
                if (need_aw) {
                    float gw[PPL];
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) gw[pp] = part[3 * pp + 0];
                    store_vec_stream<T, PPL>(gaw + (size_t)tu.u * LK + j * PPL, gw);
                }
                if (need_pts) {
                    float gp[2 * PPL];
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) {
                        gp[2 * pp + 0] = part[3 * pp + 1] * (wa[pp] * sx[pp]);
                        gp[2 * pp + 1] = part[3 * pp + 2] * (wa[pp] * sy[pp]);
                    }
                    store_vec_stream<T, 2 * PPL>(gpts + ((size_t)tu.u * LK + j * PPL) * 2, gp);
                }
            }

            tu = tu_n;
            q_local = q_local_n;
#pragma unroll
            for (int i = 0; i < 2 * PPL; ++i) xy[i] = xy_n[i];
#pragma unroll
            for (int i = 0; i < PPL; ++i) wa[i] = wa_n[i];
#pragma unroll
            for (int i = 0; i < VEC; ++i) go[i] = go_n[i];
        }

        // ---- phase C: per-row segmented reduction of the binned levels ----
        __syncthreads();
        {
            float *__restrict__ gimg_rows = gimg + st_bh_off + (size_t)base_row * a.H * a.D + j * VEC;
            const unsigned group_mask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << (g * LANES);
            for (int r = threadIdx.x / LANES; r < nrows; r += NGROUPS) {
                // the group leader pops the whole list (read + reset in one lane: no intra-group read/write hazard)
                unsigned idx = END;
                if (j == 0) {
                    idx = s_head[r];
                    s_head[r] = END;
                }
                idx = __shfl_sync(group_mask, idx, g * LANES);
                if (idx != END) {
                    float acc[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
                    do {
                        const float w = s_w[idx];
                        const unsigned nxt = s_next[idx];
                        const unsigned q = (idx / LK) % TQ;
                        const Pack<float, VEC> gq = *reinterpret_cast<const Pack<float, VEC> *>(s_go + q * DCH + j * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w, gq.v[e], acc[e]);
                        idx = nxt;
                    } while (idx != END);
                    red_add_vec<VEC>(gimg_rows + (size_t)r * a.H * a.D, acc);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int LANES, int LK, int ROUNDS, int MAXROWS>
static cudaError_t launch_binned_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    constexpr int THREADS = 512, NB = 2;
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int TQ = (THREADS / 32) * Cfg::G * ROUNDS;
    constexpr size_t kSmem = sizeof(float) * 4 * TQ * LK + sizeof(float) * TQ * LANES * Cfg::VEC +
                             sizeof(unsigned) * MAXROWS + sizeof(unsigned short) * 4 * TQ * LK;
    if (!tiled_offsets_fit(a, sizeof(T))) return cudaErrorNotSupported;
    const int stiles_per_bh = (a.Q + TQ - 1) / TQ;
    const int total_stiles = a.B * a.H * stiles_per_bh;
    const int grid = total_stiles < sm_count ? (total_stiles < 1 ? 1 : total_stiles) : sm_count;
    auto launch = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, THREADS, kSmem, st>>>(a, stiles_per_bh, total_stiles);
        return cudaGetLastError();
    };
    if (a.border) return launch(msda_bwd_binned_kernel<T, LANES, LK, true, NB, THREADS, ROUNDS, MAXROWS>);
    return launch(msda_bwd_binned_kernel<T, LANES, LK, false, NB, THREADS, ROUNDS, MAXROWS>);
}

template <typename T, int LANES, int LK>
static cudaError_t launch_tiled_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    constexpr int THREADS = 512, NB = 2;
    constexpr int G = TiledCfg<T, LANES, LK>::G;
    if (!tiled_offsets_fit(a, sizeof(T))) return cudaErrorNotSupported;
    const int tiles_per_bh = (a.Q + G - 1) / G;
    const int total_tiles = a.B * a.H * tiles_per_bh;
    const int warps = THREADS / 32;
    const int want = (total_tiles + warps - 1) / warps;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    // one wave keeps its pyramid slices AND (when grad_img is produced) the fp32 grad_img slices in L2
    const size_t per_slice_factor = (a.flags & kNeedImg) ? sizeof(T) + sizeof(float) : sizeof(T);
    const WaveSchedule ws = make_wave_schedule(a, tiles_per_bh, per_slice_factor, kBwdL2Budget);
    if (a.border)
        msda_bwd_tiled_kernel<T, LANES, LK, true, NB, THREADS><<<grid, THREADS, 0, st>>>(a, ws);
    else
        msda_bwd_tiled_kernel<T, LANES, LK, false, NB, THREADS><<<grid, THREADS, 0, st>>>(a, ws);
    return cudaGetLastError();
}

// MSDA_B200_BWD_BINNED selects the experimental binned kernels: unset/0 = plain atomics kernel (default),
// 1 = super-tiles of 256 queries binning up to 1408 rows, 2 = 256 queries / 320 rows, 3 = 128 queries / 1408 rows,
// 4 = split backward (K1 = this file's plain kernel without grad_img, K2 = msda_bwd_scatter.cu: 0.19 + 0.29 ms on the
// bench shape versus 0.47 ms fused -- K2's list walk costs ~240 warp instructions per unit).
// Measured on B200 (profiles/r1_ncu_summary.md): `red` sectors drop to 40 %, but the shared-memory footprint
// shrinks L1 (gather hit rate 66 % -> 40 %) and the net effect is within +-8 % of the plain kernel, so it is opt-in.
static int binned_variant() {
    const char *e = std::getenv("MSDA_B200_BWD_BINNED");
    return e && e[0] ? std::atoi(e) : 0;
}

cudaError_t launch_backward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK != 16 || a.L > 16) return cudaErrorNotSupported;
    const int variant = binned_variant();
    if ((a.flags & kNeedImg) && variant == 4 && dtype == 0 && a.D == 32) {
        // split backward: K1 = grad_points / grad_weights (gathers, L1 for the pyramid), K2 = grad_img (no gathers)
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
    if ((a.flags & kNeedImg) && variant != 0 && dtype == 0 && a.D == 32) {
        switch (variant) {
            case 1: return launch_binned_t<float, 8, 16, 4, 1408>(a, sm_count, st);
            case 2: return launch_binned_t<float, 8, 16, 4, 320>(a, sm_count, st);
            case 3: return launch_binned_t<float, 8, 16, 2, 1408>(a, sm_count, st);
        }
    }
    if (dtype == 0) {
        if (a.D == 32) return launch_tiled_t<float, 8, 16>(a, sm_count, st);
        if (a.D == 64) return launch_tiled_t<float, 16, 16>(a, sm_count, st);
    } else if (dtype == 1) {
        if (a.D == 32) return launch_tiled_t<__half, 4, 16>(a, sm_count, st);
        if (a.D == 64) return launch_tiled_t<__half, 8, 16>(a, sm_count, st);
    } else if (dtype == 2) {
        if (a.D == 32) return launch_tiled_t<__nv_bfloat16, 4, 16>(a, sm_count, st);
        if (a.D == 64) return launch_tiled_t<__nv_bfloat16, 8, 16>(a, sm_count, st);
    }
    return cudaErrorNotSupported;
}

}  // namespace msda
