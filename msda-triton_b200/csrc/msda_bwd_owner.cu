// msda_bwd_owner.cu -- tuned backward with the COARSE pyramid levels accumulated in shared memory by an owner warp.
//
// Why.  The tuned backward (msda_bwd_tiled.cu) is bound by the number of 128-byte `red.global.add.v4.f32` row adds an SM
// can inject into L2 (measured ~5.6 clk per row and SM; 64 rows per unit = the same 81.9 M atomic sectors the
// reference issues, kernels.py:550-553).  The persistent (b,h)-major schedule keeps a CTA on ONE (b,h) slice for
// thousands of units, and the coarse levels of one slice are tiny (benchmark pyramid: 16x16 + 8x8 = 320 rows = 40 KB
// of fp32 grad rows; DETR encoder 13x21 = 273 rows).  So the row adds of those levels do not have to leave the SM:
//
//   * warps 0..14 ("workers") run the tuned backward as before, but for the points of the coarse levels they do not
//     issue `red`s: the lane group of a unit writes, per point, ONE 32-byte record {4 corner weights, 4 accumulator row
//     offsets} into the warp's slot in shared memory;
//   * warp 15 (the "owner") is the ONLY writer of a shared-memory accumulator that holds the coarse rows of the CTA's
//     current (b,h) slice: plain LDS.128 / FFMA / STS.128, lane = (corner, 16-byte chunk), so one instruction covers
//     the four corner rows of a point; no shared-memory atomics (fp32 ATOMS is a CAS loop on sm_100a);
//   * when the CTA's tile range leaves the slice (and at the end) the owner flushes the accumulator with one `red`
//     per row -- 320 rows per slice and CTA instead of 32 rows per unit.
//
// Hazards.  A single warp executes its LDS/STS in order, but to hide the LDS latency the owner keeps the rows of up to
// four points in flight (an X/Y pair per coarse level), and two points of one level can share rows.  The workers make
// that harmless before the hand-off: (i) clamped twin corners of a point (x0c == x1c / y0c == y1c, or zeros-mode
// corners outside the level) are FOLDED into one corner, the freed corner is pointed at a trash row, so the four rows
// of a point are distinct; (ii) for the second point Y of a pair every corner that hits a row of the first point X
// gets X's weight for that row added ("merge"), and the owner stores X's rows first, Y's rows second: whichever rows
// coincide end up with old + (wX + wY) * grad_out.
//
// Hand-off.  One slot per worker warp, guarded by two mbarriers (full: worker -> owner, empty: owner -> worker).  The
// owner visits the tiles of the CTA in tile order (it knows the schedule), so no polling and no per-record queue
// bookkeeping; a worker produces a tile every ~14k clk and the owner needs ~500 clk per tile, so nobody waits.
//
// Scope: fp32, D == 32, L == 4, K == 4 (all BASELINE shapes), grad_img requested.  Everything else takes
// msda_bwd_tiled.cu.  Written from scratch; the reference has no counterpart (its backward is one Triton program per
// unit issuing 64 vector atomics, kernels.py:396-553).
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"
#include "msda_tuning.h"

namespace msda {

namespace {

constexpr int kNeedImg = 1, kNeedPts = 2, kNeedAw = 4;
constexpr size_t kOwnL2Budget = 48u << 20;   // img + grad_img bytes of one wave of (b,h) slices (as msda_bwd_tiled.cu)

constexpr int kMaxWorkers = 15;     // warps 0..W-1 run the backward proper, warp W owns the accumulator
constexpr int kMaxCP = 8;           // coarse point slots per unit (the last kMaxCP points of the unit)
constexpr int kDefaultRows = 320;   // accumulator capacity in pyramid rows (16x16 + 8x8)
constexpr int kMaxRows = 448;

// One hand-off slot: the coarse points of the 4 units of a warp tile.  [unit][coarse point][corner]
struct OwnerSlot {
    float w[4][kMaxCP][4];          // corner weight = attention weight x bilinear weight (folded / merged, see above)
    unsigned off[4][kMaxCP][4];     // byte offset of the corner's row in the accumulator (row * 128), or the trash row
};
static_assert(sizeof(OwnerSlot) == 1024, "slot layout");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
}

}  // namespace

template <bool BORDER, int kWorkers>
__global__ void __launch_bounds__((kWorkers + 1) * 32, 1)
    msda_bwd_owner_kernel(const KernelArgs a, const WaveSchedule ws, const int cap_rows) {
    using T = float;
    constexpr int kThreads = (kWorkers + 1) * 32;
    constexpr int LANES = 8, LK = 16, VEC = 4, NB = 2;
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int G = Cfg::G, PPL = Cfg::PPL;
    static_assert(G == 4 && PPL == 2, "lane layout");

    extern __shared__ __align__(16) unsigned char s_dyn[];
    unsigned char *s_acc = s_dyn;                                                         // [(cap_rows + 1)][128 B]
    OwnerSlot *s_slot = reinterpret_cast<OwnerSlot *>(s_dyn + (size_t)(cap_rows + 1) * 128);   // [kWorkers]
    __shared__ Level s_lv[8];
    __shared__ int s_cfg[4];   // {number of coarse points per unit, first coarse pyramid row, coarse rows, -}
    __shared__ __align__(8) unsigned long long s_full[kWorkers], s_empty[kWorkers];

    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;
    const bool need_img = (a.flags & kNeedImg) != 0, need_pts = (a.flags & kNeedPts) != 0,
               need_aw = (a.flags & kNeedAw) != 0;
    if (threadIdx.x == 0) {
        // coarse levels: the longest suffix of the pyramid that fits the accumulator and kMaxCP points per unit
        int lc = a.L, rows = 0;
        for (int l = a.L - 1; l >= 0; --l) {
            const int n = s_lv[l].h * s_lv[l].w;
            if (rows + n > cap_rows || (a.L - l) * a.K > kMaxCP) break;
            rows += n;
            lc = l;
        }
        const int ncp = need_img ? (a.L - lc) * a.K : 0;
        s_cfg[0] = ncp;
        s_cfg[1] = ncp ? s_lv[lc].off : 0;
        s_cfg[2] = ncp ? rows : 0;
        for (int w = 0; w < kWorkers; ++w) {
            mbar_init(&s_full[w], 1);
            mbar_init(&s_empty[w], 1);
        }
    }
    for (int i = threadIdx.x; i < (cap_rows + 1) * 8; i += kThreads)
        reinterpret_cast<float4 *>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int ncp = s_cfg[0], row0 = s_cfg[1], coarse_rows = s_cfg[2];
    const int fcs = LK - ncp;                              // first coarse point slot of a unit
    const unsigned trash = (unsigned)cap_rows * 128u;      // byte offset of the trash row

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);
    const int tiles_per_bh = ws.tiles_per_bh;

    // worker state
    unsigned empty_parity = 1;   // a fresh mbarrier counts as "previous phase complete": the first wait passes
    OwnerSlot *slot = s_slot + (warp < kWorkers ? warp : 0);
    // owner state.  lane = (corner c, 16-byte chunk jc): one LDS.128 / STS.128 covers the four corner rows of a point
    const int c = lane >> 3, jc = lane & 7;
    unsigned char *acc_lane = s_acc + jc * 16;
    unsigned full_parity = 0;    // bit w: parity of s_full[w] to wait for next
    int cur_bh = -1;
    auto flush = [&](int bh) {
        const int b = bh / a.H, h = bh - b * a.H;
        float *base = gimg + ((size_t)b * a.Npix * a.H + h) * a.D + (size_t)row0 * a.H * a.D + jc * 4;
        for (int r = c; r < coarse_rows; r += 4) {
            float4 *p = reinterpret_cast<float4 *>(acc_lane + (size_t)r * 128);
            const float4 v = *p;
            if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
                red_add_v4(base + (size_t)r * a.H * a.D, v.x, v.y, v.z, v.w);
                *p = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
    };

    for (int wave = 0; wave < ws.waves; ++wave) {
        int t_begin, t_end;
        wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);
        if (warp < kWorkers) {
            // =========================================== workers ===========================================
            int tile = t_begin + warp;
            if (tile < t_end) {
                TileUnit tu = decode_tile(tile, tiles_per_bh, g, G, a);
                LaneOperands<T, PPL, false> op;
                float go[VEC];
                load_operands<T, LANES, LK, false, false>(a, tu, j, op);
                load_vec_stream<T, VEC>(gout + (size_t)tu.u * a.D + j * VEC, go);

                for (; tile < t_end; tile += kWorkers) {
                    const int tile_n = tile + kWorkers;
                    const bool has_next = tile_n < t_end;
                    const TileUnit tu_n = decode_tile(has_next ? tile_n : tile, tiles_per_bh, g, G, a);
                    LaneOperands<T, PPL, false> op_n;
                    float go_n[VEC];
                    load_operands<T, LANES, LK, false, false>(a, tu_n, j, op_n);
                    load_vec_stream<T, VEC>(gout + (size_t)tu_n.u * a.D + j * VEC, go_n);

                    const unsigned char *__restrict__ lane_base =
                        reinterpret_cast<const unsigned char *>(img + tu.bh_off + j * VEC);
                    unsigned char *__restrict__ gimg_base = reinterpret_cast<unsigned char *>(gimg + tu.bh_off + j * 4);
                    const bool live = tu.live;   // padding queries of the last tile shadow a real one and add nothing

                    Tap<float> tap[PPL];
                    float sx[PPL], sy[PPL];
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) {
                        const Level lv = s_lv[slot_level(j * PPL + pp, a)];
                        tap[pp] = locate<float>(op.xy[2 * pp], op.xy[2 * pp + 1], lv, BORDER, align);
                        sx[pp] = align ? (float)(lv.w - 1) : (float)lv.w;
                        sy[pp] = align ? (float)(lv.h - 1) : (float)lv.h;
                    }

                    // the owner has long finished with this warp's previous tile; this wait is one successful probe
                    if (ncp) mbar_wait(&s_empty[warp], empty_parity);

                    float part[3 * LK];
#pragma unroll
                    for (int pp = 0; pp < PPL; ++pp) {
#pragma unroll
                        for (int jj0 = 0; jj0 < LANES; jj0 += NB) {
                            uint4 raw[NB][4];
                            float fx[NB], fy[NB], fw[NB];
                            unsigned o[NB][4];
                            int rr[NB][4];
                            unsigned msk[NB];
                            const int pidx0 = jj0 * PPL + pp;          // point of n = 0; n = 1 is pidx0 + PPL
                            // warp-uniform; holds for both points of the batch (first term: compile time, keeps the hand-off code
                            // out of the batches that can never be coarse)
                            const bool coarse = pidx0 >= LK - kMaxCP && pidx0 >= fcs;
#pragma unroll
                            for (int n = 0; n < NB; ++n) {
                                const int src = jj0 + n;
                                const int r00 = __shfl_sync(0xffffffffu, tap[pp].row00, src, LANES);
                                const int pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                                fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                                fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                                fw[n] = __shfl_sync(0xffffffffu, op.wa[pp], src, LANES);
                                const int sxb = (pack >> kPackDxBit) & 1, syr = pack & kPackDyMask;
                                rr[n][0] = r00;
                                rr[n][1] = r00 + sxb;
                                rr[n][2] = r00 + syr;
                                rr[n][3] = r00 + syr + sxb;
                                msk[n] = BORDER ? 0xFu : (((unsigned)pack >> kPackMaskShift) & 0xFu);
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    o[n][c] = (unsigned)rr[n][c] * row_bytes;
                                    raw[n][c] = gather_row(lane_base, o[n][c]);   // clamped rows: always in range
                                }
                            }
                            float cw[NB][4];   // corner weights (attention x bilinear, 0 for masked corners)
#pragma unroll
                            for (int n = 0; n < NB; ++n) {
                                const float dx = fx[n], dy = fy[n];
                                float bw[4];
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
                                    const bool valid = BORDER || ((msk[n] >> c) & 1u);
                                    d[c] = valid ? acc : 0.0f;
                                    cw[n][c] = valid ? fw[n] * bw[c] : 0.0f;
                                    if (need_img && !coarse) {
                                        float gv[VEC];
#pragma unroll
                                        for (int e = 0; e < VEC; ++e) gv[e] = go[e] * cw[n][c];
                                        if (live && valid)
                                            red_add_v4(reinterpret_cast<float *>(gimg_base + o[n][c]), gv[0], gv[1], gv[2],
                                                       gv[3]);
                                    }
                                }
                                const int pidx = pidx0 + n * PPL;
                                part[3 * pidx + 0] = bw[0] * d[0] + bw[1] * d[1] + bw[2] * d[2] + bw[3] * d[3];
                                part[3 * pidx + 1] = (1.0f - dy) * (d[1] - d[0]) + dy * (d[3] - d[2]);
                                part[3 * pidx + 2] = (1.0f - dx) * (d[2] - d[0]) + dx * (d[3] - d[1]);
                            }
                            if (coarse) {   // implies need_img (ncp == 0 otherwise)
                                unsigned ao[NB][4];
#pragma unroll
                                for (int n = 0; n < NB; ++n) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) ao[n][c] = (unsigned)(rr[n][c] - row0) * 128u;
                                    // fold clamped twins (identical rows) into one corner; the freed corner -> trash row
                                    if (rr[n][1] == rr[n][0]) {
                                        cw[n][0] += cw[n][1];
                                        cw[n][2] += cw[n][3];
                                        ao[n][1] = trash;
                                        ao[n][3] = trash;
                                    }
                                    if (rr[n][2] == rr[n][0]) {
                                        cw[n][0] += cw[n][2];
                                        cw[n][1] += cw[n][3];
                                        ao[n][2] = trash;
                                        ao[n][3] = trash;
                                    }
                                    if (!live) {
#pragma unroll
                                        for (int c = 0; c < 4; ++c) ao[n][c] = trash;
                                    }
                                }
                                // merge: rows of Y (n = 1) that X (n = 0) also hits take X's weight along (the owner stores
                                // X first, Y second)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
#pragma unroll
                                    for (int i = 0; i < 4; ++i)
                                        if (ao[1][k] == ao[0][i]) cw[1][k] += cw[0][i];
                                if (j < 2) {
#pragma unroll
                                    for (int n = 0; n < NB; ++n) {
                                        const int cp = pidx0 + n * PPL - (LK - kMaxCP);
                                        uint4 v;
                                        if (j == 0)
                                            v = make_uint4(__float_as_uint(cw[n][0]), __float_as_uint(cw[n][1]),
                                                           __float_as_uint(cw[n][2]), __float_as_uint(cw[n][3]));
                                        else
                                            v = make_uint4(ao[n][0], ao[n][1], ao[n][2], ao[n][3]);
                                        void *dst = j == 0 ? static_cast<void *>(&slot->w[g][cp][0])
                                                           : static_cast<void *>(&slot->off[g][cp][0]);
                                        *reinterpret_cast<uint4 *>(dst) = v;
                                    }
                                }
                            }
                        }
                    }
                    if (ncp) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_full[warp]);
                        empty_parity ^= 1u;
                    }

                    // ---- reduce over the lanes; lane j ends with its own PPL points in part[0 .. 3*PPL) ----
                    transpose_reduce<3 * LK, LANES / 2>(part, j);
                    if (tu.live) {
                        T *__restrict__ gaw_u = gaw + (size_t)tu.u * LK;
                        T *__restrict__ gpts_u = gpts + (size_t)tu.u * LK * 2;
                        if (need_aw) {
                            float gw[PPL];
#pragma unroll
                            for (int pp = 0; pp < PPL; ++pp) gw[pp] = part[3 * pp + 0];
                            store_vec_stream<T, PPL>(gaw_u + j * PPL, gw);
                        }
                        if (need_pts) {
                            float gp[2 * PPL];
#pragma unroll
                            for (int pp = 0; pp < PPL; ++pp) {
                                gp[2 * pp + 0] = part[3 * pp + 1] * (op.wa[pp] * sx[pp]);
                                gp[2 * pp + 1] = part[3 * pp + 2] * (op.wa[pp] * sy[pp]);
                            }
                            store_vec_stream<T, 2 * PPL>(gpts_u + (j * PPL) * 2, gp);
                        }
                    }

                    tu = tu_n;
                    op = op_n;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) go[i] = go_n[i];
                }
            }
        } else {
            // ============================================ owner ============================================
            if (ncp && t_begin < t_end) {
                float4 gv[4], gv_n[4];
                {
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const TileUnit tu = decode_tile(t_begin, tiles_per_bh, gg, G, a);
                        gv[gg] = __ldg(reinterpret_cast<const float4 *>(gout + (size_t)tu.u * a.D + jc * 4));
                    }
                }
                for (int tile = t_begin; tile < t_end; ++tile) {
                    const int w = (tile - t_begin) % kWorkers;
                    {
                        const int tn = tile + 1 < t_end ? tile + 1 : tile;
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            const TileUnit tu = decode_tile(tn, tiles_per_bh, gg, G, a);
                            gv_n[gg] = __ldg(reinterpret_cast<const float4 *>(gout + (size_t)tu.u * a.D + jc * 4));
                        }
                    }
                    const int bh = tile / tiles_per_bh;
                    if (bh != cur_bh) {
                        if (cur_bh >= 0) flush(cur_bh);
                        cur_bh = bh;
                    }
                    mbar_wait(&s_full[w], (full_parity >> w) & 1u);
                    full_parity ^= 1u << w;
                    const OwnerSlot *S = s_slot + w;
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        unsigned off[kMaxCP];
                        float wt[kMaxCP];
#pragma unroll
                        for (int cp = 0; cp < kMaxCP; ++cp) {
                            if (cp >= kMaxCP - 4 || ncp == kMaxCP) {   // warp-uniform
                                off[cp] = S->off[gg][cp][c];
                                wt[cp] = S->w[gg][cp][c];
                            } else {
                                off[cp] = trash;
                                wt[cp] = 0.0f;
                            }
                        }
                        const float4 q = gv[gg];
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            // X = {r, 4 + r}, Y = {2 + r, 6 + r}: the worker's gather batches (pidx, pidx + 2)
                            const int ia = r, ib = 2 + r, ic = 4 + r, id = 6 + r;
                            float4 *pa = reinterpret_cast<float4 *>(acc_lane + off[ia]);
                            float4 *pb = reinterpret_cast<float4 *>(acc_lane + off[ib]);
                            float4 *pc = reinterpret_cast<float4 *>(acc_lane + off[ic]);
                            float4 *pd = reinterpret_cast<float4 *>(acc_lane + off[id]);
                            float4 va, vb, vc = *pc, vd = *pd;
                            if (ncp == kMaxCP) {
                                va = *pa;
                                vb = *pb;
                            }
                            vc.x = fmaf(wt[ic], q.x, vc.x); vc.y = fmaf(wt[ic], q.y, vc.y);
                            vc.z = fmaf(wt[ic], q.z, vc.z); vc.w = fmaf(wt[ic], q.w, vc.w);
                            vd.x = fmaf(wt[id], q.x, vd.x); vd.y = fmaf(wt[id], q.y, vd.y);
                            vd.z = fmaf(wt[id], q.z, vd.z); vd.w = fmaf(wt[id], q.w, vd.w);
                            if (ncp == kMaxCP) {
                                va.x = fmaf(wt[ia], q.x, va.x); va.y = fmaf(wt[ia], q.y, va.y);
                                va.z = fmaf(wt[ia], q.z, va.z); va.w = fmaf(wt[ia], q.w, va.w);
                                vb.x = fmaf(wt[ib], q.x, vb.x); vb.y = fmaf(wt[ib], q.y, vb.y);
                                vb.z = fmaf(wt[ib], q.z, vb.z); vb.w = fmaf(wt[ib], q.w, vb.w);
                                *pa = va;
                            }
                            *pc = vc;
                            __syncwarp();   // X's rows land before Y's (rows they share carry the merged weight in Y)
                            if (ncp == kMaxCP) *pb = vb;
                            *pd = vd;
                            __syncwarp();   // ... and before the next round reads them
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[w]);
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) gv[gg] = gv_n[gg];
                }
            }
        }
        wave_pace_cta(ws, wave);
    }
    if (warp == kWorkers && cur_bh >= 0) flush(cur_bh);
}

// Shared memory the kernel needs for an accumulator of `cap_rows` rows (+ trash row) and the hand-off slots.
static size_t owner_smem_bytes(int cap_rows, int workers) {
    return (size_t)(cap_rows + 1) * 128 + sizeof(OwnerSlot) * workers;
}

template <bool BORDER, int WORKERS>
static cudaError_t launch_owner_t(const KernelArgs &a, const WaveSchedule &ws, int grid, int cap_rows, cudaStream_t st) {
    // opt in to > 48 KB of dynamic shared memory once per device and instantiation (benign race: setting it twice is fine)
    static bool attr_set[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(msda_bwd_owner_kernel<BORDER, WORKERS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)owner_smem_bytes(kMaxRows, WORKERS));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    msda_bwd_owner_kernel<BORDER, WORKERS>
        <<<grid, (WORKERS + 1) * 32, owner_smem_bytes(cap_rows, WORKERS), st>>>(a, ws, cap_rows);
    return cudaGetLastError();
}

cudaError_t launch_backward_owner(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    // K == 4: the owner keeps one point pair of EACH of the last two levels in flight and relies on the two pairs
    // being on different levels (no shared rows)
    if (dtype != 0 || a.D != 32 || a.L != 4 || a.K != 4) return cudaErrorNotSupported;
    if (!(a.flags & kNeedImg)) return cudaErrorNotSupported;
    if (!tiled_offsets_fit(a, sizeof(float))) return cudaErrorNotSupported;
    int cap_rows = tuning().owner_rows > 0 ? tuning().owner_rows : kDefaultRows;
    if (cap_rows > kMaxRows) cap_rows = kMaxRows;
    const int workers = tuning().owner_workers == 14 ? 14 : kMaxWorkers;
    constexpr int G = 4;
    const int tiles_per_bh = (a.Q + G - 1) / G;
    const int total_tiles = a.B * a.H * tiles_per_bh;
    const int want = (total_tiles + workers - 1) / workers;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    WaveSchedule ws = make_wave_schedule(a, tiles_per_bh, sizeof(float) + sizeof(float), kOwnL2Budget);
    const bool big_waves = (long long)ws.slices_per_wave * tiles_per_bh >= 4LL * workers * grid;
    if (ws.waves > 1 && grid == sm_count && (big_waves || pacing_forced())) {
        const cudaError_t e = acquire_pace_counter(st, &ws.pace);
        if (e != cudaSuccess) return e;
    }
    if (workers == 14)
        return a.border ? launch_owner_t<true, 14>(a, ws, grid, cap_rows, st)
                        : launch_owner_t<false, 14>(a, ws, grid, cap_rows, st);
    return a.border ? launch_owner_t<true, 15>(a, ws, grid, cap_rows, st)
                    : launch_owner_t<false, 15>(a, ws, grid, cap_rows, st);
}

}  // namespace msda
