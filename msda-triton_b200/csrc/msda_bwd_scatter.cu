// msda_bwd_scatter.cu -- grad_img as a "scatter-only" kernel.
//
// grad_img needs only (sampling_points, attention_weights, grad_out) -- the pyramid values are needed just for
// grad_points / grad_weights.  This kernel therefore does no gathers at all, which frees the whole 227 KB of shared
// memory for an in-CTA segmented reduction of the coarse pyramid levels (in the fused backward such structures push the
// pyramid out of L1 and cost more than they save: profiles/r1_ncu_summary.md section 4).  It is what msda_backward
// runs when ONLY grad_img is requested (0.29 ms versus 0.45 ms on the bench shape), and -- together with the tuned
// backward without grad_img -- the opt-in split backward (MSDA_B200_BWD_SPLIT=1).
//
// The kernel walks the persistent (b,h)-major schedule in super-tiles of TQ queries:
//   phase A: every lane resolves its PPL points.  Corners in the "binned" levels (the coarsest levels with at most
//            MAXROWS rows in total, decided on device) are pushed onto a linked list in shared memory: one native 32-bit
//            ATOMS.EXCH on the list head, then an 8-byte record {weight, next | query << 16} in natural order.  Corners
//            of the fine levels go straight to L2 as `red.v4.f32`.  The unit's grad_out row is parked in shared memory.
//   phase C: one lane group per list walks it, accumulates weight * grad_out[q] in registers and issues ONE `red.v4.f32`
//            per lane for the whole super-tile.
// A destination row of a very coarse level would collect TQ*K*4/rows records (96 for the 8x8 level) while a row of a
// finer binned level collects a handful; to keep the lane groups of phase C evenly loaded such rows get several lists
// (split by the low bits of the query index), so that every list is about kTargetList records long.
// For the benchmark pyramid levels 1-3 (1344 rows, 12 of 16 points) are binned: `red` rows per 384 queries drop from
// 24576 to 6144 + <= 1536.
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"

namespace msda {

constexpr int kTargetList = 24;   // records per list phase C aims for
constexpr int kMaxBinLevels = 8;

struct BinLevel {
    int level;        // pyramid level
    int head_base;    // first list head of this level
    int split_log2;   // lists per destination row = 1 << split_log2
    int rows;         // h * w
};

template <typename T, int LANES, int LK, bool BORDER, int THREADS, int ROUNDS, int MAXHEADS, int NBP>
__global__ void __launch_bounds__(THREADS, 1)
    msda_bwd_scatter_kernel(const KernelArgs a, const int stiles_per_bh, const int total_stiles) {
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int VEC = Cfg::VEC, G = Cfg::G, PPL = Cfg::PPL;
    constexpr int NW = THREADS / 32, TQ = NW * G * ROUNDS, NGROUPS = THREADS / LANES;
    constexpr int DCH = LANES * VEC;                       // channels per row (== D)
    constexpr unsigned END = 0xFFFFu;
    static_assert(TQ * NBP * 4 < 0xFFFF && TQ <= 0xFFFF, "record index and query index must fit in 16 bits");

    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint2 *s_rec = reinterpret_cast<uint2 *>(s_dyn);                               // [4][TQ][NBP] {weight, next | q<<16}
    float *s_go = reinterpret_cast<float *>(s_rec + 4 * TQ * NBP);                 // [TQ][DCH]    grad_out rows
    unsigned *s_head = reinterpret_cast<unsigned *>(s_go + TQ * DCH);              // [MAXHEADS]   list heads
    __shared__ Level s_lv[8];
    __shared__ BinLevel s_bl[kMaxBinLevels];
    __shared__ int s_bin[3];  // first binned point, number of binned levels, number of list heads

    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;
    if (threadIdx.x == 0) {
        // binned levels: the longest suffix (coarsest first) that fits MAXHEADS list heads and NBP points per unit
        int heads = 0, l0 = a.L, n_bl = 0;
        for (int l = a.L - 1; l >= 0 && n_bl < kMaxBinLevels; --l) {
            const int rows = s_lv[l].h * s_lv[l].w;
            int split_log2 = 0;   // lists per row so that a list holds about kTargetList of the TQ*K*4 records
            while (split_log2 < 4 && (TQ * a.K * 4) >= (kTargetList << (split_log2 + 1)) * rows) ++split_log2;
            if (heads + (rows << split_log2) > MAXHEADS || (a.L - l) * a.K > NBP) break;
            s_bl[n_bl].level = l;
            s_bl[n_bl].head_base = heads;
            s_bl[n_bl].split_log2 = split_log2;
            s_bl[n_bl].rows = rows;
            heads += rows << split_log2;
            ++n_bl;
            l0 = l;
        }
        s_bin[0] = l0 * a.K;
        s_bin[1] = n_bl;
        s_bin[2] = heads;
    }
    for (int i = threadIdx.x; i < MAXHEADS; i += THREADS) s_head[i] = END;
    __syncthreads();
    const int p0 = s_bin[0], n_bl = s_bin[1];

    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const size_t row_stride = (size_t)a.H * a.D;           // fp32 accumulation image: elements between pixel rows

    const int st_begin = (int)((long long)total_stiles * blockIdx.x / gridDim.x);
    const int st_end = (int)((long long)total_stiles * (blockIdx.x + 1) / gridDim.x);
    if (st_begin >= st_end) return;

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
            float *__restrict__ gimg_lane = gimg + tu.bh_off + j * VEC;

            // park this unit's grad_out row for phase C
            *reinterpret_cast<Pack<float, VEC> *>(s_go + q_local * DCH + j * VEC) =
                *reinterpret_cast<const Pack<float, VEC> *>(go);

            // ---- this lane's points: records for binned levels, exchange registers for the direct ones ----
            int d_row[PPL], d_pack[PPL];
            float d_w[PPL][4];
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                const int p = j * PPL + pp;
                const int l = p / a.K;
                const Level lv = s_lv[l];
                const Tap<float> t = locate<float>(xy[2 * pp], xy[2 * pp + 1], lv, BORDER, align);
                const int step_y = t.pack & kPackDyMask;
                const int step_x = (t.pack >> kPackDxBit) & 1;
                const unsigned mask = (unsigned)(t.pack >> kPackMaskShift) & 0xFu;
                const float wl = wa[pp];   // padding queries add nothing: their row adds are predicated on tu.live
                const float wy1 = wl * t.dy, wy0 = wl - wy1;
                float w4[4];
                w4[1] = wy0 * t.dx;
                w4[0] = wy0 - w4[1];
                w4[3] = wy1 * t.dx;
                w4[2] = wy1 - w4[3];
                d_row[pp] = t.row00;
                d_pack[pp] = t.pack;
#pragma unroll
                for (int c = 0; c < 4; ++c) d_w[pp][c] = w4[c];
                if (p >= p0 && tu.live) {
                    const BinLevel bl = s_bl[a.L - 1 - l];       // binned levels are stored coarsest first
                    const int r00 = t.row00 - lv.off;
                    const int rows4[4] = {r00, r00 + step_x, r00 + step_y, r00 + step_y + step_x};
                    const int sub = q_local & ((1 << bl.split_log2) - 1);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (BORDER || ((mask >> c) & 1u)) {
                            const unsigned idx = (unsigned)((c * TQ + q_local) * NBP + (p - p0));
                            const int head = bl.head_base + (rows4[c] << bl.split_log2) + sub;
                            const unsigned prev = atomicExch(&s_head[head], idx);
                            s_rec[idx] = make_uint2(__float_as_uint(w4[c]), prev | ((unsigned)q_local << 16));
                        }
                    }
                }
            }

            // ---- direct levels: one red.v4 per lane per valid corner ----
#pragma unroll
            for (int pidx = 0; pidx < LK; ++pidx) {
                if (pidx < p0) {   // warp-uniform
                    const int src = pidx / PPL, pp = pidx % PPL;
                    const int row00 = __shfl_sync(0xffffffffu, d_row[pp], src, LANES);
                    const int pack = __shfl_sync(0xffffffffu, d_pack[pp], src, LANES);
                    float w4[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) w4[c] = __shfl_sync(0xffffffffu, d_w[pp][c], src, LANES);
                    const int step_y = pack & kPackDyMask;
                    const int step_x = (pack >> kPackDxBit) & 1;
                    const unsigned mask = (unsigned)(pack >> kPackMaskShift) & 0xFu;
                    const int rows4[4] = {row00, row00 + step_x, row00 + step_y, row00 + step_y + step_x};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float gv[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) gv[e] = go[e] * w4[c];
                        if (tu.live && (BORDER || ((mask >> c) & 1u)))
                            red_add_vec<VEC>(gimg_lane + (size_t)rows4[c] * row_stride, gv);
                    }
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

        // ---- phase C: one lane group per list ----
        __syncthreads();
        {
            const unsigned group_mask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << (g * LANES);
            for (int b = 0; b < n_bl; ++b) {
                const BinLevel bl = s_bl[b];
                float *__restrict__ gimg_level = gimg + st_bh_off + (size_t)s_lv[bl.level].off * row_stride + j * VEC;
                const int n_heads = bl.rows << bl.split_log2;
                for (int hd = threadIdx.x / LANES; hd < n_heads; hd += NGROUPS) {
                    unsigned idx = END;
                    if (j == 0) {   // the group leader pops the whole list
                        idx = s_head[bl.head_base + hd];
                        s_head[bl.head_base + hd] = END;
                    }
                    idx = __shfl_sync(group_mask, idx, g * LANES);
                    if (idx == END) continue;
                    float acc[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
                    uint2 rec = s_rec[idx];
                    while (true) {
                        const unsigned nxt = rec.y & 0xFFFFu;
                        const unsigned q = rec.y >> 16;
                        const float w = __uint_as_float(rec.x);
                        if (nxt != END) rec = s_rec[nxt];            // next record is in flight during the FMAs
                        const Pack<float, VEC> gq = *reinterpret_cast<const Pack<float, VEC> *>(s_go + q * DCH + j * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w, gq.v[e], acc[e]);
                        if (nxt == END) break;
                    }
                    red_add_vec<VEC>(gimg_level + (size_t)(hd >> bl.split_log2) * row_stride, acc);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int LANES, int LK>
static cudaError_t launch_scatter_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    constexpr int THREADS = 1024, MAXHEADS = 2048, NBP = 12;
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int ROUNDS = 384 / ((THREADS / 32) * Cfg::G);   // super-tiles of 384 queries
    static_assert(ROUNDS >= 1, "super-tile smaller than one round");
    constexpr int TQ = (THREADS / 32) * Cfg::G * ROUNDS;
    constexpr size_t kSmem = sizeof(uint2) * 4 * TQ * NBP + sizeof(float) * TQ * LANES * Cfg::VEC +
                             sizeof(unsigned) * MAXHEADS;
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
    const int stiles_per_bh = (a.Q + TQ - 1) / TQ;
    const int total_stiles = a.B * a.H * stiles_per_bh;
    const int grid = total_stiles < sm_count ? (total_stiles < 1 ? 1 : total_stiles) : sm_count;
    auto launch = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, THREADS, kSmem, st>>>(a, stiles_per_bh, total_stiles);
        return cudaGetLastError();
    };
    if (a.border) return launch(msda_bwd_scatter_kernel<T, LANES, LK, true, THREADS, ROUNDS, MAXHEADS, NBP>);
    return launch(msda_bwd_scatter_kernel<T, LANES, LK, false, THREADS, ROUNDS, MAXHEADS, NBP>);
}

// grad_img only (a.gimg = fp32 accumulation image, already zero-filled).  cudaErrorNotSupported -> caller falls back.
cudaError_t launch_backward_scatter(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK != 16 || a.L > 8 || a.D != 32) return cudaErrorNotSupported;
    const unsigned long long tiles = (unsigned long long)a.B * a.H * a.Q;
    if (tiles >= (1ull << 31) || (unsigned long long)a.Npix >= (1ull << 23)) return cudaErrorNotSupported;
    if (dtype == 0) return launch_scatter_t<float, 8, 16>(a, sm_count, st);
    return cudaErrorNotSupported;
}

}  // namespace msda
