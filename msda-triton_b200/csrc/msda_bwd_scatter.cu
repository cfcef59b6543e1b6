// msda_bwd_scatter.cu -- grad_img as a separate "scatter-only" kernel (split backward, opt-in / shape-gated).
//
// grad_img needs only (sampling_points, attention_weights, grad_out) -- the pyramid values are needed just for
// grad_points / grad_weights.  Splitting the backward into
//     K1 = the tuned backward without grad_img (gathers, partials; L1 fully available for the pyramid), and
//     K2 = this kernel (no gathers at all, so the whole 227 KB of shared memory can hold binning state)
// lifts the conflict measured on the fused binned kernel, where the binning structures pushed the pyramid out of L1.
//
// K2 walks the same persistent (b,h)-major schedule in super-tiles of TQ = 512 queries:
//   phase A: every lane resolves its PPL points.  Corners in the "binned" levels (the coarsest levels with at most
//            MAXROWS rows in total, decided on device) are pushed onto a per-row linked list in shared memory
//            (native 32-bit ATOMS.EXCH on the head; weight + next index stored in natural order, so the query is
//            implied by the record index).  Corners of the fine levels go straight to L2 as `red.v4.f32`.  The unit's
//            grad_out row is parked in shared memory.
//   phase C: one lane group per destination row walks the row's list, accumulates weight * grad_out[q] in registers
//            and issues ONE `red.v4.f32` per lane for the whole super-tile.
// For the benchmark pyramid levels 1-3 (1344 rows, 12 of 16 points) are binned: `red` rows per 512 queries drop from
// 32768 to 8192 + <=1344.
#include <cstdlib>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"

namespace msda {

template <typename T, int LANES, int LK, bool BORDER, int THREADS, int ROUNDS, int MAXROWS, int NBP>
__global__ void __launch_bounds__(THREADS, 1)
    msda_bwd_scatter_kernel(const KernelArgs a, const int stiles_per_bh, const int total_stiles) {
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int VEC = Cfg::VEC, G = Cfg::G, PPL = Cfg::PPL;
    constexpr int NW = THREADS / 32, TQ = NW * G * ROUNDS, NGROUPS = THREADS / LANES;
    constexpr int DCH = LANES * VEC;                       // channels per row (== D)
    constexpr unsigned END = 0xFFFFu;
    static_assert(TQ * NBP * 4 < 0xFFFF, "record index must fit in 16 bits");

    extern __shared__ __align__(16) unsigned char s_dyn[];
    float *s_w = reinterpret_cast<float *>(s_dyn);                                 // [4][TQ][NBP] record weights
    float *s_go = s_w + 4 * TQ * NBP;                                              // [TQ][DCH]    grad_out rows
    unsigned *s_head = reinterpret_cast<unsigned *>(s_go + TQ * DCH);              // [MAXROWS]    list heads
    unsigned short *s_next = reinterpret_cast<unsigned short *>(s_head + MAXROWS); // [4][TQ][NBP] next record
    __shared__ Level s_lv[LK];
    __shared__ int s_bin[3];  // first binned point, first binned pixel row, number of binned rows

    build_level_table(s_lv, a.shapes, a.L);
    if (threadIdx.x == 0) {
        // binned levels: the longest suffix (coarsest first) with <= MAXROWS rows and <= NBP points per unit
        int rows = 0, l0 = a.L;
        for (int l = a.L - 1; l >= 0; --l) {
            const int n = s_lv[l].h * s_lv[l].w;
            if (rows + n > MAXROWS || (a.L - l) * a.K > NBP) break;
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
                const Level lv = s_lv[p / a.K];
                const Tap<float> t = locate<float>(xy[2 * pp], xy[2 * pp + 1], lv, BORDER, align);
                const int step_y = t.pack & kPackDyMask;
                const int step_x = (t.pack >> kPackDxBit) & 1;
                const unsigned mask = (unsigned)(t.pack >> kPackMaskShift) & 0xFu;
                const float wl = tu.live ? wa[pp] : 0.0f;
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
                    const int r00 = t.row00 - base_row;
                    const int rows4[4] = {r00, r00 + step_x, r00 + step_y, r00 + step_y + step_x};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (BORDER || ((mask >> c) & 1u)) {
                            const unsigned idx = (unsigned)((c * TQ + q_local) * NBP + (p - p0));
                            const unsigned prev = atomicExch(&s_head[rows4[c]], idx);
                            s_next[idx] = (unsigned short)prev;
                            s_w[idx] = w4[c];
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
                        if (BORDER || ((mask >> c) & 1u)) red_add_vec<VEC>(gimg_lane + (size_t)rows4[c] * row_stride, gv);
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

        // ---- phase C: per-row segmented reduction of the binned levels ----
        __syncthreads();
        {
            float *__restrict__ gimg_rows = gimg + st_bh_off + (size_t)base_row * row_stride + j * VEC;
            const unsigned group_mask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << (g * LANES);
            for (int r = threadIdx.x / LANES; r < nrows; r += NGROUPS) {
                unsigned idx = END;
                if (j == 0) {   // the group leader pops the whole list
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
                        const unsigned q = (idx / NBP) % TQ;
                        const Pack<float, VEC> gq = *reinterpret_cast<const Pack<float, VEC> *>(s_go + q * DCH + j * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w, gq.v[e], acc[e]);
                        idx = nxt;
                    } while (idx != END);
                    red_add_vec<VEC>(gimg_rows + (size_t)r * row_stride, acc);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int LANES, int LK>
static cudaError_t launch_scatter_t(const KernelArgs &a, int sm_count, cudaStream_t st) {
    constexpr int THREADS = 1024, MAXROWS = 1408, NBP = 12;
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int ROUNDS = 512 / ((THREADS / 32) * Cfg::G);   // super-tiles of 512 queries
    static_assert(ROUNDS >= 1, "super-tile smaller than one round");
    constexpr int TQ = (THREADS / 32) * Cfg::G * ROUNDS;
    constexpr size_t kSmem = sizeof(float) * 4 * TQ * NBP + sizeof(float) * TQ * LANES * Cfg::VEC +
                             sizeof(unsigned) * MAXROWS + sizeof(unsigned short) * 4 * TQ * NBP;
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
    if (a.border) return launch(msda_bwd_scatter_kernel<T, LANES, LK, true, THREADS, ROUNDS, MAXROWS, NBP>);
    return launch(msda_bwd_scatter_kernel<T, LANES, LK, false, THREADS, ROUNDS, MAXROWS, NBP>);
}

// grad_img only (a.gimg = fp32 accumulation image, already zero-filled).  cudaErrorNotSupported -> caller falls back.
cudaError_t launch_backward_scatter(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (a.LK != 16 || a.L > 16 || a.D != 32) return cudaErrorNotSupported;
    const unsigned long long tiles = (unsigned long long)a.B * a.H * a.Q;
    if (tiles >= (1ull << 31) || (unsigned long long)a.Npix >= (1ull << 23)) return cudaErrorNotSupported;
    if (dtype == 0) return launch_scatter_t<float, 8, 16>(a, sm_count, st);
    return cudaErrorNotSupported;
}

}  // namespace msda
