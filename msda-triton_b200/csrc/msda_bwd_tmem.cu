// msda_bwd_tmem.cu -- tuned backward with the COARSE pyramid levels accumulated in TENSOR MEMORY.
//
// Why.  The tuned backward (msda_bwd_tiled.cu) is bound by the number of 128-byte `red.global.add.v4.f32` row adds an
// SM can inject into L2 (measured ~5.6 clk per row and SM = 6.4 TB/s chip-wide; 64 rows per unit = the same 81.9 M
// atomic sectors the reference issues, kernels.py:550-553).  The persistent (b,h)-major schedule keeps a CTA on ONE
// (b,h) slice for thousands of units, and the coarse levels of one slice are tiny (benchmark pyramid: 16x16 + 8x8 =
// 320 rows; DETR encoder 13x21 = 273 rows), so their row adds do not have to leave the SM.  Shared memory is the wrong
// place for them: while the LSU is saturated with `red`s every LDS / STS / SHFL of the SM queues behind them
// (scripts/micro/lsu_contention.cu: 595 clk per dependent LDS->FFMA->STS step, and the shared-memory owner-warp
// backward built on it ran 5x slower than the plain kernel).  Tensor memory has its own datapath: a tcgen05.ld ->
// FFMA -> tcgen05.st step costs 56 clk under the same load, and REDUX (the warp broadcast used below) is not an LSU
// instruction either.
//
// How.  A warp can reach 32 TMEM lanes x 512 columns (the quarter `warp % 4`).  Lane = channel (D = 32), column = row
// of the coarse level(s): one quarter holds up to ~500 grad rows of the CTA's current (b,h) slice, accumulated in fp32.
//   * Every warp runs the tuned backward as before, but for the points of the coarse levels it issues no `red`s:
//     lane (g, j) of the warp keeps ONE record {columns, 4 folded corner weights} -- the record of coarse point j of
//     the warp's unit g.
//   * At the end of the warp tile the warp takes its TURN on the accumulator of its quarter: for each of the 32 records
//     the holder lane's five words are broadcast with REDUX, the two x-adjacent corner rows are read with ONE
//     tcgen05.ld.32x32b.x2 (rows c, c+1 and c+W, c+W+1), updated with grad_out in lane = channel layout, and stored
//     back.  With two coarse levels a record of each level is in flight at a time (disjoint columns).
//   * The 4 (3) warps of a quarter share one accumulator; turns are handed round a fixed ring with NAMED BARRIERS
//     (bar.arrive -> bar.sync, ids 1..15: the barrier unit is not behind the LSU either).  Steady state: a turn is
//     ~2k clk, a warp tile ~20k clk, so the ring never waits.
//   * When the ring crosses into another (b,h) slice, and at the end of every wave, the warp holding the turn flushes
//     the quarter: one 128-byte row add per non-zero row (320 per quarter and slice instead of 32 per unit).
//
// Hazards inside a record.  The four corner rows of a point must be distinct columns: clamped twin corners (border
// clamping: x0c == x1c or y0c == y1c; zeros-mode corners outside the level) are FOLDED into one corner by the worker
// lanes, the freed corners are masked, and a point whose two corner rows coincide reads its second pair from trash
// columns.  Levels are separated by a pad column because the x2 access of a level's last row touches the next column.
//
// Scope: fp32, D == 32, L*K == 16, K == 4 (all BASELINE shapes), grad_img requested; at most the two coarsest levels,
// as many as fit 503 columns.  Everything else takes msda_bwd_tiled.cu.  Written from scratch; the reference has no
// counterpart (its backward is one Triton program per unit issuing 64 vector atomics, kernels.py:396-553).
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"
#include "msda_tuning.h"

namespace msda {

namespace {

constexpr int kNeedImg = 1, kNeedPts = 2, kNeedAw = 4;
constexpr size_t kTmemL2Budget = 48u << 20;   // img + grad_img bytes of one wave of (b,h) slices (as msda_bwd_tiled.cu)

constexpr int kMaxWarps = 15;                 // ring edges = named barriers 1..15 (0 is __syncthreads)
constexpr int kTmemCols = 512;
// Column map of a quarter: [0, kMaxRegion) accumulator (levels + one pad column each), 4 trash columns, the broadcast
// buffer (two records x five words x four replicas), and per warp of the quarter 40 private columns that hold the
// records of the warp's current tile (8 point slots x 5 words per lane) -- tensor memory as a lane-private spill space.
constexpr int kMaxRegion = 347;
constexpr unsigned kTrashA = 348, kTrashB = 350;   // x2 trash columns of the two records in flight
constexpr unsigned kBcast = 352;
constexpr unsigned kSpill = 392, kSpillStride = 40;
constexpr int kMaxWarpsPerQuarter = 3;
static_assert(kSpill + kMaxWarpsPerQuarter * kSpillStride == kTmemCols, "column map");
constexpr int kCP = 8;                        // coarse point slots per unit = the last 8 of the 16 points

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- tensor memory (tcgen05) ----
__device__ __forceinline__ void tmem_alloc_all(unsigned *smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_all(unsigned base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kTmemCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Two adjacent columns of this thread's TMEM lane.  The loaded registers are only defined after tmem_wait_ld*(), which
// takes them as in/out operands so that no use can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld2(unsigned addr, unsigned &r0, unsigned &r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_st2(unsigned addr, unsigned r0, unsigned r1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld4(unsigned (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3])::"memory");
}
__device__ __forceinline__ void tmem_wait_ld8(unsigned (&u)[4], unsigned (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3])::"memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned addr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(unsigned (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void tmem_zero16(unsigned addr) {
    const unsigned z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(addr),
        "r"(z)
        : "memory");
}

// ---- named barriers: the ring edge into warp w is barrier 1 + w, two warps (64 threads) each ----
__device__ __forceinline__ void ring_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ring_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// One record as broadcast to the whole warp.
struct CoarseRec {
    unsigned w0;        // bits 0..8 column of corner rows 00|01, bits 9..17 column of rows 10|11, bits 18..21 corner mask
    unsigned cw[4];     // folded corner weights (attention x bilinear), fp32 bits
};

// Broadcast of the four units' records of one point slot to every lane of the warp THROUGH TENSOR MEMORY (no LSU, no
// REDUX): the lanes of group g all hold the record of unit g; every lane stores the five words, each replicated into
// four columns, into its own TMEM lane (32x32b), and the warp reads them back with the 16x128b shape, which hands
// thread t the word at (TMEM lane t/4 [+8], column t%4 of each 4-column repetition) -- lanes t/4 and t/4+8 belong to
// groups 0|1 (2|3 for the upper 16 lanes) and all four columns hold the same word, so EVERY thread receives the words
// of all four units.  (SHFL queues behind the reds in the LSU: ~1700 clk per shuffle inside this kernel; REDUX costs
// ~33 clk per word.)
__device__ __forceinline__ void bcast_store(unsigned addr, const CoarseRec &r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %2, %2, %2, %2, %3, %3, %3, %3, %4, %4, %4, %4};" ::"r"(addr),
        "r"(r.cw[0]), "r"(r.cw[1]), "r"(r.cw[2]), "r"(r.cw[3])
        : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr + 16u), "r"(r.w0) : "memory");
}
// out[0], out[1]: the records of the two units whose lane groups sit in the 16 TMEM lanes at `addr`.
__device__ __forceinline__ void bcast_load(unsigned addr, CoarseRec &g0, CoarseRec &g1) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(g0.cw[0]), "=r"(g1.cw[0]), "=r"(g0.cw[1]), "=r"(g1.cw[1]), "=r"(g0.cw[2]), "=r"(g1.cw[2]),
                   "=r"(g0.cw[3]), "=r"(g1.cw[3])
                 : "r"(addr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(g0.w0), "=r"(g1.w0) : "r"(addr + 16u) : "memory");
}
// A lane's own record to / from its private columns.
__device__ __forceinline__ void spill_store(unsigned addr, const CoarseRec &r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r.cw[0]), "r"(r.cw[1]),
                 "r"(r.cw[2]), "r"(r.cw[3])
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr + 4u), "r"(r.w0) : "memory");
}
__device__ __forceinline__ void spill_load(unsigned addr, CoarseRec &r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.cw[0]), "=r"(r.cw[1]), "=r"(r.cw[2]), "=r"(r.cw[3])
                 : "r"(addr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r.w0) : "r"(addr + 4u) : "memory");
}
__device__ __forceinline__ void rec_pin(CoarseRec &r) {   // ties the record to the preceding tcgen05.wait::ld
    asm volatile("" : "+r"(r.w0), "+r"(r.cw[0]), "+r"(r.cw[1]), "+r"(r.cw[2]), "+r"(r.cw[3])::"memory");
}

// Do the column pairs of record r come within one column of (pc, pc+1) / (pc2, pc2+1)?
__device__ __forceinline__ bool rec_near(unsigned x, unsigned y) { return x - y + 1u <= 2u; }
__device__ __forceinline__ bool rec_conflict(const CoarseRec &r, unsigned pc, unsigned pc2) {
    const unsigned c = r.w0 & 511u, c2 = (r.w0 >> 9) & 511u;
    return rec_near(c, pc) | rec_near(c, pc2) | rec_near(c2, pc) | rec_near(c2, pc2);
}

__device__ __forceinline__ void rec_load(const CoarseRec &r, unsigned lane_base, unsigned (&v)[4]) {
    tmem_ld2(lane_base + (r.w0 & 511u), v[0], v[1]);
    tmem_ld2(lane_base + ((r.w0 >> 9) & 511u), v[2], v[3]);
}
__device__ __forceinline__ void rec_update(const CoarseRec &r, unsigned (&v)[4], float q) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float nv = fmaf(__uint_as_float(r.cw[c]), q, __uint_as_float(v[c]));
        if ((r.w0 >> (18 + c)) & 1u) v[c] = __float_as_uint(nv);
    }
}
__device__ __forceinline__ void rec_store(const CoarseRec &r, unsigned lane_base, const unsigned (&v)[4]) {
    tmem_st2(lane_base + (r.w0 & 511u), v[0], v[1]);
    tmem_st2(lane_base + ((r.w0 >> 9) & 511u), v[2], v[3]);
}

#ifdef MSDA_TMEM_PROF
__device__ long long g_tmem_prof[148 * 16 * 8];   // per CTA and warp: {ring wait, turn work, flush, whole kernel} clocks
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(slot, dt) prof_acc[slot] += (dt)
#else
#define PROF_T(v)
#define PROF_ADD(slot, dt)
#endif

}  // namespace

template <bool BORDER, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
    msda_bwd_tmem_kernel(const KernelArgs a, const WaveSchedule ws, const int max_levels) {
    using T = float;
    constexpr int LANES = 8, LK = 16, VEC = 4, NB = 2;
    using Cfg = TiledCfg<T, LANES, LK>;
    constexpr int G = Cfg::G, PPL = Cfg::PPL;
    static_assert(G == 4 && PPL == 2, "lane layout");

    __shared__ Level s_lv[8];
    __shared__ int s_cfg[4];
    __shared__ unsigned s_tmem;
#ifdef MSDA_TMEM_PROF
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = clock64();
#endif

    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;
    const bool need_pts = (a.flags & kNeedPts) != 0, need_aw = (a.flags & kNeedAw) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        // levels kept in tensor memory: the longest suffix of <= max_levels levels that fits the column region
        int nlev = 0, rows = 0;
        for (int l = a.L - 1; l >= 0 && nlev < max_levels; --l) {
            const int n = s_lv[l].h * s_lv[l].w;
            if (rows + n + nlev + 1 > kMaxRegion || s_lv[l].w < 2) break;
            rows += n;
            ++nlev;
        }
        if (rows < 16) nlev = 0;   // the flush walks the region in 16-column pieces
        s_cfg[0] = nlev;
        s_cfg[1] = nlev ? s_lv[a.L - nlev].off : 0;                  // first pyramid row of the region
        s_cfg[2] = nlev == 2 ? s_lv[a.L - 1].off : 0x7fffffff;       // first row of the second level (after the pad)
        s_cfg[3] = rows + (nlev == 2 ? 1 : 0);                       // columns in use
    }
    if (warp == 0) tmem_alloc_all(&s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const int nlev = s_cfg[0], row0 = s_cfg[1], off_b = s_cfg[2], used_cols = s_cfg[3];
    const int pad_col = nlev == 2 ? off_b - row0 : -1;
    const int ncp = nlev * a.K;                       // coarse points per unit (0, 4 or 8)
    const int fcs = LK - ncp;                         // first coarse point slot
    const unsigned tmem_base = s_tmem;
    const unsigned lane_base = tmem_base + ((unsigned)((warp & 3) * 32) << 16);
    const unsigned spill_base = lane_base + kSpill + (unsigned)(warp >> 2) * kSpillStride;
    if (nlev && warp < 4) {                           // the first warp of each quarter clears it
        for (int c = 0; c < kTmemCols; c += 16) tmem_zero16(lane_base + c);
        tmem_wait_st();
    }

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    float *__restrict__ gimg = static_cast<float *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int j = lane % LANES, g = lane / LANES;
    const bool align = a.align != 0;
    const unsigned row_bytes = (unsigned)(a.H * a.D) * (unsigned)sizeof(T);
    const int tiles_per_bh = ws.tiles_per_bh;

    // ring of the quarter: warps q, q+4, q+8(, q+12)
    static_assert(WARPS >= 4 && WARPS <= kMaxWarps && WARPS <= 4 * kMaxWarpsPerQuarter, "barriers / private columns");
    const int ring_last = (warp & 3) + 4 * ((WARPS - 1 - (warp & 3)) / 4);
    const int ring_next = warp + 4 < WARPS ? warp + 4 : (warp & 3);
    int total_turns = 0;
    if (nlev) {
        for (int wave = 0; wave < ws.waves; ++wave) {
            int tb, te;
            wave_range(ws, wave, blockIdx.x, gridDim.x, tb, te);
            total_turns += (te - tb + WARPS - 1) / WARPS;
        }
    }
    int turn = 0;

    // Adds the quarter's accumulator to grad_img[b, :, h, :] of slice `bh` and clears it.  Whole warp, lane = channel.
    auto flush = [&](int bh) {
        const int b = bh / a.H, h = bh - b * a.H;
        float *base = gimg + ((size_t)b * a.Npix * a.H + h) * a.D + lane;
        const size_t row_elems = (size_t)a.H * a.D;
        for (int c0 = 0; c0 < used_cols; c0 += 16) {
            const int c = c0 + 16 <= used_cols ? c0 : used_cols - 16;   // last piece overlaps: re-reads cleared columns
            unsigned v[16];
            tmem_ld16(lane_base + c, v);
            tmem_wait_ld16(v);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int col = c + k;
                const int row = row0 + col - (col > pad_col && pad_col >= 0 ? 1 : 0);
                const float f = __uint_as_float(v[k]);
                if (col != pad_col && f != 0.0f) red_add_v1(base + (size_t)row * row_elems, f);
            }
            __syncwarp();   // the tcgen05 / bar instructions below are .aligned: reconverge after the predicated adds
            tmem_zero16(lane_base + c);
            tmem_wait_st();
        }
    };

    for (int wave = 0; wave < ws.waves; ++wave) {
        int t_begin, t_end;
        wave_range(ws, wave, blockIdx.x, gridDim.x, t_begin, t_end);
        const int rounds = (t_end - t_begin + WARPS - 1) / WARPS;

        int tile = t_begin + warp;
        TileUnit tu = decode_tile(tile < t_end ? tile : t_begin, tiles_per_bh, g, G, a);
        LaneOperands<T, PPL, false> op;
        float go[VEC];
        if (tile < t_end) {
            load_operands<T, LANES, LK, false, false>(a, tu, j, op);
            load_vec_stream<T, VEC>(gout + (size_t)tu.u * a.D + j * VEC, go);
        }

        for (int r = 0; r < rounds; ++r, tile += WARPS) {
            const bool valid = tile < t_end;
            const int bh = tile / tiles_per_bh;

            if (valid) {
                const int tile_n = tile + WARPS;
                const bool has_next = tile_n < t_end;
                const TileUnit tu_n = decode_tile(has_next ? tile_n : tile, tiles_per_bh, g, G, a);
                LaneOperands<T, PPL, false> op_n;
                float go_n[VEC];
                load_operands<T, LANES, LK, false, false>(a, tu_n, j, op_n);
                load_vec_stream<T, VEC>(gout + (size_t)tu_n.u * a.D + j * VEC, go_n);

                const unsigned char *__restrict__ lane_img =
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

                float part[3 * LK];
#pragma unroll
                for (int pp = 0; pp < PPL; ++pp) {
#pragma unroll
                    for (int jj0 = 0; jj0 < LANES; jj0 += NB) {
                        uint4 raw[NB][4];
                        float fx[NB], fy[NB], fw[NB];
                        unsigned o[NB][4];
                        int r00[NB], sxb[NB], syr[NB];
                        unsigned msk[NB];
                        const int pidx0 = jj0 * PPL + pp;   // point of n = 0; n = 1 is pidx0 + PPL (same level: K == 4)
                        // warp-uniform; first term is a compile-time bound that keeps the record code out of the fine batches
                        const bool coarse = pidx0 >= LK - kCP && pidx0 >= fcs;
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const int src = jj0 + n;
                            r00[n] = __shfl_sync(0xffffffffu, tap[pp].row00, src, LANES);
                            const int pack = __shfl_sync(0xffffffffu, tap[pp].pack, src, LANES);
                            fx[n] = __shfl_sync(0xffffffffu, tap[pp].dx, src, LANES);
                            fy[n] = __shfl_sync(0xffffffffu, tap[pp].dy, src, LANES);
                            fw[n] = __shfl_sync(0xffffffffu, op.wa[pp], src, LANES);
                            sxb[n] = (pack >> kPackDxBit) & 1;
                            syr[n] = pack & kPackDyMask;
                            msk[n] = BORDER ? 0xFu : (((unsigned)pack >> kPackMaskShift) & 0xFu);
                            const int rr[4] = {r00[n], r00[n] + sxb[n], r00[n] + syr[n], r00[n] + syr[n] + sxb[n]};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                o[n][c] = (unsigned)rr[c] * row_bytes;
                                raw[n][c] = gather_row(lane_img, o[n][c]);   // clamped rows: always in range
                            }
                        }
#pragma unroll
                        for (int n = 0; n < NB; ++n) {
                            const float dx = fx[n], dy = fy[n];
                            float bw[4];
                            bw[1] = (1.0f - dy) * dx;
                            bw[0] = (1.0f - dy) - bw[1];
                            bw[3] = dy * dx;
                            bw[2] = dy - bw[3];
                            float d[4], cw[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float v[VEC];
                                widen_row<T, VEC>(raw[n][c], v);
                                float acc = 0.0f;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) acc = fmaf(go[e], v[e], acc);
                                const bool ok = BORDER || ((msk[n] >> c) & 1u);
                                d[c] = ok ? acc : 0.0f;
                                cw[c] = fw[n] * bw[c];
                                if (!coarse) {
                                    float gv[VEC];
#pragma unroll
                                    for (int e = 0; e < VEC; ++e) gv[e] = go[e] * cw[c];
                                    if (live && ok)
                                        red_add_v4(reinterpret_cast<float *>(gimg_base + o[n][c]), gv[0], gv[1], gv[2], gv[3]);
                                }
                            }
                            const int pidx = pidx0 + n * PPL;
                            part[3 * pidx + 0] = bw[0] * d[0] + bw[1] * d[1] + bw[2] * d[2] + bw[3] * d[3];
                            part[3 * pidx + 1] = (1.0f - dy) * (d[1] - d[0]) + dy * (d[3] - d[2]);
                            part[3 * pidx + 2] = (1.0f - dx) * (d[2] - d[0]) + dx * (d[3] - d[1]);

                            if (coarse) {
                                // ---- the point's record: distinct columns, folded weights ----
                                unsigned m = live ? msk[n] : 0u;
#pragma unroll
                                for (int c = 0; c < 4; ++c) cw[c] = ((m >> c) & 1u) ? cw[c] : 0.0f;
                                if (sxb[n] == 0) {          // x twins: corners 01 / 11 are the rows of 00 / 10
                                    cw[0] += cw[1];
                                    cw[2] += cw[3];
                                    m = (m | (m >> 1)) & 0x5u;
                                }
                                if (syr[n] == 0) {          // y twins: corners 10 / 11 are the rows of 00 / 01
                                    cw[0] += cw[2];
                                    cw[1] += cw[3];
                                    m = (m | (m >> 2)) & 0x3u;
                                }
                                const int cp = pidx - (LK - kCP);                       // 0..7
                                const bool second = r00[n] >= off_b;                    // second coarse level (after the pad)
                                const unsigned col = (unsigned)(r00[n] - row0 + (second ? 1 : 0));
                                const unsigned trash = (cp < 4) ? kTrashA : kTrashB;    // cp < 4: first record in flight
                                const unsigned col2 = syr[n] ? col + (unsigned)syr[n] : trash;
                                CoarseRec mine;   // parked in this lane's private columns until the warp's turn
                                mine.w0 = col | (col2 << 9) | (m << 18);
#pragma unroll
                                for (int c = 0; c < 4; ++c) mine.cw[c] = __float_as_uint(cw[c]);
                                spill_store(spill_base + 5u * (unsigned)cp, mine);
                            }
                        }
                    }
                }

                // grad_out of the warp's four units in lane = channel layout for the accumulator turn (L1 hits; the
                // latency is covered by the butterfly below)
                float goc[G];
                if (nlev) {
#pragma unroll
                    for (int gg = 0; gg < G; ++gg) {
                        const TileUnit t2 = decode_tile(tile, tiles_per_bh, gg, G, a);
                        goc[gg] = __ldg(gout + (size_t)t2.u * a.D + lane);
                    }
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
                __syncwarp();   // reconverge after the predicated stores: the turn's instructions are .aligned
                if (nlev) {
                    // ================================ this warp's turn on the quarter ================================
                    tmem_wait_st();   // this tile's records are in the private columns
                    PROF_T(pt0);
                    if (!(warp < 4 && turn == 0)) ring_wait(1 + warp);
                    tc_fence_after_sync();
                    PROF_T(pt1);
                    PROF_ADD(0, pt1 - pt0);
                    // previous holder of the turn worked on another (b,h) slice: its rows go out first
                    const int prev_tile = warp >= 4 ? tile - 4 : tile - warp - WARPS + ring_last;
                    if (prev_tile >= t_begin) {
                        const int bh_prev = prev_tile / tiles_per_bh;
                        if (bh_prev != bh) flush(bh_prev);
                    }
                    // Point slots i (first level, two levels only) and 4 + i (last level) are handled together, unit by
                    // unit.  Their records reach all lanes through the broadcast buffer; for slot pair i + 1 the lanes'
                    // own records are read back during unit 0 of pair i, stored replicated after it, and read transposed
                    // after unit 1, so the round trips hide behind the read-modify-writes (every tcgen05.wait covers
                    // all earlier accesses of the warp).
                    const unsigned lo16 = lane_base, hi16 = lane_base + (16u << 16);
                    CoarseRec wa[G], wb[G], wa_n[G], wb_n[G], own_a, own_b;
                    own_a.w0 = 0u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) own_a.cw[c] = 0u;
                    if (nlev == 2) spill_load(spill_base, own_a);
                    spill_load(spill_base + 20u, own_b);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    rec_pin(own_a);
                    rec_pin(own_b);
                    if (nlev == 2) bcast_store(lane_base + kBcast, own_a);
                    bcast_store(lane_base + kBcast + 20u, own_b);
                    tmem_wait_st();
                    if (nlev == 2) {
                        bcast_load(lo16 + kBcast, wa[0], wa[1]);
                        bcast_load(hi16 + kBcast, wa[2], wa[3]);
                    }
                    bcast_load(lo16 + kBcast + 20u, wb[0], wb[1]);
                    bcast_load(hi16 + kBcast + 20u, wb[2], wb[3]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int gg = 0; gg < G; ++gg) {
                        rec_pin(wa[gg]);
                        rec_pin(wb[gg]);
                    }
                    // Stores of the previous step stay in flight while the next step loads -- unless the two steps share
                    // columns (then the stores are waited for first).  pa / pb: columns of the previous step's records.
                    unsigned pa = 1000u, pa2 = 1000u, pb = 1000u, pb2 = 1000u;
                    PROF_T(pq);
                    PROF_ADD(4, pq - pt1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i + 1 < 4) {   // own records of the next slot pair; defined after unit 0's tcgen05.wait::ld
                            if (nlev == 2) spill_load(spill_base + 5u * (unsigned)(i + 1), own_a);
                            spill_load(spill_base + 5u * (unsigned)(4 + i + 1), own_b);
                        }
#pragma unroll
                        for (int gg = 0; gg < G; ++gg) {
                            const CoarseRec ra = wa[gg], rb = wb[gg];
                            const float q = goc[gg];
                            unsigned va[4], vb[4];
                            if (nlev == 2) {
                                const bool hit = rec_conflict(ra, pa, pa2) | rec_conflict(rb, pb, pb2);
                                if (__any_sync(0xffffffffu, hit)) tmem_wait_st();
                                rec_load(ra, lane_base, va);
                                rec_load(rb, lane_base, vb);
                                tmem_wait_ld8(va, vb);
                                rec_update(ra, va, q);
                                rec_update(rb, vb, q);
                                tmem_wait_st();
                                rec_store(ra, lane_base, va);
                                rec_store(rb, lane_base, vb);
                                pa = ra.w0 & 511u;
                                pa2 = (ra.w0 >> 9) & 511u;
                            } else {
                                PROF_T(q0);
                                const bool hit = rec_conflict(rb, pb, pb2);
                                if (__any_sync(0xffffffffu, hit)) tmem_wait_st();
                                rec_load(rb, lane_base, vb);
                                tmem_wait_ld4(vb);
                                PROF_T(q1);
                                rec_update(rb, vb, q);
                                tmem_wait_st();
                                PROF_T(q2);
                                rec_store(rb, lane_base, vb);
                                PROF_T(q3);
                                PROF_ADD(5, q1 - q0);
                                PROF_ADD(6, q2 - q1);
                                PROF_ADD(7, q3 - q2);
                            }
                            pb = rb.w0 & 511u;
                            pb2 = (rb.w0 >> 9) & 511u;
                            if (gg == 0 && i + 1 < 4) {   // the buffer is free: this pair's records are in wa / wb
                                rec_pin(own_a);
                                rec_pin(own_b);
                                if (nlev == 2) bcast_store(lane_base + kBcast, own_a);
                                bcast_store(lane_base + kBcast + 20u, own_b);
                            }
                            if (gg == 1 && i + 1 < 4) {   // unit 1's tcgen05.wait::st covered the broadcast stores
                                if (nlev == 2) {
                                    bcast_load(lo16 + kBcast, wa_n[0], wa_n[1]);
                                    bcast_load(hi16 + kBcast, wa_n[2], wa_n[3]);
                                }
                                bcast_load(lo16 + kBcast + 20u, wb_n[0], wb_n[1]);
                                bcast_load(hi16 + kBcast + 20u, wb_n[2], wb_n[3]);
                            }
                        }
                        if (i + 1 < 4) {
#pragma unroll
                            for (int gg = 0; gg < G; ++gg) {   // loaded before the last units' tcgen05.wait::ld
                                rec_pin(wa_n[gg]);
                                rec_pin(wb_n[gg]);
                                wa[gg] = wa_n[gg];
                                wb[gg] = wb_n[gg];
                            }
                        }
                    }
                    tmem_wait_st();
                    PROF_T(pt2);
                    PROF_ADD(1, pt2 - pt1);
                    // nobody after me in this wave: the quarter's rows go out now
                    const int next_tile = warp + 4 < WARPS ? tile + 4 : tile - warp + WARPS + (warp & 3);
                    if (next_tile >= t_end) flush(bh);
                    PROF_T(pt3);
                    PROF_ADD(2, pt3 - pt2);
                    tc_fence_before_sync();
                    if (!(warp + 4 >= WARPS && turn == total_turns - 1)) ring_pass(1 + ring_next);
                    ++turn;
                }
            } else if (nlev) {
                // no tile in the last round of the wave: take the turn and pass it on
                if (!(warp < 4 && turn == 0)) ring_wait(1 + warp);
                if (!(warp + 4 >= WARPS && turn == total_turns - 1)) ring_pass(1 + ring_next);
                ++turn;
            }
        }
        wave_pace_cta(ws, wave);
    }

#ifdef MSDA_TMEM_PROF
    if (lane == 0 && blockIdx.x < 148) {
        prof_acc[3] = clock64() - prof_t0;
        for (int i = 0; i < 8; ++i) g_tmem_prof[(blockIdx.x * 16 + warp) * 8 + i] = prof_acc[i];
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_all(tmem_base);
}

#ifdef MSDA_TMEM_PROF
extern "C" int msda_debug_tmem_prof(long long *host_dst) {
    return (int)cudaMemcpyFromSymbol(host_dst, g_tmem_prof, sizeof(long long) * 148 * 16 * 8);
}
#endif

template <bool BORDER, int WARPS>
static cudaError_t launch_tmem_t(const KernelArgs &a, int sm_count, int max_levels, cudaStream_t st) {
    constexpr int G = 4;
    const int tiles_per_bh = (a.Q + G - 1) / G;
    const int total_tiles = a.B * a.H * tiles_per_bh;
    const int want = (total_tiles + WARPS - 1) / WARPS;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    WaveSchedule ws = make_wave_schedule(a, tiles_per_bh, sizeof(float) + sizeof(float), kTmemL2Budget);
    const bool big_waves = (long long)ws.slices_per_wave * tiles_per_bh >= 4LL * WARPS * grid;
    if (ws.waves > 1 && grid == sm_count && (big_waves || pacing_forced())) {
        const cudaError_t e = acquire_pace_counter(st, &ws.pace);
        if (e != cudaSuccess) return e;
    }
    msda_bwd_tmem_kernel<BORDER, WARPS><<<grid, WARPS * 32, 0, st>>>(a, ws, max_levels);
    return cudaGetLastError();
}

cudaError_t launch_backward_tmem(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st) {
    if (dtype != 0 || a.D != 32 || a.LK != 16 || a.K != 4 || a.L > 8) return cudaErrorNotSupported;
    if (!(a.flags & kNeedImg)) return cudaErrorNotSupported;
    if (!tiled_offsets_fit(a, sizeof(float))) return cudaErrorNotSupported;
    int max_levels = tuning().tmem_levels;
    if (max_levels < 0 || max_levels > 2) max_levels = 2;
    // 12 warps x 168 registers run the main loop as fast as 16 x 128 (the kernel is bound by the row adds, not by
    // occupancy), leave room for the turn's broadcast registers, and make rings of 3
    return a.border ? launch_tmem_t<true, 12>(a, sm_count, max_levels, st)
                    : launch_tmem_t<false, 12>(a, sm_count, max_levels, st);
}

}  // namespace msda
