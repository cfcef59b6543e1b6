// msda_tiled.cuh -- shared pieces of the tuned ("tiled") forward/backward kernels.
//
// The tuned kernels cover the shapes BASELINE.json names: L*K == 16 sampling points per unit and a pixel row of
// 64..256 bytes per head (D=32 fp32 -> 128 B -> 8 lanes x 128-bit, or 4 lanes x 256-bit in the forward; D=32 bf16/fp16
// -> 64 B -> 4 lanes).
//
// Scheduling: the grid is PERSISTENT, one CTA per SM.  Units are ordered (b, h, q) with q fastest and cut into
// contiguous ranges, one range per CTA, so at any moment every warp of an SM gathers from the SAME (b,h) slice of
// the pyramid.  For the benchmark pyramid the three coarse levels of one (b,h) slice are 168 KB and stay resident
// in that SM's L1 while the SM walks its ~2k units (measured: 66 % L1 hit rate, L2 traffic 0.92 GB instead of the
// 2.62 GB of corner rows); only the finest level streams from L2.  (The reference's grid is (q, b, h) with one tiny
// program per unit, kernels.py:365, so co-resident programs touch unrelated slices.)
//
// Each warp works on G = 32/LANES consecutive queries of one (b,h) per iteration ("warp tile").  The sampling
// points / weights of the NEXT warp tile are fetched before the current one is processed, so the only exposed
// latency per iteration is the gathers themselves.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "msda_common.cuh"
#include "msda_tuning.h"

namespace msda {

template <typename T, int LANES, int LK, int VECB = 16> struct TiledCfg {
    static constexpr int VEC = VECB / (int)sizeof(T);   // elements per lane load (128-bit; 256-bit with VECB = 32)
    static constexpr int G = 32 / LANES;              // units per warp iteration
    static constexpr int PPL = LK / LANES;            // sampling points resolved by each lane
    static_assert(LANES * PPL == LK, "LK must be a multiple of LANES");
    static_assert(32 % LANES == 0, "LANES must divide the warp");
};

// Host-side eligibility of the 32-bit offset arithmetic below (shapes live on the device, so bound w by Npix).
inline bool tiled_offsets_fit(const KernelArgs &a, size_t elem_size, int subs = 1) {
    const unsigned long long row_bytes = (unsigned long long)a.H * a.D * elem_size;
    const unsigned long long tiles = (unsigned long long)a.B * a.H * a.Q * subs;  // upper bound on the warp-tile count
    return (unsigned long long)a.Npix * row_bytes < (1ull << 28) && tiles < (1ull << 31);
}

// L2-sized waves.  With many images (B=64 encoder training: 512 (b,h) slices of 2.8 MB) a plain contiguous split
// would have every SM on a different slice and the concurrently gathered set (148 slices, 420 MB) thrashes the
// 126 MB L2.  The (b,h) slices are therefore processed in waves of `slices_per_wave` slices whose pyramid (and, in
// the backward, grad_img) bytes fit L2 together; inside a wave the warp tiles are split contiguously over the CTAs
// exactly as before, so every SM still sees one (b,h) slice at a time in its L1.
struct WaveSchedule {
    int tiles_per_bh;      // warp tiles per (b,h) slice
    int slices;            // B*H
    int slices_per_wave;
    int waves;
    unsigned *pace;        // arrival counter of this launch (zeroed on the stream before it), nullptr = no pacing
    int pace_slack;        // waves a CTA may run ahead of the slowest CTA
};

// Wave pacing.  The CTAs of the persistent grid walk the waves independently; over many waves (B=64: 64 waves) the
// faster ones drift ahead, the set of images gathered from at the same time grows past L2, and the laggards -- whose
// image is being evicted by the leaders -- fall back further (measured on the B=64 encoder shape: 6x the algorithmic
// DRAM traffic in the backward, +28 % time per image against B=16).  So a CTA announces every wave it has completed
// on a counter, and no warp starts a wave while any CTA of the grid is more than `pace_slack` waves behind it.
//   * Forward (wave_pace_warp): no CTA-wide barrier -- the warps of a CTA count their arrivals per wave in shared
//     memory (ring of kPaceRing counters), the last one publishes the CTA's arrival, and each warp polls the global
//     counter on its own (one L2 read per wave in the common case).  A __syncthreads() per wave costs the forward
//     10 % (the warps' software pipelines drain and refill together): B=16 76 -> 87 us per image.
//   * Backward (wave_pace_cta): the CTA does synchronise at the end of a wave; keeping its warps on ONE (b,h) slice
//     is worth more than the drained pipelines there (B=16: 270 -> 258 us per image, B=64: 351 -> 258 us; the
//     warp-level variant only reaches 311 us at B=64).
//   * This is a PERFORMANCE HINT, not a correctness barrier: the wait is bounded (kPaceTimeoutCycles) and the first
//     time-out switches the pacing off for the rest of the launch, so a grid that is not fully co-resident (another
//     kernel holding SMs, a partitioned GPU) only loses the pacing and ~130 us; it cannot deadlock.
constexpr long long kPaceTimeoutCycles = 1ll << 18;   // ~130 us at 1.97 GHz
constexpr int kPaceRing = 8;                          // > pace_slack + 2 waves between any two warps of a CTA
constexpr int kPaceMaxSlack = 4;

// Bounded wait until every CTA of the grid has completed `wave + 1 - pace_slack` waves.
__device__ __forceinline__ void pace_wait(const WaveSchedule &w, int wave) {
    const int must_have_finished = wave + 1 - w.pace_slack;
    if (must_have_finished <= 0) return;
    const unsigned target = (unsigned)must_have_finished * gridDim.x;
    const long long t0 = clock64();
    while (true) {
        unsigned seen;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(w.pace) : "memory");
        if (seen >= target) break;
        if (clock64() - t0 > kPaceTimeoutCycles) {
            // the grid is evidently not co-resident (or a CTA is badly delayed): switch the pacing off for the rest of
            // this launch -- the top bit makes every later comparison succeed at once
            atomicOr(w.pace, 0x80000000u);
            break;
        }
        __nanosleep(256);
    }
}

// s_arrivals: kPaceRing zero-initialised shared counters.  Called by every warp (all lanes) after its last tile of `wave`.
__device__ __forceinline__ void wave_pace_warp(const WaveSchedule &w, int wave, unsigned *s_arrivals, int lane,
                                               int nwarps) {
    if (w.pace == nullptr || wave + 1 >= w.waves) return;   // uniform across the grid
    if (lane == 0) {
        unsigned *mine = s_arrivals + (wave % kPaceRing);
        if (atomicAdd(mine, 1u) == (unsigned)nwarps - 1u) {   // last warp of this CTA to finish the wave
            *mine = 0;                                        // reused kPaceRing waves later
            atomicAdd(w.pace, 1u);
        }
        pace_wait(w, wave);
    }
    __syncwarp();
}

// Called by every thread of the CTA after the last tile of `wave`.
__device__ __forceinline__ void wave_pace_cta(const WaveSchedule &w, int wave) {
    if (w.pace == nullptr || wave + 1 >= w.waves) return;   // uniform across the grid
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(w.pace, 1u);
        pace_wait(w, wave);
    }
    __syncthreads();
}

// min_tiles_per_wave: with few queries per image (decoder: 900 queries against a 22k-pixel pyramid) one image is less
// than one tile per warp and the persistent grid idles at every wave boundary; a wave is then grown, up to
// kWaveHardCapBytes of L2, until it holds that many tiles (B=8 x Q=900 fp32 forward: 78.8 -> 53 us).
constexpr unsigned long long kWaveHardCapBytes = 96ull << 20;

inline WaveSchedule make_wave_schedule(const KernelArgs &a, int tiles_per_bh, size_t elem_size, size_t l2_budget_bytes,
                                       long long min_tiles_per_wave = 0) {
    WaveSchedule w;
    w.tiles_per_bh = tiles_per_bh;
    w.slices = a.B * a.H;
    // A wave is a whole number of IMAGES (all H head slices of a pixel share one H*D*e-byte span of memory, so a
    // wave then reads full DRAM pages); measured on the B=64 encoder shape, waves of 8 = H slices run the forward
    // in 6.7 ms versus 9.7 ms for 32 slices and 11.7 ms without waves.
    const unsigned long long image_bytes = (unsigned long long)a.Npix * a.H * a.D * elem_size;
    long long images = (long long)(l2_budget_bytes / (image_bytes ? image_bytes : 1));
    if (images < 1) images = 1;
    while (images * a.H * tiles_per_bh < min_tiles_per_wave && images < a.B &&
           (unsigned long long)(images + 1) * image_bytes <= kWaveHardCapBytes)
        ++images;
    long long max_slices = images * a.H;
    if (tuning().slices_per_wave > 0) max_slices = tuning().slices_per_wave;   // tuning knob
    if (max_slices > w.slices) max_slices = w.slices;
    w.waves = (int)((w.slices + max_slices - 1) / max_slices);
    w.slices_per_wave = (int)max_slices;
    w.pace = nullptr;
    const int slack = tuning().pace_slack;   // tuning knob
    w.pace_slack = slack < 0 ? 0 : (slack > kPaceMaxSlack ? kPaceMaxSlack : slack);
    return w;
}

// [begin, end) of the warp tiles CTA `cta` of `ctas` owns in wave `wave`.
__device__ __forceinline__ void wave_range(const WaveSchedule &w, int wave, int cta, int ctas, int &begin, int &end) {
    const int s0 = wave * w.slices_per_wave;
    const int ns = min(w.slices_per_wave, w.slices - s0);
    const long long wt = (long long)ns * w.tiles_per_bh;
    const long long base = (long long)s0 * w.tiles_per_bh;
    begin = (int)(base + wt * cta / ctas);
    end = (int)(base + wt * (cta + 1) / ctas);
}

// Decodes a warp tile (G consecutive queries of one (b,h)) into this lane-group's unit.
struct TileUnit {
    long long u;      // unit index (b*Q + q)*H + h of the (possibly shadowed) query
    size_t bh_off;    // element offset of img[b, 0, h, 0]
    int p0;           // first sampling point of this tile's sub-unit (0 unless the unit is split, see below)
    bool live;        // false for the padding queries of the last tile of a (b,h)
};

// `subs` > 1 (backward only): a unit with more sampling points than the instantiation has slots is processed as
// `subs` SUB-UNITS of `slots` points each -- the backward has no reduction across points, so the sub-units only share
// the grad_out row.  A warp tile belongs to ONE sub-unit (so the dead slots of a padded instantiation are uniform
// across the warp) and the sub-units of a query tile are consecutive tiles, i.e. they run on neighbouring warps of
// the CTA at the same time: the row adds into the few, hot rows of the coarsest level (which all sit in the last
// sub-unit) stay interleaved with the rest of the traffic.  (Sub-unit-major order, all queries of sub-unit 0 first,
// measured 0.81 ms against 0.75 ms for the generic kernel on a 5-level pyramid: same-address serialisation in L2.)
__device__ __forceinline__ TileUnit decode_tile(int tile, int tiles_per_bh, int g, int G, const KernelArgs &a,
                                                int subs = 1, int slots = 0) {
    const int bh = tile / tiles_per_bh;
    int qt = tile - bh * tiles_per_bh;
    int sub = 0;
    if (subs > 1) {
        const int qt_full = qt;
        qt = qt_full / subs;
        sub = qt_full - qt * subs;
    }
    const int b = bh / a.H;
    const int h = bh - b * a.H;
    const int q_raw = qt * G + g;
    TileUnit t;
    t.live = q_raw < a.Q;
    t.p0 = sub * slots;
    const int q = t.live ? q_raw : a.Q - 1;
    t.u = ((long long)b * a.Q + q) * a.H + h;
    t.bh_off = ((size_t)b * a.Npix * a.H + h) * a.D;
    return t;
}

// Level of point slot p (dead slots of padded instantiations fall back to the last level; their weight is 0).
__device__ __forceinline__ int slot_level(int p, const KernelArgs &a) {
    const int l = p / a.K;
    return l < a.L ? l : a.L - 1;
}

// One resolved sampling point as exchanged between the lanes of a group (4 registers + the attention weight).
//   off  : byte offset of the (y0,x0) corner row inside the (b,h) slice           (< 2^28, see tiled_offsets_fit)
//   pack : bits 0..23 = byte step to the y1 rows / 16, bit 24 = x1 differs from x0, bits 25..28 = corner validity
struct TileTap {
    unsigned off;
    unsigned pack;
    float dx, dy;
};

template <bool BORDER>
__device__ __forceinline__ TileTap resolve_tap(float px, float py, const Level lv, bool align, unsigned row_bytes) {
    const Tap<float> t = locate<float>(px, py, lv, BORDER, align);
    TileTap r;
    r.off = (unsigned)t.row00 * row_bytes;
    const unsigned step_rows = (unsigned)t.pack & (unsigned)kPackDyMask;
    r.pack = ((step_rows * row_bytes) >> 4) | ((unsigned)t.pack & ~(unsigned)kPackDyMask);
    r.dx = t.dx;
    r.dy = t.dy;
    return r;
}

// Byte offsets of the four corner rows from an exchanged tap.
__device__ __forceinline__ void corner_offsets(unsigned off, unsigned pack, unsigned row_bytes, unsigned (&o)[4]) {
    const unsigned sx = (pack & (1u << kPackDxBit)) ? row_bytes : 0u;
    const unsigned sy = (pack & (unsigned)kPackDyMask) << 4;
    o[0] = off;
    o[1] = off + sx;
    o[2] = off + sy;
    o[3] = off + sy + sx;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-lane operands of one unit: either the materialised sampling points / attention weights (the operator of
// frontend.py:145) or, for the fused module core, the raw query projection and the reference point, from which the
// lane group derives them on the fly (frontend.py:253-284: softmax over L*K, ref + offset / level shape).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kNeedRef = 16;   // msda_module_backward: grad of the reference points requested

// Elements per aligned chunk of a lane's (offset x, offset y, logit) triples in the fused module core.
template <typename T, int PPL> struct FusedChunk {
    static constexpr int kBytes = ((3 * PPL * (int)sizeof(T)) % 8 == 0) ? 8 : 4;
    static constexpr int kElems = kBytes / (int)sizeof(T);
    static_assert(kElems >= 1 && (3 * PPL) % kElems == 0, "unsupported lane layout for the fused module core");
};

template <typename T, int PPL, bool FUSED> struct LaneOperands {
    float xy[2 * PPL];    // sampling points of this lane's PPL points (x, y)
    float wa[PPL];        // attention weights
    float raw[FUSED ? 3 * PPL : 1];   // FUSED: (offset x, offset y, logit) triples as loaded
    float ref[FUSED ? 4 : 1];         // FUSED: reference point of the unit's query
};

// Issues the (streaming) loads for unit `tu` -- nothing here depends on the loaded values, so the call can sit one
// warp tile ahead of its use.  LK is the number of point SLOTS of the instantiation (LANES * PPL); with PADDED the
// unit really has a.LK <= LK points, the remaining slots are dead (weight 0, never gathered).
template <typename T, int LANES, int LK, bool FUSED, bool PADDED = false>
__device__ __forceinline__ void load_operands(const KernelArgs &a, const TileUnit &tu, int j,
                                              LaneOperands<T, LK / LANES, FUSED> &op) {
    constexpr int PPL = LK / LANES;
    if constexpr (!FUSED) {
        const T *__restrict__ pts = static_cast<const T *>(a.pts) + (size_t)tu.u * a.LK * 2;
        const T *__restrict__ aw = static_cast<const T *>(a.aw) + (size_t)tu.u * a.LK;
        if constexpr (!PADDED) {
            load_vec_stream<T, 2 * PPL>(pts + (tu.p0 + j * PPL) * 2, op.xy);
            load_vec_stream<T, PPL>(aw + tu.p0 + j * PPL, op.wa);
        } else {
#pragma unroll
            for (int pp = 0; pp < PPL; ++pp) {
                const int p = tu.p0 + j * PPL + pp;
                float xy2[2] = {0.0f, 0.0f}, w1[1] = {0.0f};
                if (p < a.LK) {
                    load_vec_stream<T, 2>(pts + 2 * p, xy2);
                    load_vec_stream<T, 1>(aw + p, w1);
                }
                op.xy[2 * pp] = xy2[0];
                op.xy[2 * pp + 1] = xy2[1];
                op.wa[pp] = w1[0];
            }
        }
    } else {
        // a lane's 3*PPL elements start at a multiple of 3*PPL*sizeof(T) bytes: 8-byte chunks when that is a multiple
        // of 8 (D = 32), else 4-byte chunks (D = 64: 12 bytes per lane)
        constexpr int E8 = FusedChunk<T, PPL>::kElems;
        static_assert(!PADDED, "the fused module core is instantiated for exact L*K only");
        const T *__restrict__ proj = static_cast<const T *>(a.proj) + ((size_t)tu.u * LK + j * PPL) * 3;
#pragma unroll
        for (int c = 0; c < 3 * PPL / E8; ++c) {
            float tmp[E8];
            load_vec_stream<T, E8>(proj + c * E8, tmp);
#pragma unroll
            for (int e = 0; e < E8; ++e) op.raw[c * E8 + e] = tmp[e];
        }
        const T *__restrict__ ref = static_cast<const T *>(a.ref) + (size_t)(tu.u / a.H) * a.ref_dim;
        float r2[2];
        load_vec<T, 2>(ref, r2);
        op.ref[0] = r2[0];
        op.ref[1] = r2[1];
        op.ref[2] = op.ref[3] = 0.0f;
        if (a.ref_dim == 4) {
            load_vec<T, 2>(ref + 2, r2);
            op.ref[2] = r2[0];
            op.ref[3] = r2[1];
        }
    }
}

// FUSED only: turns (raw, ref) into (xy, wa).  Softmax runs over the LK logits of the unit, which are spread over the
// LANES lanes of the group (PPL each): two butterfly reductions.  Same arithmetic order as the torch module in fp32:
// exp(x - max) / sum;  ref + off / shape  (x is divided by the level HEIGHT and y by the WIDTH, as the reference
// does, frontend.py:272-276);  ref_xy + off * ref_wh / (2 K)  for 4-d reference points (frontend.py:278-282).
template <typename T, int LANES, int LK>
__device__ __forceinline__ void derive_operands(const KernelArgs &a, const Level *s_lv, int j,
                                                LaneOperands<T, LK / LANES, true> &op) {
    constexpr int PPL = LK / LANES;
    float m = op.raw[2];
#pragma unroll
    for (int pp = 1; pp < PPL; ++pp) m = fmaxf(m, op.raw[3 * pp + 2]);
#pragma unroll
    for (int s = LANES / 2; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    float sum = 0.0f;
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        op.wa[pp] = expf(op.raw[3 * pp + 2] - m);
        sum += op.wa[pp];
    }
#pragma unroll
    for (int s = LANES / 2; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
#pragma unroll
    for (int pp = 0; pp < PPL; ++pp) {
        op.wa[pp] = __fdiv_rn(op.wa[pp], sum);
        const Level lv = s_lv[(j * PPL + pp) / a.K];
        if (a.ref_dim == 2) {
            op.xy[2 * pp + 0] = op.ref[0] + __fdiv_rn(op.raw[3 * pp + 0], (float)lv.h);
            op.xy[2 * pp + 1] = op.ref[1] + __fdiv_rn(op.raw[3 * pp + 1], (float)lv.w);
        } else {
            const float two_k = (float)(2 * a.K);
            op.xy[2 * pp + 0] = op.ref[0] + __fdiv_rn(__fmul_rn(op.raw[3 * pp + 0], op.ref[2]), two_k);
            op.xy[2 * pp + 1] = op.ref[1] + __fdiv_rn(__fmul_rn(op.raw[3 * pp + 1], op.ref[3]), two_k);
        }
    }
}

// Deterministic backward by exact row adds (msda_bwd_detq.cu): the quantum of (b, h, level) entry `idx`.
//   R = slmax[idx]: the largest row counter of the slice-level; a row's counter is the sum of t_i = ceil(|c_i| / U) + 1 over
//   its contributions c_i, U = amax / 4096, hence  sum |c_i| <= U (R - n)  for a row with n contributions.
//   The kernel rounds c to c' = a multiple of q with |c' - c| <= q/2 (<= q for the at most three contributions per row
//   that exceed 2^22 q, where fp32 spacing is 2q).  With q <= U/2, which R <= 2^22 guarantees:
//       sum |c'_i| <= U (R - n) + n q/2 + 3 q/2 <= U R        (n >= 3; two addends commute whatever they are)
//   and  q = 2^(ilogb(U R) + 1 - 24) > U R 2^-24  keeps every partial sum of every row, in any order, an integer multiple
//   of q below 2^24 q: exactly representable, no add ever rounds.  Rows beyond R = 2^22 -- 10^4 and more contributions
//   of full size -- use the cruder  sum |c'| <= 2 sum |c|  and one more bit of head room.
// Returns 0 ("do not quantise") when nothing lands on the slice-level or the inputs are not finite.
__device__ __forceinline__ float row_quantum(const KernelArgs &a, int idx) {
    const unsigned long long r = a.q_slmax[idx];
    const float amax = __uint_as_float(a.q_amax[0]) * __uint_as_float(a.q_amax[1]);
    if (r == 0ull || !(amax > 0.0f) || !(amax < 3.0e38f)) return 0.0f;
    const float bound = (float)r * (amax * (1.0f / 4096.0f));
    if (!(bound > 1.0e-30f) || !(bound < 1.0e30f)) return 0.0f;
    return scalbnf(1.0f, ilogbf(bound) + (r <= (1ull << 22) ? 1 : 2) - 24);
}

// Read-only gather of one lane's slice of a corner row (16 bytes, or 8 bytes for the 16-bit-storage backward that runs
// 8 lanes x 4 channels), kept raw until it is consumed.
template <int BYTES> struct RawSlice;
template <> struct RawSlice<16> { using type = uint4; };
template <> struct RawSlice<8> { using type = uint2; };
struct alignas(32) Raw256 { uint4 lo, hi; };
template <> struct RawSlice<32> { using type = Raw256; };   // sm_100: LDG.E.256

// Plain read-only gather of one lane's slice of a corner row.
template <int BYTES>
__device__ __forceinline__ typename RawSlice<BYTES>::type gather_slice(const unsigned char *__restrict__ lane_base,
                                                                       unsigned byte_off) {
    if constexpr (BYTES == 32) {
        Raw256 r;
        asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
            : "l"(lane_base + byte_off));
        return r;
    } else {
        return __ldg(reinterpret_cast<const typename RawSlice<BYTES>::type *>(lane_base + byte_off));
    }
}

// The same gather served through L1 WITHOUT being written into it (LDG.E.NA).  Used for the pyramid levels that cannot
// stay resident in L1 anyway (see streamed_points): their lines then neither spend an L1 fill wavefront nor evict the
// coarse levels that do fit.  The choice is a COMPILE-TIME property of the point slot (template parameter of the
// kernels): as a run-time branch the two flavours were joined by one predicated MOV per loaded register (20 % of the
// forward's instructions), and as a predicated pair in one asm statement the second load waits for the first one's
// destination registers (7x slower).
template <int BYTES>
__device__ __forceinline__ typename RawSlice<BYTES>::type gather_slice_na(const unsigned char *__restrict__ lane_base,
                                                                          unsigned byte_off) {
    if constexpr (BYTES == 32) {
        Raw256 r;
        asm("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
            : "l"(lane_base + byte_off));
        return r;
    } else if constexpr (BYTES == 16) {
        uint4 r;
        asm("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
            : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
            : "l"(lane_base + byte_off));
        return r;
    } else {
        uint2 r;
        asm("ld.global.nc.L1::no_allocate.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(lane_base + byte_off));
        return r;
    }
}

__device__ __forceinline__ uint4 gather_row(const unsigned char *__restrict__ lane_base, unsigned byte_off) {
    return gather_slice<16>(lane_base, byte_off);
}

template <typename T> __device__ __forceinline__ void widen_word(unsigned r, float &lo, float &hi) {
    if constexpr (std::is_same<T, __half>::value) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r));
        lo = f.x;
        hi = f.y;
    } else {
        lo = __uint_as_float(r << 16);            // bf16 -> fp32 is a 16-bit shift
        hi = __uint_as_float(r & 0xffff0000u);
    }
}

// Widens a raw slice to fp32: VEC = 4 (fp32 x 16 B, or 16-bit x 8 B) or 8 (16-bit x 16 B).
template <typename T, int VEC> __device__ __forceinline__ void widen_row(const uint4 raw, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(raw.x);
        v[1] = __uint_as_float(raw.y);
        v[2] = __uint_as_float(raw.z);
        v[3] = __uint_as_float(raw.w);
    } else {
        widen_word<T>(raw.x, v[0], v[1]);
        widen_word<T>(raw.y, v[2], v[3]);
        widen_word<T>(raw.z, v[4], v[5]);
        widen_word<T>(raw.w, v[6], v[7]);
    }
}
template <typename T, int VEC> __device__ __forceinline__ void widen_row(const Raw256 raw, float (&v)[VEC]) {
    static_assert(VEC * sizeof(T) == 32, "32-byte slices: eight fp32 or sixteen 16-bit channels");
    constexpr int HALF = VEC / 2;
    float lo[HALF], hi[HALF];
    widen_row<T, HALF>(raw.lo, lo);
    widen_row<T, HALF>(raw.hi, hi);
#pragma unroll
    for (int e = 0; e < HALF; ++e) {
        v[e] = lo[e];
        v[HALF + e] = hi[e];
    }
}
template <typename T, int VEC> __device__ __forceinline__ void widen_row(const uint2 raw, float (&v)[VEC]) {
    static_assert(sizeof(T) == 2 && VEC == 4, "8-byte slices hold four 16-bit channels");
    widen_word<T>(raw.x, v[0], v[1]);
    widen_word<T>(raw.y, v[2], v[3]);
}

// Transposing butterfly over the LANES lanes of a group (backward kernels): every step halves the values a lane keeps,
// so reducing N partials costs N*(1 - 1/LANES) shuffles instead of N*log2(LANES), and lane j ends with exactly the
// points it loaded in part[0 .. N/LANES).  Call with STEP = LANES / 2.
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

// How many LEADING point slots (whole levels, finest first) the forward gathers with no-allocate loads.  The level
// shapes live on the device, so the host models the pyramid from Npix as levels shrinking 4x each (exact for the
// benchmark pyramid, within 5 % for 800x1333 at strides 8..64); walking up from the coarsest level, levels are kept
// while the rows of one (b,h) slice of all kept levels fit `keep_bytes` (default 100 KB) of L1: benchmark pyramid, fp32:
// 8x8 + 16x16 = 40 KB kept, 32x32 and 64x64 streamed; DETR pyramid: 13x21 kept.  Measured forward, keep = off / 40 KB /
// 176 KB: bench 0.143 / 0.115 / 0.119 ms, DETR encoder 0.166 / 0.141 / 0.145 ms, DETR with local sampling points
// 0.176 / 0.141 / 0.139 ms.  A wrong guess only costs performance.  Result in {0, K, 2K, ...}.
inline int streamed_points(const KernelArgs &a, size_t row_bytes_per_head) {
    const long long keep = tuning().l1_keep_kb < 0 ? -1 : (long long)tuning().l1_keep_kb * 1024;   // tuning knob
    if (keep < 0 || a.L < 1) return 0;
    double norm = 0.0, w = 1.0;
    for (int l = 0; l < a.L; ++l, w *= 0.25) norm += w;
    int first_kept = a.L;
    double cum = 0.0;
    for (int l = a.L - 1; l >= 0; --l) {
        double frac = 1.0;
        for (int i = 0; i < l; ++i) frac *= 0.25;
        cum += (double)a.Npix * frac / norm * (double)row_bytes_per_head;
        if (cum > (double)keep) break;
        first_kept = l;
    }
    return first_kept * a.K;
}

}  // namespace msda
