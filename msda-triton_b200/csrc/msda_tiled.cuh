// msda_tiled.cuh -- shared pieces of the tuned ("tiled") forward/backward kernels.
//
// The tuned kernels cover the shapes BASELINE.json names: L*K == 16 sampling points per unit and a pixel row of
// 64..256 bytes per head (D=32 fp32 -> 128 B -> 8 lanes x 128-bit; D=32 bf16/fp16 -> 64 B -> 4 lanes).
//
// Scheduling: the grid is PERSISTENT, one CTA per SM.  Units are ordered (b, h, q) with q fastest and cut into
// contiguous ranges, one range per CTA, so at any moment every warp of an SM gathers from the SAME (b,h) slice of
// the pyramid.  For the benchmark pyramid the three coarse levels of one (b,h) slice are 168 KB and stay resident
// in that SM's L1 while the SM walks its ~2k units; only the finest level streams from L2.  (The reference's grid
// is (q, b, h) with one tiny program per unit, kernels.py:365, so co-resident programs touch unrelated slices.)
#pragma once
#include <type_traits>

#include "msda_common.cuh"

namespace msda {

constexpr int kTiledThreads = 512;

template <typename T, int LANES, int LK> struct TiledCfg {
    static constexpr int VEC = 16 / (int)sizeof(T);   // elements per 128-bit lane load
    static constexpr int G = 32 / LANES;              // units per warp iteration
    static constexpr int PPL = LK / LANES;            // sampling points resolved by each lane
    static_assert(LANES * PPL == LK, "LK must be a multiple of LANES");
    static_assert(32 % LANES == 0, "LANES must divide the warp");
};

// Decodes a warp tile (G consecutive queries of one (b,h)) into this lane-group's unit.
struct TileUnit {
    long long u;      // unit index (b*Q + q)*H + h of the (possibly shadowed) query
    size_t bh_off;    // element offset of img[b, 0, h, 0]
    bool live;        // false for the padding queries of the last tile of a (b,h)
};

__device__ __forceinline__ TileUnit decode_tile(long long tile, int tiles_per_bh, int g, int G, const KernelArgs &a) {
    const long long bh = tile / tiles_per_bh;
    const int qt = (int)(tile - bh * tiles_per_bh);
    const int b = (int)(bh / a.H);
    const int h = (int)(bh - (long long)b * a.H);
    const int q_raw = qt * G + g;
    TileUnit t;
    t.live = q_raw < a.Q;
    const int q = t.live ? q_raw : a.Q - 1;
    t.u = ((long long)b * a.Q + q) * a.H + h;
    t.bh_off = ((size_t)b * a.Npix * a.H + h) * a.D;
    return t;
}

// 128-bit read-only gather of one corner row slice (VEC storage elements, kept raw until they are consumed).
template <typename T> __device__ __forceinline__ uint4 gather_row(const T *__restrict__ p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

// Widens the raw 128 bits to fp32 (VEC = 4 for fp32 storage, 8 for fp16 / bf16 storage).
template <typename T, int VEC> __device__ __forceinline__ void widen_row(const uint4 raw, float (&v)[VEC]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(raw.x);
        v[1] = __uint_as_float(raw.y);
        v[2] = __uint_as_float(raw.z);
        v[3] = __uint_as_float(raw.w);
    } else {
        const unsigned r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __half>::value) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r[i]));
                v[2 * i] = f.x;
                v[2 * i + 1] = f.y;
            } else {
                // bf16 -> fp32 is a 16-bit shift
                v[2 * i] = __uint_as_float(r[i] << 16);
                v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
            }
        }
    }
}

}  // namespace msda
