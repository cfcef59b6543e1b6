// msda_common.cuh -- shared device helpers for the sm_100a MSDA kernels.
//
// Semantics follow the reference's device helpers (rziga/msda-triton, src/msda_triton/kernels.py):
//   level table ............ kernels.py:44-64   (exclusive prefix sum of h*w, done ONCE per CTA here)
//   coordinate math ........ kernels.py:139-169 (un-normalise, floor, zeros-mode validity, clamp-in-float-then-int)
//   addressing ............. kernels.py:180-203
// The code is written from scratch for CUDA; nothing here is a translation of the Triton source.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

// ---------------------------------------------------------------------------------------------------------------
// Kernel argument block (passed by value; 64-bit sizes where products can exceed 2^31: B=64 encoder tensors).
// ---------------------------------------------------------------------------------------------------------------
struct KernelArgs {
    const void *img;           // [B, Npix, H, D]
    const long long *shapes;   // [L, 2] int64 (h, w), device
    const void *pts;           // [B, Q, H, L, K, 2]
    const void *aw;            // [B, Q, H, L, K]
    void *out;                 // fwd: [B, Q, H, D]
    const void *gout;          // bwd: [B, Q, H, D]
    void *gimg;                // bwd: accumulation image (float for f32/f16/bf16 storage, double for f64)
    void *gpts;                // bwd: [B, Q, H, L, K, 2]
    void *gaw;                 // bwd: [B, Q, H, L, K]
    // fused module core (frontend.py:253-289): raw projection + reference points instead of pts / aw
    const void *proj;          // [B, Q, H, L, K, 3]  (offset x, offset y, attention logit)
    const void *ref;           // [B, Q, ref_dim]     reference points, ref_dim = 2 (x,y) or 4 (cx,cy,w,h)
    void *gproj;               // bwd: [B, Q, H, L, K, 3]
    float *gref;               // bwd: [B, Q, ref_dim] fp32, zero-filled by the library, accumulated with atomics
    int ref_dim;
    // deterministic backward by exact (quantised) row adds, msda_bwd_detq.cu: per (b, h, level) the largest row bound
    // in units of amax / 4096, and {max |grad_out|, max |attention weight|} as float bits
    const unsigned long long *q_slmax;
    const unsigned *q_amax;
    long long units;           // B*Q*H  (one unit = one output row (b,q,h))
    int B, Q, H, D, L, K, Npix;
    int LK;                    // L*K
    int lanes;                 // lanes cooperating on one unit (power of two, <= 32)
    int chunks;                // ceil(D / (lanes*VEC))
    int border;                // padding_mode == border
    int align;                 // align_corners
    int flags;                 // msda_bwd_flags
};

// One pyramid level as the kernels see it: {h, w, first pixel row, unused}.
struct __align__(16) Level {
    int h, w, off, pad;
};

// Builds the level table in shared memory once per CTA (the reference redoes this in every one of its B*Q*H
// programs, kernels.py:290).  Works for any L: a single thread runs the exclusive prefix sum, L is tiny.
// Returns false (for every thread of the CTA) when the device-side shapes do not describe the pyramid the caller
// announced -- a level with h <= 0 or w <= 0, or sum h*w > Npix (fewer rows than the image holds are harmless).  The reference trusts the table blindly
// (kernels.py:52-62); here a mismatch would turn into out-of-bounds gathers and row adds, so the kernels return
// without touching memory instead (the check is one compare per CTA; s_lv[0].pad carries the verdict).
__device__ __forceinline__ bool build_level_table(Level *s_lv, const long long *__restrict__ shapes, int L, int Npix) {
    if (threadIdx.x == 0) {
        long long run = 0;
        bool ok = true;
        for (int l = 0; l < L; ++l) {
            const long long h = shapes[2 * l + 0];
            const long long w = shapes[2 * l + 1];
            ok = ok && h > 0 && w > 0 && h <= Npix && w <= Npix;
            s_lv[l].h = (int)h;
            s_lv[l].w = (int)w;
            s_lv[l].off = (int)run;
            s_lv[l].pad = 0;
            run += ok ? h * w : 0;
        }
        s_lv[0].pad = (ok && run <= (long long)Npix) ? 1 : 0;
    }
    __syncthreads();
    return s_lv[0].pad != 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Type traits: storage type T -> compute type CT (fp32 for f32/f16/bf16 storage, fp64 for f64), grad_img
// accumulation type (== CT), and vector packs.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct Traits;
template <> struct Traits<float> {
    using CT = float;
    static constexpr int kMaxVec = 4;
    static __device__ __forceinline__ float to_ct(float v) { return v; }
    static __device__ __forceinline__ float from_ct(float v) { return v; }
};
template <> struct Traits<double> {
    using CT = double;
    static constexpr int kMaxVec = 2;
    static __device__ __forceinline__ double to_ct(double v) { return v; }
    static __device__ __forceinline__ double from_ct(double v) { return v; }
};
template <> struct Traits<__half> {
    using CT = float;
    static constexpr int kMaxVec = 8;
    static __device__ __forceinline__ float to_ct(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_ct(float v) { return __float2half_rn(v); }
};
template <> struct Traits<__nv_bfloat16> {
    using CT = float;
    static constexpr int kMaxVec = 8;
    static __device__ __forceinline__ float to_ct(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_ct(float v) { return __float2bfloat16_rn(v); }
};

template <typename T, int N> struct __align__(sizeof(T) * N) Pack {
    T v[N];
};

// Vector load of N storage elements (N*sizeof(T) in {2,4,8,16} bytes, pointer aligned to that), widened to CT.
template <typename T, int N>
__device__ __forceinline__ void load_vec(const T *__restrict__ p, typename Traits<T>::CT (&dst)[N]) {
    const Pack<T, N> raw = *reinterpret_cast<const Pack<T, N> *>(p);
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = Traits<T>::to_ct(raw.v[i]);
}

template <typename T, int N>
__device__ __forceinline__ void store_vec(T *__restrict__ p, const typename Traits<T>::CT (&src)[N]) {
    Pack<T, N> raw;
#pragma unroll
    for (int i = 0; i < N; ++i) raw.v[i] = Traits<T>::from_ct(src[i]);
    *reinterpret_cast<Pack<T, N> *>(p) = raw;
}

// Streaming ("cache streaming", evict-first) variants for operands that are touched exactly once per launch --
// sampling points, attention weights, grad_out, out, grad_points, grad_weights -- so they neither displace the
// pyramid rows in L1 nor, for large batches, in L2 (ld.global.cs / st.global.cs).
template <int BYTES> struct RawWord;
template <> struct RawWord<2> { using type = unsigned short; };
template <> struct RawWord<4> { using type = unsigned; };
template <> struct RawWord<8> { using type = uint2; };
template <> struct RawWord<16> { using type = uint4; };

template <typename T, int N>
__device__ __forceinline__ void load_vec_stream(const T *__restrict__ p, typename Traits<T>::CT (&dst)[N]) {
    constexpr int kBytes = (int)sizeof(T) * N;
    if constexpr (kBytes > 16) {
        // wider than one 128-bit access: consecutive 16-byte pieces
        constexpr int E16 = 16 / (int)sizeof(T);
        static_assert(N % E16 == 0, "wide streaming loads must be a whole number of 16-byte pieces");
#pragma unroll
        for (int c = 0; c < N / E16; ++c) {
            typename Traits<T>::CT piece[E16];
            load_vec_stream<T, E16>(p + c * E16, piece);
#pragma unroll
            for (int e = 0; e < E16; ++e) dst[c * E16 + e] = piece[e];
        }
    } else {
        using W = typename RawWord<kBytes>::type;
        union {
            W w;
            Pack<T, N> pack;
        } u;
        u.w = __ldcs(reinterpret_cast<const W *>(p));
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = Traits<T>::to_ct(u.pack.v[i]);
    }
}

template <typename T, int N>
__device__ __forceinline__ void store_vec_stream(T *__restrict__ p, const typename Traits<T>::CT (&src)[N]) {
    constexpr int kBytes = (int)sizeof(T) * N;
    if constexpr (kBytes > 16) {
        constexpr int E16 = 16 / (int)sizeof(T);
        static_assert(N % E16 == 0, "wide streaming stores must be a whole number of 16-byte pieces");
#pragma unroll
        for (int c = 0; c < N / E16; ++c) {
            typename Traits<T>::CT piece[E16];
#pragma unroll
            for (int e = 0; e < E16; ++e) piece[e] = src[c * E16 + e];
            store_vec_stream<T, E16>(p + c * E16, piece);
        }
    } else {
        using W = typename RawWord<kBytes>::type;
        union {
            W w;
            Pack<T, N> pack;
        } u;
#pragma unroll
        for (int i = 0; i < N; ++i) u.pack.v[i] = Traits<T>::from_ct(src[i]);
        __stcs(reinterpret_cast<W *>(p), u.w);
    }
}

// Round-to-nearest mul / sub that the compiler may NOT contract into an FMA: the reference writes
// `x * w - 0.5` (kernels.py:145) and its CPU-interpreted golden vectors and our oracle evaluate it as two
// rounded operations; keeping the same two roundings makes the cell index floor(x) bit-identical.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float floor_ct(float a) { return floorf(a); }
__device__ __forceinline__ double floor_ct(double a) { return floor(a); }
__device__ __forceinline__ float clamp_ct(float v, float hi) { return fminf(fmaxf(v, 0.0f), hi); }
__device__ __forceinline__ double clamp_ct(double v, double hi) { return fmin(fmax(v, 0.0), hi); }

// ---------------------------------------------------------------------------------------------------------------
// One bilinear tap (sampling point) resolved against its level.
//   row00 : pixel-row index (incl. level offset) of the (y0,x0) corner after clamping
//   pack  : bits 0..23  = (y1c - y0c) * w   (0 or w: row step to the lower corners)
//           bit  24     = (x1c - x0c)       (0 or 1: row step to the right corners)
//           bits 25..28 = validity of corners 00,01,10,11 (zeros mode; all ones in border mode)
//   dx,dy : fractional offsets x - floor(x), y - floor(y)  (NOT clipped in border mode: kernels.py:235-237)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPackDxBit = 24;
constexpr int kPackMaskShift = 25;
constexpr int kPackDyMask = (1 << 24) - 1;

template <typename CT> struct Tap {
    int row00;
    int pack;
    CT dx, dy;
};

template <typename CT>
__device__ __forceinline__ Tap<CT> locate(CT px, CT py, const Level lv, bool border, bool align) {
    const CT wm1 = (CT)(lv.w - 1), hm1 = (CT)(lv.h - 1);
    CT x, y;
    if (align) {
        x = mul_rn(px, wm1);
        y = mul_rn(py, hm1);
    } else {
        x = sub_rn(mul_rn(px, (CT)lv.w), (CT)0.5);
        y = sub_rn(mul_rn(py, (CT)lv.h), (CT)0.5);
    }
    const CT x0 = floor_ct(x), y0 = floor_ct(y);
    const CT x1 = x0 + (CT)1, y1 = y0 + (CT)1;
    unsigned mask = 0xFu;
    if (!border) {
        const bool x0m = ((CT)0 <= x0) && (x0 <= wm1);
        const bool x1m = ((CT)0 <= x1) && (x1 <= wm1);
        const bool y0m = ((CT)0 <= y0) && (y0 <= hm1);
        const bool y1m = ((CT)0 <= y1) && (y1 <= hm1);
        mask = (unsigned)(y0m && x0m) | ((unsigned)(y0m && x1m) << 1) | ((unsigned)(y1m && x0m) << 2) |
               ((unsigned)(y1m && x1m) << 3);
    }
    const int x0c = (int)clamp_ct(x0, wm1), x1c = (int)clamp_ct(x1, wm1);
    const int y0c = (int)clamp_ct(y0, hm1), y1c = (int)clamp_ct(y1, hm1);
    Tap<CT> t;
    t.row00 = lv.off + y0c * lv.w + x0c;
    t.pack = ((y1c - y0c) * lv.w) | ((x1c - x0c) << kPackDxBit) | ((int)mask << kPackMaskShift);
    t.dx = x - x0;
    t.dy = y - y0;
    return t;
}

__device__ __forceinline__ float shfl_ct(float v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
__device__ __forceinline__ double shfl_ct(double v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
__device__ __forceinline__ float shfl_xor_ct(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor_ct(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ---------------------------------------------------------------------------------------------------------------
// Relaxed, result-less vector reductions into global memory (grad_img scatter).
// nvcc 12.9 lowers these to REDG.E.ADD.F32x4 / F32x2 on sm_100a.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_v2(float *p, float a, float b) {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v1(float *p, float a) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add_v1(double *p, double a) {
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}

template <int N> __device__ __forceinline__ void red_add_vec(float *p, const float (&v)[N]) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 4) red_add_v4(p + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (N == 2) {
        red_add_v2(p, v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) red_add_v1(p + i, v[i]);
    }
}
template <int N> __device__ __forceinline__ void red_add_vec(double *p, const double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) red_add_v1(p + i, v[i]);
}

}  // namespace msda
