// msda_bwd_detq.cu -- deterministic grad_img WITHOUT a sort: exact (quantised) fp32 row adds.
//
// The reference adds into grad_img with tl.atomic_add (src/msda_triton/kernels.py:549-553): the fp32 summation order, and
// with it the low bits of grad_img, changes from run to run.  The sorted-segment path (msda_bwd_det.cu) fixes the order
// at the price of a 20 M-pair radix sort and a gather pass (1.43 ms on the bench shape, 3x the atomic mode).  This
// path keeps the atomics and removes the ROUNDING instead: floating-point addition is associative as long as no add
// rounds, and no add rounds when every addend is an integer multiple of a quantum q and every partial sum stays below
// 2^24 q.  So
//   1. amax    : max |grad_out| per unit (kept) and overall, max |attention weight| overall   -- max is order-independent
//   2. bounds  : every valid bilinear corner adds  ceil(w * g_unit / amax * 4096) + 1  (an INTEGER) to a 64-bit counter of
//                its destination row, w = |attention weight| x bilinear weight: an upper bound, in units of amax/4096,
//                of the sum of |values| the row will receive; integer atomics are exact in any order.  The same
//                kernel keeps, per (b, h, level), the largest row counter.
//   3. backward: the regular tuned kernel (msda_bwd_tiled.cu, QUANT), which rounds each value to a multiple of
//                q(b,h,level) = 2^(ilogb(bound) + 2 - 24) before the usual red.global.add.v4.f32.
// Result: bit-identical grad_img on every run, one pass over the pyramid, grad_points / grad_weights from the same
// kernel.  Cost of the rounding: each added value moves by at most q/2 <= 2^-22 x (the largest row bound of its
// slice-level), i.e. about the error of ONE fp32 rounding at the magnitude of the largest row sum.
//
// Scope: fp32, D == 32, L*K == 16 (the tuned kernel's shapes, all BASELINE configs); everything else keeps
// msda_bwd_det.cu.  MSDA_B200_DET_VARIANT=0 forces the sorted-segment path.
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tuning.h"

namespace msda {

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace layout: amax[2] (+ padding to 256 B) | slmax[B*H*L] u64 | rowsum[B*Npix*H] u64 | gmax[units] float
struct DetqLayout {
    size_t off_amax, off_slmax, off_rowsum, off_gmax, zero_bytes, total;
};
DetqLayout layout(const KernelArgs &a) {
    DetqLayout l;
    l.off_amax = 0;
    l.off_slmax = 256;
    l.off_rowsum = l.off_slmax + align_up((size_t)a.B * a.H * a.L * 8, 256);
    l.off_gmax = l.off_rowsum + align_up((size_t)a.B * a.Npix * a.H * 8, 256);
    l.zero_bytes = l.off_gmax;                                  // amax, slmax and rowsum start from zero
    l.total = l.off_gmax + align_up((size_t)a.units * 4, 256);
    return l;
}

int grid_for(long long work_items, int per_cta, int sm_count) {
    long long want = (work_items + per_cta - 1) / per_cta;
    const long long cap = (long long)sm_count * 16;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// 1. gmax[u] = max_c |grad_out[u, c]|, amax[0] = max over everything, amax[1] = max |attention weight|.
//    D == 32: 8 lanes x 4 channels per unit.  Non-negative floats order like their bit patterns, so atomicMax on the
//    bits is an exact, order-independent maximum (NaN bits compare above +inf: a NaN poisons amax, see row_quantum()).
__global__ void __launch_bounds__(256) detq_amax_kernel(const KernelArgs a, float *__restrict__ gmax,
                                                        unsigned *__restrict__ amax) {
    const float *__restrict__ gout = static_cast<const float *>(a.gout);
    const float *__restrict__ aw = static_cast<const float *>(a.aw);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    unsigned best_g = 0u, best_w = 0u;
    const long long n4 = a.units * 8;   // float4 pieces of grad_out
    for (long long i = tid; i < ((n4 + 31) / 32) * 32; i += stride) {
        float m = 0.0f;
        if (i < n4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(gout) + i);
            m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
            if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) m = __uint_as_float(0x7fc00000u);
        }
        unsigned b = __float_as_uint(m);
        b = max(b, __shfl_xor_sync(0xffffffffu, b, 1));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, 2));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, 4));
        if (i < n4 && (threadIdx.x & 7) == 0) gmax[i >> 3] = __uint_as_float(b);
        best_g = max(best_g, b);
    }
    const long long nw = a.units * a.LK;
    for (long long i = tid; i < nw; i += stride) {
        const float v = __ldg(aw + i);
        best_w = max(best_w, v != v ? 0x7fc00000u : __float_as_uint(fabsf(v)));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        best_g = max(best_g, __shfl_xor_sync(0xffffffffu, best_g, s));
        best_w = max(best_w, __shfl_xor_sync(0xffffffffu, best_w, s));
    }
    if ((threadIdx.x & 31) == 0) {
        if (best_g) atomicMax(amax + 0, best_g);
        if (best_w) atomicMax(amax + 1, best_w);
    }
}

// 2 + 3. Row bounds and their per-(b, h, level) maxima.  One CTA = one (b,h) slice x one chunk of queries, one thread
//    per sampling point.  The last `cap` rows of the pyramid (the coarse levels, where thousands of corners hit the same
//    row; the whole pyramid when it fits) are counted in shared memory first and reach the global 64-bit counters as
//    one add per row and CTA; rows in front of them go to the global counters directly.  The add RETURNS the previous
//    value: row counters only grow, so the largest "previous + mine" any CTA sees IS the row's final value, and the
//    per-level maxima need no second pass over the rows.
constexpr int kBoundsThreads = 512;

__global__ void __launch_bounds__(kBoundsThreads) detq_bounds_kernel(const KernelArgs a, const float *__restrict__ gmax,
                                                                    const unsigned *__restrict__ amax,
                                                                    unsigned long long *__restrict__ rowsum,
                                                                    unsigned long long *__restrict__ slmax,
                                                                    const int chunks, const int cap) {
    extern __shared__ unsigned s_cnt[];          // [cap] row bounds of this CTA, rows Npix - cap .. Npix - 1
    __shared__ Level s_lv[8];
    __shared__ unsigned long long s_max[8];
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;
    const float m = __uint_as_float(amax[0]) * __uint_as_float(amax[1]);
    if (!(m > 0.0f) || !(m < 3.0e38f)) return;    // nothing to add, or non-finite inputs: the backward runs unquantised
    const float scale = 4096.0f / m;
    for (int i = threadIdx.x; i < cap; i += kBoundsThreads) s_cnt[i] = 0u;
    if (threadIdx.x < 8) s_max[threadIdx.x] = 0ull;
    __syncthreads();
    const float *__restrict__ pts = static_cast<const float *>(a.pts);
    const float *__restrict__ aw = static_cast<const float *>(a.aw);
    const bool border = a.border != 0, align = a.align != 0;
    const int bh = blockIdx.x / chunks, chunk = blockIdx.x - bh * chunks;
    const int b = bh / a.H, h = bh - b * a.H;
    const int q_lo = (int)((long long)a.Q * chunk / chunks), q_hi = (int)((long long)a.Q * (chunk + 1) / chunks);
    const int first_smem_row = a.Npix - cap;
    unsigned long long *__restrict__ rows_bh = rowsum + ((unsigned long long)b * a.Npix * a.H + h);   // + row * H
    unsigned long long best[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) best[l] = 0ull;
    auto note = [&](int row, unsigned long long total) {
        int lvl = 0;
        while (lvl + 1 < a.L && row >= s_lv[lvl + 1].off) ++lvl;
#pragma unroll
        for (int l = 0; l < 8; ++l)
            if (l == lvl && total > best[l]) best[l] = total;
    };
    const int n_local = (q_hi - q_lo) * a.LK;
    for (int i = threadIdx.x; i < n_local; i += kBoundsThreads) {
        const int q = q_lo + i / a.LK, p = i % a.LK;
        const long long u = ((long long)b * a.Q + q) * a.H + h;
        const long long gi = u * a.LK + p;
        const float2 xy = __ldg(reinterpret_cast<const float2 *>(pts) + gi);
        const Tap<float> t = locate<float>(xy.x, xy.y, s_lv[p / a.K], border, align);
        const int step_y = t.pack & kPackDyMask, step_x = (t.pack >> kPackDxBit) & 1;
        const unsigned mask = (unsigned)(t.pack >> kPackMaskShift) & 0xFu;
        const int rows[4] = {t.row00, t.row00 + step_x, t.row00 + step_y, t.row00 + step_y + step_x};
        const float wg = fabsf(__ldg(aw + gi)) * gmax[u] * scale;
        // same bilinear weights as the backward kernel up to an ulp; the "+ 1" below covers that and more
        const float bw1 = (1.0f - t.dy) * t.dx, bw3 = t.dy * t.dx;
        const float bw[4] = {(1.0f - t.dy) - bw1, bw1, t.dy - bw3, bw3};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (!((mask >> c) & 1u)) continue;
            const unsigned units = (unsigned)ceilf(fminf(fabsf(bw[c]) * wg, 8192.0f)) + 1u;   // <= 8193
            if (rows[c] >= first_smem_row) {
                atomicAdd(s_cnt + (rows[c] - first_smem_row), units);
            } else {
                const unsigned long long old = atomicAdd(rows_bh + (size_t)rows[c] * a.H, (unsigned long long)units);
                note(rows[c], old + units);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cap; i += kBoundsThreads) {
        const unsigned v = s_cnt[i];
        if (v == 0u) continue;
        const int row = first_smem_row + i;
        const unsigned long long old = atomicAdd(rows_bh + (size_t)row * a.H, (unsigned long long)v);
        note(row, old + v);
    }
#pragma unroll
    for (int l = 0; l < 8; ++l)
        if (l < a.L && best[l]) atomicMax(s_max + l, best[l]);
    __syncthreads();
    if (threadIdx.x < a.L && s_max[threadIdx.x]) atomicMax(slmax + (size_t)bh * a.L + threadIdx.x, s_max[threadIdx.x]);
}

}  // namespace

size_t detq_workspace_bytes(const KernelArgs &a) { return layout(a).total; }

cudaError_t launch_backward_detq(const KernelArgs &a, int dtype, void *workspace, int sm_count, cudaStream_t st) {
    if (!quant_backward_supported(a, dtype) || !(a.flags & 1)) return cudaErrorNotSupported;
    if (tuning().det_variant == 0) return cudaErrorNotSupported;
    const DetqLayout l = layout(a);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    unsigned *amax = reinterpret_cast<unsigned *>(ws + l.off_amax);
    unsigned long long *slmax = reinterpret_cast<unsigned long long *>(ws + l.off_slmax);
    unsigned long long *rowsum = reinterpret_cast<unsigned long long *>(ws + l.off_rowsum);
    float *gmax = reinterpret_cast<float *>(ws + l.off_gmax);
    cudaError_t e = cudaMemsetAsync(ws, 0, l.zero_bytes, st);
    if (e != cudaSuccess) return e;
    detq_amax_kernel<<<grid_for(a.units * 8, 256 * 4, sm_count), 256, 0, st>>>(a, gmax, amax);
    // shared-memory counters: as much of the pyramid's tail as fits; a CTA's chunk stays below 2^32 / (64 * 8193) queries
    constexpr int kMaxCap = 48 * 1024;    // rows, 4 bytes each
    const int cap = a.Npix < kMaxCap ? a.Npix : kMaxCap;
    int chunks = (a.Q + 4095) / 4096;
    const long long slices = (long long)a.B * a.H;
    while (slices * chunks < 2LL * sm_count && (a.Q + chunks - 1) / chunks > 256) chunks *= 2;   // fill the machine
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(detq_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxCap * 4);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    detq_bounds_kernel<<<(unsigned)(slices * chunks), kBoundsThreads, (size_t)cap * 4, st>>>(a, gmax, amax, rowsum, slmax,
                                                                                           chunks, cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    KernelArgs k = a;
    k.q_slmax = slmax;
    k.q_amax = amax;
    return launch_backward_tiled_quant(k, dtype, sm_count, st);
}

}  // namespace msda
