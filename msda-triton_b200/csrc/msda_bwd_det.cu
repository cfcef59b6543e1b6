// msda_bwd_det.cu -- deterministic grad_img: sorted-segment reduction instead of atomics (MSDA_BWD_DETERMINISTIC).
//
// The reference's backward adds into grad_img with tl.atomic_add (src/msda_triton/kernels.py:549-553), so the fp32
// summation order -- and therefore the low bits of grad_img -- changes from run to run.  This path produces
// bit-identical grad_img on every run:
//   1. keys   : every bilinear corner (unit, point, corner) emits key = destination row (b, pixel, h), value = its own
//               index, and its scalar weight (attention weight x bilinear weight) into a side array indexed by that
//               value; invalid (zeros-mode, out-of-range) corners get the sentinel key B*Npix*H (one past the last
//               row), so only ceil(log2(rows+1)) key bits need sorting.
//   2. sort   : stable LSD radix sort by key (cub::DeviceRadixSort, deterministic), so inside a segment the
//               contributions are ordered by (unit, point, corner).
//   3. starts : one pass over the sorted keys records where each destination row's segment begins.
//   4. reduce : one warp per destination row walks its segment IN THAT ORDER and accumulates
//               weight[value] * grad_out[unit] in registers; the row is written once with a plain store (no
//               zero-fill, no atomics, one rounding to the storage dtype).
// (The first version recomputed the corner weight from the sampling point in step 4 and found the segment by binary
// search: 1.05 ms for the reduce on the bench shape; with the weights precomputed and the start table it is a gather
// of grad_out rows and an FMA per contribution: 0.64 ms, the whole deterministic backward 1.84 -> 1.43 ms.)
// grad_sampling_points / grad_attention_weights never needed atomics and come from the regular backward kernel.
#include <cub/device/device_radix_sort.cuh>

#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {


template <typename T>
__global__ void __launch_bounds__(256) det_keys_kernel(const KernelArgs a, unsigned *__restrict__ keys,
                                                       unsigned *__restrict__ vals,
                                                       typename Traits<T>::CT *__restrict__ wts,
                                                       const long long n_points, const unsigned sentinel) {
    using CT = typename Traits<T>::CT;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Level *s_lv = reinterpret_cast<Level *>(s_raw);
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;
    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    const bool border = a.border != 0, align = a.align != 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += stride) {
        const long long u = i / a.LK;
        const int p = (int)(i - u * a.LK);
        const int h = (int)(u % a.H);
        const long long b = u / ((long long)a.H * a.Q);
        CT xy[2];
        load_vec<T, 2>(pts + 2 * i, xy);
        const Tap<CT> t = locate<CT>(xy[0], xy[1], s_lv[p / a.K], border, align);
        const int step_y = t.pack & kPackDyMask;
        const int step_x = (t.pack >> kPackDxBit) & 1;
        const unsigned mask = (unsigned)(t.pack >> kPackMaskShift) & 0xFu;
        const int rows[4] = {t.row00, t.row00 + step_x, t.row00 + step_y, t.row00 + step_y + step_x};
        const CT w_att = Traits<T>::to_ct(aw[i]);
        uint4 k4, v4;
        unsigned *kk = reinterpret_cast<unsigned *>(&k4), *vv = reinterpret_cast<unsigned *>(&v4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned long long key = ((unsigned long long)b * a.Npix + rows[c]) * a.H + h;
            kk[c] = ((mask >> c) & 1u) ? (unsigned)key : sentinel;
            vv[c] = (unsigned)(4 * i + c);
            const CT wx = (c & 1) ? t.dx : (CT)1 - t.dx;
            const CT wy = (c & 2) ? t.dy : (CT)1 - t.dy;
            wts[4 * i + c] = w_att * (wy * wx);
        }
        reinterpret_cast<uint4 *>(keys)[i] = k4;
        reinterpret_cast<uint4 *>(vals)[i] = v4;
    }
}

// starts[r] = index of the first sorted contribution of destination row r (rows without contributions keep kNoStart).
constexpr unsigned kNoStart = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256) det_starts_kernel(const unsigned *__restrict__ keys, unsigned *__restrict__ starts,
                                                         const long long n, const unsigned n_rows) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned k = keys[i];
        if (k < n_rows && (i == 0 || keys[i - 1] != k)) starts[k] = (unsigned)i;
    }
}

// One WARP per destination row.  The warp's 32/lanes lane groups take the row's contributions round-robin (group t
// handles positions lo+t, lo+t+T, ... of the sorted segment, two at a time for memory-level parallelism); the group
// partials are then combined by a fixed butterfly, so the summation order depends only on the sorted order.
// Rows are visited in memory order: neighbouring rows own neighbouring segments of the sorted arrays (visiting them in
// a scattered order to spread the long segments of the coarsest level over the grid was 1.8x SLOWER, and fetching the
// (value, weight) pairs of a block with one coalesced load + shuffles made no difference: 0.67 vs 0.64 ms).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) det_reduce_kernel(const KernelArgs a, const unsigned *__restrict__ keys,
                                                         const unsigned *__restrict__ vals,
                                                         const typename Traits<T>::CT *__restrict__ wts,
                                                         const unsigned *__restrict__ starts, const long long n,
                                                         const long long n_rows) {
    using CT = typename Traits<T>::CT;
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    T *__restrict__ gimg = static_cast<T *>(a.gimg);
    const int lanes = a.lanes;
    const int lane = threadIdx.x & 31;
    const int j = lane & (lanes - 1);
    const int team = lane / lanes, teams = 32 / lanes;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned points4 = 4u * (unsigned)a.LK;   // contributions per unit

    auto contribution = [&](long long i, int c0, CT (&acc)[VEC]) {
        const unsigned v = vals[i];
        const CT w = wts[v];
        const unsigned u = v / points4;               // unit (b, q, h) the contribution comes from
        CT go[VEC];
        load_vec<T, VEC>(gout + (size_t)u * a.D + c0, go);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += go[e] * w;
    };

    for (long long r = warp_global; r < n_rows; r += warps_total) {
        const unsigned first = starts[r];
        const unsigned key = (unsigned)r;
        const long long lo = first == kNoStart ? n : (long long)first;
        for (int chunk = 0; chunk < a.chunks; ++chunk) {
            const int c0 = (chunk * lanes + j) * VEC;
            const bool c_live = c0 < a.D;
            CT acc0[VEC], acc1[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc0[e] = acc1[e] = (CT)0;
            if (c_live) {
                long long i = lo + team;
                for (; i + teams < n && keys[i + teams] == key; i += 2 * teams) {   // both i and i+teams in the row
                    contribution(i, c0, acc0);
                    contribution(i + teams, c0, acc1);
                }
                if (i < n && keys[i] == key) contribution(i, c0, acc0);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                CT v = acc0[e] + acc1[e];
                for (int m = lanes; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                acc0[e] = v;
            }
            if (c_live && team == 0) store_vec<T, VEC>(gimg + (size_t)r * a.D + c0, acc0);
        }
    }
}

static int grid_for(long long work_items, int per_cta, int sm_count) {
    long long want = (work_items + per_cta - 1) / per_cta;
    const long long cap = (long long)sm_count * 32;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

size_t det_sort_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                    (const unsigned *)nullptr, (unsigned *)nullptr, n);
    return bytes;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// layout: keys_in | keys_out | vals_in | vals_out | weights (8 bytes per contribution reserved: fp64 problems) |
//         row starts | cub temp
size_t det_workspace_bytes(const KernelArgs &a) {
    const long long n = a.units * a.LK * 4;
    const long long n_rows = (long long)a.B * a.Npix * a.H;
    return 4 * align_up((size_t)n * sizeof(unsigned), 256) + align_up((size_t)n * sizeof(double), 256) +
           align_up((size_t)n_rows * sizeof(unsigned), 256) + align_up(det_sort_temp_bytes(n), 256);
}

bool det_supported(const KernelArgs &a) {
    const unsigned long long n = (unsigned long long)a.units * a.LK * 4;
    const unsigned long long rows = (unsigned long long)a.B * a.Npix * a.H;
    return n < 0xFFFFFFFFull && rows < 0xFFFFFFFFull;
}

template <typename T>
static cudaError_t launch_det_t(const KernelArgs &a, int vec, void *workspace, int sm_count, cudaStream_t st) {
    const long long n_points = a.units * a.LK, n = n_points * 4;
    const long long n_rows = (long long)a.B * a.Npix * a.H;
    const size_t arr = align_up((size_t)n * sizeof(unsigned), 256);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    unsigned *keys_in = reinterpret_cast<unsigned *>(ws), *keys_out = reinterpret_cast<unsigned *>(ws + arr);
    unsigned *vals_in = reinterpret_cast<unsigned *>(ws + 2 * arr), *vals_out = reinterpret_cast<unsigned *>(ws + 3 * arr);
    using CT = typename Traits<T>::CT;
    CT *wts = reinterpret_cast<CT *>(ws + 4 * arr);
    const size_t wts_bytes = align_up((size_t)n * sizeof(double), 256);
    unsigned *starts = reinterpret_cast<unsigned *>(ws + 4 * arr + wts_bytes);
    const size_t starts_bytes = align_up((size_t)n_rows * sizeof(unsigned), 256);
    void *temp = ws + 4 * arr + wts_bytes + starts_bytes;
    size_t temp_bytes = det_sort_temp_bytes(n);
    const size_t smem = sizeof(Level) * (size_t)a.L;

    int key_bits = 1;
    while (key_bits < 32 && (1ull << key_bits) <= (unsigned long long)n_rows) ++key_bits;   // keys in [0, n_rows]
    det_keys_kernel<T><<<grid_for(n_points, 256, sm_count), 256, smem, st>>>(a, keys_in, vals_in, wts, n_points,
                                                                           (unsigned)n_rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, key_bits, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(starts, 0xFF, (size_t)n_rows * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    det_starts_kernel<<<grid_for(n, 256, sm_count), 256, 0, st>>>(keys_out, starts, n, (unsigned)n_rows);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int grid = grid_for(n_rows, 256 / 32, sm_count);   // one warp per destination row
    switch (vec) {
        case 8:
            if constexpr (Traits<T>::kMaxVec >= 8) det_reduce_kernel<T, 8><<<grid, 256, 0, st>>>(a, keys_out, vals_out, wts, starts, n, n_rows);
            break;
        case 4:
            if constexpr (Traits<T>::kMaxVec >= 4) det_reduce_kernel<T, 4><<<grid, 256, 0, st>>>(a, keys_out, vals_out, wts, starts, n, n_rows);
            break;
        case 2: det_reduce_kernel<T, 2><<<grid, 256, 0, st>>>(a, keys_out, vals_out, wts, starts, n, n_rows); break;
        default: det_reduce_kernel<T, 1><<<grid, 256, 0, st>>>(a, keys_out, vals_out, wts, starts, n, n_rows); break;
    }
    return cudaGetLastError();
}

// a.gimg must point at the grad_img tensor in STORAGE dtype (rows are written once, no accumulation image).
cudaError_t launch_backward_det(const KernelArgs &a, int dtype, int vec, void *workspace, int sm_count, cudaStream_t st) {
    switch (dtype) {
        case 0: return launch_det_t<float>(a, vec, workspace, sm_count, st);
        case 1: return launch_det_t<__half>(a, vec, workspace, sm_count, st);
        case 2: return launch_det_t<__nv_bfloat16>(a, vec, workspace, sm_count, st);
        case 3: return launch_det_t<double>(a, vec, workspace, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace msda
