// msda_fwd.cu -- forward kernels: out[b,q,h,:] = sum_{l,k} w * bilinear(img_l[b,:,h,:], p).
//
// Replaces the reference's Triton forward (src/msda_triton/kernels.py:267-348, launched from :351-379).
//
// Work decomposition: one "unit" = one output row (b,q,h).  `lanes` lanes of a warp cooperate on a
// unit; each lane owns VEC consecutive channels and gathers them with ONE vector load per bilinear corner
// (128-bit when D*sizeof(T) allows).  The coordinate math of the L*K sampling points is NOT replicated across the
// lanes: lane j resolves points j, j+lanes, ... and the results are exchanged with warp shuffles.
// Accumulation is in registers in the compute type (fp32, or fp64 for fp64 storage); there is no shared-memory
// reduction and no tensor-core use -- the op is a gather (about 0.6 flop per byte).
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

// ---------------------------------------------------------------------------------------------------------------
// Generic kernel: any D (multiple of VEC), any L, K, both padding modes / align settings at run time.
// Units are taken in natural (b,q,h) order with a grid-stride loop.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const KernelArgs a) {
    using CT = typename Traits<T>::CT;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Level *s_lv = reinterpret_cast<Level *>(s_raw);
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    T *__restrict__ out = static_cast<T *>(a.out);

    const int lanes = a.lanes;
    const int j = threadIdx.x & (lanes - 1);               // lane within the unit's group
    const int group = threadIdx.x / lanes;                 // group within the CTA
    const int groups_per_cta = blockDim.x / lanes;
    const bool border = a.border != 0, align = a.align != 0;
    const size_t row_stride = (size_t)a.H * a.D;           // elements between consecutive pixel rows
    const int LK = a.LK;

    for (long long ubase = (long long)blockIdx.x * groups_per_cta; ubase < a.units;
         ubase += (long long)gridDim.x * groups_per_cta) {
        const long long u_raw = ubase + group;
        const bool live = u_raw < a.units;
        const long long u = live ? u_raw : a.units - 1;    // dead groups shadow the last unit (keeps shuffles full-warp)
        const int h = (int)(u % a.H);
        const long long b = u / ((long long)a.H * a.Q);
        const T *__restrict__ img_bh = img + ((size_t)b * a.Npix * a.H + h) * a.D;
        const T *__restrict__ pts_u = pts + (size_t)u * LK * 2;
        const T *__restrict__ aw_u = aw + (size_t)u * LK;

        for (int chunk = 0; chunk < a.chunks; ++chunk) {
            const int c0 = (chunk * lanes + j) * VEC;
            const bool c_live = c0 < a.D;
            CT acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = (CT)0;

            for (int base = 0; base < LK; base += lanes) {
                // --- this lane resolves point (base + j) ---
                Tap<CT> t;
                t.row00 = 0;
                t.pack = 0;
                t.dx = t.dy = (CT)0;
                CT w_att = (CT)0;
                const int p = base + j;
                if (p < LK) {
                    const Level lv = s_lv[p / a.K];
                    CT xy[2];
                    load_vec<T, 2>(pts_u + 2 * p, xy);
                    w_att = Traits<T>::to_ct(aw_u[p]);
                    t = locate<CT>(xy[0], xy[1], lv, border, align);
                }
                // --- all lanes of the group consume the points one by one ---
                const int n = min(lanes, LK - base);
                for (int i = 0; i < n; ++i) {
                    const int row00 = __shfl_sync(0xffffffffu, t.row00, i, lanes);
                    const int pack = __shfl_sync(0xffffffffu, t.pack, i, lanes);
                    const CT dx = shfl_ct(t.dx, i, lanes);
                    const CT dy = shfl_ct(t.dy, i, lanes);
                    const CT wa = shfl_ct(w_att, i, lanes);
                    if (!c_live) continue;
                    const int step_y = pack & kPackDyMask;
                    const int step_x = (pack >> kPackDxBit) & 1;
                    const unsigned mask = (unsigned)(pack >> kPackMaskShift) & 0xFu;
                    const CT w00 = wa * (((CT)1 - dy) * ((CT)1 - dx));
                    const CT w01 = wa * (((CT)1 - dy) * dx);
                    const CT w10 = wa * (dy * ((CT)1 - dx));
                    const CT w11 = wa * (dy * dx);
                    const T *__restrict__ p00 = img_bh + (size_t)row00 * row_stride + c0;
                    CT v[VEC];
                    if (mask & 1u) {
                        load_vec<T, VEC>(p00, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] += w00 * v[e];
                    }
                    if (mask & 2u) {
                        load_vec<T, VEC>(p00 + (size_t)step_x * row_stride, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] += w01 * v[e];
                    }
                    if (mask & 4u) {
                        load_vec<T, VEC>(p00 + (size_t)step_y * row_stride, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] += w10 * v[e];
                    }
                    if (mask & 8u) {
                        load_vec<T, VEC>(p00 + (size_t)(step_y + step_x) * row_stride, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] += w11 * v[e];
                    }
                }
            }
            if (live && c_live) store_vec<T, VEC>(out + (size_t)u * a.D + c0, acc);
        }
    }
}

template <typename T, int VEC> static cudaError_t launch_generic(const KernelArgs &a, int sm_count, cudaStream_t st) {
    const int threads = 256;
    const int groups_per_cta = threads / a.lanes;
    long long want = (a.units + groups_per_cta - 1) / groups_per_cta;
    const long long cap = (long long)sm_count * 16;  // grid-stride beyond a few waves
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    const size_t smem = sizeof(Level) * (size_t)a.L;
    msda_fwd_generic_kernel<T, VEC><<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename T> static cudaError_t dispatch_vec(const KernelArgs &a, int vec, int sm_count, cudaStream_t st) {
    switch (vec) {
        case 8:
            if constexpr (Traits<T>::kMaxVec >= 8) return launch_generic<T, 8>(a, sm_count, st);
            break;
        case 4:
            if constexpr (Traits<T>::kMaxVec >= 4) return launch_generic<T, 4>(a, sm_count, st);
            break;
        case 2:
            return launch_generic<T, 2>(a, sm_count, st);
        case 1:
            return launch_generic<T, 1>(a, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_forward_generic(const KernelArgs &a, int dtype, int vec, int sm_count, cudaStream_t st) {
    switch (dtype) {
        case 0: return dispatch_vec<float>(a, vec, sm_count, st);
        case 1: return dispatch_vec<__half>(a, vec, sm_count, st);
        case 2: return dispatch_vec<__nv_bfloat16>(a, vec, sm_count, st);
        case 3: return dispatch_vec<double>(a, vec, sm_count, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace msda
