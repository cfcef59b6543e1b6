// msda_bwd.cu -- backward kernels: one pass that recomputes the forward sampling and produces
//   grad_attention_weights[p] = sum_d go[d] * sample_p[d]                                   (kernels.py:494)
//   grad_sampling_points[p]   = sum_d go[d] * w_p * scale * d(sample_p[d])/d(x|y)           (kernels.py:510-524)
//   grad_img[corner rows]    += go * w_p * bilinear_weight                                  (kernels.py:543-553)
// Replaces the reference's Triton backward (src/msda_triton/kernels.py:396-553, launched from :556-592).
//
// grad_img is scattered with relaxed, result-less vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4),
// always into an fp32 (fp64 for fp64 storage) accumulation image, so 16-bit storage does not lose the small
// contributions the reference's fp16 atomics drop.  The gradient w.r.t. a sampling point is w.r.t. the NORMALISED
// [0,1] coordinate, hence the (w-1)|w scale factor.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

constexpr int kNeedImg = 1, kNeedPts = 2, kNeedAw = 4;

template <typename T, int VEC>
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const KernelArgs a) {
    using CT = typename Traits<T>::CT;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Level *s_lv = reinterpret_cast<Level *>(s_raw);
    if (!build_level_table(s_lv, a.shapes, a.L, a.Npix)) return;

    const T *__restrict__ img = static_cast<const T *>(a.img);
    const T *__restrict__ pts = static_cast<const T *>(a.pts);
    const T *__restrict__ aw = static_cast<const T *>(a.aw);
    const T *__restrict__ gout = static_cast<const T *>(a.gout);
    CT *__restrict__ gimg = static_cast<CT *>(a.gimg);
    T *__restrict__ gpts = static_cast<T *>(a.gpts);
    T *__restrict__ gaw = static_cast<T *>(a.gaw);

    const int lanes = a.lanes;
    const int j = threadIdx.x & (lanes - 1);
    const int group = threadIdx.x / lanes;
    const int groups_per_cta = blockDim.x / lanes;
    const bool border = a.border != 0, align = a.align != 0;
    const bool need_img = (a.flags & kNeedImg) != 0, need_pts = (a.flags & kNeedPts) != 0,
               need_aw = (a.flags & kNeedAw) != 0;
    const size_t row_stride = (size_t)a.H * a.D;
    const int LK = a.LK;

    for (long long ubase = (long long)blockIdx.x * groups_per_cta; ubase < a.units;
         ubase += (long long)gridDim.x * groups_per_cta) {
        const long long u_raw = ubase + group;
        const bool live = u_raw < a.units;
        const long long u = live ? u_raw : a.units - 1;
        const int h = (int)(u % a.H);
        const long long b = u / ((long long)a.H * a.Q);
        const size_t bh_off = ((size_t)b * a.Npix * a.H + h) * a.D;
        const T *__restrict__ img_bh = img + bh_off;
        CT *__restrict__ gimg_bh = gimg + bh_off;
        const T *__restrict__ pts_u = pts + (size_t)u * LK * 2;
        const T *__restrict__ aw_u = aw + (size_t)u * LK;
        const T *__restrict__ go_u = gout + (size_t)u * a.D;
        const bool scatter = need_img && live;  // dead groups shadow the last unit and must not add twice

        // grad_out slice of chunk 0 stays in registers for the whole unit
        CT go0[VEC];
        {
            const int c0 = j * VEC;
            if (c0 < a.D) {
                load_vec<T, VEC>(go_u + c0, go0);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) go0[e] = (CT)0;
            }
        }

        for (int base = 0; base < LK; base += lanes) {
            Tap<CT> t;
            t.row00 = 0;
            t.pack = 0;
            t.dx = t.dy = (CT)0;
            CT w_att = (CT)0, sx = (CT)0, sy = (CT)0;
            const int p = base + j;
            if (p < LK) {
                const Level lv = s_lv[p / a.K];
                CT xy[2];
                load_vec<T, 2>(pts_u + 2 * p, xy);
                w_att = Traits<T>::to_ct(aw_u[p]);
                t = locate<CT>(xy[0], xy[1], lv, border, align);
                sx = align ? (CT)(lv.w - 1) : (CT)lv.w;
                sy = align ? (CT)(lv.h - 1) : (CT)lv.h;
            }
            CT my_gw = (CT)0, my_gx = (CT)0, my_gy = (CT)0;

            const int n = min(lanes, LK - base);
            for (int i = 0; i < n; ++i) {
                const int row00 = __shfl_sync(0xffffffffu, t.row00, i, lanes);
                const int pack = __shfl_sync(0xffffffffu, t.pack, i, lanes);
                const CT dx = shfl_ct(t.dx, i, lanes);
                const CT dy = shfl_ct(t.dy, i, lanes);
                const CT wa = shfl_ct(w_att, i, lanes);
                const int step_y = pack & kPackDyMask;
                const int step_x = (pack >> kPackDxBit) & 1;
                const unsigned mask = (unsigned)(pack >> kPackMaskShift) & 0xFu;
                const CT b00 = ((CT)1 - dy) * ((CT)1 - dx), b01 = ((CT)1 - dy) * dx;
                const CT b10 = dy * ((CT)1 - dx), b11 = dy * dx;
                const size_t r00 = (size_t)row00 * row_stride;
                const size_t r01 = r00 + (size_t)step_x * row_stride;
                const size_t r10 = r00 + (size_t)step_y * row_stride;
                const size_t r11 = r10 + (size_t)step_x * row_stride;

                CT d00 = (CT)0, d01 = (CT)0, d10 = (CT)0, d11 = (CT)0;  // per-corner <go, v> over my channels
                for (int chunk = 0; chunk < a.chunks; ++chunk) {
                    const int c0 = (chunk * lanes + j) * VEC;
                    if (c0 >= a.D) continue;
                    CT go[VEC];
                    if (chunk == 0) {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) go[e] = go0[e];
                    } else {
                        load_vec<T, VEC>(go_u + c0, go);
                    }
                    CT v[VEC], g[VEC];
                    if (mask & 1u) {
                        load_vec<T, VEC>(img_bh + r00 + c0, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) d00 += go[e] * v[e];
                        if (scatter) {
                            const CT s = wa * b00;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g[e] = go[e] * s;
                            red_add_vec<VEC>(gimg_bh + r00 + c0, g);
                        }
                    }
                    if (mask & 2u) {
                        load_vec<T, VEC>(img_bh + r01 + c0, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) d01 += go[e] * v[e];
                        if (scatter) {
                            const CT s = wa * b01;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g[e] = go[e] * s;
                            red_add_vec<VEC>(gimg_bh + r01 + c0, g);
                        }
                    }
                    if (mask & 4u) {
                        load_vec<T, VEC>(img_bh + r10 + c0, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) d10 += go[e] * v[e];
                        if (scatter) {
                            const CT s = wa * b10;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g[e] = go[e] * s;
                            red_add_vec<VEC>(gimg_bh + r10 + c0, g);
                        }
                    }
                    if (mask & 8u) {
                        load_vec<T, VEC>(img_bh + r11 + c0, v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) d11 += go[e] * v[e];
                        if (scatter) {
                            const CT s = wa * b11;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) g[e] = go[e] * s;
                            red_add_vec<VEC>(gimg_bh + r11 + c0, g);
                        }
                    }
                }
                // masked corners contribute value 0 (kernels.py:227-231), which is what d?? == 0 encodes
                CT s_w = b00 * d00 + b01 * d01 + b10 * d10 + b11 * d11;
                CT s_x = ((CT)1 - dy) * (d01 - d00) + dy * (d11 - d10);
                CT s_y = ((CT)1 - dx) * (d10 - d00) + dx * (d11 - d01);
                for (int m = lanes >> 1; m > 0; m >>= 1) {
                    s_w += shfl_xor_ct(s_w, m);
                    s_x += shfl_xor_ct(s_x, m);
                    s_y += shfl_xor_ct(s_y, m);
                }
                if (j == i) {
                    my_gw = s_w;
                    my_gx = s_x;
                    my_gy = s_y;
                }
            }
            if (live && p < LK) {
                if (need_aw) gaw[(size_t)u * LK + p] = Traits<T>::from_ct(my_gw);
                if (need_pts) {
                    CT g2[2] = {my_gx * (w_att * sx), my_gy * (w_att * sy)};
                    store_vec<T, 2>(gpts + ((size_t)u * LK + p) * 2, g2);
                }
            }
        }
    }
}

template <typename T, int VEC> static cudaError_t launch_generic(const KernelArgs &a, int sm_count, cudaStream_t st) {
    const int threads = 256;
    const int groups_per_cta = threads / a.lanes;
    long long want = (a.units + groups_per_cta - 1) / groups_per_cta;
    const long long cap = (long long)sm_count * 16;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    const size_t smem = sizeof(Level) * (size_t)a.L;
    msda_bwd_generic_kernel<T, VEC><<<grid, threads, smem, st>>>(a);
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

cudaError_t launch_backward_generic(const KernelArgs &a, int dtype, int vec, int sm_count, cudaStream_t st) {
    switch (dtype) {
        case 0: return dispatch_vec<float>(a, vec, sm_count, st);
        case 1: return dispatch_vec<__half>(a, vec, sm_count, st);
        case 2: return dispatch_vec<__nv_bfloat16>(a, vec, sm_count, st);
        case 3: return dispatch_vec<double>(a, vec, sm_count, st);
    }
    return cudaErrorInvalidValue;
}


// ---------------------------------------------------------------------------------------------------------------
// 16-bit storage epilogue: grad_img[T] = round(accumulation image[fp32]).
// `permuted_lanes` != 0: the image was produced by the tuned kernels, whose rows store channel 8j + 4h + e at position
// 4*lanes*h + 4j + e (see red_add_row in msda_bwd_tiled.cu); 0: natural channel order (generic kernels).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void round_grad_img_kernel(T *__restrict__ dst, const float *__restrict__ src, long long n, int D,
                                      int permuted_lanes) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long s = i;
        if (permuted_lanes) {
            const long long row = i / D;
            const int c = (int)(i - row * D);
            s = row * D + 4 * permuted_lanes * ((c >> 2) & 1) + 4 * (c >> 3) + (c & 3);
        }
        dst[i] = Traits<T>::from_ct(src[s]);
    }
}

// Natural channel order, everything 16-byte aligned: eight elements per thread and step -- two streaming 128-bit loads,
// one 128-bit store.  (The element-wise kernel above ran the B=8 x 22 223-pixel decoder pyramid, 182 MB of fp32 in,
// 91 MB out, in 117 us = 2.3 TB/s; this one is bound by HBM.)
template <typename T>
__global__ void __launch_bounds__(256) round_grad_img_vec8_kernel(T *__restrict__ dst, const float *__restrict__ src,
                                                                  long long n8) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(src) + 2 * i);
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(src) + 2 * i + 1);
        Pack<T, 8> o;
        o.v[0] = Traits<T>::from_ct(a.x);
        o.v[1] = Traits<T>::from_ct(a.y);
        o.v[2] = Traits<T>::from_ct(a.z);
        o.v[3] = Traits<T>::from_ct(a.w);
        o.v[4] = Traits<T>::from_ct(b.x);
        o.v[5] = Traits<T>::from_ct(b.y);
        o.v[6] = Traits<T>::from_ct(b.z);
        o.v[7] = Traits<T>::from_ct(b.w);
        reinterpret_cast<Pack<T, 8> *>(dst)[i] = o;
    }
}

// The same with the tuned kernels' permuted rows (D = 8 * lanes channels per head row; channel 8j + 4h + e sits at position
// 4*lanes*h + 4j + e): output channels 8j .. 8j+7 are the two 16-byte pieces at positions 4j and 4*lanes + 4j.  Optionally
// the pass also leaves the COLUMN SUMS of what it reads -- sum over (b, pixel) of grad_value[b, pixel, h, c], fp32, one per
// (h, c) -- in `colsum`: that is the bias gradient of the projection that produced `value` (frontend.py:259,
// img_input_proj), which torch otherwise computes with a reduction kernel over the rounded tensor (B=8 x 22 223 pixels x
// 256 columns: ~100 us for 91 MB; here it rides on data that is in registers anyway).
// Requires blockDim.x % (HD / 8) == 0 (so that a thread keeps its columns over the grid stride) and HD <= 2048.
template <typename T, bool COLSUM>
__global__ void __launch_bounds__(256) round_grad_img_perm_vec8_kernel(T *__restrict__ dst, const float *__restrict__ src,
                                                                       long long n8, int lanes, float *__restrict__ colsum,
                                                                       int HD) {
    __shared__ float s_sum[COLSUM ? 2048 : 1];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (COLSUM) {
        for (int c = threadIdx.x; c < HD; c += blockDim.x) s_sum[c] = 0.0f;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const long long row = i / lanes;          // head row = D channels = `lanes` chunks of 8
        const int j = (int)(i - row * lanes);
        const float *base = src + row * (8LL * lanes);
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(base + 4 * j));
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(base + 4 * lanes + 4 * j));
        Pack<T, 8> o;
        o.v[0] = Traits<T>::from_ct(a.x);
        o.v[1] = Traits<T>::from_ct(a.y);
        o.v[2] = Traits<T>::from_ct(a.z);
        o.v[3] = Traits<T>::from_ct(a.w);
        o.v[4] = Traits<T>::from_ct(b.x);
        o.v[5] = Traits<T>::from_ct(b.y);
        o.v[6] = Traits<T>::from_ct(b.z);
        o.v[7] = Traits<T>::from_ct(b.w);
        reinterpret_cast<Pack<T, 8> *>(dst)[i] = o;
        if constexpr (COLSUM) {
            acc[0] += a.x, acc[1] += a.y, acc[2] += a.z, acc[3] += a.w;
            acc[4] += b.x, acc[5] += b.y, acc[6] += b.z, acc[7] += b.w;
        }
    }
    if constexpr (COLSUM) {
        // the thread's chunk column never changes: stride is a multiple of blockDim.x, which is a multiple of HD / 8
        const int col = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % (HD / 8)) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&s_sum[col + e], acc[e]);
        __syncthreads();
        for (int c = threadIdx.x; c < HD; c += blockDim.x) red_add_v1(colsum + c, s_sum[c]);
    }
}

bool round_colsum_supported(int dtype, int D, int HD, long long n) {
    return (dtype == 1 || dtype == 2) && D % 8 == 0 && HD % 8 == 0 && HD <= 2048 && 256 % (HD / 8) == 0 && n % 8 == 0;
}

cudaError_t launch_round_grad_img(void *dst, const float *src, long long n, int dtype, int D, int permuted_lanes,
                                  cudaStream_t st, float *colsum, int HD) {
    if (n <= 0) return cudaSuccess;
    const bool vec_ok = n % 8 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(src) % 16 == 0 && (dtype == 1 || dtype == 2);
    if (colsum && !(vec_ok && permuted_lanes > 0 && permuted_lanes * 8 == D && round_colsum_supported(dtype, D, HD, n)))
        return cudaErrorNotSupported;
    if (vec_ok && permuted_lanes > 0 && permuted_lanes * 8 == D) {
        const long long n8 = n / 8;
        long long want8 = (n8 + 255) / 256;
        const int grid8 = (int)(want8 > 148 * 16 ? 148 * 16 : want8);
        if (colsum) {
            const cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)HD, st);
            if (e != cudaSuccess) return e;
            if (dtype == 1)
                round_grad_img_perm_vec8_kernel<__half, true>
                    <<<grid8, 256, 0, st>>>(static_cast<__half *>(dst), src, n8, permuted_lanes, colsum, HD);
            else
                round_grad_img_perm_vec8_kernel<__nv_bfloat16, true>
                    <<<grid8, 256, 0, st>>>(static_cast<__nv_bfloat16 *>(dst), src, n8, permuted_lanes, colsum, HD);
        } else if (dtype == 1) {
            round_grad_img_perm_vec8_kernel<__half, false>
                <<<grid8, 256, 0, st>>>(static_cast<__half *>(dst), src, n8, permuted_lanes, nullptr, 0);
        } else {
            round_grad_img_perm_vec8_kernel<__nv_bfloat16, false>
                <<<grid8, 256, 0, st>>>(static_cast<__nv_bfloat16 *>(dst), src, n8, permuted_lanes, nullptr, 0);
        }
        return cudaGetLastError();
    }
    if (permuted_lanes == 0 && vec_ok) {
        const long long n8 = n / 8;
        long long want8 = (n8 + 255) / 256;
        const int grid8 = (int)(want8 > 148 * 16 ? 148 * 16 : want8);
        if (dtype == 1)
            round_grad_img_vec8_kernel<__half><<<grid8, 256, 0, st>>>(static_cast<__half *>(dst), src, n8);
        else
            round_grad_img_vec8_kernel<__nv_bfloat16><<<grid8, 256, 0, st>>>(static_cast<__nv_bfloat16 *>(dst), src, n8);
        return cudaGetLastError();
    }
    const int threads = 256;
    long long want = (n + threads - 1) / threads;
    const int grid = (int)(want > 148 * 32 ? 148 * 32 : want);
    if (dtype == 1)
        round_grad_img_kernel<__half><<<grid, threads, 0, st>>>(static_cast<__half *>(dst), src, n, D, permuted_lanes);
    else if (dtype == 2)
        round_grad_img_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(static_cast<__nv_bfloat16 *>(dst), src, n, D,
                                                                      permuted_lanes);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace msda
