// msda_api.cu -- the extern "C" surface of libmsda_b200.so (declared in include/msda_b200.h).
//
// Validation, kernel selection and launch only: no allocation, no synchronisation, no default-stream use.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/msda_b200.h"
#include "msda_common.cuh"
#include "msda_launch.h"
#include "msda_tiled.cuh"
#include "msda_tuning.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

// Immutable per-device facts, looked up once.
struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
};
DeviceInfo g_dev[64];
std::once_flag g_dev_once[64];

int device_info(DeviceInfo *out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) return fail(MSDA_ERR_UNSUPPORTED_DEVICE, "device ordinal %d out of range", dev);
    std::call_once(g_dev_once[dev], [dev]() {
        cudaDeviceGetAttribute(&g_dev[dev].sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_dev[dev].cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    });
    *out = g_dev[dev];
    if (out->cc_major != 10)
        return fail(MSDA_ERR_UNSUPPORTED_DEVICE,
                    "libmsda_b200 is built for sm_100a only; current device has compute capability major %d",
                    out->cc_major);
    return MSDA_OK;
}

size_t dtype_size(int dtype) {
    switch (dtype) {
        case MSDA_DTYPE_F32: return 4;
        case MSDA_DTYPE_F16: return 2;
        case MSDA_DTYPE_BF16: return 2;
        case MSDA_DTYPE_F64: return 8;
    }
    return 0;
}

int validate(const msda_problem *p) {
    if (!p) return fail(MSDA_ERR_NULL_POINTER, "msda_problem is NULL");
    if (dtype_size(p->dtype) == 0) return fail(MSDA_ERR_BAD_DTYPE, "unknown dtype code %d", p->dtype);
    if (p->padding_mode != MSDA_PAD_ZEROS && p->padding_mode != MSDA_PAD_BORDER)
        return fail(MSDA_ERR_BAD_MODE, "padding_mode must be MSDA_PAD_ZEROS or MSDA_PAD_BORDER, got %d", p->padding_mode);
    if (p->align_corners != 0 && p->align_corners != 1)
        return fail(MSDA_ERR_BAD_MODE, "align_corners must be 0 or 1, got %d", p->align_corners);
    if (p->B < 0 || p->Q < 0 || p->H <= 0 || p->D <= 0 || p->L <= 0 || p->K <= 0 || p->Npix <= 0)
        return fail(MSDA_ERR_BAD_SHAPE, "invalid sizes B=%lld Npix=%lld H=%lld D=%lld Q=%lld L=%lld K=%lld",
                    (long long)p->B, (long long)p->Npix, (long long)p->H, (long long)p->D, (long long)p->Q,
                    (long long)p->L, (long long)p->K);
    const long long kInt = 0x7fffffffLL;
    if (p->Npix > kInt || p->Q > kInt || p->H > kInt || p->D > kInt || p->B > kInt || p->L * p->K > (1 << 20) ||
        p->L > 2048)
        return fail(MSDA_ERR_BAD_SHAPE, "a dimension exceeds the supported range (each < 2^31, L <= 2048, L*K <= 2^20)");
    if (p->B * p->H > kInt) return fail(MSDA_ERR_BAD_SHAPE, "B*H must be below 2^31");
    if (p->Npix * p->H * p->D > (1LL << 46))
        return fail(MSDA_ERR_BAD_SHAPE, "one image of the pyramid is too large");
    return MSDA_OK;
}

bool aligned(const void *p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Largest power-of-two vector width (elements) such that D % vec == 0, vec*sizeof(T) <= 16 and every row pointer
// is aligned to it.
int pick_vec(const msda_problem *p, std::initializer_list<const void *> row_ptrs) {
    const size_t es = dtype_size(p->dtype);
    int vec = (int)(16 / es);
    while (vec > 1) {
        bool ok = (p->D % vec) == 0;
        for (const void *q : row_ptrs) ok = ok && aligned(q, vec * es);
        if (ok) break;
        vec >>= 1;
    }
    return vec;
}

void fill_args(msda::KernelArgs &a, const msda_problem *p, int vec) {
    std::memset(&a, 0, sizeof(a));
    a.B = (int)p->B;
    a.Q = (int)p->Q;
    a.H = (int)p->H;
    a.D = (int)p->D;
    a.L = (int)p->L;
    a.K = (int)p->K;
    a.Npix = (int)p->Npix;
    a.LK = (int)(p->L * p->K);
    a.units = p->B * p->Q * p->H;
    a.border = p->padding_mode == MSDA_PAD_BORDER;
    a.align = p->align_corners;
    const int need = (int)((p->D + vec - 1) / vec);  // lanes needed to cover D once
    int lanes = 1;
    while (lanes < need && lanes < 32) lanes <<= 1;
    a.lanes = lanes;
    a.chunks = (need + lanes - 1) / lanes;
}

// Debug / test switch: MSDA_B200_FORCE_GENERIC=1 routes every problem to the generic kernels.
bool force_generic() { return msda::tuning().force_generic != 0; }

}  // namespace

extern "C" {

int msda_abi_version(void) { return MSDA_B200_ABI_VERSION; }

const char *msda_last_error(void) { return g_err; }

void msda_reload_tuning(void) { msda::reload_tuning(); }

int msda_forward(void *out, const void *img, const int64_t *img_shapes, const void *sampling_points,
                 const void *attention_weights, const msda_problem *prob, void *stream) {
    int rc = validate(prob);
    if (rc != MSDA_OK) return rc;
    if (prob->B == 0 || prob->Q == 0) return MSDA_OK;
    if (!out || !img || !img_shapes || !sampling_points || !attention_weights)
        return fail(MSDA_ERR_NULL_POINTER, "msda_forward: NULL device pointer");
    const size_t es = dtype_size(prob->dtype);
    if (!aligned(sampling_points, 2 * es) || !aligned(attention_weights, es) || !aligned(img_shapes, 8))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_forward: sampling_points must be aligned to one (x,y) pair");
    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const int vec = pick_vec(prob, {img, out});
    msda::KernelArgs a;
    fill_args(a, prob, vec);
    a.img = img;
    a.shapes = reinterpret_cast<const long long *>(img_shapes);
    a.pts = sampling_points;
    a.aw = attention_weights;
    a.out = out;

    cudaError_t e = cudaErrorNotSupported;
    // the tuned forward reads up to 4 attention weights (fp32: 16 bytes) and 8 coordinates per lane access
    const bool tiled_ok = !force_generic() && vec * es == 16 && aligned(sampling_points, 16) &&
                          aligned(attention_weights, 16);
    if (tiled_ok) e = msda::launch_forward_tiled(a, prob->dtype, dev.sm_count, st);
    if (e == cudaErrorNotSupported) e = msda::launch_forward_generic(a, prob->dtype, vec, dev.sm_count, st);
    if (e != cudaSuccess) return fail_cuda(e, "msda_forward launch");
    return MSDA_OK;
}

size_t msda_backward_workspace_bytes(const msda_problem *prob, int flags) {
    if (validate(prob) != MSDA_OK) return 0;
    if (!(flags & MSDA_BWD_NEED_IMG) || prob->B == 0) return 0;
    if (flags & MSDA_BWD_DETERMINISTIC) {
        if (prob->Q == 0) return 0;
        msda::KernelArgs a;
        fill_args(a, prob, 1);
        // fp32 tuned shapes: exact (quantised) row adds need a few MB; everything else sorts (24 B per corner)
        if (!force_generic() && msda::tuning().det_variant != 0 && msda::quant_backward_supported(a, prob->dtype))
            return msda::detq_workspace_bytes(a);
        return msda::det_supported(a) ? msda::det_workspace_bytes(a) : 0;
    }
    if (prob->dtype == MSDA_DTYPE_F16 || prob->dtype == MSDA_DTYPE_BF16) {
        if ((flags & MSDA_BWD_VALUE_COLSUM) && msda_module_colsum_supported(prob))
            return msda_module_colsum_offset(prob) + sizeof(float) * (size_t)prob->H * prob->D;
        return sizeof(float) * (size_t)prob->B * prob->Npix * prob->H * prob->D;
    }
    return 0;
}

int msda_module_colsum_supported(const msda_problem *prob) {
    if (validate(prob) != MSDA_OK) return 0;
    const long long n = (long long)prob->B * prob->Npix * prob->H * prob->D;
    return msda::round_colsum_supported(prob->dtype, (int)prob->D, (int)(prob->H * prob->D), n) ? 1 : 0;
}

size_t msda_module_colsum_offset(const msda_problem *prob) {
    if (validate(prob) != MSDA_OK) return 0;
    const size_t accum = sizeof(float) * (size_t)prob->B * prob->Npix * prob->H * prob->D;
    return (accum + 255) / 256 * 256;
}

int msda_backward(void *grad_img, void *grad_points, void *grad_weights, const void *grad_out, const void *img,
                  const int64_t *img_shapes, const void *sampling_points, const void *attention_weights,
                  const msda_problem *prob, int flags, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = validate(prob);
    if (rc != MSDA_OK) return rc;
    if ((flags & MSDA_BWD_NEED_ALL) == 0) return MSDA_OK;
    const bool need_img = flags & MSDA_BWD_NEED_IMG, need_pts = flags & MSDA_BWD_NEED_POINTS,
               need_aw = flags & MSDA_BWD_NEED_WEIGHTS;
    const size_t es = dtype_size(prob->dtype);
    const size_t img_elems = (size_t)prob->B * prob->Npix * prob->H * prob->D;
    const bool no_units = prob->B == 0 || prob->Q == 0;  // empty tensors legitimately have NULL data pointers
    if ((need_img && !grad_img && img_elems > 0) || (!no_units && ((need_pts && !grad_points) || (need_aw && !grad_weights))))
        return fail(MSDA_ERR_NULL_POINTER, "msda_backward: a requested gradient buffer is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;

    if ((flags & MSDA_BWD_DETERMINISTIC) && need_img && img_elems > 0 && !no_units && !force_generic() &&
        msda::tuning().det_variant != 0) {
        // fp32 tuned shapes: ONE pass of the tuned backward with exact (quantised) row adds, msda_bwd_detq.cu
        msda::KernelArgs q;
        fill_args(q, prob, 4);
        if (msda::quant_backward_supported(q, prob->dtype) && grad_out && img && img_shapes && sampling_points &&
            attention_weights && aligned(img, 16) && aligned(grad_out, 16) && aligned(grad_img, 16) &&
            aligned(sampling_points, 16) && aligned(attention_weights, 8) && aligned(img_shapes, 8) &&
            (!need_pts || aligned(grad_points, 16)) && (!need_aw || aligned(grad_weights, 8))) {
            const size_t want = msda::detq_workspace_bytes(q);
            if (!workspace || workspace_bytes < want || !aligned(workspace, 256))
                return fail(MSDA_ERR_WORKSPACE, "msda_backward: deterministic mode needs a 256-byte aligned workspace of "
                                                "%zu bytes, got %zu", want, workspace_bytes);
            cudaError_t e0 = cudaMemsetAsync(grad_img, 0, img_elems * es, st);
            if (e0 != cudaSuccess) return fail_cuda(e0, "msda_backward zero-fill");
            q.img = img;
            q.shapes = reinterpret_cast<const long long *>(img_shapes);
            q.pts = sampling_points;
            q.aw = attention_weights;
            q.gout = grad_out;
            q.gimg = grad_img;
            q.gpts = grad_points;
            q.gaw = grad_weights;
            q.flags = flags & MSDA_BWD_NEED_ALL;
            e0 = msda::launch_backward_detq(q, prob->dtype, workspace, dev.sm_count, st);
            if (e0 == cudaSuccess) return MSDA_OK;
            if (e0 != cudaErrorNotSupported) return fail_cuda(e0, "msda_backward deterministic (exact row adds) path");
        }
    }
    if ((flags & MSDA_BWD_DETERMINISTIC) && need_img) {
        // grad_points / grad_weights from the regular kernel (no atomics there), grad_img by sorted segments
        if (need_pts || need_aw) {
            rc = msda_backward(nullptr, grad_points, grad_weights, grad_out, img, img_shapes, sampling_points,
                               attention_weights, prob, flags & (MSDA_BWD_NEED_POINTS | MSDA_BWD_NEED_WEIGHTS), nullptr,
                               0, stream);
            if (rc != MSDA_OK) return rc;
        }
        if (img_elems == 0) return MSDA_OK;
        if (no_units) {
            cudaError_t e0 = cudaMemsetAsync(grad_img, 0, img_elems * es, st);
            return e0 == cudaSuccess ? MSDA_OK : fail_cuda(e0, "msda_backward zero-fill");
        }
        if (!grad_out || !img_shapes || !sampling_points || !attention_weights)
            return fail(MSDA_ERR_NULL_POINTER, "msda_backward: NULL device pointer");
        if (!aligned(sampling_points, 2 * es) || !aligned(img_shapes, 8))
            return fail(MSDA_ERR_BAD_SHAPE, "msda_backward: sampling_points must be aligned to one (x,y) pair");
        const int dvec = pick_vec(prob, {grad_out, grad_img});
        msda::KernelArgs d;
        fill_args(d, prob, dvec);
        if (!msda::det_supported(d))
            return fail(MSDA_ERR_BAD_SHAPE, "msda_backward: problem too large for the deterministic mode (needs "
                                            "B*Q*H*L*K*4 < 2^32 and B*Npix*H < 2^32)");
        const size_t want = msda::det_workspace_bytes(d);
        if (!workspace || workspace_bytes < want || !aligned(workspace, 256))
            return fail(MSDA_ERR_WORKSPACE, "msda_backward: deterministic mode needs a 256-byte aligned workspace of "
                                            "%zu bytes, got %zu", want, workspace_bytes);
        d.shapes = reinterpret_cast<const long long *>(img_shapes);
        d.pts = sampling_points;
        d.aw = attention_weights;
        d.gout = grad_out;
        d.gimg = grad_img;
        cudaError_t e1 = msda::launch_backward_det(d, prob->dtype, dvec, workspace, dev.sm_count, st);
        if (e1 != cudaSuccess) return fail_cuda(e1, "msda_backward deterministic path");
        return MSDA_OK;
    }

    // grad_img accumulates in fp32 (fp64 for f64): directly in grad_img for f32/f64, in the workspace for 16-bit.
    const bool staged = need_img && (prob->dtype == MSDA_DTYPE_F16 || prob->dtype == MSDA_DTYPE_BF16);
    void *accum = grad_img;
    size_t accum_bytes = img_elems * es;
    if (staged) {
        const size_t want = img_elems * sizeof(float);
        if ((!workspace && want > 0) || workspace_bytes < want)
            return fail(MSDA_ERR_WORKSPACE, "msda_backward: workspace of %zu bytes required, got %zu", want,
                        workspace_bytes);
        if (!aligned(workspace, 16)) return fail(MSDA_ERR_WORKSPACE, "msda_backward: workspace must be 16-byte aligned");
        accum = workspace;
        accum_bytes = img_elems * sizeof(float);
    }
    if (need_img && img_elems > 0) {
        cudaError_t e = cudaMemsetAsync(accum, 0, accum_bytes, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward zero-fill");
    }
    if (prob->B == 0 || prob->Q == 0) return MSDA_OK;
    if (!grad_out || !img || !img_shapes || !sampling_points || !attention_weights)
        return fail(MSDA_ERR_NULL_POINTER, "msda_backward: NULL device pointer");
    if (!aligned(sampling_points, 2 * es) || !aligned(grad_points, 2 * es) || !aligned(img_shapes, 8))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_backward: sampling_points / grad_points must be aligned to one (x,y) pair");

    // accumulation rows are CT-typed: alignment requirement scales with sizeof(CT)/sizeof(T)
    const size_t acc_es = prob->dtype == MSDA_DTYPE_F64 ? 8 : 4;
    int vec = pick_vec(prob, {img, grad_out});
    while (vec > 1 && need_img && !aligned(accum, vec * acc_es > 16 ? 16 : vec * acc_es)) vec >>= 1;

    // kernel arguments for a given lane vector width
    auto bind = [&](int v) {
        msda::KernelArgs k;
        fill_args(k, prob, v);
        k.img = img;
        k.shapes = reinterpret_cast<const long long *>(img_shapes);
        k.pts = sampling_points;
        k.aw = attention_weights;
        k.gout = grad_out;
        k.gimg = accum;
        k.gpts = grad_points;
        k.gaw = grad_weights;
        k.flags = flags & MSDA_BWD_NEED_ALL;
        return k;
    };

    cudaError_t e = cudaErrorNotSupported;
    const bool tiled_ok = !force_generic() && vec * es == 16 && aligned(sampling_points, 16) &&
                          aligned(attention_weights, 8) && aligned(grad_points, 16) && aligned(grad_weights, 8) &&
                          aligned(accum, 16);
    if (tiled_ok) e = msda::launch_backward_tiled(bind(vec), prob->dtype, dev.sm_count, st);
    if (e == cudaErrorNotSupported) {
        // generic kernel, 16-bit storage: lanes of 4 channels make every red.v4 a whole, contiguous 16 bytes
        if (staged && vec > 4) vec = 4;
        e = msda::launch_backward_generic(bind(vec), prob->dtype, vec, dev.sm_count, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "msda_backward launch");

    if (staged) {
        e = msda::launch_round_grad_img(grad_img, static_cast<const float *>(accum), (long long)img_elems, prob->dtype,
                                        (int)prob->D, /*permuted_lanes=*/0, st);   // natural channel order on both paths
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward grad_img rounding");
    }
    return MSDA_OK;
}

int msda_module_supported(const msda_problem *prob, int ref_dim) {
    if (validate(prob) != MSDA_OK) return 0;
    if (ref_dim != 2 && ref_dim != 4) return 0;
    if (prob->dtype == MSDA_DTYPE_F64 || (prob->D != 32 && prob->D != 64) || prob->L * prob->K != 16 || prob->L > 8)
        return 0;
    msda::KernelArgs a;
    fill_args(a, prob, 1);
    return msda::tiled_offsets_fit(a, dtype_size(prob->dtype)) ? 1 : 0;
}

int msda_module_forward(void *out, const void *value, const int64_t *img_shapes, const void *proj, const void *ref,
                        int ref_dim, const msda_problem *prob, void *stream) {
    int rc = validate(prob);
    if (rc != MSDA_OK) return rc;
    if (!msda_module_supported(prob, ref_dim))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_module_forward: unsupported problem (needs fp32/fp16/bf16, D in {32, 64}, "
                                        "L*K == 16, ref_dim 2 or 4); use msda_forward on materialised operands");
    if (prob->B == 0 || prob->Q == 0) return MSDA_OK;
    if (!out || !value || !img_shapes || !proj || !ref)
        return fail(MSDA_ERR_NULL_POINTER, "msda_module_forward: NULL device pointer");
    const size_t es = dtype_size(prob->dtype);
    if (!aligned(value, 16) || !aligned(out, 16) || !aligned(proj, 8) || !aligned(ref, 2 * es) || !aligned(img_shapes, 8))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_module_forward: value/out must be 16-byte, proj 8-byte aligned");
    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;
    msda::KernelArgs a;
    fill_args(a, prob, (int)(16 / es));
    a.img = value;
    a.shapes = reinterpret_cast<const long long *>(img_shapes);
    a.proj = proj;
    a.ref = ref;
    a.ref_dim = ref_dim;
    a.out = out;
    cudaError_t e = msda::launch_module_forward_tiled(a, prob->dtype, dev.sm_count, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "msda_module_forward launch");
    return MSDA_OK;
}

int msda_module_backward(void *grad_value, void *grad_proj, float *grad_ref, const void *grad_out, const void *value,
                         const int64_t *img_shapes, const void *proj, const void *ref, int ref_dim,
                         const msda_problem *prob, int flags, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = validate(prob);
    if (rc != MSDA_OK) return rc;
    if (!msda_module_supported(prob, ref_dim))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_module_backward: unsupported problem (see msda_module_supported)");
    if (flags & MSDA_BWD_DETERMINISTIC)
        return fail(MSDA_ERR_BAD_MODE, "msda_module_backward: the deterministic mode is available through the unfused "
                                       "path only (msda_backward)");
    const bool need_img = flags & MSDA_BWD_NEED_IMG, need_proj = flags & (MSDA_BWD_NEED_POINTS | MSDA_BWD_NEED_WEIGHTS),
               need_ref = flags & MSDA_BWD_NEED_REF;
    if (!need_img && !need_proj && !need_ref) return MSDA_OK;
    const bool staged = need_img && prob->dtype != MSDA_DTYPE_F32;
    const bool want_colsum = (flags & MSDA_BWD_VALUE_COLSUM) != 0;
    if (want_colsum && !(staged && msda_module_colsum_supported(prob)))
        return fail(MSDA_ERR_BAD_MODE, "msda_module_backward: MSDA_BWD_VALUE_COLSUM needs MSDA_BWD_NEED_IMG, fp16/bf16 "
                                       "storage and msda_module_colsum_supported(prob)");
    const size_t es = dtype_size(prob->dtype);
    const size_t img_elems = (size_t)prob->B * prob->Npix * prob->H * prob->D;
    const bool no_units = prob->B == 0 || prob->Q == 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceInfo dev;
    rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;
    if ((need_img && !grad_value && img_elems > 0) || (!no_units && ((need_proj && !grad_proj) || (need_ref && !grad_ref))))
        return fail(MSDA_ERR_NULL_POINTER, "msda_module_backward: a requested gradient buffer is NULL");

    void *accum = grad_value;
    size_t accum_bytes = img_elems * es;
    if (staged) {
        const size_t want = want_colsum ? msda_module_colsum_offset(prob) + sizeof(float) * (size_t)prob->H * prob->D
                                        : img_elems * sizeof(float);
        if ((!workspace && want > 0) || workspace_bytes < want || !aligned(workspace, 16))
            return fail(MSDA_ERR_WORKSPACE, "msda_module_backward: 16-byte aligned workspace of %zu bytes required, got %zu",
                        want, workspace_bytes);
        accum = workspace;
        accum_bytes = img_elems * sizeof(float);
    }
    float *colsum = want_colsum
        ? reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + msda_module_colsum_offset(prob)) : nullptr;
    cudaError_t e;
    if (need_img && img_elems > 0) {
        // no queries: nothing will be accumulated or rounded, so the zeros go straight into grad_value
        e = (staged && no_units) ? cudaMemsetAsync(grad_value, 0, img_elems * es, st)
                                 : cudaMemsetAsync(accum, 0, accum_bytes, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_module_backward zero-fill");
    }
    if (colsum && (no_units || img_elems == 0)) {   // nothing is accumulated: the sums are zero
        e = cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)prob->H * prob->D, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_module_backward zero-fill of the column sums");
    }
    if (need_ref && !no_units) {
        e = cudaMemsetAsync(grad_ref, 0, sizeof(float) * (size_t)prob->B * prob->Q * ref_dim, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_module_backward zero-fill of grad_ref");
    }
    if (no_units) return MSDA_OK;
    if (!grad_out || !value || !img_shapes || !proj || !ref)
        return fail(MSDA_ERR_NULL_POINTER, "msda_module_backward: NULL device pointer");
    if (!aligned(value, 16) || !aligned(grad_out, 16) || !aligned(accum, 16) || !aligned(proj, 8) ||
        !aligned(grad_proj, 8) || !aligned(ref, 2 * es) || !aligned(img_shapes, 8))
        return fail(MSDA_ERR_BAD_SHAPE, "msda_module_backward: misaligned pointer");

    msda::KernelArgs a;
    fill_args(a, prob, (int)(16 / es));
    a.img = value;
    a.shapes = reinterpret_cast<const long long *>(img_shapes);
    a.proj = proj;
    a.ref = ref;
    a.ref_dim = ref_dim;
    a.gout = grad_out;
    a.gimg = accum;
    a.gproj = grad_proj;
    a.gref = grad_ref;
    a.flags = (need_img ? MSDA_BWD_NEED_IMG : 0) | (need_proj ? (MSDA_BWD_NEED_POINTS | MSDA_BWD_NEED_WEIGHTS) : 0) |
              (need_ref ? MSDA_BWD_NEED_REF : 0);
    e = msda::launch_module_backward_tiled(a, prob->dtype, dev.sm_count, st);
    if (e != cudaSuccess) return fail_cuda(e, "msda_module_backward launch");
    if (staged) {
        e = msda::launch_round_grad_img(grad_value, static_cast<const float *>(accum), (long long)img_elems, prob->dtype,
                                        (int)prob->D, (int)(prob->D / 8), st, colsum, (int)(prob->H * prob->D));
        if (e != cudaSuccess) return fail_cuda(e, "msda_module_backward grad_value rounding");
    }
    return MSDA_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------------------------
// Level table as a stand-alone kernel + the L2 roof probes
// -------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void level_table_kernel(int32_t *__restrict__ table, const long long *__restrict__ shapes, int L, int Npix) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long run = 0;
        for (int l = 0; l < L; ++l) {
            const long long h = shapes[2 * l], w = shapes[2 * l + 1];
            table[4 * l + 0] = (int32_t)h;
            table[4 * l + 1] = (int32_t)w;
            table[4 * l + 2] = (int32_t)run;
            table[4 * l + 3] = 0;
            run += h * w;
        }
        table[4 * L + 0] = (int32_t)run;
        table[4 * L + 1] = Npix;
        table[4 * L + 2] = run == (long long)Npix;
        table[4 * L + 3] = 0;
    }
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// One 8-lane group per 128-byte row, 8 independent rows in flight per lane: the access shape of the MSDA gathers.
__global__ void __launch_bounds__(512) probe_gather_kernel(float *__restrict__ sink, const float *__restrict__ buf,
                                                           long long buf_rows, long long rows, uint32_t seed) {
    // Row indices by multiplicative hashing (two integer instructions per row: the first version spent 15 on a full
    // avalanche hash and was ALU-bound at 71 % of the pipe, i.e. it measured its own index arithmetic), 16 independent
    // 128-bit loads in flight per lane, 8 lanes per 128-byte row like the kernels' gathers.
    const int j = threadIdx.x & 7;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) >> 3;
    int shift = 0;                                     // buf_rows is rounded down to a power of two
    while ((2ll << shift) <= buf_rows) ++shift;
    const unsigned down = 32u - (unsigned)shift;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = group * 16; r < rows; r += ngroups * 16) {
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t row = (((uint32_t)r + (uint32_t)i) * 2654435761U + seed) >> down;
            v[i] = __ldg(reinterpret_cast<const float4 *>(buf + (size_t)row * 32 + j * 4));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc.x += v[i].x;
            acc.y += v[i].y;
            acc.z += v[i].z;
            acc.w += v[i].w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123456.789f) sink[0] = acc.x;  // keeps the loads alive
}

__global__ void __launch_bounds__(512) probe_scatter_kernel(float *__restrict__ buf, long long buf_rows, long long rows,
                                                            uint32_t seed) {
    const int j = threadIdx.x & 7;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) >> 3;
    for (long long r = group * 8; r < rows; r += ngroups * 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t hsh = mix32((uint32_t)(r + i) * 2654435761U + seed);
            const long long row = (long long)(((unsigned long long)hsh * (unsigned long long)buf_rows) >> 32);
            msda::red_add_v4(buf + row * 32 + j * 4, 1.0f, 1.0f, 1.0f, 1.0f);
        }
    }
}

}  // namespace

extern "C" {

int msda_level_table(int32_t *table, const int64_t *img_shapes, int64_t L, int64_t Npix, void *stream) {
    if (!table || !img_shapes) return fail(MSDA_ERR_NULL_POINTER, "msda_level_table: NULL device pointer");
    if (L <= 0 || L > 2048 || Npix < 0 || Npix > 0x7fffffffLL) return fail(MSDA_ERR_BAD_SHAPE, "msda_level_table: bad L / Npix");
    level_table_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        table, reinterpret_cast<const long long *>(img_shapes), (int)L, (int)Npix);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "msda_level_table launch");
    return MSDA_OK;
}

int msda_probe_gather(float *sink, const float *buf, int64_t buf_rows, int64_t rows, uint32_t seed, void *stream) {
    if (!sink || !buf || buf_rows <= 0 || rows <= 0) return fail(MSDA_ERR_BAD_SHAPE, "msda_probe_gather: bad arguments");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;
    probe_gather_kernel<<<dev.sm_count * 2, 512, 0, static_cast<cudaStream_t>(stream)>>>(sink, buf, buf_rows, rows, seed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "msda_probe_gather launch");
    return MSDA_OK;
}

int msda_probe_scatter(float *buf, int64_t buf_rows, int64_t rows, uint32_t seed, void *stream) {
    if (!buf || buf_rows <= 0 || rows <= 0) return fail(MSDA_ERR_BAD_SHAPE, "msda_probe_scatter: bad arguments");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc != MSDA_OK) return rc;
    probe_scatter_kernel<<<dev.sm_count * 2, 512, 0, static_cast<cudaStream_t>(stream)>>>(buf, buf_rows, rows, seed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "msda_probe_scatter launch");
    return MSDA_OK;
}

}  // extern "C"
