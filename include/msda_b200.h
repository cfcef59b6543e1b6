/*
 * msda_b200.h -- C ABI of libmsda_b200.so: multiscale deformable attention (MSDA) for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's kernel layer.  The reference (rziga/msda-triton) has exactly two
 * operator entry points, both Python functions that launch Triton kernels:
 *
 *   triton_multi_scale_deformable_attention_fwd   src/msda_triton/kernels.py:351-379   -> msda_forward
 *   triton_multi_scale_deformable_attention_bwd   src/msda_triton/kernels.py:556-592   -> msda_backward
 *
 * and one piece of device-side preprocessing that every Triton program repeats:
 *
 *   load_shapes_and_level_offsets                 src/msda_triton/kernels.py:44-64     -> done inside the kernels
 *                                                                                         (once per CTA), and exposed
 *                                                                                         as msda_level_table
 *
 * Conventions
 *   - All pointers are DEVICE pointers on the current CUDA device (including img_shapes, like the reference, which
 *     reads the [L,2] int64 (h,w) table on device and never syncs: frontend.py:93-95, kernels.py:52-56).
 *   - All tensors are dense row-major ("contiguous") in the reference's layouts:
 *       img   [B, Npix, H, D]       Npix = sum_l h_l*w_l        (frontend.py:157)
 *       img_shapes [L, 2] int64, (h, w) order                     (frontend.py:158)
 *       sampling_points [B, Q, H, L, K, 2], (x, y) in [0,1]       (frontend.py:159)
 *       attention_weights [B, Q, H, L, K]                         (frontend.py:160)
 *       out / grad_out [B, Q, H, D]                               (frontend.py:165)
 *   - The caller owns every buffer.  The library never allocates device memory, never synchronises, and launches only
 *     on the stream it is given (cudaStream_t passed as void*; NULL = legacy default stream).  All entry points are
 *     therefore CUDA-graph capturable.
 *   - Return value: 0 on success, a negative MSDA_ERR_* code on invalid arguments, a positive cudaError_t value if
 *     a CUDA call failed.  msda_last_error() returns a thread-local, human-readable description of the last failure.
 *   - Thread safety: the library may be called concurrently from several host threads (autograd calls backward from
 *     its own thread).  Its state: the thread-local error string; a per-device cache of immutable device properties;
 *     the measurement / test knobs read ONCE from the environment (MSDA_B200_*, see msda_reload_tuning); and a ring of
 *     256 device words ("pace counters", msda_pace.cu) from which a launch that walks several L2-sized waves takes the
 *     next one and zeroes it on its stream.  A captured CUDA graph keeps the counter it was captured with; a replay
 *     that overlaps other multi-wave launches whose ticket wrapped onto the same word may pass a wave early or sit
 *     out one bounded wait (~130 us) -- results are unaffected (the pacing is a performance hint only).
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 1

/* storage dtype of img / sampling_points / attention_weights / out (all four share it) */
enum msda_dtype {
    MSDA_DTYPE_F32 = 0,  /* compute + accumulate fp32 */
    MSDA_DTYPE_F16 = 1,  /* fp16 storage, fp32 compute + accumulate */
    MSDA_DTYPE_BF16 = 2, /* bf16 storage, fp32 compute + accumulate (the reference rejects bf16: kernels.py:40-41) */
    MSDA_DTYPE_F64 = 3   /* compute + accumulate fp64 */
};

/* reference: padding_mode Literal["border","zeros"] (frontend.py:150) */
enum msda_padding { MSDA_PAD_ZEROS = 0, MSDA_PAD_BORDER = 1 };

/* msda_backward flags */
enum msda_bwd_flags {
    MSDA_BWD_NEED_IMG = 1,      /* produce grad_img            (ctx.needs_input_grad[0]) */
    MSDA_BWD_NEED_POINTS = 2,   /* produce grad_sampling_points (ctx.needs_input_grad[2]) */
    MSDA_BWD_NEED_WEIGHTS = 4,  /* produce grad_attention_weights (ctx.needs_input_grad[3]) */
    MSDA_BWD_NEED_ALL = 7,
    MSDA_BWD_DETERMINISTIC = 8, /* grad_img by sorted-segment reduction instead of atomics (bit-reproducible) */
    MSDA_BWD_NEED_REF = 16,     /* msda_module_backward only: produce grad_reference_points */
    MSDA_BWD_VALUE_COLSUM = 32  /* msda_module_backward only, fp16/bf16 storage: also leave sum over (b, pixel) of
                                   grad_value[b, pixel, h, c] -- the bias gradient of the projection that produced
                                   `value` (frontend.py:259, img_input_proj) -- as H*D floats at
                                   workspace + msda_module_colsum_offset(prob); see msda_module_backward */
};

enum msda_error {
    MSDA_OK = 0,
    MSDA_ERR_NULL_POINTER = -1,
    MSDA_ERR_BAD_DTYPE = -2,
    MSDA_ERR_BAD_SHAPE = -3,
    MSDA_ERR_BAD_MODE = -4,
    MSDA_ERR_WORKSPACE = -5,
    MSDA_ERR_UNSUPPORTED_DEVICE = -6
};

/* Problem description shared by every entry point (sizes follow BASELINE.json naming). */
typedef struct msda_problem {
    int64_t B;    /* batch */
    int64_t Npix; /* total pyramid pixels, sum_l h_l*w_l */
    int64_t H;    /* heads */
    int64_t D;    /* channels per head */
    int64_t Q;    /* queries */
    int64_t L;    /* pyramid levels */
    int64_t K;    /* sampling points per level */
    int32_t dtype;         /* enum msda_dtype */
    int32_t padding_mode;  /* enum msda_padding */
    int32_t align_corners; /* 0 / 1 */
    int32_t reserved;      /* must be 0 */
} msda_problem;

int msda_abi_version(void);
const char *msda_last_error(void);

/*
 * The library reads its measurement / test knobs (environment variables MSDA_B200_FORCE_GENERIC, _SLICES_PER_WAVE,
 * _PACE_SLACK, _WAVE_PACING, _FWD_VARIANT, _BWD_SPLIT, _SPLIT_SLOTS, _BWD_OWNER, _OWNER_ROWS, _OWNER_WORKERS,
 * _DET_VARIANT, _BWD_DENSE, _DENSE_PF, _BWD_SHAPE, _CARVEOUT) once, at the first call that needs them, never on the launch path.  A process that changes one of them
 * afterwards calls this to have them read again.  No reference counterpart (the reference's only knob is Triton's
 * autotuner, kernels.py:259-265).
 */
void msda_reload_tuning(void);

/*
 * Forward: out[b,q,h,:] = sum_{l,k} w[b,q,h,l,k] * bilinear(img_l[b,:,h,:], p[b,q,h,l,k])
 * Replaces kernels.py:351-379 (wrapper) + :267-348 (kernel) + :120-252 (sample_bilinear).
 */
int msda_forward(void *out, const void *img, const int64_t *img_shapes, const void *sampling_points,
                 const void *attention_weights, const msda_problem *prob, void *stream);

/*
 * Bytes of scratch msda_backward needs for `prob` and `flags` (0 if none).
 *   - fp16/bf16 storage: grad_img is accumulated in an fp32 image of B*Npix*H*D floats, then rounded once.
 *   - MSDA_BWD_DETERMINISTIC: sort keys/values + segment scratch.
 */
size_t msda_backward_workspace_bytes(const msda_problem *prob, int flags);

/*
 * Backward: grads of out w.r.t. img (scatter-add), sampling_points (w.r.t. the normalised [0,1] coordinate) and
 * attention_weights, computed in ONE pass that also recomputes the forward sampling.
 * Replaces kernels.py:556-592 (wrapper, incl. its three zero-fills) + :396-553 (kernel).
 * grad_img is zero-filled by the library on `stream`; grad_points / grad_weights are fully overwritten.
 * Pointers whose MSDA_BWD_NEED_* bit is clear may be NULL.
 */
int msda_backward(void *grad_img, void *grad_points, void *grad_weights, const void *grad_out, const void *img,
                  const int64_t *img_shapes, const void *sampling_points, const void *attention_weights,
                  const msda_problem *prob, int flags, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Fused module core -- the part of MultiscaleDeformableAttention.forward between the projections
 * (src/msda_triton/frontend.py:253-289): softmax of the attention logits over L*K, sampling_points =
 * reference + offset / level shape (2-d references; x is divided by the level HEIGHT and y by the WIDTH exactly as the
 * reference does, frontend.py:272-276) or reference_xy + offset * reference_wh / (2K) (4-d, frontend.py:278-282),
 * then the MSDA operator -- in ONE kernel, without materialising sampling_points / attention_weights.
 *   value  [B, Npix, H, D]        projected pyramid
 *   proj   [B, Q, H, L, K, 3]     query projection viewed as (offset x, offset y, attention logit) triples
 *   ref    [B, Q, ref_dim]        reference points, ref_dim = 2 or 4
 * msda_module_supported returns 1 when the fused kernels cover `prob` (fp32/fp16/bf16, D in {32, 64}, L*K == 16); otherwise
 * the caller composes the unfused pieces (softmax etc. + msda_forward).
 * Backward flags: MSDA_BWD_NEED_IMG -> grad_value, NEED_POINTS|NEED_WEIGHTS -> grad_proj, NEED_REF -> grad_ref.
 * grad_ref is an fp32 [B, Q, ref_dim] buffer regardless of the storage dtype; the library zero-fills it and grad_value.
 * Workspace: msda_backward_workspace_bytes(prob, flags) (fp32 accumulation image for 16-bit storage).
 * MSDA_BWD_VALUE_COLSUM (with NEED_IMG, fp16/bf16 storage, msda_module_colsum_supported(prob) == 1): the rounding pass
 * that turns the fp32 accumulation image into grad_value also sums it over (b, pixel); on completion the H*D fp32 column
 * sums sit at (char *)workspace + msda_module_colsum_offset(prob).  They are what autograd would compute as
 * grad_value.sum over rows for the bias of the value projection (frontend.py:259) with a separate reduction kernel.
 * msda_backward_workspace_bytes(prob, flags) includes the extra H*D floats when the flag is passed.
 */
int msda_module_supported(const msda_problem *prob, int ref_dim);
int msda_module_colsum_supported(const msda_problem *prob);
size_t msda_module_colsum_offset(const msda_problem *prob);
int msda_module_forward(void *out, const void *value, const int64_t *img_shapes, const void *proj, const void *ref,
                        int ref_dim, const msda_problem *prob, void *stream);
int msda_module_backward(void *grad_value, void *grad_proj, float *grad_ref, const void *grad_out, const void *value,
                         const int64_t *img_shapes, const void *proj, const void *ref, int ref_dim,
                         const msda_problem *prob, int flags, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Device-side level preprocessing (kernels.py:44-64): table[l] = {h_l, w_l, offset_l, 0} as int32, plus
 * table[L] = {sum_l h_l*w_l, Npix, sum==Npix, 0} so a caller can validate asynchronously.  table: (L+1)*4 int32.
 */
int msda_level_table(int32_t *table, const int64_t *img_shapes, int64_t L, int64_t Npix, void *stream);

/*
 * Measurement helpers for the L2 gather / atomic roof (SURVEY.md section 8d "L2_peak must be measured on the box").
 * Each launch issues `rows` independent random 128-byte row reads (or red.global.add.v4.f32 row adds) over a
 * buffer of `buf_rows` rows of 128 B, one 8-lane group per row, the same access shape the MSDA kernels use.
 */
int msda_probe_gather(float *sink, const float *buf, int64_t buf_rows, int64_t rows, uint32_t seed, void *stream);
int msda_probe_scatter(float *buf, int64_t buf_rows, int64_t rows, uint32_t seed, void *stream);

/*
 * Query-sharded MSDA over NVLink peer memory (no reference counterpart: the reference has no distributed code; the only
 * coupling between output rows is the atomic accumulation into grad_img, kernels.py:549-553, which is what has to be
 * summed across the ranks that share an image).  One process per GPU; the buffers the peers read are allocated in
 * symmetric memory by the host side (msda_triton/distributed.py: PeerPixelExchange), which passes every rank's device
 * pointers here.  fp32 pyramids; all sizes per image and rank.
 *
 *   msda_peer_all_gather      every rank's pyramid[b, r*chunk .. (r+1)*chunk) <- this rank r's pixel shard of image b
 *                             (pushed over NVLink); on return of the kernel this rank's pyramid holds all shards
 *   msda_peer_reduce_scatter  grad_shard[b, i] <- sum over ranks r (ascending) of partial_r[b, my_rank*chunk + i]
 *
 * Every rank of the group must issue the same sequence of calls.  The kernels spin on flags written by the peers (one
 * launch of `SM count` CTAs each) and keep their call counts on the device, so a launch has no call-dependent argument:
 * a whole step (all_gather, forward, backward, reduce_scatter) may be captured into one CUDA graph and replayed.
 */
typedef struct msda_peer_ctx {
    int32_t world, rank;
    const void *const *peer_pyramids;  /* host array [world]: device address of every rank's gathered pyramid [B, world*chunk, H, D] */
    const void *const *peer_partials;  /* host array [world]: every rank's partial grad_img [B, world*chunk, H, D] (fp32) */
    uint32_t *const *peer_flags;       /* host array [world]: every rank's flag block, uint32 [4][world], zero-initialised */
    uint32_t *counters;                /* this rank's 8 counter words (device), zero-initialised once */
} msda_peer_ctx;

int msda_peer_all_gather(const void *shard, const msda_peer_ctx *ctx, int64_t B, int64_t shard_bytes_per_image, void *stream);
int msda_peer_reduce_scatter(void *grad_shard, const msda_peer_ctx *ctx, int64_t B, int64_t shard_floats_per_image,
                             void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
