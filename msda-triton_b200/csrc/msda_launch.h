// msda_launch.h -- internal launcher declarations (not part of the C ABI; see include/msda_b200.h for that).
#pragma once
#include <cuda_runtime.h>

#include "msda_common.cuh"

namespace msda {

// dtype codes follow enum msda_dtype: 0 f32, 1 f16, 2 bf16, 3 f64.
cudaError_t launch_forward_generic(const KernelArgs &a, int dtype, int vec, int sm_count, cudaStream_t st);
cudaError_t launch_backward_generic(const KernelArgs &a, int dtype, int vec, int sm_count, cudaStream_t st);

// Tuned paths (D*sizeof(T) == 128 or 64 bytes per row, L*K == 16).  Return cudaErrorNotSupported when the problem
// does not fit so the caller falls through to the generic kernels.
cudaError_t launch_forward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);
cudaError_t launch_backward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);
// pieces of the above that live in their own translation units (compile time): point counts other than 16 (forward),
// more than 16 points per unit as sub-units (backward)
cudaError_t launch_forward_tiled_points(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);
cudaError_t launch_backward_tiled_split(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);

// Fused module core (raw projection + reference points in, see msda_tiled.cuh).  cudaErrorNotSupported when the
// problem is outside (fp32|fp16|bf16) x D in {32, 64} x L*K=16.
cudaError_t launch_module_forward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);
cudaError_t launch_module_backward_tiled(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);

// Experiments around the tuned fp32 backward, each in its own translation unit: n owner warps accumulating the coarsest
// level in registers (msda_bwd_dense.cu) and the launch shapes 2..8 of MSDA_B200_BWD_SHAPE (msda_bwd_shapes.cu).
cudaError_t launch_backward_dense(const KernelArgs &a, int sm_count, cudaStream_t st, int nown, int prefetch);
cudaError_t launch_backward_shape_variant(const KernelArgs &a, int shape, int sm_count, cudaStream_t st);

// Tuned backward with the coarse pyramid levels accumulated in tensor memory (msda_bwd_tmem.cu):
// fp32, D == 32, L*K == 16, K == 4, grad_img requested.  cudaErrorNotSupported otherwise.
cudaError_t launch_backward_tmem(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);

// Deterministic mode by exact (quantised) row adds: msda_bwd_detq.cu prepares the quanta in `workspace`, then runs the
// tuned backward (all requested gradients in one pass).  cudaErrorNotSupported outside fp32, D == 32, L*K == 16.
bool quant_backward_supported(const KernelArgs &a, int dtype);
size_t detq_workspace_bytes(const KernelArgs &a);
cudaError_t launch_backward_detq(const KernelArgs &a, int dtype, void *workspace, int sm_count, cudaStream_t st);
cudaError_t launch_backward_tiled_quant(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);

// grad_img alone, without gathers (split backward; msda_bwd_scatter.cu).  a.gimg = zero-filled fp32 accumulation image.
cudaError_t launch_backward_scatter(const KernelArgs &a, int dtype, int sm_count, cudaStream_t st);

// Deterministic grad_img (sorted-segment reduction, msda_bwd_det.cu).  a.gimg = grad_img in STORAGE dtype.
bool det_supported(const KernelArgs &a);
size_t det_workspace_bytes(const KernelArgs &a);
cudaError_t launch_backward_det(const KernelArgs &a, int dtype, int vec, void *workspace, int sm_count, cudaStream_t st);

// grad_img epilogue for 16-bit storage: rounds the fp32 accumulation image to T.
// permuted_lanes = D/8 when the image was written by the tuned kernels (channel-permuted rows), 0 for natural order.
// colsum != nullptr: additionally the fp32 column sums over (b, pixel) of the HD = H*D columns (zeroed here first); only
// for permuted rows with round_colsum_supported(); cudaErrorNotSupported otherwise.
cudaError_t launch_round_grad_img(void *dst, const float *src, long long n, int dtype, int D, int permuted_lanes,
                                  cudaStream_t st, float *colsum = nullptr, int HD = 0);
bool round_colsum_supported(int dtype, int D, int HD, long long n);

// Arrival counter for wave pacing (msda_pace.cu): *slot = a device word zeroed on `st`, or nullptr when pacing is off.
cudaError_t acquire_pace_counter(cudaStream_t st, unsigned **slot);
bool pacing_forced();   // MSDA_B200_WAVE_PACING=2: pace every multi-wave launch, whatever the wave size (tests)

}  // namespace msda
