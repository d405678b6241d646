/*
 * ganet_b200.h -- C ABI of the B200-native GANet guided-aggregation library
 * (libganet_b200.so).
 *
 * This is the drop-in boundary for the reference's native layer
 * (libs/GANet/src/GANet_cuda.cpp:5-75, pybind module "GANet" with
 * sga/lga/lga3d_cuda_{forward,backward}, all taking at::Tensor by value).
 * Here every entry point is plain C: device pointers, int64 sizes and a CUDA
 * stream handle; no torch types.  The reference-side binding a maintainer
 * would add (ctypes / a 20-line cpp_extension) is shown in INTEGRATION.md.
 *
 * Conventions (differences from the reference are deliberate and listed):
 *   - all tensors fp32, contiguous, NC(D)HW, resident on the CURRENT device;
 *   - the caller allocates every buffer, the library allocates nothing, keeps
 *     no global state and never synchronises (reference: libs/GANet/functions/
 *     GANet.py:14-16 allocates and zero-fills in Python; same ownership);
 *   - outputs are OVERWRITTEN -- they need not be zero-filled (the reference
 *     accumulates with += into buffers the caller must zero, GANet_kernel.cu
 *     :227, :1168, :1207); `ganet_lga_backward` has an explicit flag for the
 *     one place where accumulation is part of the contract;
 *   - work is enqueued on `stream` (a cudaStream_t; NULL = legacy default
 *     stream, which is what the reference always uses, GANet_kernel.cu:961);
 *   - the direction mask is 1 byte per voxel (reference: an fp32 volume,
 *     functions/GANet.py:16) and the reference's saved `temp_out` volume is
 *     not needed: backward recomputes every direction on chip / in scratch;
 *   - element counts are 64-bit (reference: `int`, silently overflows at 2^31,
 *     GANet_kernel.cu:960);
 *   - return value: 0 on success, a positive cudaError_t value if a launch
 *     failed, a negative GANET_E* code for argument errors (reference: always
 *     returns 1, GANet_cuda.cpp:11).  ganet_error_string() decodes both.
 *
 * Thread safety: re-entrant; may be called concurrently from several host
 * threads on different devices/streams (DataParallel pattern).
 */
#ifndef GANET_B200_H
#define GANET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GANET_B200_ABI_VERSION 2

typedef void *ganet_stream_t; /* cudaStream_t */

enum {
    GANET_OK = 0,
    GANET_EINVAL = -1,     /* null pointer / non-positive dimension */
    GANET_EUNSUPPORTED = -2, /* shape outside the compiled kernel range (D > 768, radius > 3) */
    GANET_EWORKSPACE = -3, /* workspace too small */
    GANET_EALIGN = -4      /* pointer not 16-byte aligned */
};

int ganet_abi_version(void);
const char *ganet_error_string(int code);

/* ------------------------------------------------------------------------
 * SGA -- semi-global guided aggregation.
 * Replaces sga_cuda_forward (GANet_cuda.cpp:39-48 -> sga_kernel_forward,
 * GANet_kernel.cu:935-998: four directional scans + Max).
 *   x, out : (N, C, D, H, W)
 *   g_down, g_up, g_right, g_left : (N, C, 5, H, W), the caller has already
 *       L1-normalised them over dim 2 (models/GANet_deep.py:265-268)
 *   mask   : (N, C, D, H, W) uint8, winning direction 0=down 1=up 2=right
 *       3=left, ties keep the lower id (GANet_kernel.cu:23-36)
 *   aggregates : optional (may be NULL).  ganet_sga_aggregate_volumes(...) * N*C*D*H*W
 *       floats that receive the four directional aggregates for ganet_sga_backward, which
 *       then skips its four recompute passes: 4 volumes, all (N,C,D,H,W), when the
 *       horizontal scans run in the standard layout (D <= 256, W % 16 == 0); otherwise 5
 *       -- down, up as (N,C,D,H,W), right, left TRANSPOSED as (N,C,D,W,H), and the H<->W
 *       transposed input.  +16 / +20 bytes per voxel of saved state -- the reference saves
 *       8: temp_out and an fp32 mask, functions/GANet.py:21.  Needs D <= 288.
 *   workspace : device scratch (transposed copies for the horizontal scans),
 *       >= ganet_sga_forward_workspace_min bytes; the (n,c) slices are processed
 *       in chunks that fit, so any size between _min and _best works
 * Values of `out` and `mask` are bit-identical to the reference CUDA build, with or
 * without `aggregates`.
 */
int ganet_sga_forward(const float *x, const float *g_down, const float *g_up,
                      const float *g_right, const float *g_left, float *out,
                      uint8_t *mask, float *aggregates, void *workspace,
                      size_t workspace_bytes, int64_t N, int64_t C, int64_t D, int64_t H,
                      int64_t W, ganet_stream_t stream);

/* Bytes of scratch the SGA calls need at least (one (n,c) slice in flight) and
 * the size at which they run all slices in one chunk. */
size_t ganet_sga_forward_workspace_min(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W);
size_t ganet_sga_forward_workspace_best(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W);
size_t ganet_sga_backward_workspace_min(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W);
size_t ganet_sga_backward_workspace_best(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W);

/* Volumes of N*C*D*H*W floats the optional `aggregates` buffer must hold for this shape (4 or
 * 5, see ganet_sga_forward).  No reference counterpart: the reference keeps temp_out + mask. */
int ganet_sga_aggregate_volumes(int64_t N, int64_t C, int64_t D, int64_t H, int64_t W);

/*
 * Replaces sga_cuda_backward (GANet_cuda.cpp:50-64 -> sga_kernel_backward,
 * GANet_kernel.cu:1000-1129).  Reproduces the reference's gradient, including
 * its first-scan-step quirks.
 *   grad_out, grad_in : (N, C, D, H, W);  gg_* : (N, C, 5, H, W), overwritten
 *   aggregates : NULL, or the buffer ganet_sga_forward filled for the same inputs
 *   max_idx : optional (N, C, H, W) int32, depth arg-max of the `right`
 *       aggregate -- what the reference leaves in its max_idx buffer
 *       (GANet_kernel.cu:1119); may be NULL
 *   workspace : device scratch, >= ganet_sga_backward_workspace_min bytes;
 *       slices are processed in chunks that fit
 */
int ganet_sga_backward(const float *x, const float *g_down, const float *g_up,
                       const float *g_right, const float *g_left, const uint8_t *mask,
                       const float *aggregates, const float *grad_out, float *grad_in,
                       float *gg_down, float *gg_up, float *gg_right, float *gg_left,
                       int32_t *max_idx, void *workspace, size_t workspace_bytes, int64_t N,
                       int64_t C, int64_t D, int64_t H, int64_t W, ganet_stream_t stream);

/* One directional aggregate without the max-combine (test / debug aid; the
 * reference exposes the `left` one as temp_out, functions/GANet.py:15,21).
 * dir: 0 down, 1 up, 2 right, 3 left.  a: (N, C, D, H, W). */
int ganet_sga_direction(const float *x, const float *g, float *a, int dir, int64_t N,
                        int64_t C, int64_t D, int64_t H, int64_t W, ganet_stream_t stream);

/* ------------------------------------------------------------------------
 * LGA -- local guided aggregation, one pass.
 * Replaces lga_cuda_forward / lga3d_cuda_forward (GANet_cuda.cpp:14-20, :31-37
 * -> lga_filtering_forward, GANet_kernel.cu:1131-1175).
 *   x, y : (B, D, H, W);  f : (B, 3*(2R+1)^2, H, W)
 * B is the batch for the 4-D op and batch*channels for the 5-D (lga3d) op.
 * y must not alias x.
 */
int ganet_lga_forward(const float *x, const float *f, float *y, int64_t B, int64_t D,
                      int64_t H, int64_t W, int radius, ganet_stream_t stream);

/*
 * Replaces lga_cuda_backward / lga3d_cuda_backward (GANet_cuda.cpp:5-12, :22-29
 * -> lga_filter_backward + lga_data_backward, GANet_kernel.cu:1177-1269).
 *   grad_x is overwritten; grad_f is overwritten when accumulate_f == 0 and
 *   accumulated (+=) when accumulate_f != 0 (Lga2Function.backward sums the
 *   filter gradient of its two passes, functions/GANet.py:197-199).
 * grad_x must not alias grad_out or x.
 */
int ganet_lga_backward(const float *x, const float *f, const float *grad_out, float *grad_x,
                       float *grad_f, int accumulate_f, int64_t B, int64_t D, int64_t H,
                       int64_t W, int radius, ganet_stream_t stream);

/* ------------------------------------------------------------------------
 * GetCostVolume (libs/GANet/modules/GANet.py:119-134; a Python loop of 2*Dm
 * slice copies in the reference).
 *   x, y : (N, C, H, W) -> cost : (N, 2C, Dm, H, W), Dm = maxdisp + 1
 */
int ganet_cost_volume_forward(const float *x, const float *y, float *cost, int64_t N,
                              int64_t C, int64_t Dm, int64_t H, int64_t W,
                              ganet_stream_t stream);
int ganet_cost_volume_backward(const float *grad_cost, float *grad_x, float *grad_y,
                               int64_t N, int64_t C, int64_t Dm, int64_t H, int64_t W,
                               ganet_stream_t stream);

/* ------------------------------------------------------------------------
 * DisparityRegression (libs/GANet/modules/GANet.py:142-148).
 *   p : (N, Dm, H, W) -> disp : (N, H, W) = sum_d d * p[:, d]
 */
int ganet_disp_regression_forward(const float *p, float *disp, int64_t N, int64_t Dm,
                                  int64_t H, int64_t W, ganet_stream_t stream);
int ganet_disp_regression_backward(const float *grad_disp, float *grad_p, int64_t N,
                                   int64_t Dm, int64_t H, int64_t W, ganet_stream_t stream);

/* DispAgg tail (partial).  Replaces, in the reference's MODELS, models/GANet_deep.py:245-247:
 * F.normalize(x, p=1, dim=1) followed by DisparityRegression -- one pass over x each way.
 *   x : (N, Dm, H, W);  disp : (N, H, W) = sum_d d*x_d / max(sum_d |x_d|, 1e-12);  norm : (N, H, W) scratch
 *   kept for backward (sum_d |x_d|, unclamped). */
int ganet_norm_disp_regression_forward(const float *x, float *disp, float *norm, int64_t N, int64_t Dm,
                                       int64_t H, int64_t W, ganet_stream_t stream);
int ganet_norm_disp_regression_backward(const float *x, const float *disp, const float *norm,
                                        const float *grad_disp, float *grad_x, int64_t N, int64_t Dm,
                                        int64_t H, int64_t W, ganet_stream_t stream);

/* ------------------------------------------------------------------------
 * SGABlock prologue / epilogue.  Replaces, in the reference's MODELS,
 * models/GANet_deep.py:264-268 (GANet11.py:246-250): torch.split of the guidance conv output into four
 * (N, C*5, H, W) blocks, .view(N, C, 5, H, W) and F.normalize(p=1, dim=2) of each -- and the autograd
 * graph of that -- by one streaming pass each way.
 *   raw    : (N, 4*C*5, H, W), the guidance conv output, direction-major (down, up, right, left)
 *   g_*    : (N, C, 5, H, W) = raw block / max(sum_j |raw block_j|, 1e-12), bit-identical to F.normalize
 *   gg_*   : gradients w.r.t. g_* (what ganet_sga_backward returns); grad_raw : (N, 4*C*5, H, W)
 * H*W must be a multiple of 4.
 */
int ganet_sga_guidance_forward(const float *raw, float *g_down, float *g_up, float *g_right,
                               float *g_left, int64_t N, int64_t C, int64_t H, int64_t W,
                               ganet_stream_t stream);
int ganet_sga_guidance_backward(const float *raw, const float *gg_down, const float *gg_up,
                                const float *gg_right, const float *gg_left, float *grad_raw,
                                int64_t N, int64_t C, int64_t H, int64_t W, ganet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GANET_B200_H */
