/*
 * afr.h -- C ABI of libafr_b200.so: the alias-free resampling hot path of
 * MDFahimAnjum/AliasFree-Diffusion-Models-PyTorch as hand-written CUDA for sm_100a.
 *
 * The reference has no FFI: its boundary is two Python free functions and the block
 * classes that call them.  Each entry point below names the
 * reference code it replaces; INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - all tensors are dense NCHW device buffers owned by the caller; the library
 *     never allocates, frees or keeps device memory;
 *   - `taps*` are HOST pointers to N*N float32 row-major filter taps (the tensor
 *     returned by circularLowpassKernel, modules/filtrs.py:20-37); they are copied
 *     into the kernel's parameter space at launch, never uploaded separately;
 *   - `dtype`: AFR_F32 or AFR_BF16 storage; arithmetic is always fp32;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); calls only enqueue
 *     work, never synchronise, and are deterministic (no atomics);
 *   - return value: AFR_OK or an afr_status; afr_last_error() gives a thread-local
 *     human-readable message.  There is no CPU fallback.
 */
#ifndef AFR_H_
#define AFR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFR_VERSION 100      /* 0.1.0 */
#define AFR_MAX_TAPS 16      /* largest supported N (filter is N x N) */

typedef enum { AFR_F32 = 0, AFR_BF16 = 1 } afr_dtype;

typedef enum {
    AFR_OK = 0,
    AFR_ERR_BAD_SHAPE = 1,     /* non-positive dims, planes overflow */
    AFR_ERR_BAD_TAPS = 2,      /* N < 1, N > AFR_MAX_TAPS or NULL taps */
    AFR_ERR_BAD_DTYPE = 3,
    AFR_ERR_NULL_POINTER = 4,
    AFR_ERR_MISALIGNED = 5,    /* buffer not aligned to its element size */
    AFR_ERR_CUDA = 6,          /* launch / driver error, see afr_last_error() */
    AFR_ERR_UNSUPPORTED = 7    /* e.g. resampling factor != 2 */
} afr_status;

/* Kernel-family selection (testing / benchmarking aid; production code leaves AUTO).
 * AUTO: N == 3 -> TMA-staged row-streaming kernel when the shape allows, else the direct
 * register-strip kernel (both in their symmetric-tap variant when the filters are
 * D4-symmetric, as every circularLowpassKernel filter is); both filters N x N with
 * N in {2,4,5,6,7,8} -> compile-time-N strip kernels; anything else -> runtime-N kernels. */
typedef enum {
    AFR_PATH_AUTO = 0,
    AFR_PATH_DIRECT = 1,    /* register-strip kernel reading global memory directly */
    AFR_PATH_TMA = 2,       /* force the TMA tile kernel (error if shape unsupported) */
    AFR_PATH_GENERIC = 3,   /* force the runtime-N shared-memory kernels */
    AFR_PATH_DIRECT_GENERAL = 4, /* DIRECT without the symmetric-tap fast path (general 3x3 taps) */
    AFR_PATH_TMA_GENERAL = 5     /* TMA without the symmetric-tap fast path */
} afr_path;

int afr_version(void);
const char *afr_last_error(void);
const char *afr_status_string(int status);
/* Sets the kernel path for subsequent calls of this process; returns the old one. */
int afr_set_path(int path);
/* Name of the kernel the last call on this thread launched (for tests/profiles). */
const char *afr_last_kernel(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
uint64_t afr_launch_count(void);

/* custom_upsample(x, sinc_filter, factor=2)            modules/filtrs.py:79-94
 * x [B,C,H,W] (in_dtype) -> u [B,C,2H,2W] (out_dtype).  Zero-stuff x2 then depthwise
 * N x N cross-correlation with 'same' zero padding; no gain compensation.  The
 * reference always returns fp32 (filtrs.py:85); out_dtype lets bf16 pipelines keep bf16. */
int afr_up2x_fwd(const void *x, void *u, int B, int C, int H, int W,
                 const float *taps, int N, int in_dtype, int out_dtype, void *stream);

/* autograd adjoint of custom_upsample: du [B,C,2H,2W] -> dx [B,C,H,W] (H, W = input dims) */
int afr_up2x_bwd(const void *du, void *dx, int B, int C, int H, int W,
                 const float *taps, int N, int du_dtype, int dx_dtype, void *stream);

/* custom_upsample writing straight into a channel slice of a larger tensor -- the second half of the
 * `torch.cat([skip_x, x], dim=1)` buffer of Up_FF / Up_FFF (modules/ddpm_utils.py:344-345, 414), which saves
 * one read and one write of the upsampled tensor.  `u` points at element (b=0, c=0) of the slice inside the
 * destination; `out_batch_stride` is the destination's batch stride in ELEMENTS (>= C*2H*2W, 16-byte
 * multiples); channel, row and column strides are dense.  N == 3, W % 4 == 0 and a 32-byte aligned slice
 * base / batch stride are required: AFR_ERR_UNSUPPORTED otherwise (callers fall back to afr_up2x_fwd + a
 * copy). */
int afr_up2x_fwd_strided(const void *x, void *u, int B, int C, int H, int W, int64_t out_batch_stride,
                         const float *taps, int N, int in_dtype, int out_dtype, void *stream);

/* adjoint of the above: du is a channel slice (batch stride `du_batch_stride` elements) of the gradient of
 * the concatenated tensor; dx [B,C,H,W] dense.  Same restrictions. */
int afr_up2x_bwd_strided(const void *du, void *dx, int B, int C, int H, int W, int64_t du_batch_stride,
                         const float *taps, int N, int dtype, void *stream);

/* custom_downsample(x, jinc_filter, factor=2)          modules/filtrs.py:71-77
 * v [B,C,H,W] -> y [B,C,ceil(H/2),ceil(W/2)], contiguous (the reference returns a
 * strided view of the full-resolution convolution; values are identical). */
int afr_down2x_fwd(const void *v, void *y, int B, int C, int H, int W,
                   const float *taps, int N, int dtype, void *stream);

/* adjoint of custom_downsample: dy [B,C,ceil(H/2),ceil(W/2)] -> dv [B,C,H,W] */
int afr_down2x_bwd(const void *dy, void *dv, int B, int C, int H, int W,
                   const float *taps, int N, int dtype, void *stream);

/* Filtered nonlinearity  custom_downsample(gelu(custom_upsample(x + residual, up)), down)
 * modules/ddpm_utils.py:123-125, 129-131 (with `x = x + residual` of :128 fused when
 * residual != NULL), 137-139.  GELU is the exact erf form.  x, residual, y: [B,C,H,W]. */
int afr_filtered_gelu_fwd(const void *x, const void *residual, void *y,
                          int B, int C, int H, int W,
                          const float *taps_up, int N_up, const float *taps_down, int N_down,
                          int dtype, void *stream);

/* Forward with the normalise + affine step of the preceding GroupNorm(1, C) folded into the load
 * (modules/ddpm_utils.py:122-125 and 127-131):  y = filtered_gelu(x * scale[b,c] + shift[b,c] (+ residual)).
 * scale_dev / shift_dev: DEVICE float32 [B*C] (gamma_c * rstd_b and beta_c - mean_b * gamma_c * rstd_b).
 * The conv zero padding applies to the normalised tensor, exactly as upstream.
 * N_up == N_down == 3 and W % 4 == 0 required (AFR_ERR_UNSUPPORTED otherwise); adjoint: ..._affine_bwd. */
int afr_filtered_gelu_affine_fwd(const void *x, const void *residual, const float *scale_dev,
                                 const float *shift_dev, void *y, int B, int C, int H, int W,
                                 const float *taps_up, int N_up, const float *taps_down, int N_down,
                                 int dtype, void *stream);

/* Statistics half of GroupNorm(1, C) (modules/ddpm_utils.py:113, 116): per sample mean / rstd over
 * (C, H, W), returned as the per-(sample, channel) affine consumed by afr_filtered_gelu_affine_fwd:
 * scale[b][c] = gamma[c] * rstd[b], shift[b][c] = beta[c] - mean[b] * scale[b][c]  (DEVICE fp32 [B*C]).
 * gamma_dev / beta_dev: DEVICE fp32 [C].  C*H*W must be a multiple of 4. */
int afr_groupnorm1_affine(const void *x, const float *gamma_dev, const float *beta_dev, float eps,
                          float *scale_dev, float *shift_dev, int B, int C, int H, int W, int dtype,
                          void *stream);

/* afr_groupnorm1_affine plus what training and the block epilogue need: `add_dev` (may be NULL) is a DEVICE fp32
 * [B*C] term folded into shift -- the time-embedding broadcast `x + emb[:, :, None, None]` that follows the last
 * GroupNorm of every Down / Up stage (modules/ddpm_utils.py:386, 416); mean_dev / rstd_dev (may be NULL) receive
 * the per-sample statistics the GroupNorm backward needs. */
int afr_groupnorm1_stats(const void *x, const float *gamma_dev, const float *beta_dev, float eps,
                         const float *add_dev, float *scale_dev, float *shift_dev, float *mean_dev,
                         float *rstd_dev, int B, int C, int H, int W, int dtype, void *stream);

/* custom_upsample / custom_downsample (and their adjoints) on channels-last memory, 3x3 filters, C % 4 == 0.
 * afr_up2x_nhwc:   x [B,H,W,C] -> u [B,2H,2W,C]  (custom_upsample; with adjoint_of_down = 1 and the DOWN filter: the
 *                  adjoint of custom_downsample for even H, W of its input, dy -> dv)
 * afr_down2x_nhwc: v [B,H,W,C] -> y [B,ceil(H/2),ceil(W/2),C]  (custom_downsample; with adjoint_of_up = 1 and the UP filter:
 *                  the adjoint of custom_upsample, du -> dx); H, W are the dims of v.
 * `*_pixel_stride` = elements between consecutive pixels: C for a dense channels-last tensor, the channel count of the
 * enclosing tensor when the operand is a channel slice -- custom_upsample writes straight into its half of the
 * torch.cat buffer (modules/ddpm_utils.py:414) and its adjoint reads the gradient slice in place. */
int afr_up2x_nhwc(const void *x, void *u, int B, int C, int H, int W, int64_t x_pixel_stride,
                  int64_t u_pixel_stride, const float *taps, int N, int adjoint_of_down, int in_dtype,
                  int out_dtype, void *stream);
int afr_down2x_nhwc(const void *v, void *y, int B, int C, int H, int W, int64_t v_pixel_stride,
                    int64_t y_pixel_stride, const float *taps, int N, int adjoint_of_up, int dtype,
                    void *stream);

/* Backward of GroupNorm(1, C) for the folded-norm training path: dz = gradient of the normalised + affine output
 * (what afr_filtered_gelu_affine_bwd / ..._nhwc_bwd return, or the incoming gradient of afr_affine_apply), x the
 * norm's input, mean_dev / rstd_dev from afr_groupnorm1_stats.  Writes dx (same layout as x), dgamma_dev and
 * dbeta_dev (DEVICE fp32 [C]).  workspace_dev: DEVICE fp32 scratch of 2*B*C + 2*B elements owned by the caller.
 * Three launches, deterministic (no atomics).  channels_last = 0: dense NCHW, H*W % 4 == 0; 1: channels-last memory,
 * C % 32 == 0.  Replaces ATen's native_group_norm_backward on this path (modules/ddpm_utils.py:113, 116 under autograd). */
int afr_groupnorm1_bwd(const void *x, const void *dz, const float *gamma_dev, const float *mean_dev,
                       const float *rstd_dev, void *dx, float *dgamma_dev, float *dbeta_dev,
                       float *workspace_dev, int B, int C, int H, int W, int dtype, int channels_last,
                       void *stream);

/* y = x * scale[b,c] + shift[b,c]: normalise + affine (+ the folded embedding) of a GroupNorm whose statistics came
 * from afr_groupnorm1_stats, as one pass (the reference runs GroupNorm, `.repeat` of the embedding and an add:
 * modules/ddpm_utils.py:385-387).  H*W must be a multiple of 4. */
int afr_affine_apply(const void *x, const float *scale_dev, const float *shift_dev, void *y,
                     int B, int C, int H, int W, int dtype, void *stream);

/* adjoint of afr_filtered_gelu_affine_fwd with respect to z = x * scale + shift (+ residual): recomputes z and u
 * from x, reads dy, writes dz.  The caller finishes with the GroupNorm backward (dz -> dx, dgamma, dbeta), which
 * needs reductions over the whole sample and stays outside this kernel.  Same restrictions as the forward. */
int afr_filtered_gelu_affine_bwd(const void *x, const void *residual, const float *scale_dev,
                                 const float *shift_dev, const void *dy, void *dz, int B, int C, int H, int W,
                                 const float *taps_up, int N_up, const float *taps_down, int N_down,
                                 int dtype, void *stream);

/* Channels-last (NHWC memory: x[b][h][w][c], what `tensor.contiguous(memory_format=torch.channels_last)` holds)
 * flavours of the fused filtered nonlinearity, so that a UNet run in channels-last -- where cuDNN needs no
 * nchw<->nhwc conversion kernels around its convolutions -- does not have to transpose for this op either.
 * Same function as afr_filtered_gelu_fwd / _bwd with the optional GroupNorm affine of ..._affine_fwd / _bwd
 * (scale_dev / shift_dev: DEVICE fp32 [B*C], or both NULL); B, C, H, W are the LOGICAL dims.  Requires 3x3
 * D4-symmetric filters (every circularLowpassKernel filter), C % 32 == 0, W % 4 == 0, H >= 2 and 16-byte aligned
 * buffers; AFR_ERR_UNSUPPORTED otherwise (callers then make the tensor NCHW-contiguous and use the entry points above). */
int afr_filtered_gelu_nhwc_fwd(const void *x, const void *residual, const float *scale_dev,
                               const float *shift_dev, void *y, int B, int C, int H, int W,
                               const float *taps_up, int N_up, const float *taps_down, int N_down,
                               int dtype, void *stream);
int afr_filtered_gelu_nhwc_bwd(const void *x, const void *residual, const float *scale_dev,
                               const float *shift_dev, const void *dy, void *dz, int B, int C, int H, int W,
                               const float *taps_up, int N_up, const float *taps_down, int N_down,
                               int dtype, void *stream);
/* afr_affine_apply for channels-last memory (C % 4 == 0). */
int afr_affine_apply_nhwc(const void *x, const float *scale_dev, const float *shift_dev, void *y,
                          int B, int C, int H, int W, int dtype, void *stream);

/* adjoint of the above wrt (x + residual): recomputes u from x, reads dy, writes dx
 * (d/dx and d/dresidual are the same tensor).  No saved 4x intermediates. */
int afr_filtered_gelu_bwd(const void *x, const void *residual, const void *dy, void *dx,
                          int B, int C, int H, int W,
                          const float *taps_up, int N_up, const float *taps_down, int N_down,
                          int dtype, void *stream);

/* Variant 4 (DoubleConv_F4, modules/ddpm_utils.py:145-197): GroupNorm sits on the 2x grid between
 * custom_upsample and the GELU, so only the second half fuses:
 *     y = custom_downsample(gelu(v * scale[b,c] + shift[b,c]), down)        v [B,C,H,W] = the 2x grid, y [B,C,H/2,W/2]
 * (lines :171-173 / :181-183; scale_dev / shift_dev as for afr_filtered_gelu_affine_fwd, or both NULL for
 * y = custom_downsample(gelu(v)) when the caller has already normalised v).  Neither norm(u) nor gelu(norm(u))
 * -- each 4x the block's activation -- is written.  N == 3, H even, W % 8 == 0 (AFR_ERR_UNSUPPORTED otherwise). */
int afr_gelu_down2x_fwd(const void *v, const float *scale_dev, const float *shift_dev, void *y,
                        int B, int C, int H, int W, const float *taps, int N, int dtype, void *stream);

/* adjoint of y = custom_downsample(gelu(v)) (the form without affine, which is what training uses):
 * dv = gelu'(v) * custom_downsample^T(dy).  v, dv [B,C,H,W]; dy [B,C,H/2,W/2]. */
int afr_gelu_down2x_bwd(const void *v, const void *dy, void *dv, int B, int C, int H, int W,
                        const float *taps, int N, int dtype, void *stream);

/* Diffusion.rotate_2d_matrix(matrix, degrees)          modules/ddpm_models.py:421-429
 * = scipy.ndimage.rotate(axes=(2,3), reshape=False, order=3, mode='grid-wrap'):
 * periodic cubic B-spline prefilter + 4x4-tap gather, entirely on device.
 * x, y: [B,C,H,W] fp32 (dtype must be AFR_F32), H*W <= 16384; x and y must not alias
 * (in-place is not supported). */
int afr_rotate_periodic_cubic(const void *x, void *y, int B, int C, int H, int W,
                              double degrees, int dtype, void *stream);

/* Fused posterior update of Algorithm 1 (modules/ddpm_models.py:374):
 * x <- ca * (x - cb * eps) + cc * noise, elementwise over n floats (fp32). noise may be NULL. */
int afr_ddpm_update(void *x, const void *eps, const void *noise, int64_t n,
                    float ca, float cb, float cc, void *stream);

/* Same update with the step's coefficients read ON THE DEVICE, so that a whole reverse step can
 * be captured once in a CUDA graph and replayed for every i:
 * table is a device array [T][3] of (ca, cb, cc) per timestep, step_dev a device int holding the
 * current timestep i (1 <= i < T). */
int afr_ddpm_update_table(void *x, const void *eps, const void *noise, int64_t n,
                          const float *table_dev, const int *step_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AFR_H_ */
