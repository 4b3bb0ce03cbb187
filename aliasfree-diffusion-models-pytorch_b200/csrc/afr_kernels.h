// Host-side launchers shared between the translation units of libafr_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "afr_common.cuh"

namespace afr {

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (kernel, device) pair: applied once per
// device the calling thread launches on (`done` is a per-kernel bit mask indexed by device ordinal;
// a lost race just sets the attribute twice).
template <class F>
inline cudaError_t ensure_dyn_smem(F kern, std::atomic<unsigned long long> &done, int bytes)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// Extra context for afr_last_error(), set by launchers when a step fails (thread local).
void set_detail(const char *fmt, ...);

// afr_generic.cu -- runtime-N kernels
cudaError_t generic_up_like(const void *in, void *out, long planes, int Hin, int Win, int Hout,
                            int Wout, const TapsG &t, int in_dtype, int out_dtype, cudaStream_t s);
cudaError_t generic_down_like(const void *in, void *out, long planes, int Hin, int Win, int Hout,
                              int Wout, const TapsG &t, int dtype, cudaStream_t s);
cudaError_t generic_fgelu(const void *x, const void *res, const void *dy, void *out, long planes,
                          int H, int W, const TapsG &tU, const TapsG &tG, const TapsG &tB,
                          bool bwd, int dtype, cudaStream_t s);

// afr_stripn.cu -- register-strip fused kernels for N in {2,4,5,6,7,8} (both filters the same size)
bool stripn_supported(int N, int H, int W, const void *x, const void *res, const void *dy, const void *out,
                      int dtype);
cudaError_t stripn_fgelu(const void *x, const void *res, const void *dy, void *out, long planes, int H, int W,
                         const TapsG &tU, const TapsG &tG, const TapsG &tB, bool bwd, int dtype, cudaStream_t s);

// standalone resamplers for the same N (shape / alignment conditions: n3_up_supported / n3_down_supported)
bool stripn_resample_supported(int N);
cudaError_t stripn_up_like(const void *in, void *out, long planes, int H, int W, const TapsG &t, int in_dtype,
                           int out_dtype, cudaStream_t s);
cudaError_t stripn_down_like(const void *in, void *out, long planes, int H, int W, const TapsG &t, int dtype,
                             cudaStream_t s);

// afr_n3.cu -- N == 3 register-strip kernels (direct and TMA-staged)
// All take stage taps already arranged for the stencil they run (see afr_api.cu).
bool n3_fgelu_supported(int H, int W, const void *const *ptrs, int nptrs, int dtype);
bool n3_prefers_plane_kernel(int H, int W, bool bwd);
bool n3_fgelu_tma_supported(long planes, int H, int W, const void *const *ptrs, int nptrs, int dtype,
                            int n_inputs);
cudaError_t n3_fgelu(const void *x, const void *res, const void *dy, const float *scale, const float *shift,
                     void *out, long planes, int H, int W, const Taps3 &kU, const Taps3 &kG, const Taps3 &kB,
                     bool bwd, int dtype, bool use_tma, bool allow_sym, cudaStream_t s, const char **kernel_name);
// channels-last ([B,H,W,C] memory, C % 32 == 0) fused forward / adjoint, symmetric taps only: cudaErrorNotSupported
// tells the caller to use the NCHW kernels (general taps)
bool nhwc_fgelu_supported(int C, int H, int W, const void *const *ptrs, int nptrs);
cudaError_t nhwc_fgelu(const void *x, const void *res, const void *dy, const float *scale, const float *shift, void *out,
                       long B, int C, int H, int W, const Taps3 &kU, const Taps3 &kG, const Taps3 &kB, bool bwd, int dtype,
                       cudaStream_t s);
// up-like with N==3: in [planes,H,W] -> out [planes,2H,2W]
bool n3_up_supported(int H, int W, const void *in, const void *out, int in_dtype, int out_dtype);
// (C > 0: the OUTPUT is a channel slice with batch stride out_bstride elements; C == 0: dense)
cudaError_t n3_up_like(const void *in, void *out, long planes, int H, int W, const Taps3 &k,
                       int in_dtype, int out_dtype, cudaStream_t s, int C = 0, long out_bstride = 0);
// down-like with N==3: in [planes,H,W] (H, W even, W % 8 == 0) -> out [planes,H/2,W/2]
bool n3_down_supported(int H, int W, const void *in, const void *out, int dtype);
// (C > 0: the INPUT is a channel slice with batch stride in_bstride elements; C == 0: dense)
cudaError_t n3_down_like(const void *in, void *out, long planes, int H, int W, const Taps3 &k,
                         int dtype, cudaStream_t s, int C = 0, long in_bstride = 0);

// afr_small.cu -- N == 3 resamplers for small planes (a CTA stages a group of whole planes); plane p = (b, c)
// lives at  base + b * bstride + c * plane_size  on the strided side (dense: C = 1, bstride = plane_size)
// shapes served by the warp-shuffle whole-plane kernels (preferred over the strip kernels of afr_n3.cu); the
// other small shapes use the shared-memory plane-group kernels only where the strip kernels cannot run
bool warp_up_shape(int H, int W);
bool warp_down_shape(int H, int W);
bool small_up_supported(int H, int W, const void *in, const void *out, long out_bstride, int out_dtype);
cudaError_t small_up_like(const void *in, void *out, long planes, int C, long out_bstride, int H, int W,
                          const Taps3 &k, int in_dtype, int out_dtype, cudaStream_t s);
bool small_down_supported(int H, int W, const void *in, const void *out, long in_bstride, int dtype);
cudaError_t small_down_like(const void *in, void *out, long planes, int C, long in_bstride, int H, int W,
                            const Taps3 &k, int dtype, cudaStream_t s);

// down-like without a row loop (latency-bound cases of the strip kernel; same support conditions as n3_down_supported)
bool flat_down_wanted(int H, int W, int dtype);
cudaError_t flat_down_like(const void *in, void *out, long planes, int C, long in_bstride, int H, int W, const Taps3 &k,
                           int dtype, cudaStream_t s);

bool flat_up_wanted(int H, int W, int in_dtype, int out_dtype);
cudaError_t flat_up_like(const void *in, void *out, long planes, int C, long out_bstride, int H, int W, const Taps3 &k,
                         int in_dtype, int out_dtype, cudaStream_t s);

// afr_nhwc_resample.cu -- N == 3 resamplers on channels-last memory; *ps = pixel stride in elements (C for a dense
// tensor, the wider tensor's channel count for a channel slice)
bool nhwc_resample_supported(int C, const void *a, const void *b, long aps, long bps, int adtype, int bdtype);
cudaError_t nhwc_up_like(const void *x, void *u, long B, int C, int H, int W, long xps, long ups, const Taps3 &k, int in_dtype,
                         int out_dtype, cudaStream_t s);
cudaError_t nhwc_down_like(const void *v, void *y, long B, int C, int H, int W, long vps, long yps, const Taps3 &k, int dtype,
                           cudaStream_t s);

// afr_actdown.cu -- variant 4: gelu (+ GroupNorm affine) fused into the N == 3 downsampler, and its adjoint
bool actdown_supported(int H, int W, const void *v, const void *y, int dtype);
cudaError_t actdown_fwd(const void *v, const float *scale, const float *shift, void *y, long planes, int H, int W,
                        const Taps3 &k, int dtype, cudaStream_t s);
cudaError_t actdown_bwd(const void *v, const void *dy, void *dv, long planes, int H, int W, const Taps3 &kflip, int dtype,
                        cudaStream_t s);

// afr_rotate.cu
cudaError_t rotate_periodic_cubic(const float *x, float *y, long planes, int H, int W,
                                  double degrees, cudaStream_t s);
cudaError_t groupnorm1_affine(const void *x, const float *gamma, const float *beta, float eps, float *scale,
                              float *shift, long B, int C, long hw, int dtype, cudaStream_t s, const float *add = nullptr,
                              float *mean_out = nullptr, float *rstd_out = nullptr);
// afr_norm.cu -- GroupNorm(1, C) backward in three launches (workspace: 2*B*C + 2*B floats), NCHW or channels-last
cudaError_t groupnorm1_bwd(const void *x, const void *dz, const float *gamma, const float *mean, const float *rstd, void *dx,
                           float *dgamma, float *dbeta, float *workspace, long B, int C, long hw, int dtype, bool nhwc,
                           cudaStream_t s);
cudaError_t affine_apply(const void *x, const float *scale, const float *shift, void *y, long planes, long hw, int dtype,
                         cudaStream_t s);
cudaError_t affine_apply_nhwc(const void *x, const float *scale, const float *shift, void *y, long B, int C, long hw, int dtype,
                              cudaStream_t s);
cudaError_t ddpm_update(float *x, const float *eps, const float *noise, long n, float ca,
                        float cb, float cc, const float *table_dev, const int *step_dev, cudaStream_t s);

}  // namespace afr
