// GroupNorm(1, C) backward for the folded-norm training path (modules/ddpm_utils.py:113, 116 under autograd).
//
// The forward never materialises the normalised tensor (afr_groupnorm1_stats + the affine folded into the
// activation kernel's load); the activation's adjoint kernel returns dz = dLoss/d(norm output).  What is left
// is the normalisation's own backward, for x [B, C, HW] (or channels-last [B, HW, C]), one group per sample:
//
//     xh      = (x - mean_b) * rstd_b
//     dgamma_c = sum_{b,hw} dz * xh            dbeta_c = sum_{b,hw} dz
//     S1_b     = sum_{c,hw} gamma_c * dz       S2_b    = sum_{c,hw} gamma_c * dz * xh
//     dx       = rstd_b * (gamma_c * dz - (S1_b + xh * S2_b) / (C * HW))
//
// Three launches, 5 n elements of traffic, no atomics (deterministic):
//   1. gn_bwd_partials_kernel   p1[b,c] = sum_hw dz, p2[b,c] = sum_hw dz * xh      (reads x and dz once)
//   2. gn_bwd_reduce_kernel     dgamma / dbeta over b, S1 / S2 over c              (B*C numbers: tiny)
//   3. gn_bwd_apply_kernel      dx                                                  (reads x and dz, writes dx)
// ATen's native_group_norm_backward takes 4-5 launches and its per-channel reduction
// (GammaBetaBackwardCUDAKernelTemplate) alone is 5.5 % of a Config-D training step
// (profiles/r01_ncu_launches_train_step.md); it also reads its operands as dense NCHW only, so a channels-last
// model pays two layout copies per norm.  Both layouts are handled here.
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

namespace {

template <typename T> __device__ __forceinline__ float4 ldq(const T *p);
template <> __device__ __forceinline__ float4 ldq<float>(const float *p) { return ld4(p); }
template <> __device__ __forceinline__ float4 ldq<bf16>(const bf16 *p) { return ld4(p); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NCHW: one CTA (128 threads) per (b, c) plane
template <typename T>
__global__ void __launch_bounds__(128)
gn_bwd_partials_nchw_kernel(const T *__restrict__ x, const T *__restrict__ dz, const float *__restrict__ mean,
                            const float *__restrict__ rstd, float *__restrict__ p1, float *__restrict__ p2, int C, int hw4)
{
    __shared__ float red[2][4];
    const long plane = blockIdx.x;
    const int b = (int)(plane / C);
    const float mu = __ldg(mean + b), rs = __ldg(rstd + b);
    const T *xp = x + plane * 4L * hw4, *dp = dz + plane * 4L * hw4;
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < hw4; i += blockDim.x) {
        const float4 xv = ldq<T>(xp + 4 * i), dv = ldq<T>(dp + 4 * i);
        s1 += (dv.x + dv.y) + (dv.z + dv.w);
        s2 += dv.x * (xv.x - mu) + dv.y * (xv.y - mu) + dv.z * (xv.z - mu) + dv.w * (xv.w - mu);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        float a = 0.f, c = 0.f;
        for (int w = 0; w < nw; ++w) { a += red[0][w]; c += red[1][w]; }
        p1[plane] = a;
        p2[plane] = c * rs;
    }
}

// NHWC: one CTA (256 threads = 8 warps) per (b, 32-channel group); lane = channel, warps stride over the pixels
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_partials_nhwc_kernel(const T *__restrict__ x, const T *__restrict__ dz, const float *__restrict__ mean,
                            const float *__restrict__ rstd, float *__restrict__ p1, float *__restrict__ p2, int C, long hw)
{
    __shared__ float red[2][8][32];
    const int cgroups = C / 32;
    const int b = (int)(blockIdx.x / cgroups), cg = (int)(blockIdx.x % cgroups);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = cg * 32 + lane;
    const float mu = __ldg(mean + b), rs = __ldg(rstd + b);
    const T *xp = x + (long)b * hw * C + c, *dp = dz + (long)b * hw * C + c;
    float s1 = 0.f, s2 = 0.f;
    for (long i = warp; i < hw; i += 8) {
        const float xv = ld1(xp + i * C), dv = ld1(dp + i * C);
        s1 += dv;
        s2 = fmaf(dv, xv - mu, s2);
    }
    red[0][warp][lane] = s1; red[1][warp][lane] = s2;
    __syncthreads();
    if (warp == 0) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[0][w][lane]; q += red[1][w][lane]; }
        p1[(long)b * C + c] = a;
        p2[(long)b * C + c] = q * rs;
    }
}

// blocks [0, C): dgamma_c, dbeta_c (sum over b);  blocks [C, C + B): S1_b, S2_b (sum over c, weighted by gamma)
__global__ void __launch_bounds__(128)
gn_bwd_reduce_kernel(const float *__restrict__ p1, const float *__restrict__ p2, const float *__restrict__ gamma,
                     float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ s1, float *__restrict__ s2,
                     int B, int C)
{
    __shared__ double red[2][4];
    double a = 0.0, q = 0.0;
    const bool per_channel = (int)blockIdx.x < C;
    if (per_channel) {
        const int c = blockIdx.x;
        for (int b = threadIdx.x; b < B; b += blockDim.x) { a += p1[(long)b * C + c]; q += p2[(long)b * C + c]; }
    } else {
        const int b = blockIdx.x - C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double g = gamma[c];
            a += g * p1[(long)b * C + c]; q += g * p2[(long)b * C + c];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tq = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += red[0][w]; tq += red[1][w]; }
        if (per_channel) { dbeta[blockIdx.x] = (float)ta; dgamma[blockIdx.x] = (float)tq; }
        else { s1[blockIdx.x - C] = (float)ta; s2[blockIdx.x - C] = (float)tq; }
    }
}

// dx = rstd_b * (gamma_c * dz - (S1_b + xh * S2_b) / n)
template <typename T, bool kNHWC>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const T *__restrict__ x, const T *__restrict__ dz, const float *__restrict__ gamma,
                    const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ s1,
                    const float *__restrict__ s2, T *__restrict__ dx, long total4, int C, long hw, float inv_n)
{
    const long sample4 = hw * C / 4;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += stride) {
        const long b = i / sample4;
        const long r = i - b * sample4;
        const float mu = __ldg(mean + b), rs = __ldg(rstd + b);
        const float k1 = __ldg(s1 + b) * inv_n, k2 = __ldg(s2 + b) * inv_n * rs;    // xh * S2 / n = (x - mu) * rs * S2 / n
        float4 g;
        if (kNHWC) {
            g = __ldg(reinterpret_cast<const float4 *>(gamma) + (int)(r % (C / 4)));
        } else {
            const float gc = __ldg(gamma + (int)(r / (hw / 4)));
            g = make_float4(gc, gc, gc, gc);
        }
        const float4 xv = ldq<T>(x + 4 * i), dv = ldq<T>(dz + 4 * i);
        float4 o;
        o.x = rs * (g.x * dv.x - k1 - (xv.x - mu) * k2);
        o.y = rs * (g.y * dv.y - k1 - (xv.y - mu) * k2);
        o.z = rs * (g.z * dv.z - k1 - (xv.z - mu) * k2);
        o.w = rs * (g.w * dv.w - k1 - (xv.w - mu) * k2);
        st4(dx + 4 * i, o);
    }
}

}  // namespace

// workspace: 2 * B * C + 2 * B floats
cudaError_t groupnorm1_bwd(const void *x, const void *dz, const float *gamma, const float *mean, const float *rstd, void *dx,
                           float *dgamma, float *dbeta, float *workspace, long B, int C, long hw, int dtype, bool nhwc,
                           cudaStream_t s)
{
    float *p1 = workspace, *p2 = workspace + B * C, *s1 = workspace + 2 * B * C, *s2 = s1 + B;
    if (B * C > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const long total4 = B * C * hw / 4;
    long grid = (total4 + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    if (grid < 1) grid = 1;
    const float inv_n = (float)(1.0 / ((double)C * (double)hw));
#define AFR_GNB(T)                                                                                                       \
    do {                                                                                                                 \
        if (nhwc) gn_bwd_partials_nhwc_kernel<T><<<(unsigned)(B * (C / 32)), 256, 0, s>>>((const T *)x, (const T *)dz, mean, rstd, p1, p2, C, hw); \
        else gn_bwd_partials_nchw_kernel<T><<<(unsigned)(B * C), 128, 0, s>>>((const T *)x, (const T *)dz, mean, rstd, p1, p2, C, (int)(hw / 4)); \
        gn_bwd_reduce_kernel<<<(unsigned)(C + B), 128, 0, s>>>(p1, p2, gamma, dgamma, dbeta, s1, s2, (int)B, C);         \
        if (nhwc) gn_bwd_apply_kernel<T, true><<<(unsigned)grid, 256, 0, s>>>((const T *)x, (const T *)dz, gamma, mean, rstd, s1, s2, (T *)dx, total4, C, hw, inv_n); \
        else gn_bwd_apply_kernel<T, false><<<(unsigned)grid, 256, 0, s>>>((const T *)x, (const T *)dz, gamma, mean, rstd, s1, s2, (T *)dx, total4, C, hw, inv_n); \
    } while (0)
    if (dtype == AFR_F32) AFR_GNB(float); else AFR_GNB(bf16);
#undef AFR_GNB
    return cudaGetLastError();
}

}  // namespace afr
