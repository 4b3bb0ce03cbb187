// Config-E rotation step on device (replaces the D2H -> scipy.ndimage.rotate -> H2D round
// trip of modules/ddpm_models.py:421-429, executed 999 times per sample), plus the fused
// Algorithm-1 posterior update of modules/ddpm_models.py:374.
//
// scipy.ndimage.rotate(order=3, mode='grid-wrap', prefilter=True) on one H x W plane:
//   1. cubic B-spline prefilter along both axes with PERIODIC boundary (pole z = sqrt(3)-2,
//      exact circulant initial values), in double;
//   2. out[o] = sum_{a,b<4} w_a(fr) w_b(fc) c[(floor(r)-1+a) mod H][(floor(c)-1+b) mod W],
//      (r, c) = R o + (ctr - R ctr),  R = [[cos, sin], [-sin, cos]];
//   3. round to float32.
// One CTA per plane; the plane (<= 16384 px) lives in shared memory as double.  The planes
// of the sampler are 32x32 x (n*3): this kernel is latency-, not bandwidth-relevant.
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

__device__ __forceinline__ void prefilter_line(double *c, int n, int stride)
{
    const double z = -0.26794919243112270647;   // sqrt(3) - 2
    if (n < 2) return;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    for (int i = 0; i < n; ++i) c[i * stride] *= gain;
    double zi = z, s = c[0];
    for (int i = 1; i < n; ++i) { s += zi * c[(n - i) * stride]; zi *= z; }
    c[0] = s / (1.0 - zi);
    for (int i = 1; i < n; ++i) c[i * stride] += z * c[(i - 1) * stride];
    zi = z; s = c[(n - 1) * stride];
    for (int i = 0; i < n - 1; ++i) { s += zi * c[i * stride]; zi *= z; }
    c[(n - 1) * stride] = s * z / (zi - 1.0);
    for (int i = n - 2; i >= 0; --i) c[i * stride] = z * (c[(i + 1) * stride] - c[i * stride]);
}

__device__ __forceinline__ void bspline3(double t, double (&w)[4])
{
    const double t1 = 1.0 - t;
    w[0] = t1 * t1 * t1 * (1.0 / 6.0);
    w[1] = 2.0 / 3.0 - 0.5 * t * t * (2.0 - t);
    w[2] = 2.0 / 3.0 - 0.5 * t1 * t1 * (2.0 - t1);
    w[3] = t * t * t * (1.0 / 6.0);
}

__device__ __forceinline__ int wrap(int i, int n)
{
    int m = i % n;
    return m < 0 ? m + n : m;
}

__global__ void __launch_bounds__(256)
rotate_kernel(const float *__restrict__ x, float *__restrict__ y, int H, int W, double cs,
              double sn, double off_r, double off_c)
{
    extern __shared__ double cbuf[];
    const int pitch = W | 1;                     // odd pitch: row-walkers hit distinct banks
    const long p = blockIdx.x;
    const float *src = x + p * (long)H * W;
    for (int e = threadIdx.x; e < H * W; e += blockDim.x)
        cbuf[(e / W) * pitch + (e % W)] = (double)__ldg(src + e);
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) prefilter_line(cbuf + j, H, pitch);
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) prefilter_line(cbuf + i * pitch, W, 1);
    __syncthreads();
    float *dst = y + p * (long)H * W;
    for (int e = threadIdx.x; e < H * W; e += blockDim.x) {
        const int i = e / W, j = e % W;
        const double rr = cs * i + sn * j + off_r, cc = -sn * i + cs * j + off_c;
        const double fr = floor(rr), fc = floor(cc);
        double wr[4], wc[4];
        bspline3(rr - fr, wr);
        bspline3(cc - fc, wc);
        int ci[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) ci[b] = wrap((int)fc - 1 + b, W);
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double *row = cbuf + wrap((int)fr - 1 + a, H) * pitch;
            double r = wc[0] * row[ci[0]];
            r = fma(wc[1], row[ci[1]], r);
            r = fma(wc[2], row[ci[2]], r);
            r = fma(wc[3], row[ci[3]], r);
            acc = fma(wr[a], r, acc);
        }
        dst[e] = (float)acc;
    }
}

cudaError_t rotate_periodic_cubic(const float *x, float *y, long planes, int H, int W,
                                  double degrees, cudaStream_t s)
{
    const double th = degrees * 3.14159265358979323846 / 180.0;
    const double cs = cos(th), sn = sin(th);
    const double ch = 0.5 * (H - 1), cw = 0.5 * (W - 1);
    const double off_r = ch - (cs * ch + sn * cw), off_c = cw - (-sn * ch + cs * cw);
    const size_t smem = sizeof(double) * (size_t)H * (W | 1);
    static std::atomic<unsigned long long> attr_done{0};
    if (cudaError_t e = ensure_dyn_smem(rotate_kernel, attr_done, 200 * 1024)) return e;
    if (planes > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const int threads = H * W >= 4096 ? 256 : 128;
    rotate_kernel<<<(unsigned)planes, threads, smem, s>>>(x, y, H, W, cs, sn, off_r, off_c);
    return cudaGetLastError();
}

// GroupNorm(1, C) statistics -> per-(sample, channel) affine for the fused activation kernel:
//   scale[b][c] = gamma[c] * rstd[b],  shift[b][c] = beta[c] - mean[b] * scale[b][c].
// One CTA per sample; two passes over the sample (mean, then centred second moment -- the second
// pass hits L2), per-thread fp32 partials, warp-shuffle + shared-memory reduction in double.
template <typename T>
__device__ __forceinline__ float4 ldv4(const T *p);
template <>
__device__ __forceinline__ float4 ldv4<float>(const float *p) { return ld4(p); }
template <>
__device__ __forceinline__ float4 ldv4<bf16>(const bf16 *p) { return ld4(p); }

__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                       // red may still be read from a previous call
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

template <typename T>
__global__ void __launch_bounds__(1024)
groupnorm1_affine_kernel(const T *__restrict__ x, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, float *__restrict__ scale,
                         float *__restrict__ shift, int C, long n_per_sample, const float *__restrict__ add,
                         float *__restrict__ mean_out, float *__restrict__ rstd_out)
{
    __shared__ double red[32];
    const long b = blockIdx.x;
    const T *src = x + b * n_per_sample;
    const long n4 = n_per_sample / 4;      // launcher guarantees n_per_sample % 4 == 0
    float s = 0.f;
    for (long i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = ldv4<T>(src + 4 * i);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const double mean = block_sum((double)s, red) / (double)n_per_sample;
    const float mf = (float)mean;
    float q = 0.f;
    for (long i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = ldv4<T>(src + 4 * i);
        const float a = v.x - mf, c = v.y - mf, d = v.z - mf, e = v.w - mf;
        q += (a * a + c * c) + (d * d + e * e);
    }
    // sum (x - mf)^2 = sum (x - mean)^2 + n (mean - mf)^2 ; the correction is ~1e-16, dropped
    const double var = block_sum((double)q, red) / (double)n_per_sample;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = __ldg(gamma + c) * rstd;
        scale[b * C + c] = sc;
        shift[b * C + c] = __ldg(beta + c) - mf * sc + (add ? __ldg(add + b * C + c) : 0.f);
    }
    if (threadIdx.x == 0) {                       // for the GroupNorm backward (training)
        if (mean_out) mean_out[b] = mf;
        if (rstd_out) rstd_out[b] = rstd;
    }
}

// y = x * scale[b, c] + shift[b, c]: the normalise + affine (+ broadcast embedding, folded into shift) epilogue
// of a GroupNorm whose statistics came from the kernel above.  One plane per blockIdx.y row, 128-bit accesses.
template <typename T>
__global__ void __launch_bounds__(256)
affine_apply_kernel(const T *__restrict__ x, const float *__restrict__ scale, const float *__restrict__ shift,
                    T *__restrict__ y, long planes, int hw4)
{
    const long total = planes * hw4;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const long p = i / hw4;
        const float a = __ldg(scale + p), b = __ldg(shift + p);
        float4 v = ldv4<T>(x + 4 * i);
        v.x = fmaf(v.x, a, b); v.y = fmaf(v.y, a, b); v.z = fmaf(v.z, a, b); v.w = fmaf(v.w, a, b);
        st4(y + 4 * i, v);
    }
}

// channels-last flavour: element i of sample b belongs to channel i % C; four consecutive elements are four channels
template <typename T>
__global__ void __launch_bounds__(256)
affine_apply_nhwc_kernel(const T *__restrict__ x, const float *__restrict__ scale, const float *__restrict__ shift,
                         T *__restrict__ y, long total4, int C4, long sample4)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += stride) {
        const long b = i / sample4;
        const int c4 = (int)((i - b * sample4) % C4);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(scale) + b * C4 + c4);
        const float4 s = __ldg(reinterpret_cast<const float4 *>(shift) + b * C4 + c4);
        float4 v = ldv4<T>(x + 4 * i);
        v.x = fmaf(v.x, a.x, s.x); v.y = fmaf(v.y, a.y, s.y); v.z = fmaf(v.z, a.z, s.z); v.w = fmaf(v.w, a.w, s.w);
        st4(y + 4 * i, v);
    }
}

cudaError_t affine_apply_nhwc(const void *x, const float *scale, const float *shift, void *y, long B, int C, long hw, int dtype,
                              cudaStream_t s)
{
    const long sample4 = hw * C / 4, total4 = B * sample4;
    long grid = (total4 + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 148 * 32) grid = 148 * 32;
    if (dtype == AFR_F32)
        affine_apply_nhwc_kernel<float><<<(unsigned)grid, 256, 0, s>>>((const float *)x, scale, shift, (float *)y, total4, C / 4, sample4);
    else
        affine_apply_nhwc_kernel<bf16><<<(unsigned)grid, 256, 0, s>>>((const bf16 *)x, scale, shift, (bf16 *)y, total4, C / 4, sample4);
    return cudaGetLastError();
}

cudaError_t affine_apply(const void *x, const float *scale, const float *shift, void *y, long planes, long hw, int dtype,
                         cudaStream_t s)
{
    const int hw4 = (int)(hw / 4);
    const long total = planes * hw4;
    long grid = (total + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 148 * 32) grid = 148 * 32;
    if (dtype == AFR_F32)
        affine_apply_kernel<float><<<(unsigned)grid, 256, 0, s>>>((const float *)x, scale, shift, (float *)y, planes, hw4);
    else
        affine_apply_kernel<bf16><<<(unsigned)grid, 256, 0, s>>>((const bf16 *)x, scale, shift, (bf16 *)y, planes, hw4);
    return cudaGetLastError();
}

cudaError_t groupnorm1_affine(const void *x, const float *gamma, const float *beta, float eps, float *scale,
                              float *shift, long B, int C, long hw, int dtype, cudaStream_t s, const float *add,
                              float *mean_out, float *rstd_out)
{
    if (B > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const long n = (long)C * hw;
    // few samples: one big CTA per sample (latency); many samples: 256-thread CTAs (occupancy)
    const int threads = (B < 2 * 148 && n >= 8192) ? 1024 : 256;
    if (dtype == AFR_F32)
        groupnorm1_affine_kernel<float><<<(unsigned)B, threads, 0, s>>>((const float *)x, gamma, beta, eps, scale, shift, C, n,
                                                                        add, mean_out, rstd_out);
    else
        groupnorm1_affine_kernel<bf16><<<(unsigned)B, threads, 0, s>>>((const bf16 *)x, gamma, beta, eps, scale, shift, C, n,
                                                                       add, mean_out, rstd_out);
    return cudaGetLastError();
}

// x <- ca * (x - cb * eps) + cc * noise      (modules/ddpm_models.py:374)
__global__ void __launch_bounds__(256)
ddpm_update_kernel(float *__restrict__ x, const float *__restrict__ eps,
                   const float *__restrict__ noise, long n4, long n, float ca, float cb, float cc,
                   const float *__restrict__ table, const int *__restrict__ step)
{
    if (table) {                                   // graph-replay flavour: coefficients live on the device
        const int i = *step;
        ca = table[3 * i]; cb = table[3 * i + 1]; cc = table[3 * i + 2];
    }
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = reinterpret_cast<float4 *>(x)[i];
        const float4 e = __ldg(reinterpret_cast<const float4 *>(eps) + i);
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (noise) z = __ldg(reinterpret_cast<const float4 *>(noise) + i);
        a.x = fmaf(cc, z.x, ca * fmaf(-cb, e.x, a.x));
        a.y = fmaf(cc, z.y, ca * fmaf(-cb, e.y, a.y));
        a.z = fmaf(cc, z.z, ca * fmaf(-cb, e.z, a.z));
        a.w = fmaf(cc, z.w, ca * fmaf(-cb, e.w, a.w));
        reinterpret_cast<float4 *>(x)[i] = a;
    }
    // tail (n not a multiple of 4)
    for (long i = 4 * n4 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float z = noise ? noise[i] : 0.f;
        x[i] = fmaf(cc, z, ca * fmaf(-cb, eps[i], x[i]));
    }
}

cudaError_t ddpm_update(float *x, const float *eps, const float *noise, long n, float ca, float cb,
                        float cc, const float *table_dev, const int *step_dev, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps) |
                       reinterpret_cast<uintptr_t>(noise)) % 16) == 0;
    const long n4 = vec ? n / 4 : 0;
    long grid = ((vec ? n4 : n) + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 148 * 16) grid = 148 * 16;
    ddpm_update_kernel<<<(unsigned)grid, 256, 0, s>>>(x, eps, noise, n4, n, ca, cb, cc, table_dev, step_dev);
    return cudaGetLastError();
}

}  // namespace afr
