// Variant 4 (modules/ddpm_utils.py:145-197): the GroupNorm sits on the 2x grid BETWEEN the upsampler and the
// GELU, so the one-kernel up -> GELU -> down fusion does not apply.  What can be fused is the second half:
//
//   gelu_down3_kernel      y  = down2x( gelu(v * a + b) ; k_down )           v = custom_upsample(...) output, 2x grid
//                          (a, b) = optional per-(sample, channel) affine = the normalise + affine step of the
//                          GroupNorm (inference; statistics from afr_groupnorm1_affine), so neither norm(u)
//                          nor gelu(norm(u)) -- both 4x the block's activation -- is ever written
//   gelu_up3_bwd_kernel    dv = gelu'(v) * up-like(dy ; flip(k_down))         adjoint of the unaffine form, which
//                          is what training uses (GroupNorm itself stays with autograd)
//
// Same register-strip structure as down3_kernel / up3_kernel (afr_n3.cu): a thread owns V output columns and
// walks down the rows, the odd input row is carried to the next output row ALREADY ACTIVATED, so every input
// sample is activated once per thread (9/8 of the samples in total because of the one-column halo).
// Zero padding applies to gelu(.), not to v: samples outside the plane are 0 without passing the activation.
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

namespace {

// activate NV values in place: pairs on the packed-FMA chain, the odd one out on the scalar form
template <int NV>
__device__ __forceinline__ void gelu_all(float (&v)[NV])
{
#pragma unroll
    for (int c = 0; c + 1 < NV; c += 2) gelu_erf_x2(v[c], v[c + 1]);
    if (NV & 1) v[NV - 1] = gelu_erf(v[NV - 1]);
}

template <typename T, int V, bool kAff>
__device__ __forceinline__ void load_act_row(const T *plane, int row, int H, int W, int c0, bool has_l, float a, float b,
                                             float (&v)[2 * V + 1])
{
    if (row < 0 || row >= H) {
#pragma unroll
        for (int c = 0; c < 2 * V + 1; ++c) v[c] = 0.f;
        return;
    }
    const T *p = plane + (long)row * W + c0;
    float w[2 * V + 1];
#pragma unroll
    for (int h = 0; h < V / 4; ++h) {
        float t[8];
        ld8(p + 8 * h, t);
#pragma unroll
        for (int c = 0; c < 8; ++c) w[8 * h + c] = t[c];
    }
    w[2 * V] = has_l ? ld1(p - 1) : 0.f;                       // the halo column rides in the last slot
    if (kAff) {
#pragma unroll
        for (int c = 0; c < 2 * V + 1; ++c) w[c] = fmaf(w[c], a, b);
    }
    gelu_all<2 * V + 1>(w);
    v[0] = has_l ? w[2 * V] : 0.f;
#pragma unroll
    for (int c = 0; c < 2 * V; ++c) v[c + 1] = w[c];
}

template <typename T, int V, bool kAff>
__global__ void __launch_bounds__(256)
gelu_down3_kernel(const T *__restrict__ in, const float *__restrict__ scale, const float *__restrict__ shift,
                  T *__restrict__ out, long planes, int H, int W, int Ho, int Wo, int strips, int nseg, int R,
                  const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = V * s, i0 = seg * R, i1 = min(Ho, i0 + R);
    const T *plane = in + p * (long)H * W;
    T *dst = out + p * (long)Ho * Wo + j;
    const bool has_l = (j > 0);
    float a = 1.f, b = 0.f;
    if (kAff) { a = __ldg(scale + p); b = __ldg(shift + p); }

    float vp[2 * V + 1], ve[2 * V + 1], vo[2 * V + 1];
    load_act_row<T, V, kAff>(plane, 2 * i0 - 1, H, W, 2 * j, has_l, a, b, vp);
    for (int i = i0; i < i1; ++i) {
        if (sizeof(T) == 2 && 2 * i + 5 < H) {                  // bf16: same L2 prefetch as down3_kernel
            const T *pf = plane + (long)(2 * i + 4) * W + 2 * j;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + W));
        }
        load_act_row<T, V, kAff>(plane, 2 * i, H, W, 2 * j, has_l, a, b, ve);
        load_act_row<T, V, kAff>(plane, 2 * i + 1, H, W, 2 * j, has_l, a, b, vo);
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float acc = k.k[0][0] * vp[2 * q];
            acc = fmaf(k.k[0][1], vp[2 * q + 1], acc);
            acc = fmaf(k.k[0][2], vp[2 * q + 2], acc);
            acc = fmaf(k.k[1][0], ve[2 * q], acc);
            acc = fmaf(k.k[1][1], ve[2 * q + 1], acc);
            acc = fmaf(k.k[1][2], ve[2 * q + 2], acc);
            acc = fmaf(k.k[2][0], vo[2 * q], acc);
            acc = fmaf(k.k[2][1], vo[2 * q + 1], acc);
            acc = fmaf(k.k[2][2], vo[2 * q + 2], acc);
            o[q] = acc;
        }
        if (V == 8) {
            float o8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o8[q] = o[q % V];
            st8(dst + (long)i * Wo, o8);
        } else {
            st4(dst + (long)i * Wo, make_float4(o[0], o[1], o[2], o[3]));
        }
#pragma unroll
        for (int c = 0; c < 2 * V + 1; ++c) vp[c] = vo[c];
    }
}

// adjoint of y = down(gelu(v)):  dv = gelu'(v) * up-like(dy; kf), kf = flip(k_down) (pads are 1/1 for N == 3).
// thread = 4 dy columns -> 8 dv columns of rows 2i, 2i+1; dy rows i (carried) and i+1.
template <typename T>
__global__ void __launch_bounds__(256)
gelu_up3_bwd_kernel(const T *__restrict__ v, const T *__restrict__ dy, T *__restrict__ dv, long planes, int Hy, int Wy,
                    int strips, int nseg, int R, const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = 4 * s, i0 = seg * R, i1 = min(Hy, i0 + R);
    const T *src = dy + p * (long)Hy * Wy + j;
    const int W2 = 2 * Wy;
    const T *vin = v + p * 4L * Hy * Wy + 2 * j;
    T *dst = dv + p * 4L * Hy * Wy + 2 * j;
    const bool has_r = (j + 4 < Wy);

    float xa[5], xb[5];
    {
        float4 c = ld4(src + (long)i0 * Wy);
        xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
        xb[4] = has_r ? ld1(src + (long)i0 * Wy + 4) : 0.f;
    }
    for (int i = i0; i < i1; ++i) {
#pragma unroll
        for (int c = 0; c < 5; ++c) xa[c] = xb[c];
        float ve[8], vo[8];
        ld8(vin + (long)(2 * i) * W2, ve);
        ld8(vin + (long)(2 * i + 1) * W2, vo);
        if (i + 1 < Hy) {
            float4 c = ld4(src + (long)(i + 1) * Wy);
            xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
            xb[4] = has_r ? ld1(src + (long)(i + 1) * Wy + 4) : 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < 5; ++c) xb[c] = 0.f;
        }
        float e[8], o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            e[2 * c] = k.k[1][1] * xa[c];
            e[2 * c + 1] = fmaf(k.k[1][2], xa[c + 1], k.k[1][0] * xa[c]);
            o[2 * c] = fmaf(k.k[2][1], xb[c], k.k[0][1] * xa[c]);
            float t = k.k[0][0] * xa[c];
            t = fmaf(k.k[0][2], xa[c + 1], t);
            t = fmaf(k.k[2][0], xb[c], t);
            o[2 * c + 1] = fmaf(k.k[2][2], xb[c + 1], t);
        }
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            gelu_grad_scaled_mul_x2(AFR_KAPPA * ve[c], AFR_KAPPA * ve[c + 1], e[c], e[c + 1]);
            gelu_grad_scaled_mul_x2(AFR_KAPPA * vo[c], AFR_KAPPA * vo[c + 1], o[c], o[c + 1]);
        }
        T *r0 = dst + (long)(2 * i) * W2;
        st8(r0, e);
        st8(r0 + W2, o);
    }
}

inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline size_t esize(int dtype) { return dtype == AFR_F32 ? 4 : 2; }
inline int pick_rows(int H) { return H < 8 ? H : 8; }

}  // namespace

// v [planes, H, W] (the 2x grid) -> y [planes, H/2, W/2]
bool actdown_supported(int H, int W, const void *v, const void *y, int dtype)
{
    return H >= 2 && (H % 2) == 0 && W >= 8 && (W % 8) == 0 && aligned_to(v, 8 * esize(dtype)) &&
           aligned_to(y, 4 * esize(dtype));
}

cudaError_t actdown_fwd(const void *v, const float *scale, const float *shift, void *y, long planes, int H, int W,
                        const Taps3 &k, int dtype, cudaStream_t s)
{
    const int Ho = H / 2, Wo = W / 2;
    const bool wide = dtype == AFR_BF16 && (Wo % 8) == 0 && aligned_to(v, 16) && aligned_to(y, 16);
    const int V = wide ? 8 : 4;
    const int strips = Wo / V, R = pick_rows(Ho), nseg = (Ho + R - 1) / R;
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
#define AFR_AD(T, VV, AFF)                                                                                         \
    gelu_down3_kernel<T, VV, AFF><<<(unsigned)grid, 256, 0, s>>>((const T *)v, scale, shift, (T *)y, planes, H, W, Ho, Wo, \
                                                                 strips, nseg, R, k)
    if (dtype == AFR_F32) { if (scale) AFR_AD(float, 4, true); else AFR_AD(float, 4, false); }
    else if (wide) { if (scale) AFR_AD(bf16, 8, true); else AFR_AD(bf16, 8, false); }
    else { if (scale) AFR_AD(bf16, 4, true); else AFR_AD(bf16, 4, false); }
#undef AFR_AD
    return cudaGetLastError();
}

cudaError_t actdown_bwd(const void *v, const void *dy, void *dv, long planes, int H, int W, const Taps3 &kflip, int dtype,
                        cudaStream_t s)
{
    const int Hy = H / 2, Wy = W / 2;
    const int strips = Wy / 4, R = pick_rows(Hy), nseg = (Hy + R - 1) / R;
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (dtype == AFR_F32)
        gelu_up3_bwd_kernel<float><<<(unsigned)grid, 256, 0, s>>>((const float *)v, (const float *)dy, (float *)dv, planes, Hy, Wy,
                                                                  strips, nseg, R, kflip);
    else
        gelu_up3_bwd_kernel<bf16><<<(unsigned)grid, 256, 0, s>>>((const bf16 *)v, (const bf16 *)dy, (bf16 *)dv, planes, Hy, Wy,
                                                                 strips, nseg, R, kflip);
    return cudaGetLastError();
}

}  // namespace afr
