// Shared device helpers for the alias-free resampling kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/afr.h"

namespace afr {

typedef __nv_bfloat16 bf16;

// N x N taps of one stage, carried in kernel parameter space (no device-side filter
// tensor, no per-call H2D copy -- the reference uploads and repeats its filter on
// every call, modules/filtrs.py:73-74, 91-92).  `pad` is the low-side padding of the
// cross-correlation this stage performs (see afr_api.cu for the forward/adjoint table).
struct TapsG {
    float k[AFR_MAX_TAPS * AFR_MAX_TAPS];
    int n;
    int pad;
};

struct Taps3 {
    float k[3][3];
};

// Offset of plane p = (b, c) of a channel slice that lives inside a larger NCHW tensor: batch stride
// `bstride`, dense channels of `plane` elements each.  C == 0 means a dense [planes, plane] tensor.
__device__ __forceinline__ long strided_base(unsigned p, int C, long bstride, long plane)
{
    return C > 0 ? (long)(p / (unsigned)C) * bstride + (long)(p % (unsigned)C) * plane : (long)p * plane;
}

// ---- scalar / vector loads and stores with fp32 conversion --------------------
__device__ __forceinline__ float ld1(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const bf16 *p)
{
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
}
__device__ __forceinline__ void st1(float *p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16 *p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements; pointer must be aligned to 4 elements.
__device__ __forceinline__ float4 ld4(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ float4 ld4(const bf16 *p)
{
    uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
}
__device__ __forceinline__ void st4(float *p, float4 v)
{
    __stcs(reinterpret_cast<float4 *>(p), v);   // streaming: written once, not re-read by us
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void st4(bf16 *p, float4 v)
{
    uint2 r;
    r.x = pack_bf16x2(v.x, v.y);
    r.y = pack_bf16x2(v.z, v.w);
    __stcs(reinterpret_cast<uint2 *>(p), r);
}

// 8 consecutive elements (pointer aligned to 8 elements): one 128-bit access for bf16
__device__ __forceinline__ void ld8(const float *p, float (&v)[8])
{
    const float4 a = ld4(p), b = ld4(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16 *p, float (&v)[8])
{
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void st8(float *p, const float (&v)[8])
{
    st4(p, make_float4(v[0], v[1], v[2], v[3]));
    st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void st8(bf16 *p, const float (&v)[8])
{
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
    __stcs(reinterpret_cast<uint4 *>(p), r);
}

// shared-memory flavours (no __ldg)
__device__ __forceinline__ float lds1(const float *p) { return *p; }
__device__ __forceinline__ float lds1(const bf16 *p)
{
    return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short *>(p)) << 16);
}
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 lds4(const bf16 *p)
{
    uint2 r = *reinterpret_cast<const uint2 *>(p);
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
}

// ---- exact-erf GELU in fp32 with one MUFU ------------------------------------------
// The fused filtered-GELU kernels are issue-bound (4 GELUs per output element), so the cost
// of GELU sets their speed.  erff() costs ~25 instructions; the forms below cost 10 (value)
// and 17 (derivative), and the pair versions share one packed-FMA Horner chain (FFMA2,
// fma.rn.f32x2 -- new on sm_100) between two arguments.
//
// value:       gelu(x) = relu(x) - t * 2^P(t),   t = |x|,  P(t) ~ log2(Phi(-t))
//              degree-5 P fitted on [0, 5.5] (tools/fit_gelu_log.py); its leading coefficient is
//              negative and P is monotonically decreasing beyond the fit interval, so no clamp
//              is needed: max abs error 7.1e-7 vs erf-GELU in double over |x| <= 3e38 and, with
//              P(0) pinned to -1, relative error <= 6e-6 for -1 < x < 0 (tiny outputs stay
//              accurate).  torch's own fp32 erf path is ~2e-7.  NaN propagates; gelu(+-inf) = NaN.
// derivative:  gelu'(x) = x < 0 ? m : 1 - m,      m = E * S(t),  t = min(|x|, 5.5),
//              E = exp(-t^2/2),  S(t) = Phi(-t) exp(t^2/2) - t/sqrt(2 pi) ~ degree-8 polynomial
//              (weight E): max abs error 1.1e-6.
typedef unsigned long long f32x2;   // two packed floats in one 64-bit register pair

__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 splat2(float c) { return pack2(c, c); }

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.EX2, rel. error 2^-22
    return y;
}

__device__ __forceinline__ float relu_nan(float x)
{
    float y;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));   // keeps NaN, unlike fmaxf
    return y;
}

#define AFR_P5 -4.881020591e-04f
#define AFR_P4 7.198718040e-03f
#define AFR_P3 -5.214663124e-02f
#define AFR_P2 -4.595958447e-01f
#define AFR_P1 -1.151000542e+00f
#define AFR_P0 -1.0f                 /* pinned: relative error -> 0 as x -> 0- */

__device__ __forceinline__ float gelu_erf(float x)
{
    const float t = fabsf(x);
    float p = AFR_P5;
    p = fmaf(p, t, AFR_P4);
    p = fmaf(p, t, AFR_P3);
    p = fmaf(p, t, AFR_P2);
    p = fmaf(p, t, AFR_P1);
    p = fmaf(p, t, AFR_P0);
    return fmaf(-t, ex2_approx(p), relu_nan(x));
}

// two GELUs, in place, sharing one FFMA2 Horner chain.  The chain runs in s = -|x| (odd
// coefficients sign-flipped) so that the final  relu(x) + s * 2^P  is one more packed FMA.
__device__ __forceinline__ void gelu_erf_x2(float &a, float &b)
{
    const f32x2 sv = pack2(-fabsf(a), -fabsf(b));
    f32x2 p = splat2(-AFR_P5);
    p = fma2(p, sv, splat2(AFR_P4));
    p = fma2(p, sv, splat2(-AFR_P3));
    p = fma2(p, sv, splat2(AFR_P2));
    p = fma2(p, sv, splat2(-AFR_P1));
    p = fma2(p, sv, splat2(AFR_P0));
    float pa, pb;
    unpack2(p, pa, pb);
    unpack2(fma2(sv, pack2(ex2_approx(pa), ex2_approx(pb)), pack2(relu_nan(a), relu_nan(b))), a, b);
}

// ---- bf16-grade variants (kernels whose tensors are bf16) ---------------------------------------
// The fused kernels are issue-bound, and bf16 storage halves the bytes but not the instructions, so the
// bf16 kernels trade the ~1e-6 accuracy of the forms above (which bf16 rounding, 2^-9, cannot show) for
// fewer packed FMAs: a degree-4 P (tools/fit_gelu_log.py 4 4.5; its leading coefficient is positive, so
// t is clamped to 4.5, where t*Phi(-t) = 1.5e-5): max abs error 1.4e-4 and relative error <= 9.1e-4
// (2^-10.1) for |x| < 1 -- both well inside the 2^-8 bound together with the output rounding -- and a
// degree-4 S~ for the derivative (max abs error 2.5e-4 = 2^-12, clamp 4.5).
#define AFR_LOW_T 4.5f
#define AFR_Q4 2.371110529e-03f
#define AFR_Q3 -3.545632926e-02f
#define AFR_Q2 -4.825676429e-01f
#define AFR_Q1 -1.141123898e+00f

__device__ __forceinline__ float gelu_erf_low(float x)
{
    const float t = fminf(fabsf(x), AFR_LOW_T);
    float p = AFR_Q4;
    p = fmaf(p, t, AFR_Q3);
    p = fmaf(p, t, AFR_Q2);
    p = fmaf(p, t, AFR_Q1);
    p = fmaf(p, t, AFR_P0);
    return fmaf(-t, ex2_approx(p), relu_nan(x));
}

__device__ __forceinline__ void gelu_erf_low_x2(float &a, float &b)
{
    const f32x2 sv = pack2(fmaxf(-fabsf(a), -AFR_LOW_T), fmaxf(-fabsf(b), -AFR_LOW_T));
    f32x2 p = splat2(AFR_Q4);
    p = fma2(p, sv, splat2(-AFR_Q3));
    p = fma2(p, sv, splat2(AFR_Q2));
    p = fma2(p, sv, splat2(-AFR_Q1));
    p = fma2(p, sv, splat2(AFR_P0));
    float pa, pb;
    unpack2(p, pa, pb);
    unpack2(fma2(sv, pack2(ex2_approx(pa), ex2_approx(pb)), pack2(relu_nan(a), relu_nan(b))), a, b);
}

#define AFR_GELU_T 5.5f
#define AFR_S8 3.792973455e-05f
#define AFR_S7 -6.157795600e-04f
#define AFR_S6 4.401968577e-03f
#define AFR_S5 -1.883073095e-02f
#define AFR_S4 5.615702028e-02f
#define AFR_S3 -1.299775150e-01f
#define AFR_S2 2.492807958e-01f
#define AFR_S1 -7.978182994e-01f
#define AFR_S0 4.999990023e-01f
#define AFR_NHALF_LOG2E -0.72134752044448170368f

__device__ __forceinline__ float gelu_erf_grad(float x)
{
    const float t = fminf(fabsf(x), AFR_GELU_T);
    const float e = ex2_approx(t * t * AFR_NHALF_LOG2E);        // exp(-t^2/2)
    float s = AFR_S8;
    s = fmaf(s, t, AFR_S7);
    s = fmaf(s, t, AFR_S6);
    s = fmaf(s, t, AFR_S5);
    s = fmaf(s, t, AFR_S4);
    s = fmaf(s, t, AFR_S3);
    s = fmaf(s, t, AFR_S2);
    s = fmaf(s, t, AFR_S1);
    s = fmaf(s, t, AFR_S0);
    const float m = e * s;
    return x < 0.f ? m : 1.0f - m;
}

// ---- derivative in the kappa-scaled variable (N == 3 adjoint kernels) -------------------
// The adjoint kernels receive the up-filter taps pre-multiplied by kappa = sqrt(log2(e)/2), so
// they hold w = kappa*u and exp(-u^2/2) = 2^(-w^2) needs no scaling multiply; S is re-expressed
// in v = min(|w|, kappa*5.5).  gelu'(u) = 1/2 + copysign(1/2 - m, u),  m = 2^(-v^2) * S~(v)
// (the AFR_W* coefficients below are those of -S~ so that 1/2 - m is a single FMA)
// (1/2 - m > 0 for all v, so the sign transfer is exact).  Packed f32x2 arithmetic throughout.
#define AFR_KAPPA 0.8493218003f
#define AFR_VT 4.6712699f
#define AFR_W8 -1.400882242e-04f
#define AFR_W7 1.931609655e-03f
#define AFR_W6 -1.172771243e-02f
#define AFR_W5 4.260943937e-02f
#define AFR_W4 -1.079232386e-01f
#define AFR_W3 2.121540929e-01f
#define AFR_W2 -3.455765616e-01f
#define AFR_W1 9.393592619e-01f
#define AFR_W0 -4.999990023e-01f

__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float copysign_bits(float mag_nonneg, float sgn)
{
    return __uint_as_float(__float_as_uint(mag_nonneg) | (__float_as_uint(sgn) & 0x80000000u));
}

// (ga, gb) <- (gelu'(wa/kappa) * ga, gelu'(wb/kappa) * gb)
__device__ __forceinline__ void gelu_grad_scaled_mul_x2(float wa, float wb, float &ga, float &gb)
{
    const float va = fminf(fabsf(wa), AFR_VT), vb = fminf(fabsf(wb), AFR_VT);
    const f32x2 v = pack2(va, vb);
    float qa, qb;
    unpack2(mul2(v, v), qa, qb);
    const f32x2 e = pack2(ex2_approx(-qa), ex2_approx(-qb));
    f32x2 s = splat2(AFR_W8);
    s = fma2(s, v, splat2(AFR_W7));
    s = fma2(s, v, splat2(AFR_W6));
    s = fma2(s, v, splat2(AFR_W5));
    s = fma2(s, v, splat2(AFR_W4));
    s = fma2(s, v, splat2(AFR_W3));
    s = fma2(s, v, splat2(AFR_W2));
    s = fma2(s, v, splat2(AFR_W1));
    s = fma2(s, v, splat2(AFR_W0));
    float ha, hb;
    unpack2(fma2(e, s, splat2(0.5f)), ha, hb);                          // 1/2 - m   (s = -S~)
    const f32x2 r = add2(pack2(copysign_bits(ha, wa), copysign_bits(hb, wb)), splat2(0.5f));
    unpack2(mul2(pack2(ga, gb), r), ga, gb);
}

// bf16-grade derivative: degree-4 S~ (tools/fit_gelu.py re-run with weight E, degree 4, T = 4.5)
#define AFR_VT_LOW 3.8219481f          /* kappa * 4.5 */
#define AFR_X4 -1.884529142e-02f
#define AFR_X3 1.151408768e-01f
#define AFR_X2 -2.969387650e-01f
#define AFR_X1 9.305840670e-01f
#define AFR_X0 -4.997464587e-01f

__device__ __forceinline__ void gelu_grad_scaled_mul_low_x2(float wa, float wb, float &ga, float &gb)
{
    const float va = fminf(fabsf(wa), AFR_VT_LOW), vb = fminf(fabsf(wb), AFR_VT_LOW);
    const f32x2 v = pack2(va, vb);
    float qa, qb;
    unpack2(mul2(v, v), qa, qb);
    const f32x2 e = pack2(ex2_approx(-qa), ex2_approx(-qb));
    f32x2 s = splat2(AFR_X4);
    s = fma2(s, v, splat2(AFR_X3));
    s = fma2(s, v, splat2(AFR_X2));
    s = fma2(s, v, splat2(AFR_X1));
    s = fma2(s, v, splat2(AFR_X0));
    float ha, hb;
    unpack2(fma2(e, s, splat2(0.5f)), ha, hb);
    const f32x2 r = add2(pack2(copysign_bits(ha, wa), copysign_bits(hb, wb)), splat2(0.5f));
    unpack2(mul2(pack2(ga, gb), r), ga, gb);
}

__device__ __forceinline__ float gelu_grad_scaled_low(float w)
{
    const float v = fminf(fabsf(w), AFR_VT_LOW);
    const float e = ex2_approx(-v * v);
    float s = AFR_X4;
    s = fmaf(s, v, AFR_X3);
    s = fmaf(s, v, AFR_X2);
    s = fmaf(s, v, AFR_X1);
    s = fmaf(s, v, AFR_X0);
    return 0.5f + copysign_bits(fmaf(e, s, 0.5f), w);
}

__device__ __forceinline__ float gelu_grad_scaled(float w)
{
    const float v = fminf(fabsf(w), AFR_VT);
    const float e = ex2_approx(-v * v);
    float s = AFR_W8;
    s = fmaf(s, v, AFR_W7);
    s = fmaf(s, v, AFR_W6);
    s = fmaf(s, v, AFR_W5);
    s = fmaf(s, v, AFR_W4);
    s = fmaf(s, v, AFR_W3);
    s = fmaf(s, v, AFR_W2);
    s = fmaf(s, v, AFR_W1);
    s = fmaf(s, v, AFR_W0);
    return 0.5f + copysign_bits(fmaf(e, s, 0.5f), w);
}

}  // namespace afr
