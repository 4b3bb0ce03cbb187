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
// gelu(x)  = relu(x) - t*E*R(t)
// gelu'(x) = 0.5 + sign(x)*(0.5 - E*R(t)) + x*E/sqrt(2 pi)
// with t = min(|x|, 5.5), E = exp(-t^2/2) and R(t) = Phi(-t) exp(t^2/2) (half the scaled
// complementary error function), a slowly varying function fitted by a degree-9
// polynomial (tools/fit_gelu.py).  Max abs error vs erf-GELU evaluated in double:
// 6.9e-7 (value), 8.7e-7 (derivative); beyond the clamp the true tail is < 1.1e-7.
// This replaces erff() (about 2x the instructions, two MUFU ops) -- the fused
// filtered-GELU kernel is issue-bound, not HBM-bound, so the cost of GELU sets its speed.
__device__ __forceinline__ float gelu_poly_R(float t)
{
    float r = -6.306644367e-06f;
    r = fmaf(r, t, 1.262593763e-04f);
    r = fmaf(r, t, -1.118502270e-03f);
    r = fmaf(r, t, 5.906143764e-03f);
    r = fmaf(r, t, -2.138766862e-02f);
    r = fmaf(r, t, 5.864728463e-02f);
    r = fmaf(r, t, -1.313082752e-01f);
    r = fmaf(r, t, 2.496296917e-01f);
    r = fmaf(r, t, -3.989106504e-01f);
    r = fmaf(r, t, 4.999995630e-01f);
    return r;
}

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

__device__ __forceinline__ float gelu_erf(float x)
{
    float t = fminf(fabsf(x), 5.5f);
    float e = ex2_approx(t * t * -0.72134752044448170368f);   // exp(-t^2/2)
    float w = t * e;
    return fmaf(-w, gelu_poly_R(t), relu_nan(x));
}

__device__ __forceinline__ float gelu_erf_grad(float x)
{
    float t = fminf(fabsf(x), 5.5f);
    float e = ex2_approx(t * t * -0.72134752044448170368f);
    float er = e * gelu_poly_R(t);
    float h = copysignf(0.5f - er, x);
    return fmaf(x * e, 0.39894228040143267794f, 0.5f + h);
}

}  // namespace afr
