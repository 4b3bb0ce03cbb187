// Throughput microbenchmark of candidate GELU evaluations (compute only, values in registers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o gelu_ubench gelu_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../aliasfree-diffusion-models-pytorch_b200/csrc/afr_common.cuh"
using namespace afr;

__device__ __forceinline__ float gelu_log6(float x)
{   // gelu = relu(x) - t * 2^P(t)
    float t = fminf(fabsf(x), 5.5f);
    float p = -8.7e-5f;
    p = fmaf(p, t, 2.9e-3f); p = fmaf(p, t, -3.6e-2f); p = fmaf(p, t, 1.1e-1f);
    p = fmaf(p, t, -7.6e-1f); p = fmaf(p, t, -1.15f); p = fmaf(p, t, -1.0f);
    return fmaf(-t, ex2_approx(p), relu_nan(x));
}
__device__ __forceinline__ float gelu_erff(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678f)); }

__device__ __forceinline__ void gelu_log6_x2(float &x0, float &x1)
{
    float t0 = fminf(fabsf(x0), 5.5f), t1 = fminf(fabsf(x1), 5.5f);
    f32x2 t = pack2(t0, t1);
    f32x2 p = pack2(-8.7e-5f, -8.7e-5f);
    p = fma2(p, t, pack2(2.9e-3f, 2.9e-3f)); p = fma2(p, t, pack2(-3.6e-2f, -3.6e-2f));
    p = fma2(p, t, pack2(1.1e-1f, 1.1e-1f)); p = fma2(p, t, pack2(-7.6e-1f, -7.6e-1f));
    p = fma2(p, t, pack2(-1.15f, -1.15f)); p = fma2(p, t, pack2(-1.0f, -1.0f));
    float p0, p1; unpack2(p, p0, p1);
    x0 = fmaf(-t0, ex2_approx(p0), relu_nan(x0));
    x1 = fmaf(-t1, ex2_approx(p1), relu_nan(x1));
}

// all-half2 GELU pair: x (half2) -> gelu (half2); Horner in HFMA2, one MUFU.EX2.F16x2
__device__ __forceinline__ __half2 gelu_h2(__half2 x)
{
    const __half2 t = __habs2(x);
    __half2 p = __float2half2_rn(AFR_P5);
    p = __hfma2(p, t, __float2half2_rn(AFR_P4));
    p = __hfma2(p, t, __float2half2_rn(AFR_P3));
    p = __hfma2(p, t, __float2half2_rn(AFR_P2));
    p = __hfma2(p, t, __float2half2_rn(AFR_P1));
    p = __hfma2(p, t, __float2half2_rn(AFR_P0));
    const __half2 e = h2exp2(p);
    return __hfma2(__hneg2(t), e, __hmax2(x, __float2half2_rn(0.f)));
}

__global__ void __launch_bounds__(256) bench_h2(float *out, int iters, float seed)
{
    __half2 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __floats2half2_rn(seed + 0.01f * (threadIdx.x + i * 37), seed - 0.02f * i);
    const __half2 a = __floats2half2_rn(0.37f, -0.21f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __hadd2(gelu_h2(v[i]), a);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __low2float(v[i]) + __high2float(v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
__global__ void __launch_bounds__(256) bench(float *out, int iters, float seed)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + 0.01f * (threadIdx.x + i * 37);
    for (int it = 0; it < iters; ++it) {
        if (V == 3) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { gelu_log6_x2(v[i], v[i + 1]); v[i] += 0.37f; v[i + 1] -= 0.21f; }
        } else if (V == 5) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) { gelu_erf_x2(v[i], v[i + 1]); v[i] += 0.37f; v[i + 1] -= 0.21f; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float g = V == 0 ? gelu_erf(v[i]) : V == 1 ? gelu_log6(v[i]) : V == 2 ? gelu_erff(v[i]) : gelu_erf_grad(v[i]);
                v[i] = g + (i & 1 ? 0.37f : -0.21f);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain FFMA / FFMA2 issue-rate probes: 8 independent chains per thread
template <int V>
__global__ void __launch_bounds__(256) fma_probe(float *out, int iters, float a, float b)
{
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = a * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);              // reg, reg, reg
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], 1.0001f, 0.5f);     // immediates
        } else {
            f32x2 aa = pack2(a, a), bb = pack2(b, b);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                f32x2 r = fma2(pack2(v[i], v[i + 1]), aa, bb);
                unpack2(r, v[i], v[i + 1]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    float *out; cudaMalloc(&out, blocks * threads * sizeof(float));
    const double n = (double)blocks * threads * iters * 8;
    const char *names[] = {"shipping gelu_erf (scalar)", "log-domain deg6 clamped", "erff (CUDA libm)", "log6 clamped f32x2 pairs", "gelu' exp*poly8 (scalar)"};
    float ms;
    ms = time_ms([&] { bench<0><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", names[0], n / ms / 1e6);
    ms = time_ms([&] { bench<1><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", names[1], n / ms / 1e6);
    ms = time_ms([&] { bench<2><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", names[2], n / ms / 1e6);
    ms = time_ms([&] { bench<3><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", names[3], n / ms / 1e6);
    ms = time_ms([&] { bench<4><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", names[4], n / ms / 1e6);
    ms = time_ms([&] { bench<5><<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", "shipping gelu_erf_x2 (FFMA2)", n / ms / 1e6);
    ms = time_ms([&] { bench_h2<<<blocks, threads>>>(out, iters, 0.1f); }); printf("%-24s %8.1f Ggelu/s\n", "all-half2 (HFMA2, EX2.F16x2)", n / ms / 1e6);
    const double nf = (double)blocks * threads * iters * 16;
    ms = time_ms([&] { fma_probe<0><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); }); printf("FFMA reg,reg,reg         %8.1f GFMA/s\n", nf / ms / 1e6);
    ms = time_ms([&] { fma_probe<1><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); }); printf("FFMA imm                 %8.1f GFMA/s\n", nf / ms / 1e6);
    ms = time_ms([&] { fma_probe<2><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); }); printf("FFMA2 (f32x2)            %8.1f GFMA/s\n", nf / ms / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
