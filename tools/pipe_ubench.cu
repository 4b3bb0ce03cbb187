// Issue / pipe interference microbenchmark for the fused-kernel instruction mix (sm_100a).
// Each warp runs ITER iterations of an unrolled body made of independent packed-FMA chains
// (FFMA2), scalar FMA-pipe ops, ALU ops (FMNMX) and MUFU.EX2, with W warps per SM sub-partition;
// reports cycles per iteration per warp and per SMSP.   nvcc -arch=sm_100a -O3 -cudart shared -o pipe_ubench pipe_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmx(float a, float b) { float d; asm volatile("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ex2(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// NP packed chains, NS scalar FMA chains, NA scalar FADD chains, NX ALU chains, NM MUFU chains; DEP: packed ops per
// chain per iteration (dependent), so the body has NP*DEP FFMA2 + NS*DEP FFMA + NA FADD + NX FMNMX + NM MUFU
template <int NP, int DEP, int NS, int NA, int NX, int NM>
__global__ void __launch_bounds__(128) body(float *out, const float *in, int iters, long long *cyc)
{
    float s0 = in[threadIdx.x & 7], s1 = in[8 + (threadIdx.x & 7)];
    u64 p[NP > 0 ? NP : 1]; float f[NS > 0 ? NS : 1], a[NA > 0 ? NA : 1], x[NX > 0 ? NX : 1], m[NM > 0 ? NM : 1];
    for (int i = 0; i < NP; ++i) p[i] = pk(s0 + i, s1 - i);
    for (int i = 0; i < NS; ++i) f[i] = s0 * i;
    for (int i = 0; i < NA; ++i) a[i] = s1 * i;
    for (int i = 0; i < NX; ++i) x[i] = s1 + i;
    for (int i = 0; i < NM; ++i) m[i] = s0 - i;
    const u64 c1 = pk(s0, s0), c2 = pk(s1, s1);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < DEP; ++d) {
#pragma unroll
            for (int i = 0; i < NP; ++i) p[i] = fma2(p[i], c1, c2);
#pragma unroll
            for (int i = 0; i < NS; ++i) f[i] = ffma(f[i], s0, s1);
        }
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i] = fadd(a[i], s1);
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = fmx(x[i], s0);
#pragma unroll
        for (int i = 0; i < NM; ++i) m[i] = ex2(m[i]);
    }
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < NP; ++i) r += lo(p[i]);
    for (int i = 0; i < NS; ++i) r += f[i];
    for (int i = 0; i < NA; ++i) r += a[i];
    for (int i = 0; i < NX; ++i) r += x[i];
    for (int i = 0; i < NM; ++i) r += m[i];
    if (r == 123.456f) out[0] = r;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int NP, int DEP, int NS, int NA, int NX, int NM>
void run(const char *name, float *out, float *in, long long *cyc)
{
    const int iters = 20000;
    const double ghz = 1.965;     // B200 SM clock under load (bench.py samples it)
    printf("%-34s FFMA2 %2d FFMA %2d FADD %2d FMNMX %2d MUFU %2d |", name, NP * DEP, NS * DEP, NA, NX, NM);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w : {1, 2, 3, 5, 8}) {          // CTAs of 4 warps per SM -> w warps per SMSP
        int blocks = 148 * w;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            body<NP, DEP, NS, NA, NX, NM><<<blocks, 128>>>(out, in, iters, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        // SM cycles an SMSP spends per warp-iteration (all w warps of the SMSP run `iters` iterations)
        printf("  w=%d %7.1f", w, best * 1e-3 * ghz * 1e9 / ((double)iters * w));
    }
    printf("\n");
}

int main()
{
    float *out, *in; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&in, 64); cudaMalloc(&cyc, sizeof(long long) * 148 * 8 * 4);
    float h[16]; for (int i = 0; i < 16; ++i) h[i] = 0.5f + 0.001f * i; cudaMemcpy(in, h, 64, cudaMemcpyHostToDevice);
    //   NP DEP NS NA NX NM
    run<8, 1, 0, 0, 0, 0>("pure FFMA2 x8", out, in, cyc);
    run<0, 1, 16, 0, 0, 0>("pure FFMA x16", out, in, cyc);
    run<1, 8, 0, 0, 0, 0>("FFMA2 1 chain x8 dep (latency)", out, in, cyc);
    run<0, 8, 1, 0, 0, 0>("FFMA 1 chain x8 dep (latency)", out, in, cyc);
    run<0, 1, 0, 0, 0, 8>("pure MUFU x8", out, in, cyc);
    run<0, 1, 0, 0, 16, 0>("pure FMNMX x16", out, in, cyc);
    run<8, 1, 8, 0, 0, 0>("FFMA2 x8 + FFMA x8", out, in, cyc);
    run<8, 1, 0, 0, 8, 0>("FFMA2 x8 + FMNMX x8", out, in, cyc);
    run<8, 1, 0, 0, 0, 4>("FFMA2 x8 + MUFU x4", out, in, cyc);
    run<0, 1, 16, 0, 0, 4>("FFMA x16 + MUFU x4", out, in, cyc);
    run<8, 1, 0, 8, 8, 0>("FFMA2 x8 + FADD x8 + FMNMX x8", out, in, cyc);
    // one fused-kernel row: 8 pair chains x 6 FFMA2, 40 scalar FMA-pipe, 16 FMNMX (+~24 other ALU), 16 MUFU
    run<8, 6, 0, 40, 40, 16>("row mix packed (sym kernel)", out, in, cyc);
    run<0, 6, 16, 40, 40, 16>("row mix scalar GELU", out, in, cyc);
    run<8, 5, 0, 40, 40, 16>("row mix packed, degree 4", out, in, cyc);
    run<8, 6, 0, 0, 40, 16>("row mix packed, no scalar FMA", out, in, cyc);
    run<8, 6, 0, 40, 40, 0>("row mix packed, no MUFU", out, in, cyc);
    run<8, 6, 0, 40, 0, 16>("row mix packed, no ALU", out, in, cyc);
    return 0;
}
