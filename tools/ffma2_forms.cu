// Does the operand form of FFMA2 change how it shares the issue port?  8 independent FFMA2 chains + 8 FMNMX per
// iteration, FFMA2 written (a) with plain 64-bit register operands, (b) with a -|x| source modifier folded in,
// (c) with a 32-bit broadcast (.F32) coefficient operand, (d) with the coefficient in a uniform register / constant
// bank, (e) b+c together (the shape the fused kernels use).  Event-timed, w warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmx(float a, float b) { float d; asm volatile("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

struct Coef { float c[8]; };

template <int MODE, int NX>
__global__ void __launch_bounds__(128) body(float *out, const float *in, int iters, const __grid_constant__ Coef K)
{
    float xa[8], xb[8], q[8];
    u64 p[8];
    for (int i = 0; i < 8; ++i) { xa[i] = in[i] + threadIdx.x; xb[i] = in[8 + i] - threadIdx.x; p[i] = pk(xa[i], xb[i]); q[i] = xa[i]; }
    const float r0 = in[3], r1 = in[5];
    const u64 c1 = pk(in[1] + threadIdx.x, in[2] - threadIdx.x), c2 = pk(in[4] + threadIdx.x, in[6]);   // per-thread: R registers
    float rr0 = in[3] + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) p[i] = fma2(p[i], c1, c2);
            if (MODE == 1) p[i] = fma2(p[i], pk(-fabsf(xa[i]), -fabsf(xb[i])), c2);
            if (MODE == 2) p[i] = fma2(p[i], c1, pk(r0, r0));
            if (MODE == 3) p[i] = fma2(p[i], c1, pk(K.c[i & 3], K.c[i & 3]));
            if (MODE == 4) p[i] = fma2(p[i], pk(-fabsf(xa[i]), -fabsf(xb[i])), pk(r1, r1));
            if (MODE == 5) p[i] = fma2(p[i], pk(-fabsf(xa[i]), -fabsf(xb[i])), pk(K.c[i & 3], K.c[i & 3]));
            if (MODE == 6) p[i] = fma2(p[i], c1, pk(0.0071987f, 0.0071987f));
            if (MODE == 7) {     // -|x| materialised with integer ORs (ALU pipe), register operands only
                const float sa = __uint_as_float(__float_as_uint(xa[i]) | 0x80000000u), sb = __uint_as_float(__float_as_uint(xb[i]) | 0x80000000u);
                p[i] = fma2(p[i], pk(sa, sb), pk(rr0, rr0));
                xa[i] = sb; xb[i] = sa;       // keep the ORs inside the loop
            }
            if (MODE == 8) p[i] = fma2(p[i], c1, pk(rr0, rr0));
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) q[i] = fmx(q[i], r0);
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float a, b; upk(p[i], a, b); r += a + b + q[i]; }
    if (r == 123.456f) out[0] = r;
}

template <int MODE, int NX>
void run(const char *name, float *out, float *in)
{
    const int iters = 20000; const double ghz = 1.965;
    Coef K; for (int i = 0; i < 8; ++i) K.c[i] = 0.25f + i;
    printf("%-44s FFMA2 8 FMNMX %d |", name, NX);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w : {1, 2, 4, 8}) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            body<MODE, NX><<<148 * w, 128>>>(out, in, iters, K);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("  w=%d %6.1f", w, best * 1e-3 * ghz * 1e9 / ((double)iters * w));
    }
    printf("\n");
}

int main()
{
    float *out, *in; cudaMalloc(&out, 64); cudaMalloc(&in, 64);
    float h[16]; for (int i = 0; i < 16; ++i) h[i] = 0.5f + 0.001f * i; cudaMemcpy(in, h, 64, cudaMemcpyHostToDevice);
    run<0, 0>("plain registers", out, in);            run<0, 8>("plain registers + FMNMX", out, in);
    run<1, 0>("-|x| modifier", out, in);              run<1, 8>("-|x| modifier + FMNMX", out, in);
    run<2, 0>("broadcast .F32 register", out, in);    run<2, 8>("broadcast .F32 register + FMNMX", out, in);
    run<3, 0>("broadcast constant / UR", out, in);    run<3, 8>("broadcast constant / UR + FMNMX", out, in);
    run<4, 0>("-|x| and .F32 register", out, in);     run<4, 8>("-|x| and .F32 register + FMNMX", out, in);
    run<5, 0>("-|x| and constant / UR", out, in);     run<5, 8>("-|x| and constant / UR + FMNMX", out, in);
    run<6, 0>("R operands, immediate coefficient", out, in);  run<6, 8>("R operands, immediate coefficient + FMNMX", out, in);
    run<8, 0>("R operands, .F32 R coefficient", out, in);     run<8, 8>("R operands, .F32 R coefficient + FMNMX", out, in);
    run<7, 0>("LOP3 -|x| (16 LOP3), R operands", out, in);    run<7, 8>("LOP3 -|x| (16 LOP3), R operands + FMNMX", out, in);
    return 0;
}
