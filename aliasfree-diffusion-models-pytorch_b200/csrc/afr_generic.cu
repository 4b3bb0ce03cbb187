// Runtime-N kernels: any filter size 1..AFR_MAX_TAPS, any H/W (odd sizes, unaligned
// rows).  These are the correctness backstop; the canonical N==3 shapes take the
// register-strip / TMA kernels in afr_n3.cu.
//
// Two stencil shapes cover forward and adjoint of both resamplers (afr_api.cu builds
// the taps/pad for each use):
//   "up-like"   out[Y][X] = sum_{a,b} k[a][b] * in[(Y+a-pad)/2][(X+b-pad)/2]
//                           over (Y+a-pad), (X+b-pad) even and inside the input
//   "down-like" out[i][j] = sum_{a,b} k[a][b] * in[2i+a-pad][2j+b-pad], zero outside
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

// ---- standalone up-like: 4 outputs per thread along X ---------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
up_like_generic_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes,
                       int Hin, int Win, int Hout, int Wout, const __grid_constant__ TapsG t)
{
    const int wq = (Wout + 3) >> 2;
    const long total = planes * (long)Hout * wq;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % wq);
        const long rest = idx / wq;
        const int Y = (int)(rest % Hout);
        const long p = rest / Hout;
        const TI *src = in + p * (long)Hin * Win;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int a = (t.pad + Y) & 1; a < t.n; a += 2) {
            const int zy = Y + a - t.pad;
            if (zy < 0) continue;
            const int i = zy >> 1;
            if (i >= Hin) break;
            const TI *row = src + (long)i * Win;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int X = 4 * q + e;
                for (int b = (t.pad + X) & 1; b < t.n; b += 2) {
                    const int zx = X + b - t.pad;
                    if (zx < 0) continue;
                    const int c = zx >> 1;
                    if (c >= Win) break;
                    acc[e] = fmaf(t.k[a * t.n + b], ld1(row + c), acc[e]);
                }
            }
        }
        TO *dst = out + (p * Hout + Y) * (long)Wout + 4 * q;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (4 * q + e < Wout) st1(dst + e, acc[e]);
    }
}

// ---- standalone down-like: one output per thread -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
down_like_generic_kernel(const T *__restrict__ in, T *__restrict__ out, long planes,
                         int Hin, int Win, int Hout, int Wout, const __grid_constant__ TapsG t)
{
    const long total = planes * (long)Hout * Wout;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wout);
        const long rest = idx / Wout;
        const int i = (int)(rest % Hout);
        const long p = rest / Hout;
        const T *src = in + p * (long)Hin * Win;
        float acc = 0.f;
        for (int a = 0; a < t.n; ++a) {
            const int vy = 2 * i + a - t.pad;
            if (vy < 0 || vy >= Hin) continue;
            for (int b = 0; b < t.n; ++b) {
                const int vx = 2 * j + b - t.pad;
                if (vx < 0 || vx >= Win) continue;
                acc = fmaf(t.k[a * t.n + b], ld1(src + (long)vy * Win + vx), acc);
            }
        }
        st1(out + idx, acc);
    }
}

// ---- fused filtered GELU, forward and adjoint, shared-memory staged -------------------
// One CTA = one TILE x TILE block of outputs of one plane.
//   stage 0: x (and dy for the adjoint) tiles with halo -> shared, zero filled
//   stage 1: mid[Y][X] on the 2x grid:  fwd  gelu(u),  adjoint  gelu'(u) * dg,
//            forced to 0 outside [0,2H)x[0,2W)   (the reference zero-pads the GELU
//            OUTPUT, modules/ddpm_utils.py:124-125 -> F.conv2d padding='same')
//   stage 2: down-like stencil over mid -> out
// u  = up-like(x,  tU);  dg = up-like(dy, tG)  (adjoint only);  out = down-like(mid, tB)
constexpr int GT = 16;   // output tile edge

// kN > 0: every stage has exactly kN x kN taps (loops fully unrolled, taps become immediate
// constant-bank operands); kN == 0: sizes are read from the TapsG structs at run time.
// The mid tile is stored with even and odd columns de-interleaved (column X lives at
// (X & 1) * MWH + (X >> 1)), so that the stride-2 reads of the final stage are conflict-free.
template <typename T, bool kBwd, int kN>
__global__ void __launch_bounds__(256)
fgelu_generic_kernel(const T *__restrict__ x, const T *__restrict__ res, const T *__restrict__ dy,
                     T *__restrict__ out, int H, int W, int tiles_y, int tiles_x,
                     const __grid_constant__ TapsG tU, const __grid_constant__ TapsG tG,
                     const __grid_constant__ TapsG tB)
{
    extern __shared__ float smem[];
    const int nU = kN ? kN : tU.n, nG = kN ? kN : tG.n, nB = kN ? kN : tB.n;
    const int H2 = 2 * H, W2 = 2 * W;
    long bid = blockIdx.x;
    const int tx = (int)(bid % tiles_x); bid /= tiles_x;
    const int ty = (int)(bid % tiles_y);
    const long p = bid / tiles_y;
    const int i0 = ty * GT, j0 = tx * GT;

    // mid region needed by the GT x GT outputs
    const int MY0 = 2 * i0 - tB.pad, MX0 = 2 * j0 - tB.pad;
    const int MH = 2 * GT - 2 + nB, MW = MH;
    const int MWH = (MW + 1) / 2 + 1;                  // half pitch of the de-interleaved mid rows
    // input rows/cols feeding that mid region through the widest up-like stage
    const int padU = tU.pad, hiU = nU - 1 - tU.pad;
    const int padG = kBwd ? tG.pad : 0, hiG = kBwd ? nG - 1 - tG.pad : 0;
    const int lo = max(padU, padG), hi = max(hiU, hiG);
    const int XY0 = (MY0 - lo) >> 1, XX0 = (MX0 - lo) >> 1;          // arithmetic shift = floor
    const int XH = ((MY0 + MH - 1 + hi) >> 1) - XY0 + 1, XW = ((MX0 + MW - 1 + hi) >> 1) - XX0 + 1;

    float *xs = smem;                       // [XH][XW]
    float *ds = xs + XH * XW;               // [XH][XW] (adjoint only)
    float *ms = ds + (kBwd ? XH * XW : 0);  // [MH][2][MWH]

    const T *xp = x + p * (long)H * W;
    const T *rp = res ? res + p * (long)H * W : nullptr;
    const T *dp = kBwd ? dy + p * (long)H * W : nullptr;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;       // 8 warps: one tile row each per pass
    for (int rr = wrp; rr < XH; rr += 8) {
        const int r = XY0 + rr;
        const bool row_in = (r >= 0 && r < H);
        for (int cc = lane; cc < XW; cc += 32) {
            const int c = XX0 + cc;
            float v = 0.f, d = 0.f;
            if (row_in && c >= 0 && c < W) {
                v = ld1(xp + (long)r * W + c);
                if (rp) v += ld1(rp + (long)r * W + c);
                if (kBwd) d = ld1(dp + (long)r * W + c);
            }
            xs[rr * XW + cc] = v;
            if (kBwd) ds[rr * XW + cc] = d;
        }
    }
    __syncthreads();

    for (int my = wrp; my < MH; my += 8) {
        const int Y = MY0 + my;
        const bool row_in = (Y >= 0 && Y < H2);
        const int aU = (tU.pad + Y) & 1, aG = (tG.pad + Y) & 1;
        for (int mx = lane; mx < MW; mx += 32) {
            const int X = MX0 + mx;
            float m = 0.f;
            if (row_in && X >= 0 && X < W2) {
                float u = 0.f;
                const int bU = (tU.pad + X) & 1;
#pragma unroll
                for (int a2 = 0; a2 < (kN ? (kN + 1) / 2 : AFR_MAX_TAPS / 2); ++a2) {
                    const int a = aU + 2 * a2;
                    if (a >= nU) break;
                    const int r = ((Y + a - tU.pad) >> 1) - XY0;      // zero rows outside the plane are in xs
#pragma unroll
                    for (int b2 = 0; b2 < (kN ? (kN + 1) / 2 : AFR_MAX_TAPS / 2); ++b2) {
                        const int b = bU + 2 * b2;
                        if (b >= nU) break;
                        const int c = ((X + b - tU.pad) >> 1) - XX0;
                        u = fmaf(tU.k[a * nU + b], xs[r * XW + c], u);
                    }
                }
                if (kBwd) {
                    float g = 0.f;
                    const int bG = (tG.pad + X) & 1;
#pragma unroll
                    for (int a2 = 0; a2 < (kN ? (kN + 1) / 2 : AFR_MAX_TAPS / 2); ++a2) {
                        const int a = aG + 2 * a2;
                        if (a >= nG) break;
                        const int r = ((Y + a - tG.pad) >> 1) - XY0;
#pragma unroll
                        for (int b2 = 0; b2 < (kN ? (kN + 1) / 2 : AFR_MAX_TAPS / 2); ++b2) {
                            const int b = bG + 2 * b2;
                            if (b >= nG) break;
                            const int c = ((X + b - tG.pad) >> 1) - XX0;
                            g = fmaf(tG.k[a * nG + b], ds[r * XW + c], g);
                        }
                    }
                    m = gelu_erf_grad(u) * g;
                } else {
                    m = gelu_erf(u);
                }
            }
            ms[(my * 2 + (mx & 1)) * MWH + (mx >> 1)] = m;
        }
    }
    __syncthreads();

    const int li = threadIdx.x / GT, lj = threadIdx.x % GT;
    const int i = i0 + li, j = j0 + lj;
    if (i < H && j < W) {
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < (kN ? kN : AFR_MAX_TAPS); ++a) {
            if (a >= nB) break;
            const float *row = ms + (2 * li + a) * 2 * MWH;
#pragma unroll
            for (int b = 0; b < (kN ? kN : AFR_MAX_TAPS); ++b) {
                if (b >= nB) break;
                acc = fmaf(tB.k[a * nB + b], row[(b & 1) * MWH + lj + (b >> 1)], acc);
            }
        }
        st1(out + (p * H + i) * (long)W + j, acc);
    }
}

// ---- host launchers --------------------------------------------------------------
static inline int grid_for(long total, int block)
{
    long g = (total + block - 1) / block;
    const long cap = 148L * 32;          // grid-stride beyond 32 waves of 148 SMs
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename TI, typename TO>
static cudaError_t launch_up_like(const void *in, void *out, long planes, int Hin, int Win,
                                  int Hout, int Wout, const TapsG &t, cudaStream_t s)
{
    long total = planes * (long)Hout * ((Wout + 3) / 4);
    up_like_generic_kernel<TI, TO><<<grid_for(total, 256), 256, 0, s>>>(
        (const TI *)in, (TO *)out, planes, Hin, Win, Hout, Wout, t);
    return cudaGetLastError();
}

cudaError_t generic_up_like(const void *in, void *out, long planes, int Hin, int Win, int Hout,
                            int Wout, const TapsG &t, int in_dtype, int out_dtype, cudaStream_t s)
{
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32)
        return launch_up_like<float, float>(in, out, planes, Hin, Win, Hout, Wout, t, s);
    if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16)
        return launch_up_like<bf16, bf16>(in, out, planes, Hin, Win, Hout, Wout, t, s);
    if (in_dtype == AFR_BF16 && out_dtype == AFR_F32)
        return launch_up_like<bf16, float>(in, out, planes, Hin, Win, Hout, Wout, t, s);
    return launch_up_like<float, bf16>(in, out, planes, Hin, Win, Hout, Wout, t, s);
}

cudaError_t generic_down_like(const void *in, void *out, long planes, int Hin, int Win, int Hout,
                              int Wout, const TapsG &t, int dtype, cudaStream_t s)
{
    long total = planes * (long)Hout * Wout;
    if (dtype == AFR_F32)
        down_like_generic_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(
            (const float *)in, (float *)out, planes, Hin, Win, Hout, Wout, t);
    else
        down_like_generic_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>(
            (const bf16 *)in, (bf16 *)out, planes, Hin, Win, Hout, Wout, t);
    return cudaGetLastError();
}

template <typename T, bool kBwd, int kN>
static cudaError_t launch_fgelu(const void *x, const void *res, const void *dy, void *out,
                                long planes, int H, int W, const TapsG &tU, const TapsG &tG,
                                const TapsG &tB, cudaStream_t s)
{
    const int tiles_y = (H + GT - 1) / GT, tiles_x = (W + GT - 1) / GT;
    const long blocks = planes * tiles_y * tiles_x;
    if (blocks > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const int MH = 2 * GT - 2 + tB.n, MWH = (MH + 1) / 2 + 1;
    const int lo = kBwd ? (tU.pad > tG.pad ? tU.pad : tG.pad) : tU.pad;
    const int hiU = tU.n - 1 - tU.pad, hiG = kBwd ? tG.n - 1 - tG.pad : 0;
    const int hi = hiU > hiG ? hiU : hiG;
    const int XH = (MH - 1 + lo + hi) / 2 + 2;     // upper bound of the kernel's XH / XW
    const size_t smem = sizeof(float) * ((size_t)XH * XH * (kBwd ? 2 : 1) + (size_t)MH * 2 * MWH);
    fgelu_generic_kernel<T, kBwd, kN><<<(unsigned)blocks, 256, smem, s>>>(
        (const T *)x, (const T *)res, (const T *)dy, (T *)out, H, W, tiles_y, tiles_x, tU, tG, tB);
    return cudaGetLastError();
}

template <typename T, bool kBwd>
static cudaError_t dispatch_n(const void *x, const void *res, const void *dy, void *out, long planes,
                              int H, int W, const TapsG &tU, const TapsG &tG, const TapsG &tB,
                              cudaStream_t s)
{
    // unrolled instances when every stage has the same small size (the usual case: one
    // `kernel_size` for both filters); anything else takes the run-time-N instance
    const bool same = (tU.n == tB.n) && (!kBwd || tG.n == tU.n);
    switch (same ? tU.n : 0) {
#define AFR_CASE(N) case N: return launch_fgelu<T, kBwd, N>(x, res, dy, out, planes, H, W, tU, tG, tB, s);
        AFR_CASE(2) AFR_CASE(4) AFR_CASE(5) AFR_CASE(6) AFR_CASE(7) AFR_CASE(8)
#undef AFR_CASE
    default: return launch_fgelu<T, kBwd, 0>(x, res, dy, out, planes, H, W, tU, tG, tB, s);
    }
}

cudaError_t generic_fgelu(const void *x, const void *res, const void *dy, void *out, long planes,
                          int H, int W, const TapsG &tU, const TapsG &tG, const TapsG &tB,
                          bool bwd, int dtype, cudaStream_t s)
{
    if (dtype == AFR_F32)
        return bwd ? dispatch_n<float, true>(x, res, dy, out, planes, H, W, tU, tG, tB, s)
                   : dispatch_n<float, false>(x, res, dy, out, planes, H, W, tU, tG, tB, s);
    return bwd ? dispatch_n<bf16, true>(x, res, dy, out, planes, H, W, tU, tG, tB, s)
               : dispatch_n<bf16, false>(x, res, dy, out, planes, H, W, tU, tG, tB, s);
}

}  // namespace afr
