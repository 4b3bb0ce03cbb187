// N == 3 standalone resamplers for channels-last memory ([B, H, W, C], C % 4 == 0).
//
// In a channels-last UNet the six custom_upsample / custom_downsample sites (and the torch.cat that follows an
// upsample) were the last places that forced NCHW copies: ~40 direct_copy launches and 5 % of a reverse step at 4096
// images (profiles/r02_ncu_launches_ddpm_step_b4096_channels_last.md).  These kernels take the tensors as they are.
// "Flat" formulation (no row loop over the plane, parallelism from the grid): a thread owns V channels (one 128-bit
// vector: 4 fp32 or, when C % 8 == 0, 8 bf16) of
//   up-like    one INPUT pixel (i, j): loads x[i][j], x[i][j+1], x[i+1][j], x[i+1][j+1] and writes the 2 x 2 output
//              block (2i..2i+1, 2j..2j+1) -- consecutive lanes are consecutive channel groups, so every access of a
//              warp is one or more whole 128-byte lines;
//   down-like  a PH x PW block of OUTPUT pixels: walks the (2 PH + 1) x (2 PW + 1) input neighbourhood row by row
//              (2 x 2 outputs: 25 loads for 4 outputs instead of 36, and 1.3x instead of 1.6x the input through L2).
// The flat index is decomposed with multiply-high divisions (host-made magic numbers): three runtime integer
// divisions were a fifth of the 1 x 1 kernel's instructions.
// Either side may be a channel slice of a wider tensor (pixel stride > C): custom_upsample writes straight into its
// half of the concatenated tensor and its adjoint reads the gradient slice in place.  `taps` arrive already arranged
// for the stencil (flipped for the adjoints, see afr_api.cu).
#include <cstdlib>

#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

namespace {

// n / d for n < 2^31: umulhi(n, m) >> s with m = ceil(2^(31 + ceil(log2 d)) / d)
struct FastDiv {
    unsigned m, s, d;
};

FastDiv make_fastdiv(unsigned d)
{
    FastDiv f{0u, 0u, d};
    if (d > 1) {
        unsigned l = 0;
        while ((1ull << l) < d) ++l;
        const unsigned p = 31 + l;
        f.m = (unsigned)(((1ull << p) + d - 1) / d);
        f.s = p - 32;
    }
    return f;
}

__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv &f) { return f.d == 1 ? n : (__umulhi(n, f.m) >> f.s); }

template <int V, typename T> __device__ __forceinline__ void ldv(const T *p, float (&v)[V])
{
    if constexpr (V == 8) {
        ld8(p, v);
    } else {
        const float4 a = ld4(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
}
template <int V, typename T> __device__ __forceinline__ void ldv_if(bool ok, const T *p, float (&v)[V])
{
    if (ok) {
        ldv<V>(p, v);
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = 0.f;
    }
}
template <int V, typename T> __device__ __forceinline__ void stv(T *p, const float (&v)[V])
{
    if constexpr (V == 8) st8(p, v);
    else st4(p, make_float4(v[0], v[1], v[2], v[3]));
}

// x: [B, H, W, (xps)] -> u: [B, 2H, 2W, (ups)], C / V channel vectors per pixel
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256)
up3_nhwc_kernel(const TI *__restrict__ x, TO *__restrict__ u, unsigned total, FastDiv dCV, FastDiv dW, FastDiv dH, int H, int W,
                long xps, long ups, const __grid_constant__ Taps3 k)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned pix = fdiv(idx, dCV);
    const int c = V * (int)(idx - pix * dCV.d);
    const unsigned row = fdiv(pix, dW);                          // b * H + i
    const int j = (int)(pix - row * dW.d);
    const unsigned b = fdiv(row, dH);
    const int i = (int)(row - b * dH.d);
    const bool has_r = j + 1 < W, has_b = i + 1 < H;
    const TI *p = x + (long)pix * xps + c;
    float a[V], ar[V], ab[V], abr[V], o[V];
    ldv<V>(p, a);
    ldv_if<V>(has_r, p + xps, ar);
    ldv_if<V>(has_b, p + (long)W * xps, ab);
    ldv_if<V>(has_r && has_b, p + (long)(W + 1) * xps, abr);
    TO *q = u + (((long)b * 2 * H + 2 * i) * (2L * W) + 2 * j) * ups + c;
    const long orow = 2L * W * ups;
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = k.k[1][1] * a[e];
    stv<V>(q, o);                                                                        // (2i,   2j)
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = fmaf(k.k[1][2], ar[e], k.k[1][0] * a[e]);
    stv<V>(q + ups, o);                                                                  // (2i,   2j+1)
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = fmaf(k.k[2][1], ab[e], k.k[0][1] * a[e]);
    stv<V>(q + orow, o);                                                                 // (2i+1, 2j)
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = fmaf(k.k[2][2], abr[e], fmaf(k.k[2][0], ab[e], fmaf(k.k[0][2], ar[e], k.k[0][0] * a[e])));
    stv<V>(q + orow + ups, o);                                                           // (2i+1, 2j+1)
}

// v: [B, H, W, (vps)] -> y: [B, Ho, Wo, (yps)], Ho = ceil(H / 2), Wo = ceil(W / 2); a thread makes PH x PW outputs
// Minimum CTAs per SM: 8 bf16 channels x 1 output at 5 (48 registers instead of 56: 0.72-0.86 of the HBM peak instead of
// 0.68-0.80; 6 CTAs = 40 registers spills and drops to 0.54-0.69); 8 x 2 x 2 at 2 (128 registers).
template <typename T, int V, int PH, int PW, int MINB = (V == 8 && PH * PW == 1) ? 5 : ((PH * PW * V >= 32) ? 2 : 1)>
__global__ void __launch_bounds__(256, MINB)
down3_nhwc_kernel(const T *__restrict__ v, T *__restrict__ y, unsigned total, FastDiv dCV, FastDiv dWb, FastDiv dHb, int H, int W,
                  int Ho, int Wo, long vps, long yps, const __grid_constant__ Taps3 k)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned pb = fdiv(idx, dCV);
    const int c = V * (int)(idx - pb * dCV.d);
    const unsigned rowb = fdiv(pb, dWb);                         // b * Hb + ib
    const int j0 = PW * (int)(pb - rowb * dWb.d);
    const unsigned b = fdiv(rowb, dHb);
    const int i0 = PH * (int)(rowb - b * dHb.d);
    const T *base = v + ((long)b * H * W + (2 * j0 - 1)) * vps + c;
    float acc[PH][PW][V];
#pragma unroll
    for (int a = 0; a < PH; ++a)
#pragma unroll
        for (int q = 0; q < PW; ++q)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[a][q][e] = 0.f;
#pragma unroll
    for (int rr = 0; rr < 2 * PH + 1; ++rr) {
        const int r = 2 * i0 - 1 + rr;
        const bool row_ok = (unsigned)r < (unsigned)H;
        const T *rp = base + (long)r * W * vps;
        float in[2 * PW + 1][V];
#pragma unroll
        for (int cc = 0; cc < 2 * PW + 1; ++cc)
            ldv_if<V>(row_ok && (unsigned)(2 * j0 - 1 + cc) < (unsigned)W, rp + cc * vps, in[cc]);
#pragma unroll
        for (int a = 0; a < PH; ++a) {
            const int ta = rr - 2 * a;                           // tap row this input row meets in output row i0 + a
            if (ta < 0 || ta > 2) continue;
#pragma unroll
            for (int q = 0; q < PW; ++q)
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[a][q][e] = fmaf(k.k[ta][t], in[2 * q + t][e], acc[a][q][e]);
        }
    }
#pragma unroll
    for (int a = 0; a < PH; ++a)
#pragma unroll
        for (int q = 0; q < PW; ++q)
            if (i0 + a < Ho && j0 + q < Wo) stv<V>(y + ((((long)b * Ho + i0 + a) * Wo) + j0 + q) * yps + c, acc[a][q]);
}

bool vec8_ok(int C, const void *a, const void *b, long aps, long bps)
{
    static const bool off = [] { const char *e = getenv("AFR_NHWC_V8"); return e && e[0] == '0'; }();
    return !off && (C % 8) == 0 && (aps % 8) == 0 && (bps % 8) == 0 && (reinterpret_cast<uintptr_t>(a) % 16) == 0 &&
           (reinterpret_cast<uintptr_t>(b) % 16) == 0;
}

// outputs per thread of the down-like kernel: AFR_NHWC_DOWN = 11 | 12 | 21 | 22 (default: by size)
int down_block_default(long outputs, int Ho, int Wo, bool is_bf16)
{
    // fp32: 2 x 2 outputs per thread (1.04-1.10 of the copy peak against 0.96-1.02) unless the tensor is tiny; bf16: the
    // blocks need 77-128 registers with 8-channel vectors and lose more to occupancy than they save (0.43-0.60 vs 0.72-0.86;
    // profiles/r02_nhwc_down_ab.txt)
    return (!is_bf16 && Ho >= 2 && Wo >= 2 && outputs >= (1L << 20)) ? 22 : 11;
}

// read on every call: tests and tools/nhwc_down_ab.py switch it inside one process
int down_block_override()
{
    const char *e = getenv("AFR_NHWC_DOWN");
    return e ? atoi(e) : 0;
}

}  // namespace

bool nhwc_resample_supported(int C, const void *a, const void *b, long aps, long bps, int adtype, int bdtype)
{
    const long ea = adtype == AFR_F32 ? 4 : 2, eb = bdtype == AFR_F32 ? 4 : 2;
    return C >= 4 && (C % 4) == 0 && (reinterpret_cast<uintptr_t>(a) % (4 * ea)) == 0 &&
           (reinterpret_cast<uintptr_t>(b) % (4 * eb)) == 0 && (aps % 4) == 0 && (bps % 4) == 0 && aps >= C && bps >= C;
}

cudaError_t nhwc_up_like(const void *x, void *u, long B, int C, int H, int W, long xps, long ups, const Taps3 &k, int in_dtype,
                         int out_dtype, cudaStream_t s)
{
    const bool v8 = in_dtype == AFR_BF16 && out_dtype == AFR_BF16 && vec8_ok(C, x, u, xps, ups);
    const int CV = C / (v8 ? 8 : 4);
    const long total = B * H * W * CV;
    if (total >= 0x7fffff00L) return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)((total + 255) / 256);
    const FastDiv dCV = make_fastdiv((unsigned)CV), dW = make_fastdiv((unsigned)W), dH = make_fastdiv((unsigned)H);
#define AFR_UN(TI, TO, V) \
    up3_nhwc_kernel<TI, TO, V><<<grid, 256, 0, s>>>((const TI *)x, (TO *)u, (unsigned)total, dCV, dW, dH, H, W, xps, ups, k)
    if (v8) AFR_UN(bf16, bf16, 8);
    else if (in_dtype == AFR_F32 && out_dtype == AFR_F32) AFR_UN(float, float, 4);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) AFR_UN(bf16, bf16, 4);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) AFR_UN(bf16, float, 4);
    else AFR_UN(float, bf16, 4);
#undef AFR_UN
    return cudaGetLastError();
}

cudaError_t nhwc_down_like(const void *v, void *y, long B, int C, int H, int W, long vps, long yps, const Taps3 &k, int dtype,
                           cudaStream_t s)
{
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const bool v8 = dtype == AFR_BF16 && vec8_ok(C, v, y, vps, yps);
    const int CV = C / (v8 ? 8 : 4);
    int blk = down_block_override();
    if (blk != 11 && blk != 12 && blk != 21 && blk != 22) blk = down_block_default(B * Ho * Wo * CV, Ho, Wo, dtype == AFR_BF16);
    const int PH = blk / 10, PW = blk % 10;
    const int Hb = (Ho + PH - 1) / PH, Wb = (Wo + PW - 1) / PW;
    const long total = B * Hb * Wb * CV;
    if (total >= 0x7fffff00L) return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)((total + 255) / 256);
    const FastDiv dCV = make_fastdiv((unsigned)CV), dWb = make_fastdiv((unsigned)Wb), dHb = make_fastdiv((unsigned)Hb);
#define AFR_DN(T, V, PH_, PW_)                                                                                            \
    down3_nhwc_kernel<T, V, PH_, PW_><<<grid, 256, 0, s>>>((const T *)v, (T *)y, (unsigned)total, dCV, dWb, dHb, H, W, Ho, Wo, \
                                                           vps, yps, k)
#define AFR_DN_P(T, V)                                                                                                    \
    do {                                                                                                                  \
        if (blk == 22) AFR_DN(T, V, 2, 2);                                                                                \
        else if (blk == 21) AFR_DN(T, V, 2, 1);                                                                           \
        else if (blk == 12) AFR_DN(T, V, 1, 2);                                                                           \
        else AFR_DN(T, V, 1, 1);                                                                                          \
    } while (0)
    if (dtype == AFR_F32) AFR_DN_P(float, 4);
    else if (v8) AFR_DN_P(bf16, 8);
    else AFR_DN_P(bf16, 4);
#undef AFR_DN_P
#undef AFR_DN
    return cudaGetLastError();
}

}  // namespace afr
