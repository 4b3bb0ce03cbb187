// N == 3 standalone resamplers for channels-last memory ([B, H, W, C], C % 4 == 0).
//
// In a channels-last UNet the six custom_upsample / custom_downsample sites (and the torch.cat that follows an
// upsample) were the last places that forced NCHW copies: ~40 direct_copy launches and 5 % of a reverse step at 4096
// images (profiles/r02_ncu_launches_ddpm_step_b4096_channels_last.md).  These kernels take the tensors as they are.
// "Flat" formulation (no row loop, every load issued up front, parallelism from the grid): a thread owns 4 channels
// (one 128-bit vector) of
//   up-like    one INPUT pixel (i, j): loads x[i][j], x[i][j+1], x[i+1][j], x[i+1][j+1] and writes the 2 x 2 output
//              block (2i..2i+1, 2j..2j+1) -- consecutive lanes are consecutive channel groups, so every access of a
//              warp is one or more whole 128-byte lines;
//   down-like  one OUTPUT pixel (i, j): loads the 3 x 3 input neighbourhood of (2i, 2j) and writes one vector.
// Either side may be a channel slice of a wider tensor (pixel stride > C): custom_upsample writes straight into its
// half of the concatenated tensor and its adjoint reads the gradient slice in place.  `taps` arrive already arranged
// for the stencil (flipped for the adjoints, see afr_api.cu).
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

namespace {

template <typename T> __device__ __forceinline__ float4 ldc4(const T *p);
template <> __device__ __forceinline__ float4 ldc4<float>(const float *p) { return ld4(p); }
template <> __device__ __forceinline__ float4 ldc4<bf16>(const bf16 *p) { return ld4(p); }

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4mul(float k, float4 a) { return make_float4(k * a.x, k * a.y, k * a.z, k * a.w); }
__device__ __forceinline__ float4 f4fma(float k, float4 a, float4 c)
{
    return make_float4(fmaf(k, a.x, c.x), fmaf(k, a.y, c.y), fmaf(k, a.z, c.z), fmaf(k, a.w, c.w));
}

// x: [B, H, W, (xps)] -> u: [B, 2H, 2W, (ups)], C4 = C / 4 channel vectors per pixel
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
up3_nhwc_kernel(const TI *__restrict__ x, TO *__restrict__ u, unsigned total, int C4, int H, int W, long xps, long ups,
                const __grid_constant__ Taps3 k)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned pix = idx / (unsigned)C4;
    const int c = 4 * (int)(idx - pix * (unsigned)C4);
    const unsigned row = pix / (unsigned)W;                      // b * H + i
    const int j = (int)(pix - row * (unsigned)W);
    const int i = (int)(row % (unsigned)H);
    const bool has_r = j + 1 < W, has_b = i + 1 < H;
    const TI *p = x + (long)pix * xps + c;
    const float4 a = ldc4<TI>(p);
    const float4 ar = has_r ? ldc4<TI>(p + xps) : f4zero();
    const float4 ab = has_b ? ldc4<TI>(p + (long)W * xps) : f4zero();
    const float4 abr = (has_r && has_b) ? ldc4<TI>(p + (long)(W + 1) * xps) : f4zero();
    const unsigned b = row / (unsigned)H;
    TO *o = u + (((long)b * 2 * H + 2 * i) * (2L * W) + 2 * j) * ups + c;
    const long orow = 2L * W * ups;
    st4(o, f4mul(k.k[1][1], a));                                                        // (2i,   2j)
    st4(o + ups, f4fma(k.k[1][2], ar, f4mul(k.k[1][0], a)));                            // (2i,   2j+1)
    st4(o + orow, f4fma(k.k[2][1], ab, f4mul(k.k[0][1], a)));                           // (2i+1, 2j)
    st4(o + orow + ups, f4fma(k.k[2][2], abr, f4fma(k.k[2][0], ab, f4fma(k.k[0][2], ar, f4mul(k.k[0][0], a)))));
}

// v: [B, H, W, (vps)] -> y: [B, Ho, Wo, (yps)], Ho = ceil(H / 2), Wo = ceil(W / 2)
template <typename T>
__global__ void __launch_bounds__(256)
down3_nhwc_kernel(const T *__restrict__ v, T *__restrict__ y, unsigned total, int C4, int H, int W, int Ho, int Wo, long vps,
                  long yps, const __grid_constant__ Taps3 k)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned pix = idx / (unsigned)C4;
    const int c = 4 * (int)(idx - pix * (unsigned)C4);
    const unsigned row = pix / (unsigned)Wo;                     // b * Ho + i
    const int j = (int)(pix - row * (unsigned)Wo);
    const unsigned b = row / (unsigned)Ho;
    const int i = (int)(row - b * (unsigned)Ho);
    const T *base = v + ((long)b * H * W) * vps + c;
    float4 acc = f4zero();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int r = 2 * i + a - 1;
        if ((unsigned)r >= (unsigned)H) continue;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int col = 2 * j + q - 1;
            if ((unsigned)col >= (unsigned)W) continue;
            acc = f4fma(k.k[a][q], ldc4<T>(base + ((long)r * W + col) * vps), acc);
        }
    }
    st4(y + (long)pix * yps + c, acc);
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
    const long total = B * H * W * (C / 4);
    if (total >= 0xffffff00L) return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)((total + 255) / 256);
#define AFR_UN(TI, TO) up3_nhwc_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI *)x, (TO *)u, (unsigned)total, C / 4, H, W, xps, ups, k)
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32) AFR_UN(float, float);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) AFR_UN(bf16, bf16);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) AFR_UN(bf16, float);
    else AFR_UN(float, bf16);
#undef AFR_UN
    return cudaGetLastError();
}

cudaError_t nhwc_down_like(const void *v, void *y, long B, int C, int H, int W, long vps, long yps, const Taps3 &k, int dtype,
                           cudaStream_t s)
{
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const long total = B * Ho * Wo * (C / 4);
    if (total >= 0xffffff00L) return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == AFR_F32)
        down3_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float *)v, (float *)y, (unsigned)total, C / 4, H, W, Ho, Wo, vps, yps, k);
    else
        down3_nhwc_kernel<bf16><<<grid, 256, 0, s>>>((const bf16 *)v, (bf16 *)y, (unsigned)total, C / 4, H, W, Ho, Wo, vps, yps, k);
    return cudaGetLastError();
}

}  // namespace afr
