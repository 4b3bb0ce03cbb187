// N == 3 standalone resamplers for SMALL planes (the inner UNet levels: 4x4 ... 16x16; custom_upsample on
// (128,4,4), (64,8,8), (32,16,16) and custom_downsample on (64,16,16), (128,8,8) per image in variant 1/3).
//
// The register-strip kernels of afr_n3.cu give one thread a 4-column strip of ONE plane: on a 4x4 plane that
// is 16-byte loads at a 64-byte stride and 32-byte stores at a 256-byte stride per lane -- every request
// touches 32 different lines, and up2x reached only 0.49 (4x4) / 0.61 (8x8) of the HBM peak, down2x 0.52,
// with W = 4 planes on the generic backstop (0.24).  Here a CTA owns a GROUP of consecutive planes, which is
// one contiguous block of memory on both sides:
//   1. the block of input planes is read with fully coalesced 128-bit loads and scattered into a
//      zero-bordered shared-memory layout (the border is the convolution's zero padding: no bounds checks),
//   2. every thread produces V consecutive output elements of the contiguous output block (V = 16 bytes)
//      from shared memory and writes them with one coalesced 128-bit store.
// Output (up) / input (down) may live inside a larger tensor with its own batch stride, which is how
// custom_upsample writes straight into the channel slice of the torch.cat buffer (modules/ddpm_utils.py:414)
// and how its adjoint reads the gradient slice back.
#include <cstdlib>

#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

namespace {

template <typename T> struct Vec { static constexpr int n = 16 / (int)sizeof(T); };

template <typename T>
__device__ __forceinline__ void store_vec(T *p, const float (&v)[Vec<T>::n]);
template <>
__device__ __forceinline__ void store_vec<float>(float *p, const float (&v)[4])
{
    st4(p, make_float4(v[0], v[1], v[2], v[3]));
}
template <>
__device__ __forceinline__ void store_vec<bf16>(bf16 *p, const float (&v)[8])
{
    st8(p, v);
}

// stage `n4` groups of 4 consecutive input elements (a whole number of planes) into the bordered layout
template <typename T>
__device__ __forceinline__ void stage_planes(const T *__restrict__ src, float *__restrict__ sm, int n4, int W,
                                             int HW, int pitch, int plane_elems, int border)
{
    for (int q = threadIdx.x; q < n4; q += blockDim.x) {
        const float4 v = ld4(src + 4 * q);
        const int e = 4 * q, pl = e / HW, rem = e - pl * HW, r = rem / W, c = rem - r * W;   // W % 4 == 0: one row
        float *d = sm + pl * plane_elems + (r + border) * pitch + c + border;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}

// ---- up-like: [planes, H, W] -> [planes, 2H, 2W] ------------------------------------------------------
// shared layout per plane: (H + 1) x (W + 1) with a zero row below and a zero column to the right
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
up3_group_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int C, long out_bstride, int H, int W,
                 int PG, const __grid_constant__ Taps3 k)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int V = Vec<TO>::n;
    const int HW = H * W, pitch = W + 1, pe = (H + 1) * pitch;
    const long p0 = (long)blockIdx.x * PG;
    const int np = (int)min((long)PG, planes - p0);
    for (int q = threadIdx.x; q < np * pe; q += blockDim.x) sm[q] = 0.f;
    __syncthreads();
    stage_planes<TI>(in + p0 * HW, sm, np * HW / 4, W, HW, pitch, pe, 0);
    __syncthreads();
    const int W2 = 2 * W, per_plane = 4 * HW / V, per_row = W2 / V;
    for (int q = threadIdx.x; q < np * per_plane; q += blockDim.x) {
        const int pl = q / per_plane, rem = q - pl * per_plane, Y = rem / per_row, X0 = (rem - Y * per_row) * V;
        const int i = Y >> 1, c0 = X0 >> 1;
        const float *ra = sm + pl * pe + i * pitch + c0, *rb = ra + pitch;
        float xa[V / 2 + 1], o[V];
#pragma unroll
        for (int c = 0; c <= V / 2; ++c) xa[c] = ra[c];
        if (Y & 1) {
            float xb[V / 2 + 1];
#pragma unroll
            for (int c = 0; c <= V / 2; ++c) xb[c] = rb[c];
#pragma unroll
            for (int c = 0; c < V / 2; ++c) {
                o[2 * c] = fmaf(k.k[2][1], xb[c], k.k[0][1] * xa[c]);
                float t = k.k[0][0] * xa[c];
                t = fmaf(k.k[0][2], xa[c + 1], t);
                t = fmaf(k.k[2][0], xb[c], t);
                o[2 * c + 1] = fmaf(k.k[2][2], xb[c + 1], t);
            }
        } else {
#pragma unroll
            for (int c = 0; c < V / 2; ++c) {
                o[2 * c] = k.k[1][1] * xa[c];
                o[2 * c + 1] = fmaf(k.k[1][2], xa[c + 1], k.k[1][0] * xa[c]);
            }
        }
        const long p = p0 + pl;
        TO *dst = out + (p / C) * out_bstride + (p % C) * 4L * HW + (long)Y * W2 + X0;
        store_vec<TO>(dst, o);
    }
}

// ---- down-like: [planes, H, W] -> [planes, ceil(H/2), W/2] -------------------------------------------------
// shared layout per plane: (H + 2) x (W + 2), zero border all round; a thread makes V consecutive elements of
// the contiguous output block (they may span rows and planes when W/2 < V)
template <typename T>
__global__ void __launch_bounds__(256)
down3_group_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int C, long in_bstride, int H, int W,
                   int PG, const __grid_constant__ Taps3 k)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int V = Vec<T>::n;
    const int HW = H * W, pitch = W + 2, pe = (H + 2) * pitch, Ho = (H + 1) / 2, Wo = W / 2, HWo = Ho * Wo;
    const long p0 = (long)blockIdx.x * PG;
    const int np = (int)min((long)PG, planes - p0);
    for (int q = threadIdx.x; q < np * pe; q += blockDim.x) sm[q] = 0.f;
    __syncthreads();
    if (in_bstride == (long)C * HW) {                       // dense input: one contiguous block
        stage_planes<T>(in + p0 * HW, sm, np * HW / 4, W, HW, pitch, pe, 1);
    } else {                                                // a channel slice of a larger tensor: plane by plane
        const int n4 = HW / 4;
        for (int q = threadIdx.x; q < np * n4; q += blockDim.x) {
            const int pl = q / n4, e = 4 * (q - pl * n4), r = e / W, c = e - r * W;
            const long p = p0 + pl;
            const float4 v = ld4(in + (p / C) * in_bstride + (p % C) * (long)HW + e);
            float *d = sm + pl * pe + (r + 1) * pitch + c + 1;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();
    const int total = np * HWo;
    T *dst = out + p0 * HWo;
    for (int q = threadIdx.x * V; q < total; q += blockDim.x * V) {
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int idx = min(q + e, total - 1);
            const int pl = idx / HWo, rem = idx - pl * HWo, i = rem / Wo, j = rem - i * Wo;
            const float *r0 = sm + pl * pe + (2 * i) * pitch + 2 * j;        // = input (2i-1, 2j-1) in the bordered layout
            float acc = k.k[0][0] * r0[0];
            acc = fmaf(k.k[0][1], r0[1], acc);
            acc = fmaf(k.k[0][2], r0[2], acc);
            acc = fmaf(k.k[1][0], r0[pitch], acc);
            acc = fmaf(k.k[1][1], r0[pitch + 1], acc);
            acc = fmaf(k.k[1][2], r0[pitch + 2], acc);
            acc = fmaf(k.k[2][0], r0[2 * pitch], acc);
            acc = fmaf(k.k[2][1], r0[2 * pitch + 1], acc);
            acc = fmaf(k.k[2][2], r0[2 * pitch + 2], acc);
            o[e] = acc;
        }
        if (q + V <= total) {
            store_vec<T>(dst + q, o);
        } else {
            for (int e = 0; q + e < total; ++e) st1(dst + q + e, o[e]);
        }
    }
}

// ---- planes that fit a warp: one thread per 4-column row strip, neighbours by warp shuffle ------------
// No shared memory and no row loop: a lane loads its strip(s) with one or two 128-bit loads that are
// contiguous across the warp, takes the row below / above and the column next to its strip from the lanes
// that hold them (a plane occupies L = rows * S consecutive lanes, L divides 32, so every neighbour inside
// the plane is inside the warp and plane borders are compile-time lane masks), and writes one contiguous
// piece of the output.  up: lane = (plane, input row r, strip s) -> output rows 2r, 2r+1, columns 8s..8s+7.
// down: lane = (plane, output row i, strip s) loads input rows 2i, 2i+1 -> outputs (i, 2s), (i, 2s+1).
template <typename TI, typename TO, int H, int S>
__global__ void __launch_bounds__(256)
up3_warp_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int C, long out_bstride,
                const __grid_constant__ Taps3 k)
{
    constexpr int W = 4 * S, L = H * S, HW = H * W, W2 = 2 * W;
    static_assert(32 % L == 0, "a plane must occupy a divisor of a warp");
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned p = g / L;
    const int within = (int)(g % L), r = within / S, s = within % S;
    const bool valid = p < (unsigned long)planes;
    if (!valid) p = (unsigned)planes - 1;                     // idle lanes shadow the last plane, stores off
    const float4 c = ld4(in + (long)p * HW + r * W + 4 * s);
    const unsigned full = 0xffffffffu;
    float xa[5] = {c.x, c.y, c.z, c.w, __shfl_down_sync(full, c.x, 1)};
    float xb[5] = {__shfl_down_sync(full, c.x, S), __shfl_down_sync(full, c.y, S), __shfl_down_sync(full, c.z, S),
                   __shfl_down_sync(full, c.w, S), __shfl_down_sync(full, c.x, S + 1)};
    if (s == S - 1) { xa[4] = 0.f; xb[4] = 0.f; }
    if (r == H - 1) {
#pragma unroll
        for (int q = 0; q < 5; ++q) xb[q] = 0.f;
    }
    float e[8], o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        e[2 * q] = k.k[1][1] * xa[q];
        e[2 * q + 1] = fmaf(k.k[1][2], xa[q + 1], k.k[1][0] * xa[q]);
        o[2 * q] = fmaf(k.k[2][1], xb[q], k.k[0][1] * xa[q]);
        float t = k.k[0][0] * xa[q];
        t = fmaf(k.k[0][2], xa[q + 1], t);
        t = fmaf(k.k[2][0], xb[q], t);
        o[2 * q + 1] = fmaf(k.k[2][2], xb[q + 1], t);
    }
    if (!valid) return;
    TO *dst = out + strided_base(p, C, out_bstride, 4L * HW) + (2 * r) * W2 + 8 * s;
    st8(dst, e);
    st8(dst + W2, o);
}

template <typename T>
__device__ __forceinline__ void store2(T *p, float a, float b);
template <>
__device__ __forceinline__ void store2<float>(float *p, float a, float b)
{
    __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b));
}
template <>
__device__ __forceinline__ void store2<bf16>(bf16 *p, float a, float b)
{
    __stcs(reinterpret_cast<unsigned int *>(p), pack_bf16x2(a, b));
}

template <typename T, int HO, int S>
__global__ void __launch_bounds__(256)
down3_warp_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int C, long in_bstride,
                  const __grid_constant__ Taps3 k)
{
    constexpr int W = 4 * S, H = 2 * HO, L = HO * S, HW = H * W, WO = 2 * S;
    static_assert(32 % L == 0, "a plane must occupy a divisor of a warp");
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned p = g / L;
    const int within = (int)(g % L), i = within / S, s = within % S;
    const bool valid = p < (unsigned long)planes;
    if (!valid) p = (unsigned)planes - 1;
    const T *src = in + strided_base(p, C, in_bstride, HW) + (2 * i) * W + 4 * s;
    const float4 a = ld4(src), b = ld4(src + W);
    const unsigned full = 0xffffffffu;
    float up[5], ra[5], rb[5];                                 // columns 4s-1 .. 4s+3 of rows 2i-1, 2i, 2i+1
    up[1] = __shfl_up_sync(full, b.x, S); up[2] = __shfl_up_sync(full, b.y, S);
    up[3] = __shfl_up_sync(full, b.z, S); up[4] = __shfl_up_sync(full, b.w, S);
    up[0] = __shfl_up_sync(full, b.w, S + 1);
    ra[0] = __shfl_up_sync(full, a.w, 1); rb[0] = __shfl_up_sync(full, b.w, 1);
    ra[1] = a.x; ra[2] = a.y; ra[3] = a.z; ra[4] = a.w;
    rb[1] = b.x; rb[2] = b.y; rb[3] = b.z; rb[4] = b.w;
    if (s == 0) { up[0] = 0.f; ra[0] = 0.f; rb[0] = 0.f; }
    if (i == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q) up[q] = 0.f;
    }
    float o[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float acc = k.k[0][0] * up[2 * q];
        acc = fmaf(k.k[0][1], up[2 * q + 1], acc);
        acc = fmaf(k.k[0][2], up[2 * q + 2], acc);
        acc = fmaf(k.k[1][0], ra[2 * q], acc);
        acc = fmaf(k.k[1][1], ra[2 * q + 1], acc);
        acc = fmaf(k.k[1][2], ra[2 * q + 2], acc);
        acc = fmaf(k.k[2][0], rb[2 * q], acc);
        acc = fmaf(k.k[2][1], rb[2 * q + 1], acc);
        acc = fmaf(k.k[2][2], rb[2 * q + 2], acc);
        o[q] = acc;
    }
    if (!valid) return;
    store2<T>(out + (long)p * (HO * WO) + i * WO + 2 * s, o[0], o[1]);
}

// ---- down-like without a row loop ("flat"): one thread per (plane, output row, V output columns) -----------
// The strip kernel of afr_n3.cu walks down the rows with the odd row carried, which exposes one memory latency per
// output row; short planes are latency-bound there (ncu: long_scoreboard).  Here every thread issues ALL its loads up front -- three input rows, 2V columns as 128-bit
// loads plus the left halo -- and the grid is (planes x rows x strips) threads, so the loads in flight are bounded
// by occupancy alone.  Row 2i-1 is read by two threads (L1 / L2 hit): 1.5x the load instructions, the same DRAM bytes.
template <typename T, int V, int ROWS, bool kPow2>
__global__ void __launch_bounds__(256)
down3_flat_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int H, int W, int Ho, int Wo, int strips,
                  int lg_strips, int lg_per_plane, int C, long in_bstride, const __grid_constant__ Taps3 k)
{
    // 32-bit index arithmetic (the launcher guarantees planes * strips * rgroups < 2^32); with power-of-two strip and
    // row-group counts (every power-of-two plane) the thread -> (plane, row group, strip) map is two shifts.  The first
    // version of this kernel spent 2/3 of its instructions on 64-bit index and address arithmetic (ncu: 267
    // instructions per thread for 36 FMAs), which is what a bf16 tensor -- half the bytes per element -- ran into.
    const unsigned rgroups = (unsigned)(Ho + ROWS - 1) / ROWS;     // ROWS consecutive output rows per thread
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned pu, rg;
    int s;
    if (kPow2) {
        pu = idx >> lg_per_plane;
        const unsigned rem = idx & ((1u << lg_per_plane) - 1u);
        rg = rem >> lg_strips;
        s = (int)(rem & ((1u << lg_strips) - 1u));
    } else {
        const unsigned per_plane = (unsigned)strips * rgroups;
        pu = idx / per_plane;
        const unsigned rem = idx - pu * per_plane;
        rg = rem / (unsigned)strips;
        s = (int)(rem - rg * (unsigned)strips);
    }
    if (pu >= (unsigned long)planes) return;
    const int i0 = (int)rg * ROWS;
    const long p = pu;
    const int j = V * s;
    const T *plane = in + strided_base(pu, C, in_bstride, (long)H * W);
    const bool has_l = (j > 0);
    const int off0 = (2 * i0 - 1) * W + 2 * j;                     // a plane has < 2^30 elements: int offsets
    float v[2 * ROWS + 1][2 * V + 1];
#pragma unroll
    for (int r = 0; r < 2 * ROWS + 1; ++r) {
        const int row = 2 * i0 - 1 + r;
        if ((unsigned)row >= (unsigned)H) {
#pragma unroll
            for (int c = 0; c < 2 * V + 1; ++c) v[r][c] = 0.f;
        } else {
            const T *q = plane + (off0 + r * W);
#pragma unroll
            for (int h = 0; h < V / 4; ++h) {
                float w[8];
                ld8(q + 8 * h, w);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[r][8 * h + c + 1] = w[c];
            }
            v[r][0] = has_l ? ld1(q - 1) : 0.f;
        }
    }
    T *dst = out + p * (long)Ho * Wo + (long)i0 * Wo + j;
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        if (i0 + rr >= Ho) break;
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float acc = k.k[0][0] * v[2 * rr][2 * q];
            acc = fmaf(k.k[0][1], v[2 * rr][2 * q + 1], acc);
            acc = fmaf(k.k[0][2], v[2 * rr][2 * q + 2], acc);
            acc = fmaf(k.k[1][0], v[2 * rr + 1][2 * q], acc);
            acc = fmaf(k.k[1][1], v[2 * rr + 1][2 * q + 1], acc);
            acc = fmaf(k.k[1][2], v[2 * rr + 1][2 * q + 2], acc);
            acc = fmaf(k.k[2][0], v[2 * rr + 2][2 * q], acc);
            acc = fmaf(k.k[2][1], v[2 * rr + 2][2 * q + 1], acc);
            acc = fmaf(k.k[2][2], v[2 * rr + 2][2 * q + 2], acc);
            o[q] = acc;
        }
        if (V == 8) {
            float o8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o8[q] = o[q % V];
            st8(dst + (long)rr * Wo, o8);
        } else {
            st4(dst + (long)rr * Wo, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// ---- up-like without a row loop: one thread per (plane, input row, VI input columns) -> 2 x 2 VI outputs -----
// Same idea as down3_flat_kernel: both input rows (i and i+1; the latter is re-read by the thread below, an L1 / L2
// hit) are requested up front and the grid supplies the parallelism.  VI = 4 stores 8 outputs per row: one 16-byte
// store for bf16, TWO for fp32 -- each of which then covers only every other 16 bytes across the warp (half sectors).
// VI = 2 (fp32 outputs) makes the fp32 store one float4 per lane, contiguous across the warp.
template <int VI> __device__ __forceinline__ void ld_cols(const float *p, float (&x)[VI + 1])
{
    if constexpr (VI == 4) {
        const float4 v = ld4(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        x[0] = v.x; x[1] = v.y;
    }
}
template <int VI> __device__ __forceinline__ void ld_cols(const bf16 *p, float (&x)[VI + 1])
{
    if constexpr (VI == 4) {
        const float4 v = ld4(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
        const uint32_t r = __ldg(reinterpret_cast<const uint32_t *>(p));
        x[0] = __uint_as_float(r << 16); x[1] = __uint_as_float(r & 0xffff0000u);
    }
}

template <typename TI, typename TO, int VI>
__global__ void __launch_bounds__(256)
up3_flat_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int H, int W, int strips, int C,
                long out_bstride, const __grid_constant__ Taps3 k)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned per_plane = (unsigned)strips * (unsigned)H;
    const unsigned pu = idx / per_plane, rem = idx - pu * per_plane;
    if (pu >= (unsigned long)planes) return;
    const unsigned i = rem / (unsigned)strips;
    const int s = (int)(rem - i * (unsigned)strips), j = VI * s;
    const TI *src = in + (long)pu * H * W + (long)i * W + j;
    const bool has_r = (j + VI < W), has_b = ((int)i + 1 < H);
    float xa[VI + 1], xb[VI + 1];
#pragma unroll
    for (int c = 0; c <= VI; ++c) xa[c] = xb[c] = 0.f;
    ld_cols<VI>(src, xa);
    if (has_b) ld_cols<VI>(src + W, xb);
    if (has_r) xa[VI] = ld1(src + VI);
    if (has_r && has_b) xb[VI] = ld1(src + W + VI);
    float e[2 * VI], o[2 * VI];
#pragma unroll
    for (int c = 0; c < VI; ++c) {
        e[2 * c] = k.k[1][1] * xa[c];
        e[2 * c + 1] = fmaf(k.k[1][2], xa[c + 1], k.k[1][0] * xa[c]);
        o[2 * c] = fmaf(k.k[2][1], xb[c], k.k[0][1] * xa[c]);
        float t = k.k[0][0] * xa[c];
        t = fmaf(k.k[0][2], xa[c + 1], t);
        t = fmaf(k.k[2][0], xb[c], t);
        o[2 * c + 1] = fmaf(k.k[2][2], xb[c + 1], t);
    }
    const int W2 = 2 * W;
    TO *dst = out + strided_base(pu, C, out_bstride, 4L * H * W) + (long)(2 * i) * W2 + 2 * j;
    if constexpr (VI == 4) {
        st8(dst, e);
        st8(dst + W2, o);
    } else {
        st4(dst, make_float4(e[0], e[1], e[2], e[3]));
        st4(dst + W2, make_float4(o[0], o[1], o[2], o[3]));
    }
}

int group_planes(int plane_floats, long planes)
{
    // ~24 KB of staged planes per CTA, a multiple of 8 planes (keeps every CTA's output block 16-byte aligned),
    // but never so many that the grid drops below ~4 CTAs per SM
    int pg = (24 * 1024 / 4) / plane_floats;
    pg = pg < 8 ? 8 : pg / 8 * 8;
    while (pg > 8 && (planes + pg - 1) / pg < 4 * 148) pg -= 8;
    return pg;
}

int small_max_hw()      // AFR_SMALL_MAX_HW: largest input plane (elements) that takes these kernels; 0 disables
{
    static const int v = []() { const char *e = getenv("AFR_SMALL_MAX_HW"); return e ? atoi(e) : 256; }();
    return v;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

static bool warp_disabled()
{
    static const bool v = []() { const char *e = getenv("AFR_NO_WARP_PLANE"); return e && atoi(e) != 0; }();
    return v;
}

// up: input planes 4x4, 8x8, 4x8, 2x4 ... (H * W / 4 lanes per plane must divide a warp)
bool warp_up_shape(int H, int W)
{
    if (warp_disabled()) return false;
    return (H == 4 && W == 4) || (H == 8 && W == 8) || (H == 8 && W == 4) || (H == 4 && W == 8) || (H == 2 && W == 4) ||
           (H == 2 && W == 8) || (H == 16 && W == 8) || (H == 16 && W == 4);
}
// down: input planes 4x4, 8x8, 16x16, ... (H even; (H/2) * W / 4 lanes per plane)
bool warp_down_shape(int H, int W)
{
    if (warp_disabled()) return false;
    return (H == 4 && W == 4) || (H == 8 && W == 8) || (H == 16 && W == 16) || (H == 8 && W == 4) || (H == 4 && W == 8) ||
           (H == 16 && W == 8) || (H == 8 && W == 16) || (H == 2 && W == 4) || (H == 32 && W == 4) || (H == 16 && W == 4);
}

template <typename TI, typename TO>
static bool launch_up_warp(const void *in, void *out, long planes, int C, long obs, int H, int W, const Taps3 &k, cudaStream_t s)
{
#define AFR_UW(HH, SS)                                                                                          \
    if (H == HH && W == 4 * SS) {                                                                               \
        const long threads = planes * (HH * SS);                                                                \
        up3_warp_kernel<TI, TO, HH, SS><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((const TI *)in, (TO *)out, planes, C, obs, k); \
        return true;                                                                                            \
    }
    AFR_UW(4, 1) AFR_UW(8, 2) AFR_UW(8, 1) AFR_UW(4, 2) AFR_UW(2, 1) AFR_UW(2, 2) AFR_UW(16, 2) AFR_UW(16, 1)
#undef AFR_UW
    return false;
}

template <typename T>
static bool launch_down_warp(const void *in, void *out, long planes, int C, long ibs, int H, int W, const Taps3 &k, cudaStream_t s)
{
#define AFR_DW(HO, SS)                                                                                          \
    if (H == 2 * HO && W == 4 * SS) {                                                                           \
        const long threads = planes * (HO * SS);                                                                \
        down3_warp_kernel<T, HO, SS><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((const T *)in, (T *)out, planes, C, ibs, k); \
        return true;                                                                                            \
    }
    AFR_DW(2, 1) AFR_DW(4, 2) AFR_DW(8, 4) AFR_DW(4, 1) AFR_DW(2, 2) AFR_DW(8, 2) AFR_DW(4, 4) AFR_DW(1, 1) AFR_DW(16, 1) AFR_DW(8, 1)
#undef AFR_DW
    return false;
}

bool small_up_supported(int H, int W, const void *in, const void *out, long out_bstride, int out_dtype)
{
    const long es = out_dtype == AFR_F32 ? 4 : 2;
    return H >= 1 && W >= 4 && (W % 4) == 0 && H * W <= small_max_hw() && H * W <= 1024 && aligned16(in) &&
           aligned16(out) && (out_bstride * es) % 16 == 0;
}

cudaError_t small_up_like(const void *in, void *out, long planes, int C, long out_bstride, int H, int W,
                          const Taps3 &k, int in_dtype, int out_dtype, cudaStream_t s)
{
    if (planes * 32 > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (warp_up_shape(H, W)) {
        const int Cw = out_bstride == 4L * H * W * C ? 0 : C;     // 0: dense output, no per-thread division
        bool ok;
        if (in_dtype == AFR_F32 && out_dtype == AFR_F32) ok = launch_up_warp<float, float>(in, out, planes, Cw, out_bstride, H, W, k, s);
        else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) ok = launch_up_warp<bf16, bf16>(in, out, planes, Cw, out_bstride, H, W, k, s);
        else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) ok = launch_up_warp<bf16, float>(in, out, planes, Cw, out_bstride, H, W, k, s);
        else ok = launch_up_warp<float, bf16>(in, out, planes, Cw, out_bstride, H, W, k, s);
        if (ok) return cudaGetLastError();
    }
    const int pe = (H + 1) * (W + 1);
    const int pg = group_planes(pe, planes);
    const long grid = (planes + pg - 1) / pg;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)pg * pe * sizeof(float);
#define AFR_UPG(TI, TO) up3_group_kernel<TI, TO><<<(unsigned)grid, 256, smem, s>>>((const TI *)in, (TO *)out, planes, C, out_bstride, H, W, pg, k)
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32) AFR_UPG(float, float);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) AFR_UPG(bf16, bf16);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) AFR_UPG(bf16, float);
    else AFR_UPG(float, bf16);
#undef AFR_UPG
    return cudaGetLastError();
}

bool small_down_supported(int H, int W, const void *in, const void *out, long in_bstride, int dtype)
{
    const long es = dtype == AFR_F32 ? 4 : 2;
    (void)es;
    return H >= 1 && W >= 4 && (W % 4) == 0 && H * W <= small_max_hw() && H * W <= 1024 && aligned16(in) &&
           aligned16(out) && (in_bstride % 4) == 0;            // planes are read in groups of 4 elements
}

cudaError_t small_down_like(const void *in, void *out, long planes, int C, long in_bstride, int H, int W,
                            const Taps3 &k, int dtype, cudaStream_t s)
{
    if (planes * 32 > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (warp_down_shape(H, W)) {
        const int Cw = in_bstride == (long)H * W * C ? 0 : C;
        const bool ok = dtype == AFR_F32 ? launch_down_warp<float>(in, out, planes, Cw, in_bstride, H, W, k, s)
                                         : launch_down_warp<bf16>(in, out, planes, Cw, in_bstride, H, W, k, s);
        if (ok) return cudaGetLastError();
    }
    const int pe = (H + 2) * (W + 2);
    const int pg = group_planes(pe, planes);
    const long grid = (planes + pg - 1) / pg;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)pg * pe * sizeof(float);
    if (dtype == AFR_F32)
        down3_group_kernel<float><<<(unsigned)grid, 256, smem, s>>>((const float *)in, (float *)out, planes, C, in_bstride, H, W, pg, k);
    else
        down3_group_kernel<bf16><<<(unsigned)grid, 256, smem, s>>>((const bf16 *)in, (bf16 *)out, planes, C, in_bstride, H, W, pg, k);
    return cudaGetLastError();
}

}  // namespace afr

namespace afr {

// AFR_DOWN_FLAT: 0 = never, 1 = fp32 always, bf16 on planes up to 32 x 32 (default), 2 = always.
// Measured (B200, L2 left clean before each repetition): fp32 0.86 -> 1.00 on 32x32 planes and 0.98 -> 1.06 of the
// copy-measured HBM peak on 64x64 (a read-dominated kernel can exceed a copy's rate); bf16 0.52 -> 0.57 on 32x32 but
// 0.69 -> 0.61-0.65 on 64x64, where the strip kernel's 8-output variant amortises its conversions better.
static int down_flat_mode()
{
    static const int v = []() { const char *e = getenv("AFR_DOWN_FLAT"); return e ? atoi(e) : 1; }();
    return v;
}

// same shape / alignment conditions as the strip kernel (n3_down_supported): W % 8 == 0, 8-element aligned input
bool flat_down_wanted(int H, int W, int dtype)
{
    const int m = down_flat_mode();
    return m == 2 || (m == 1 && (dtype == AFR_F32 || (long)H * W <= 1024));
}

// AFR_UP_FLAT: 0 = never, 1 = every plane size when the OUTPUT is fp32 (default), 2 = always.  Measured (B200, fp32, of
// the HBM peak, 4x4 ... 256x256 planes): 0.68 / 0.73 / 0.87 / 0.87 / 0.92 / 0.89 / 0.85 with the warp-shuffle, 4-column
// flat and row-walking strip kernels -> 0.92 / 0.97 / 0.99 / 0.99 / 0.99 / 0.96 / 0.94 with two input columns per thread:
// the 8 fp32 outputs per row of a 4-column thread are two 16-byte stores that each cover only every other 16 bytes
// across the warp (half sectors); one float4 per lane is contiguous.  bf16 outputs (8 outputs = one 16-byte store
// already): 2 columns 0.56, 4 columns -0 .. -4 % against the strip / warp kernels, so they stay there.
#ifndef AFR_UP_FLAT_F32_COLS
#define AFR_UP_FLAT_F32_COLS 2
#endif

static int up_flat_mode()
{
    static const int v = []() { const char *e = getenv("AFR_UP_FLAT"); return e ? atoi(e) : 1; }();
    return v;
}

bool flat_up_wanted(int H, int W, int in_dtype, int out_dtype)
{
    (void)H; (void)W; (void)in_dtype;
    const int m = up_flat_mode();
    return m == 2 || (m == 1 && out_dtype == AFR_F32);
}

cudaError_t flat_up_like(const void *in, void *out, long planes, int C, long out_bstride, int H, int W, const Taps3 &k,
                         int in_dtype, int out_dtype, cudaStream_t s)
{
    // AFR_UP_FLAT_COLS = 2 | 4: input columns per thread (default: 2 for fp32 outputs, 4 for bf16 outputs)
    static const int want_cols = []() { const char *e = getenv("AFR_UP_FLAT_COLS"); return e ? atoi(e) : 0; }();
    const int vi = (want_cols == 2 || want_cols == 4) ? want_cols : (out_dtype == AFR_F32 ? AFR_UP_FLAT_F32_COLS : 4);
    const int strips = W / vi;
    const long total = planes * (long)strips * H;
    const long grid = (total + 255) / 256;
    if (total >= 0xffffff00L || planes > 0x7fffffffL) return cudaErrorInvalidConfiguration;
#define AFR_UF2(TI, TO, VI) up3_flat_kernel<TI, TO, VI><<<(unsigned)grid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, H, W, strips, C, out_bstride, k)
#define AFR_UF(TI, TO) do { if (vi == 2) AFR_UF2(TI, TO, 2); else AFR_UF2(TI, TO, 4); } while (0)
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32) AFR_UF(float, float);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) AFR_UF(bf16, bf16);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) AFR_UF(bf16, float);
    else AFR_UF(float, bf16);
#undef AFR_UF
#undef AFR_UF2
    return cudaGetLastError();
}

cudaError_t flat_down_like(const void *in, void *out, long planes, int C, long in_bstride, int H, int W, const Taps3 &k,
                           int dtype, cudaStream_t s)
{
    const int Ho = (H + 1) / 2, Wo = W / 2;
    static const bool want_wide = []() { const char *e = getenv("AFR_DOWN_FLAT_WIDE"); return e && atoi(e) != 0; }();
    static const int want_rows = []() { const char *e = getenv("AFR_DOWN_FLAT_ROWS"); return e ? atoi(e) : 0; }();
    const bool wide = want_wide && dtype == AFR_BF16 && (Wo % 8) == 0 && (reinterpret_cast<uintptr_t>(in) % 16) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) % 16) == 0;
    const int V = wide ? 8 : 4;
    const int strips = Wo / V;
    int rows = want_rows > 0 ? want_rows : 1;
    if (rows != 1 && rows != 2) rows = 2;
    const long total = planes * (long)strips * ((Ho + rows - 1) / rows);
    const long grid = (total + 255) / 256;
    if (total >= 0xffffff00L || planes > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const int rgroups = (Ho + rows - 1) / rows;
    auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    const bool pow2 = (strips & (strips - 1)) == 0 && (rgroups & (rgroups - 1)) == 0;
    const int lgs = lg2(strips), lgp = lg2(strips * rgroups);
#define AFR_FLAT2(T, VV, RR, PP) down3_flat_kernel<T, VV, RR, PP><<<(unsigned)grid, 256, 0, s>>>((const T *)in, (T *)out, planes, H, W, Ho, Wo, strips, lgs, lgp, C, in_bstride, k)
#define AFR_FLAT(T, VV, RR) do { if (pow2) AFR_FLAT2(T, VV, RR, true); else AFR_FLAT2(T, VV, RR, false); } while (0)
    if (dtype == AFR_F32) { if (rows == 2) AFR_FLAT(float, 4, 2); else AFR_FLAT(float, 4, 1); }
    else if (wide) { if (rows == 2) AFR_FLAT(bf16, 8, 2); else AFR_FLAT(bf16, 8, 1); }
    else { if (rows == 2) AFR_FLAT(bf16, 4, 2); else AFR_FLAT(bf16, 4, 1); }
#undef AFR_FLAT2
#undef AFR_FLAT
    return cudaGetLastError();
}

}  // namespace afr
