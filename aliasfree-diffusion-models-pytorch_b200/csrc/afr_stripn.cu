// Fused filtered GELU for filter sizes other than 3 (compile-time N in {2,4,5,6,7,8}; the reference's
// circularLowpassKernel defaults to N = 6, modules/filtrs.py:20): the register-strip formulation of
// afr_n3.cu generalised to N x N taps with the 'same' padding split of even N.
//
// A thread owns 4 adjacent output columns j..j+3 of one plane and walks down the rows.  Step i
// produces the two mid rows Y = 2i + C0 and Y + 1 (C0 = N - 2 - padB) -- the last two rows output
// row i needs -- at the thread's OWN mid columns X = 2j .. 2j+7:
//   u[Y][X] = sum_{a,b} kU[a][b] x[(Y+a-padU)/2][(X+b-padU)/2]     over even (Y+a-padU), (X+b-padU)
// from a register window of input rows (one new row per step), applies the activation (16 GELUs per
// step = 4 per output, packed FFMA2 pairs), fetches the padB left and N-2-padB right halo columns
// of every mid row from lanes -1 / +1 by shuffle, and adds the row's contribution to the K =
// floor((N-1)/2)+1 output rows it belongs to (K x 4 accumulators carried in registers):
//   acc[r][q] += sum_b kB[a][b] mid[Y][2(j+q)+b-padB],   a = Y - 2r + padB.
// Row i is complete after step i and is stored; a strip starts K-1 steps early (nothing stored) to
// fill the accumulators.  Mid values outside [0,2H) x [0,2W) are forced to zero (the reference
// zero-pads the GELU OUTPUT).
//
// Thread mapping.  When the strips of a plane row divide a warp (W = 4 .. 128, powers of two) they are
// packed densely and every neighbour strip is the neighbour lane.  Otherwise the strips of ALL planes
// form one stream that is cut into tiles of 30: a warp computes one tile plus the strip before and
// after it (lanes 0 and 31, "ghosts": they produce their mid columns for the neighbours' halos and
// store nothing), so 30 of 32 lanes do useful work for any width and no lane ever has to rebuild a
// halo column on its own.  A halo that would cross a plane edge is zero by definition.
//
// Adjoint: mid = gelu'(up(x; kU)) * up(dy; kG), out = down(mid; kB) with kG = flip(k_down) (pad ph),
// kB = flip(k_up) (pad ph); kU arrives scaled by kappa (afr_common.cuh).
#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

__host__ __device__ constexpr int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int mod2(int v) { return ((v % 2) + 2) % 2; }

// Index geometry of one up-like stage (N taps, low padding PAD) feeding mid rows 2i+C0+e, e = 0/1,
// and own mid columns 2j+m, m = 0..7.  Row / column offsets are relative to i / j and exact
// (the parity classes pa / pb select the taps that hit a non-zero sample of the zero-stuffed grid).
template <int N, int PAD, int C0>
struct UpGeom {
    static __host__ __device__ constexpr int pa(int e) { return mod2(C0 + e - PAD); }
    static __host__ __device__ constexpr int pb(int m) { return mod2(m - PAD); }
    static __host__ __device__ constexpr int roff(int e, int a) { return (C0 + e + a - PAD) / 2; }
    static __host__ __device__ constexpr int coff(int m, int b) { return (m + b - PAD) / 2; }
    static __host__ __device__ constexpr int rlo()
    {
        int v = 1 << 20;
        for (int e = 0; e < 2; ++e)
            for (int a = pa(e); a < N; a += 2) v = imin(v, roff(e, a));
        return v;
    }
    static __host__ __device__ constexpr int rhi()
    {
        int v = -(1 << 20);
        for (int e = 0; e < 2; ++e)
            for (int a = pa(e); a < N; a += 2) v = imax(v, roff(e, a));
        return v;
    }
    static __host__ __device__ constexpr int clo()
    {
        int v = 1 << 20;
        for (int m = 0; m < 8; ++m)
            for (int b = pb(m); b < N; b += 2) v = imin(v, coff(m, b));
        return v;
    }
    static __host__ __device__ constexpr int chi()
    {
        int v = -(1 << 20);
        for (int m = 0; m < 8; ++m)
            for (int b = pb(m); b < N; b += 2) v = imax(v, coff(m, b));
        return v;
    }
    static constexpr int NR = rhi() - rlo() + 1, NC = chi() - clo() + 1;
};

// one input row, columns j+CLO .. j+CHI (zero outside the plane), optional fused residual
template <typename T, bool kRes, int CLO, int CHI>
__device__ __forceinline__ void load_window_row(const T *__restrict__ xp, const T *__restrict__ rp, int row,
                                                int H, int W, int j, float (&v)[CHI - CLO + 1])
{
    static_assert(CLO <= 0 && CHI >= 3, "window must contain the strip's own columns");
    if (row < 0 || row >= H) {
#pragma unroll
        for (int c = 0; c < CHI - CLO + 1; ++c) v[c] = 0.f;
        return;
    }
    const long off = (long)row * W + j;
    float4 c4 = ld4(xp + off);
    if (kRes) {
        const float4 q4 = ld4(rp + off);
        c4.x += q4.x; c4.y += q4.y; c4.z += q4.z; c4.w += q4.w;
    }
    v[0 - CLO] = c4.x; v[1 - CLO] = c4.y; v[2 - CLO] = c4.z; v[3 - CLO] = c4.w;
#pragma unroll
    for (int c = CLO; c <= CHI; ++c) {
        if (c >= 0 && c < 4) continue;
        float t = 0.f;
        if (j + c >= 0 && j + c < W) {
            t = ld1(xp + off + c);
            if (kRes) t += ld1(rp + off + c);
        }
        v[c - CLO] = t;
    }
}

// up-like value of one stage at own column m of mid row parity e, from the register window
template <class G, int N, int E, int M>
__device__ __forceinline__ float up_from_window(const float (&w)[G::NR][G::NC], const TapsG &t)
{
    float u = 0.f;
#pragma unroll
    for (int a = G::pa(E); a < N; a += 2)
#pragma unroll
        for (int b = G::pb(M); b < N; b += 2)
            u = fmaf(t.k[a * N + b], w[G::roff(E, a) - G::rlo()][G::coff(M, b) - G::clo()], u);
    return u;
}

template <class G, int N, int E, int M = 0>
struct UpRow {
    static __device__ __forceinline__ void run(const float (&w)[G::NR][G::NC], const TapsG &t, float (&u)[8])
    {
        u[M] = up_from_window<G, N, E, M>(w, t);
        UpRow<G, N, E, M + 1>::run(w, t, u);
    }
};
template <class G, int N, int E>
struct UpRow<G, N, E, 8> {
    static __device__ __forceinline__ void run(const float (&)[G::NR][G::NC], const TapsG &, float (&)[8]) {}
};

template <typename T, bool kBwd, bool kRes, int N>
__global__ void __launch_bounds__(128)
fgelu_strip_kernel(const T *__restrict__ x, const T *__restrict__ res, const T *__restrict__ dy,
                   T *__restrict__ out, long planes, int H, int W, int strips, int nseg, int R, int tiles,
                   const __grid_constant__ TapsG tU, const __grid_constant__ TapsG tG,
                   const __grid_constant__ TapsG tB)
{
    constexpr int PL = (N - 1) / 2, PH = N - 1 - PL;
    constexpr int PU = PL, PG = PH, PB = kBwd ? PH : PL;
    constexpr int C0 = N - 2 - PB;                 // step i makes mid rows 2i+C0, 2i+C0+1
    constexpr int K = (N - 1) / 2 + 1;             // output rows a mid row contributes to
    constexpr int HL = PB, HR = imax(0, N - 2 - PB);
    constexpr int NM = imax(N + 6, HL + 8);        // mid columns 2j-PB .. , own ones at [HL, HL+8)
    using GU = UpGeom<N, PU, C0>;
    using GG = UpGeom<N, PG, C0>;

    // tiles == 0: strips of a plane row are packed densely over the lanes (strips divides 32);
    // tiles > 0: tile t of the strip stream (plane-major) sits in lanes 1..30 of one warp
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned seg_threads = tiles ? (unsigned)tiles * 32u : (unsigned)planes * (unsigned)strips;
    bool valid = idx < seg_threads * (unsigned)nseg;
    if (!valid) idx = 0;                           // idle lanes shadow thread 0, stores off
    const int seg = (int)(idx / seg_threads);
    const unsigned rem = idx - (unsigned)seg * seg_threads;
    long g = tiles ? (long)(rem >> 5) * 30 + lane - 1 : (long)rem;     // position in the strip stream
    const bool col_ok = (g >= 0) && (g < planes * strips);            // stream ends: mid == 0
    if (tiles) valid = valid && col_ok && lane >= 1 && lane <= 30;
    if (!col_ok) g = 0;
    const long pu = g / strips;
    const int s = (int)(g - pu * strips);
    const int j = 4 * s, i0 = seg * R, i1 = min(H, i0 + R);
    const long pbase = (long)pu * H * W;
    const T *xp = x + pbase, *rp = kRes ? res + pbase : nullptr, *dp = kBwd ? dy + pbase : nullptr;
    T *op = out + pbase + j;
    const bool first = (s == 0), last = (s == strips - 1);   // halo beyond the plane edge is zero

    float xw[GU::NR][GU::NC];
    float dw[GG::NR][GG::NC];                      // adjoint only (never materialised otherwise)
    float acc[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q] = 0.f;

    const int ifirst = i0 - (K - 1);
    // rows ifirst+rlo .. ifirst+rhi-1 (the step loads the last row of its window itself)
#pragma unroll
    for (int r = 1; r < GU::NR; ++r)
        load_window_row<T, kRes, GU::clo(), GU::chi()>(xp, rp, ifirst + GU::rlo() + r - 1, H, W, j, xw[r]);
    if constexpr (kBwd) {
#pragma unroll
        for (int r = 1; r < GG::NR; ++r)
            load_window_row<T, false, GG::clo(), GG::chi()>(dp, nullptr, ifirst + GG::rlo() + r - 1, H, W, j, dw[r]);
    }

    const int nsteps = R + K - 1;
    for (int st = 0; st < nsteps; ++st) {
        const int i = ifirst + st;
        // slide the input windows down by one row
#pragma unroll
        for (int r = 0; r + 1 < GU::NR; ++r)
#pragma unroll
            for (int c = 0; c < GU::NC; ++c) xw[r][c] = xw[r + 1][c];
        load_window_row<T, kRes, GU::clo(), GU::chi()>(xp, rp, i + GU::rhi(), H, W, j, xw[GU::NR - 1]);
        if constexpr (kBwd) {
#pragma unroll
            for (int r = 0; r + 1 < GG::NR; ++r)
#pragma unroll
                for (int c = 0; c < GG::NC; ++c) dw[r][c] = dw[r + 1][c];
            load_window_row<T, false, GG::clo(), GG::chi()>(dp, nullptr, i + GG::rhi(), H, W, j, dw[GG::NR - 1]);
        }
        float m[2][NM];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int Y = 2 * i + C0 + e;
            const bool row_in = (Y >= 0) && (Y < 2 * H) && col_ok;
            float u[8], g[8];
            if (e == 0) UpRow<GU, N, 0>::run(xw, tU, u); else UpRow<GU, N, 1>::run(xw, tU, u);
            if constexpr (kBwd) {
                if (e == 0) UpRow<GG, N, 0>::run(dw, tG, g); else UpRow<GG, N, 1>::run(dw, tG, g);
#pragma unroll
                for (int c = 0; c < 8; c += 2) gelu_grad_scaled_mul_x2(u[c], u[c + 1], g[c], g[c + 1]);
            } else {
#pragma unroll
                for (int c = 0; c < 8; c += 2) gelu_erf_x2(u[c], u[c + 1]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float v = kBwd ? g[c] : u[c];
                if (HL + c < NM) m[e][HL + c] = row_in ? v : 0.f;
            }
            // halo columns from the neighbour strips
#pragma unroll
            for (int t = 0; t < HL; ++t) {
                const float v = __shfl_up_sync(0xffffffffu, m[e][HL + 8 - HL + t], 1);
                m[e][t] = first ? 0.f : v;
            }
#pragma unroll
            for (int t = 0; t < HR; ++t) {
                const float v = __shfl_down_sync(0xffffffffu, m[e][HL + t], 1);
                if (HL + 8 + t < NM) m[e][HL + 8 + t] = last ? 0.f : v;
            }
        }
        // contributions of the two new mid rows to the K live output rows i .. i+K-1
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int a = N - 2 + e - 2 * k;
                if (a < 0 || a >= N) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < N; ++b)
                        acc[k][q] = fmaf(tB.k[a * N + b], m[e][2 * q + b], acc[k][q]);
            }
        if (valid && i >= i0 && i < i1) st4(op + (long)i * W, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
#pragma unroll
        for (int k = 0; k + 1 < K; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[k][q] = acc[k + 1][q];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[K - 1][q] = 0.f;
    }
}

// ---------------------------------------------------------------------------------
// standalone resamplers for compile-time N (custom_upsample / custom_downsample with N != 3)
// ---------------------------------------------------------------------------------
// up-like: thread = 4 input columns -> 8 output columns of the two output rows 2i, 2i+1, from the
// same register window of input rows (UpGeom with C0 = 0); 128-bit stores
template <typename TI, typename TO, int N, int PAD>
__global__ void __launch_bounds__(256)
upn_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int H, int W, int strips, int nseg,
           int R, const __grid_constant__ TapsG t)
{
    using G = UpGeom<N, PAD, 0>;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = 4 * s, i0 = seg * R, i1 = min(H, i0 + R);
    const TI *src = in + p * (long)H * W;
    TO *dst = out + p * 4L * H * W + 2 * j;
    const int W2 = 2 * W;
    float w[G::NR][G::NC];
#pragma unroll
    for (int r = 1; r < G::NR; ++r)
        load_window_row<TI, false, G::clo(), G::chi()>(src, nullptr, i0 + G::rlo() + r - 1, H, W, j, w[r]);
    for (int i = i0; i < i1; ++i) {
#pragma unroll
        for (int r = 0; r + 1 < G::NR; ++r)
#pragma unroll
            for (int c = 0; c < G::NC; ++c) w[r][c] = w[r + 1][c];
        // bf16 input: pull the row after next into L2 (+18 %; costs 15 % with fp32 input, so bf16 only)
        if (sizeof(TI) == 2 && i + G::rhi() + 2 < H) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (long)(i + G::rhi() + 2) * W + j));
        load_window_row<TI, false, G::clo(), G::chi()>(src, nullptr, i + G::rhi(), H, W, j, w[G::NR - 1]);
        float e[8], o[8];
        UpRow<G, N, 0>::run(w, t, e);
        UpRow<G, N, 1>::run(w, t, o);
        TO *r0 = dst + (long)(2 * i) * W2;
        st8(r0, e);
        st8(r0 + W2, o);
    }
}

// down-like: thread = 4 output columns; every input row 2i+C0+e is read once (8 own columns as
// 128-bit loads plus the halo scalars) and added into the K output rows it belongs to
template <typename T, int N, int PAD>
__global__ void __launch_bounds__(256)
downn_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int H, int W, int Ho, int Wo,
             int strips, int nseg, int R, const __grid_constant__ TapsG t)
{
    constexpr int C0 = N - 2 - PAD, K = (N - 1) / 2 + 1, HL = PAD, HR = imax(0, N - 2 - PAD);
    constexpr int NM = imax(N + 6, HL + 8);
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = 4 * s, i0 = seg * R, i1 = min(Ho, i0 + R);
    const T *plane = in + p * (long)H * W;
    T *dst = out + p * (long)Ho * Wo + j;
    float acc[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[k][q] = 0.f;
    for (int i = i0 - (K - 1); i < i1; ++i) {
        if (sizeof(T) == 2 && 2 * i + C0 + 5 < H && 2 * i + C0 + 4 >= 0) {   // bf16 only, as in down3_kernel
            const T *pf = plane + (long)(2 * i + C0 + 4) * W + 2 * j;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + W));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int Y = 2 * i + C0 + e;
            float m[NM];
            if (Y >= 0 && Y < H) {
                const T *row = plane + (long)Y * W + 2 * j;
                float own[8];
                ld8(row, own);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (HL + c < NM) m[HL + c] = own[c];
#pragma unroll
                for (int c = 0; c < HL; ++c) m[c] = (2 * j - HL + c >= 0) ? ld1(row - HL + c) : 0.f;
#pragma unroll
                for (int c = 0; c < HR; ++c)
                    if (HL + 8 + c < NM) m[HL + 8 + c] = (2 * j + 8 + c < W) ? ld1(row + 8 + c) : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < NM; ++c) m[c] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int a = N - 2 + e - 2 * k;
                if (a < 0 || a >= N) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < N; ++b) acc[k][q] = fmaf(t.k[a * N + b], m[2 * q + b], acc[k][q]);
            }
        }
        if (i >= i0) st4(dst + (long)i * Wo, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
#pragma unroll
        for (int k = 0; k + 1 < K; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[k][q] = acc[k + 1][q];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[K - 1][q] = 0.f;
    }
}

// ---- host side --------------------------------------------------------------------
static inline bool aligned_to(const void *p, size_t a) { return !p || (reinterpret_cast<uintptr_t>(p) % a) == 0; }

bool stripn_supported(int N, int H, int W, const void *x, const void *res, const void *dy, const void *out,
                      int dtype)
{
    if (!(N == 2 || N == 4 || N == 5 || N == 6 || N == 7 || N == 8)) return false;
    if (H < 1 || W < 4 || (W % 4) != 0) return false;
    const size_t a = 4 * (dtype == AFR_F32 ? 4 : 2);       // 4-element vector accesses
    return aligned_to(x, a) && aligned_to(res, a) && aligned_to(dy, a) && aligned_to(out, a);
}

template <typename T, bool kBwd, bool kRes, int N>
static cudaError_t launch_strip(const void *x, const void *res, const void *dy, void *out, long planes, int H,
                                int W, const TapsG &tU, const TapsG &tG, const TapsG &tB, cudaStream_t s)
{
    constexpr int K = (N - 1) / 2 + 1;
    const int strips = W / 4;
    const long nstrips = planes * strips;
    if (nstrips >= 0x7fffffffL / 2) { set_detail("tensor too large for the strip kernel's 32-bit indexing"); return cudaErrorInvalidConfiguration; }
    const int tiles = (strips <= 32 && 32 % strips == 0) ? 0 : (int)((nstrips + 29) / 30);
    const long seg_threads = tiles ? tiles * 32L : nstrips;
    // whole-height strips unless that leaves the GPU under-filled; every extra segment costs K-1 dry steps
    int nseg = 1;
    while (seg_threads * nseg < 148L * 1024 && (H + nseg - 1) / nseg > 4 * K) nseg *= 2;
    const int R = (H + nseg - 1) / nseg;
    nseg = (H + R - 1) / R;
    const long total = seg_threads * nseg;
    if (total >= 0x7fffffffL) { set_detail("tensor too large for the strip kernel's 32-bit indexing"); return cudaErrorInvalidConfiguration; }
    const int block = 128;
    fgelu_strip_kernel<T, kBwd, kRes, N><<<(unsigned)((total + block - 1) / block), block, 0, s>>>(
        (const T *)x, (const T *)res, (const T *)dy, (T *)out, planes, H, W, strips, nseg, R, tiles, tU, tG, tB);
    return cudaGetLastError();
}

cudaError_t stripn_fgelu(const void *x, const void *res, const void *dy, void *out, long planes, int H, int W,
                         const TapsG &tU_in, const TapsG &tG, const TapsG &tB, bool bwd, int dtype, cudaStream_t s)
{
    TapsG tU = tU_in;
    if (bwd)                                        // the adjoint evaluates gelu' in w = kappa * u
        for (int i = 0; i < tU.n * tU.n; ++i) tU.k[i] *= AFR_KAPPA;
#define AFR_SN(T, N)                                                                                         \
    case N:                                                                                                  \
        if (bwd) return res ? launch_strip<T, true, true, N>(x, res, dy, out, planes, H, W, tU, tG, tB, s)   \
                            : launch_strip<T, true, false, N>(x, res, dy, out, planes, H, W, tU, tG, tB, s); \
        return res ? launch_strip<T, false, true, N>(x, res, dy, out, planes, H, W, tU, tG, tB, s)           \
                   : launch_strip<T, false, false, N>(x, res, dy, out, planes, H, W, tU, tG, tB, s);
#define AFR_SN_ALL(T) switch (tU.n) { AFR_SN(T, 2) AFR_SN(T, 4) AFR_SN(T, 5) AFR_SN(T, 6) AFR_SN(T, 7) AFR_SN(T, 8) default: break; }
    if (dtype == AFR_F32) { AFR_SN_ALL(float) } else { AFR_SN_ALL(bf16) }
#undef AFR_SN_ALL
#undef AFR_SN
    set_detail("strip kernel: unsupported N=%d", tU.n);
    return cudaErrorInvalidConfiguration;
}

static inline int rows_per_segment(long planes, int strips, int H, int K, int *nseg_out)
{
    int nseg = 1;
    while (planes * strips * nseg < 148L * 2048 && (H + nseg - 1) / nseg > 4 * K + 4) nseg *= 2;
    const int R = (H + nseg - 1) / nseg;
    *nseg_out = (H + R - 1) / R;
    return R;
}

bool stripn_resample_supported(int N) { return N == 2 || N == 4 || N == 5 || N == 6 || N == 7 || N == 8; }

// in [planes,H,W] -> out [planes,2H,2W]; `t.pad` must be pl or ph of N
template <typename TI, typename TO>
static cudaError_t launch_upn(const void *in, void *out, long planes, int H, int W, const TapsG &t, cudaStream_t s)
{
    const int strips = W / 4;
    int nseg;
    const int R = rows_per_segment(planes, strips, H, 1, &nseg);
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const int pl = (t.n - 1) / 2;
#define AFR_UPN(N)                                                                                              \
    case N:                                                                                                     \
        if (t.pad == pl)                                                                                        \
            upn_kernel<TI, TO, N, (N - 1) / 2><<<(unsigned)grid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, H, W, \
                                                                              strips, nseg, R, t);             \
        else                                                                                                    \
            upn_kernel<TI, TO, N, N - 1 - (N - 1) / 2><<<(unsigned)grid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, \
                                                                                      H, W, strips, nseg, R, t); \
        break;
    switch (t.n) { AFR_UPN(2) AFR_UPN(4) AFR_UPN(5) AFR_UPN(6) AFR_UPN(7) AFR_UPN(8) default: return cudaErrorInvalidConfiguration; }
#undef AFR_UPN
    return cudaGetLastError();
}

cudaError_t stripn_up_like(const void *in, void *out, long planes, int H, int W, const TapsG &t, int in_dtype,
                           int out_dtype, cudaStream_t s)
{
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32) return launch_upn<float, float>(in, out, planes, H, W, t, s);
    if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) return launch_upn<bf16, bf16>(in, out, planes, H, W, t, s);
    if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) return launch_upn<bf16, float>(in, out, planes, H, W, t, s);
    return launch_upn<float, bf16>(in, out, planes, H, W, t, s);
}

// in [planes,H,W] (W % 8 == 0) -> out [planes,ceil(H/2),W/2]
template <typename T>
static cudaError_t launch_downn(const void *in, void *out, long planes, int H, int W, const TapsG &t, cudaStream_t s)
{
    const int Ho = (H + 1) / 2, Wo = W / 2, strips = Wo / 4;
    int nseg;
    const int R = rows_per_segment(planes, strips, Ho, (t.n - 1) / 2 + 1, &nseg);
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const int pl = (t.n - 1) / 2;
#define AFR_DNN(N)                                                                                              \
    case N:                                                                                                     \
        if (t.pad == pl)                                                                                        \
            downn_kernel<T, N, (N - 1) / 2><<<(unsigned)grid, 256, 0, s>>>((const T *)in, (T *)out, planes, H, W, Ho, \
                                                                           Wo, strips, nseg, R, t);            \
        else                                                                                                    \
            downn_kernel<T, N, N - 1 - (N - 1) / 2><<<(unsigned)grid, 256, 0, s>>>((const T *)in, (T *)out, planes, H, \
                                                                                   W, Ho, Wo, strips, nseg, R, t); \
        break;
    switch (t.n) { AFR_DNN(2) AFR_DNN(4) AFR_DNN(5) AFR_DNN(6) AFR_DNN(7) AFR_DNN(8) default: return cudaErrorInvalidConfiguration; }
#undef AFR_DNN
    return cudaGetLastError();
}

cudaError_t stripn_down_like(const void *in, void *out, long planes, int H, int W, const TapsG &t, int dtype,
                             cudaStream_t s)
{
    return dtype == AFR_F32 ? launch_downn<float>(in, out, planes, H, W, t, s)
                            : launch_downn<bf16>(in, out, planes, H, W, t, s);
}

}  // namespace afr
