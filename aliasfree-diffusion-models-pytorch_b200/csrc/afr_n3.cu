// N == 3 kernels (the canonical filter of the reference: kernel_size=3, Train.ipynb:101-104).
//
// Register-strip formulation.  A thread owns 4 adjacent output columns j..j+3 of one plane
// and walks down R rows.  Everything that lives on the 2x grid (u, gelu(u), and for the
// adjoint dg and gelu'(u)*dg) exists only in registers:
//
//   per output row i the thread needs "mid" rows 2i-1, 2i, 2i+1 at columns 2j-1 .. 2j+7;
//   row 2i+1 of this iteration is row 2(i+1)-1 of the next, so it is carried and only two
//   new mid rows are evaluated per step; their column 2j-1 is the left neighbour's column
//   2j+7 and arrives by warp shuffle, so a step costs 16 GELUs for 4 outputs -- the ideal 4
//   per output.  x rows are carried the same way (one new row of 6 values per step), and
//   GELUs are evaluated in pairs on packed f32x2 FMAs (afr_common.cuh).
//
// Forward:  mid = gelu(up(x; kU)),                      out = down(mid; kB), kB = k_down
// Adjoint:  mid = gelu'(up(x; kU)) * up(dy; kG),        out = down(mid; kB),
//           kG = flip(k_down), kB = flip(k_up)          (pads are 1/1 for N == 3);
//           kU arrives pre-scaled by kappa (afr_common.cuh) so that the kernel holds kappa*u
// mid is forced to zero outside [0,2H) x [0,2W): only column 2j-1 at j == 0 and row
// 2i-1 at i == 0 can fall outside, which is what `first_col` / the i0 == 0 start handle.
//
// Two data paths feed the same core:
//   direct  rows are read from global memory (128-bit loads + two scalars, L1-prefetched),
//           used for 4x4 planes and shapes TMA cannot describe;
//   tma     row-streaming: a CTA owns P planes x one column tile for the whole plane height; one
//           elected thread keeps a 2-stage ring of cp.async.bulk.tensor boxes
//           [Tw + 2*halo, R+1 rows, P planes] in flight (out-of-bounds zero fill = the conv's
//           zero padding, mbarrier complete_tx), threads read rows from the staged chunk.
#include <cuda.h>
#include <cstdlib>
#include <memory>
#include <new>
#include <type_traits>

#include "afr_common.cuh"
#include "afr_kernels.h"

namespace afr {

// ---------------------------------------------------------------------------------
// mid-row arithmetic
// ---------------------------------------------------------------------------------
// 9 columns m = 0..8  <->  X = 2j-1+m.  xa/xb hold input columns j-1 .. j+4.
// Column 0 is kept separate from columns 1..8: in shuffle mode it usually comes from lane-1.
template <int M0>
__device__ __forceinline__ void up_even_cols(const float (&xa)[6], const Taps3 &k, float (&u)[9])
{
#pragma unroll
    for (int m = M0; m < (M0 == 0 ? 1 : 9); ++m) {
        if (m & 1) u[m] = k.k[1][1] * xa[(m + 1) / 2];
        else u[m] = fmaf(k.k[1][2], xa[m / 2 + 1], k.k[1][0] * xa[m / 2]);
    }
}

template <int M0>
__device__ __forceinline__ void up_odd_cols(const float (&xa)[6], const float (&xb)[6],
                                            const Taps3 &k, float (&u)[9])
{
#pragma unroll
    for (int m = M0; m < (M0 == 0 ? 1 : 9); ++m) {
        if (m & 1) {
            u[m] = fmaf(k.k[2][1], xb[(m + 1) / 2], k.k[0][1] * xa[(m + 1) / 2]);
        } else {
            float t = k.k[0][0] * xa[m / 2];
            t = fmaf(k.k[0][2], xa[m / 2 + 1], t);
            t = fmaf(k.k[2][0], xb[m / 2], t);
            u[m] = fmaf(k.k[2][2], xb[m / 2 + 1], t);
        }
    }
}

// Activation on the mid rows.  fwd: u <- gelu(u) in place (g is unused, pass u twice).
// bwd: g <- gelu'(u) * g.  Columns 1..8 form packed pairs; column 0 (X = 2j-1) is separate
// because in shuffle mode it is normally taken from the left neighbour's column 8.
// kLow: the bf16-grade polynomials of afr_common.cuh (kernels whose tensors are bf16).  The adjoint's
// degree-4 S~ is a pure saving (4 packed FMAs per pair; its clamp exists in the fp32 form too).  The
// forward's degree-4 P needs a clamp the degree-5 form does not (one FMNMX per element for one FFMA2 per
// pair): AFR_BF16_LOW_FWD selects it (A/B builds).
#ifndef AFR_BF16_LOW_FWD
#define AFR_BF16_LOW_FWD 0
#endif
template <typename TO, bool kBwd> struct LowMath {
    static constexpr bool value = sizeof(TO) == 2 && (kBwd || AFR_BF16_LOW_FWD);
};
template <bool kBwd, bool kLow>
__device__ __forceinline__ void act_pair(float &ua, float &ub, float &ga, float &gb)
{
    if (kBwd) {
        if (kLow) gelu_grad_scaled_mul_low_x2(ua, ub, ga, gb);
        else gelu_grad_scaled_mul_x2(ua, ub, ga, gb);
    } else {
        if (kLow) gelu_erf_low_x2(ua, ub);
        else gelu_erf_x2(ua, ub);
    }
}

template <bool kBwd, bool kLow>
__device__ __forceinline__ void act_cols_1_8(float (&u)[9], float (&g)[9])
{
#pragma unroll
    for (int c = 1; c < 9; c += 2) act_pair<kBwd, kLow>(u[c], u[c + 1], g[c], g[c + 1]);
}

template <bool kBwd, bool kLow>
__device__ __forceinline__ void act_col0_pair(float (&ue)[9], float (&uo)[9], float (&ge)[9], float (&go)[9])
{
    act_pair<kBwd, kLow>(ue[0], uo[0], ge[0], go[0]);
}

// ---------------------------------------------------------------------------------
// row sources
// ---------------------------------------------------------------------------------
// Optional per-plane affine x * a + b folded into the load (the normalise + affine step of the
// GroupNorm that precedes the activation, modules/ddpm_utils.py:122, 127).  Zero padding applies
// to the normalised tensor, so `b` is added only to samples inside the plane: bc / bl / br are the
// shift for the 4 centre columns, the left and the right neighbour (0 where that column is outside).
struct Affine {
    float a, bc, bl, br;
};

template <typename T, bool kRes, bool kAff = false>
struct GlobalRows {
    const T *x;   // plane base, already offset to column j
    const T *r;   // residual plane (kRes), same offset
    int H, W, j;
    Affine f;
    __device__ __forceinline__ void load(int row, float (&v)[6]) const
    {
        if (row < 0 || row >= H) {
#pragma unroll
            for (int c = 0; c < 6; ++c) v[c] = 0.f;
            return;
        }
        const long off = (long)row * W;
        float4 c4 = ld4(x + off);
        float l = (j > 0) ? ld1(x + off - 1) : 0.f;
        float rr = (j + 4 < W) ? ld1(x + off + 4) : 0.f;
        if (kAff) {
            c4.x = fmaf(c4.x, f.a, f.bc); c4.y = fmaf(c4.y, f.a, f.bc);
            c4.z = fmaf(c4.z, f.a, f.bc); c4.w = fmaf(c4.w, f.a, f.bc);
            l = fmaf(l, f.a, f.bl); rr = fmaf(rr, f.a, f.br);
        }
        if (kRes) {
            float4 q4 = ld4(r + off);
            c4.x += q4.x; c4.y += q4.y; c4.z += q4.z; c4.w += q4.w;
            if (j > 0) l += ld1(r + off - 1);
            if (j + 4 < W) rr += ld1(r + off + 4);
        }
        v[0] = l; v[1] = c4.x; v[2] = c4.y; v[3] = c4.z; v[4] = c4.w; v[5] = rr;
    }
};

// The two halo columns (j-1 and j+4) are the neighbouring strips' c4.w / c4.x.  Read from the staged tile
// as scalars (lane stride 16 bytes) each is a 4-way bank conflict: ncu counts 48 % of the kernel's shared
// wavefronts as conflict replays.  AFR_HALO_SHFL=1 takes them from lane -1 / lane +1 by shuffle instead
// (only lanes whose neighbour strip is not in the adjacent lane read the staged halo, conflict-free).
// Measured on B200 (round 2, tools/ab_sweep.py): the shuffle form is 0-1.5 % SLOWER on every shape, fp32 and
// bf16, forward and adjoint -- the LSU pipe is at 11 % and the kernel is issue-bound, so the replays are
// hidden while two extra SHFL + two predicated LDS per row are not.  The scalar loads stay the default.
#ifndef AFR_HALO_SHFL
#define AFR_HALO_SHFL 0
#endif
template <typename T, bool kRes, bool kAff = false>
struct TileRows {
    const T *x;   // shared tile, pointing at (tile row 0, this thread's column j)
    const T *r;
    int pitch, row0;   // row0 = global row index of tile row 0
    int H;             // plane height (kAff: rows outside the plane must stay exactly zero)
    Affine f;
    bool lds_l, lds_r; // this lane reads its left / right halo column from the tile (see above)
    __device__ __forceinline__ void halo(const T *p, const float4 &c4, float &l, float &rr) const
    {
#if AFR_HALO_SHFL
        l = __shfl_up_sync(0xffffffffu, c4.w, 1);
        rr = __shfl_down_sync(0xffffffffu, c4.x, 1);
        if (lds_l) l = lds1(p - 1);
        if (lds_r) rr = lds1(p + 4);
#else
        l = lds1(p - 1); rr = lds1(p + 4);
#endif
    }
    __device__ __forceinline__ void load(int row, float (&v)[6]) const
    {
        const int off = (row - row0) * pitch;
        float4 c4 = lds4(x + off);
        float l, rr;
        halo(x + off, c4, l, rr);
        if (kAff) {                                   // staged zeros (TMA OOB fill) times a stay zero
            const bool in = (unsigned)row < (unsigned)H;
            const float bc = in ? f.bc : 0.f, bl = in ? f.bl : 0.f, br = in ? f.br : 0.f;
            c4.x = fmaf(c4.x, f.a, bc); c4.y = fmaf(c4.y, f.a, bc);
            c4.z = fmaf(c4.z, f.a, bc); c4.w = fmaf(c4.w, f.a, bc);
            l = fmaf(l, f.a, bl); rr = fmaf(rr, f.a, br);
        }
        if (kRes) {
            const float4 q4 = lds4(r + off);
            float ql, qr;
            halo(r + off, q4, ql, qr);
            c4.x += q4.x; c4.y += q4.y; c4.z += q4.z; c4.w += q4.w;
            l += ql; rr += qr;
        }
        v[0] = l; v[1] = c4.x; v[2] = c4.y; v[3] = c4.z; v[4] = c4.w; v[5] = rr;
    }
};

__device__ __forceinline__ Affine make_affine(const float *scale, const float *shift, long p, bool plane_ok,
                                              int j, int W)
{
    Affine f = {0.f, 0.f, 0.f, 0.f};
    if (scale && plane_ok) {
        const float b = __ldg(shift + p);
        f.a = __ldg(scale + p);
        f.bc = (j >= 0 && j < W) ? b : 0.f;
        f.bl = (j > 0 && j <= W) ? b : 0.f;
        f.br = (j + 4 >= 0 && j + 4 < W) ? b : 0.f;
    }
    return f;
}

// ---------------------------------------------------------------------------------
// the strip core
// ---------------------------------------------------------------------------------
// Per-thread state carried from one output row to the next.  Two sets are used ping-pong so that
// "row i+1 becomes row i" costs no moves; members a variant does not touch never become registers.
struct RowSet {
    float x[6], d[6];      // input row (x / dy), columns j-1 .. j+4
    float m[9];            // general taps: mid row 2i-1, columns 2j-1 .. 2j+7
    float hx[5], hd[5];    // symmetric taps: x[c] + x[c+1] (and dy likewise), c = 1..4 used
    float r[4];            // symmetric taps: contribution of mid row 2i-1 to the 4 outputs
};

// One output row, general 3x3 taps.  In: A.x / A.d = input rows i, A.m = mid row 2i-1.  Loads rows
// i+1 into B.x / B.d, evaluates mid rows 2i (local) and 2i+1 (-> B.m), stores out[i][j..j+3] if `store`.
//
// Column 0 of every mid row (X = 2j-1) is the same sample as column 8 of the thread one strip to
// the left, so it is fetched from lane-1 with a shuffle instead of being recomputed: 16 GELUs per
// step instead of 18.  Every lane of the warp therefore runs every step (lanes without real work
// run on safe coordinates with `store` off).  `own0` marks the lanes whose left neighbour is not
// lane-1 (first strip of a later column tile, or lane 0 when the strip count does not divide 32);
// they compute column 0 themselves in a short divergent branch.
struct StepK {             // general-tap kernels
    Taps3 kU, kG, kB;
};

// Where a step's four outputs go: a pointer (NCHW: four adjacent columns, one 128-bit store) or an NHWC
// position (the four columns are `cs` channels apart: four scalar stores, each coalesced across the warp's
// 32 channels).
template <typename T> struct NhwcOut {
    T *p;
    int cs;
    __device__ __forceinline__ NhwcOut operator+(long d) const { return NhwcOut{p + d, cs}; }
};
template <class OUT> struct OutElem;
template <typename T> struct OutElem<T *> { typedef T type; };
template <typename T> struct OutElem<NhwcOut<T>> { typedef T type; };
template <typename T> __device__ __forceinline__ void store4(T *p, float4 v) { st4(p, v); }
__device__ __forceinline__ void st1cs(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st1cs(bf16 *p, float v) { st1(p, v); }
template <typename T> __device__ __forceinline__ void store4(const NhwcOut<T> &o, float4 v)
{
    st1cs(o.p, v.x); st1cs(o.p + o.cs, v.y); st1cs(o.p + 2 * o.cs, v.z); st1cs(o.p + 3 * o.cs, v.w);
}

template <bool kBwd, class SX, class SD, class OUT>
__device__ __forceinline__ void strip_step(const SX &sx, const SD &sd, OUT out_row, int i,
                                           bool store, bool first_col, bool any0, bool own0, const StepK &K,
                                           const RowSet &A, RowSet &B)
{
    constexpr bool kLow = LowMath<typename OutElem<OUT>::type, kBwd>::value;
    const Taps3 &kU = K.kU, &kG = K.kG, &kB = K.kB;
    const float (&xa)[6] = A.x, (&da)[6] = A.d, (&mp)[9] = A.m;
    float (&xb)[6] = B.x, (&db)[6] = B.d, (&mo)[9] = B.m;
    sx.load(i + 1, xb);
    if (kBwd) sd.load(i + 1, db);
    float me[9];
    if (kBwd) {
        float ue[9], uo[9];
        up_even_cols<1>(xa, kU, ue);
        up_odd_cols<1>(xa, xb, kU, uo);
        up_even_cols<1>(da, kG, me);
        up_odd_cols<1>(da, db, kG, mo);
        act_cols_1_8<true, kLow>(ue, me);
        act_cols_1_8<true, kLow>(uo, mo);
        if (any0 && own0) {      // any0 is warp-uniform: the common shapes skip with one uniform branch
            up_even_cols<0>(xa, kU, ue);
            up_odd_cols<0>(xa, xb, kU, uo);
            up_even_cols<0>(da, kG, me);
            up_odd_cols<0>(da, db, kG, mo);
            act_col0_pair<true, kLow>(ue, uo, me, mo);
        }
    } else {
        up_even_cols<1>(xa, kU, me);
        up_odd_cols<1>(xa, xb, kU, mo);
        act_cols_1_8<false, kLow>(me, me);
        act_cols_1_8<false, kLow>(mo, mo);
        if (any0 && own0) {      // any0 is warp-uniform: the common shapes skip with one uniform branch
            up_even_cols<0>(xa, kU, me);
            up_odd_cols<0>(xa, xb, kU, mo);
            act_col0_pair<false, kLow>(me, mo, me, mo);
        }
    }
    const float le = __shfl_up_sync(0xffffffffu, me[8], 1), lo = __shfl_up_sync(0xffffffffu, mo[8], 1);
    if (first_col) { me[0] = 0.f; mo[0] = 0.f; }             // own0 and first_col exclude each other (j > 0 vs j == 0)
    if (!(own0 || first_col)) { me[0] = le; mo[0] = lo; }
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float acc = kB.k[0][0] * mp[2 * q];
        acc = fmaf(kB.k[0][1], mp[2 * q + 1], acc);
        acc = fmaf(kB.k[0][2], mp[2 * q + 2], acc);
        acc = fmaf(kB.k[1][0], me[2 * q], acc);
        acc = fmaf(kB.k[1][1], me[2 * q + 1], acc);
        acc = fmaf(kB.k[1][2], me[2 * q + 2], acc);
        acc = fmaf(kB.k[2][0], mo[2 * q], acc);
        acc = fmaf(kB.k[2][1], mo[2 * q + 1], acc);
        acc = fmaf(kB.k[2][2], mo[2 * q + 2], acc);
        o[q] = acc;
    }
    if (store) store4(out_row, make_float4(o[0], o[1], o[2], o[3]));
}

// The same step for D4-symmetric taps [[a,b,a],[b',c,b'],[a,b,a]] (every filter circularLowpassKernel
// designs, modules/filtrs.py:20-37: a radial profile times an outer product of a symmetric window).
// Each of the four polyphase classes of the 2x grid then carries ONE tap value s_ph, so
//   u = s_ph * xi,  xi = x | x+x_right | x+x_below | sum of the 2x2 block          (adds only)
// and the multiplications move out of the stencil:
//   forward  gelu(s*xi) = s * ghat_ph(xi),  ghat_ph(xi) = relu(xi) - |xi| 2^P(s|xi|)   (s >= 0):
//            P(s t) is a polynomial in t with per-phase coefficients p_k s^k (host), and the factor s
//            is folded into the down taps (every down tap reads exactly one phase class);
//   adjoint  dg = sG_ph * delta (delta = the same sums of dy) with sG folded into the down taps,
//            w = (kappa s_ph) * xi feeds the unchanged derivative.
// With top/bottom-symmetric down taps the odd mid row 2i+1 contributes the same 3-tap sum r[q] to
// output rows i and i+1, so r is what is carried (4 registers instead of a 9-value mid row) and an
// output costs 7 FMA-pipe operations instead of 9.  Per output row a thread issues 12 + 28 FMA-pipe
// operations around the 16 GELUs instead of 36 + 36.
enum { PH_EE = 0, PH_EO = 1, PH_OE = 2, PH_OO = 3 };   // (mid row parity, mid column parity)
struct SymK {
    float p[4][5];         // forward: Horner coefficients (degree 5 .. 1) of P(s_ph t) in -t; P(0) = -1
                           // (bf16 kernels: degree 4 .. 1 in p[.][1..4], p[.][0] unused)
    float sw[4];           // adjoint: kappa * s_ph
    float dn[4];           // folded down taps by the phase class they read
    float cl[4];           // bf16 forward: -T / s_ph, the clamp of -|xi| that keeps s_ph |xi| <= T
};

template <bool kLow>
__device__ __forceinline__ void gelu_hat_x2(float &a, float &b, const float (&ca)[5], const float (&cb)[5],
                                            float cla, float clb)
{
    f32x2 sv, p;
    if (kLow) {
        sv = pack2(fmaxf(-fabsf(a), cla), fmaxf(-fabsf(b), clb));
        p = pack2(ca[1], cb[1]);
    } else {
        sv = pack2(-fabsf(a), -fabsf(b));
        p = fma2(pack2(ca[0], cb[0]), sv, pack2(ca[1], cb[1]));
    }
    p = fma2(p, sv, pack2(ca[2], cb[2]));
    p = fma2(p, sv, pack2(ca[3], cb[3]));
    p = fma2(p, sv, pack2(ca[4], cb[4]));
    p = fma2(p, sv, splat2(AFR_P0));
    float pa, pb;
    unpack2(p, pa, pb);
    unpack2(fma2(sv, pack2(ex2_approx(pa), ex2_approx(pb)), pack2(relu_nan(a), relu_nan(b))), a, b);
}

// xi (and delta) of mid rows 2i (e) and 2i+1 (o), columns M0 .. (M0 == 0 ? 0 : 8)
template <int M0>
__device__ __forceinline__ void sym_sums(const float (&xa)[6], const float (&ha)[5], const float (&xb)[6],
                                         const float (&hb)[5], float (&e)[9], float (&o)[9])
{
#pragma unroll
    for (int m = M0; m < (M0 == 0 ? 1 : 9); ++m) {
        if (m & 1) {
            e[m] = xa[(m + 1) / 2];
            o[m] = xa[(m + 1) / 2] + xb[(m + 1) / 2];
        } else if (m == 0) {                         // only the column-0 branch needs h[0]
            const float h0a = xa[0] + xa[1];
            e[m] = h0a;
            o[m] = h0a + (xb[0] + xb[1]);
        } else {
            e[m] = ha[m / 2];
            o[m] = ha[m / 2] + hb[m / 2];
        }
    }
}

template <bool kBwd, class SX, class SD, class OUT>
__device__ __forceinline__ void strip_step(const SX &sx, const SD &sd, OUT out_row, int i,
                                           bool store, bool first_col, bool any0, bool own0, const SymK &K,
                                           const RowSet &A, RowSet &B)
{
    constexpr bool kLow = LowMath<typename OutElem<OUT>::type, kBwd>::value;
    sx.load(i + 1, B.x);
    if (kBwd) sd.load(i + 1, B.d);
#pragma unroll
    for (int c = 1; c < 5; ++c) {
        B.hx[c] = B.x[c] + B.x[c + 1];
        if (kBwd) B.hd[c] = B.d[c] + B.d[c + 1];
    }
    float me[9], mo[9];
    float c0e = 0.f, c0o = 0.f;      // column 0 when it is not the left neighbour's: 0 at a plane edge, else own
    if (kBwd) {
        float ue[9], uo[9];
        sym_sums<1>(A.x, A.hx, B.x, B.hx, ue, uo);
        sym_sums<1>(A.d, A.hd, B.d, B.hd, me, mo);
#pragma unroll
        for (int m = 1; m < 9; ++m) {
            ue[m] *= K.sw[(m & 1) ? PH_EE : PH_EO];
            uo[m] *= K.sw[(m & 1) ? PH_OE : PH_OO];
        }
        act_cols_1_8<true, kLow>(ue, me);
        act_cols_1_8<true, kLow>(uo, mo);
        if (any0 && own0) {      // any0 is warp-uniform: the common shapes skip with one uniform branch
            sym_sums<0>(A.x, A.hx, B.x, B.hx, ue, uo);
            sym_sums<0>(A.d, A.hd, B.d, B.hd, me, mo);
            ue[0] *= K.sw[PH_EO];
            uo[0] *= K.sw[PH_OO];
            act_col0_pair<true, kLow>(ue, uo, me, mo);
            c0e = me[0]; c0o = mo[0];
        }
    } else {
        sym_sums<1>(A.x, A.hx, B.x, B.hx, me, mo);
        gelu_hat_x2<kLow>(me[1], me[3], K.p[PH_EE], K.p[PH_EE], K.cl[PH_EE], K.cl[PH_EE]);
        gelu_hat_x2<kLow>(me[5], me[7], K.p[PH_EE], K.p[PH_EE], K.cl[PH_EE], K.cl[PH_EE]);
        gelu_hat_x2<kLow>(me[2], me[4], K.p[PH_EO], K.p[PH_EO], K.cl[PH_EO], K.cl[PH_EO]);
        gelu_hat_x2<kLow>(me[6], me[8], K.p[PH_EO], K.p[PH_EO], K.cl[PH_EO], K.cl[PH_EO]);
        gelu_hat_x2<kLow>(mo[1], mo[3], K.p[PH_OE], K.p[PH_OE], K.cl[PH_OE], K.cl[PH_OE]);
        gelu_hat_x2<kLow>(mo[5], mo[7], K.p[PH_OE], K.p[PH_OE], K.cl[PH_OE], K.cl[PH_OE]);
        gelu_hat_x2<kLow>(mo[2], mo[4], K.p[PH_OO], K.p[PH_OO], K.cl[PH_OO], K.cl[PH_OO]);
        gelu_hat_x2<kLow>(mo[6], mo[8], K.p[PH_OO], K.p[PH_OO], K.cl[PH_OO], K.cl[PH_OO]);
        if (any0 && own0) {      // any0 is warp-uniform: the common shapes skip with one uniform branch
            sym_sums<0>(A.x, A.hx, B.x, B.hx, me, mo);
            gelu_hat_x2<kLow>(me[0], mo[0], K.p[PH_EO], K.p[PH_OO], K.cl[PH_EO], K.cl[PH_OO]);
            c0e = me[0]; c0o = mo[0];
        }
    }
    const float le = __shfl_up_sync(0xffffffffu, me[8], 1), lo = __shfl_up_sync(0xffffffffu, mo[8], 1);
    const bool take = !(own0 || first_col);                  // own0 and first_col exclude each other
    me[0] = take ? le : c0e;
    mo[0] = take ? lo : c0o;
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float ro = K.dn[PH_OE] * mo[2 * q + 1];
        ro = fmaf(K.dn[PH_OO], mo[2 * q] + mo[2 * q + 2], ro);
        B.r[q] = ro;
        float acc = fmaf(K.dn[PH_EO], me[2 * q] + me[2 * q + 2], A.r[q]);
        acc = fmaf(K.dn[PH_EE], me[2 * q + 1], acc);
        o[q] = acc + ro;
    }
    if (store) store4(out_row, make_float4(o[0], o[1], o[2], o[3]));
}

template <class KT> struct IsSym { static constexpr bool value = false; };
template <> struct IsSym<SymK> { static constexpr bool value = true; };

// state before the first step of a strip: input row i0 in A, nothing above it yet
template <bool kBwd, class KT, class SX, class SD>
__device__ __forceinline__ void strip_begin(const SX &sx, const SD &sd, int i0, RowSet &A)
{
    sx.load(i0, A.x);
    if (kBwd) sd.load(i0, A.d);
    if (IsSym<KT>::value) {
#pragma unroll
        for (int c = 1; c < 5; ++c) {
            A.hx[c] = A.x[c] + A.x[c + 1];
            if (kBwd) A.hd[c] = A.d[c] + A.d[c + 1];
        }
    }
}

__device__ __forceinline__ void clear_carry(RowSet &A)     // mid row -1 lies outside the 2x grid
{
#pragma unroll
    for (int c = 0; c < 9; ++c) A.m[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) A.r[c] = 0.f;
}

// R steps starting at row i0 (R is warp-uniform); rows >= i1 are computed but not stored.
template <bool kBwd, class KT, class SX, class SD, typename TO>
__device__ __forceinline__ void strip_core(const SX &sx, const SD &sd, TO *__restrict__ out, int W,
                                           int i0, int i1, int R, bool valid, bool first_col, bool own0,
                                           const KT &K)
{
    RowSet S0, S1;
#pragma unroll
    for (int c = 0; c < 6; ++c) { S0.d[c] = 0.f; S1.d[c] = 0.f; }
    // The carried mid row 2*i0-1 comes from one dry step at row i0-1 (nothing stored).  The branch
    // is warp-uniform (a warp that holds both a first and a later row segment runs it on all lanes
    // and zeroes the carry below) so that the warp stays converged for the shuffles; planes that fit
    // one segment (4x4, 8x8) skip it altogether.
    clear_carry(S1);
    if (__any_sync(0xffffffffu, i0 > 0)) {
        strip_begin<kBwd, KT>(sx, sd, i0 - 1, S1);
        strip_step<kBwd>(sx, sd, out, i0 - 1, false, first_col, true, own0, K, S1, S0);
    } else {
        strip_begin<kBwd, KT>(sx, sd, i0, S0);
    }
    if (i0 == 0) clear_carry(S0);
    const int iend = i0 + R;
    int i = i0;
    for (; i + 1 < iend; i += 2) {
        strip_step<kBwd>(sx, sd, out + (long)i * W, i, valid && i < i1, first_col, true, own0, K, S0, S1);
        strip_step<kBwd>(sx, sd, out + (long)(i + 1) * W, i + 1, valid && i + 1 < i1, first_col, true, own0, K, S1, S0);
    }
    if (i < iend) strip_step<kBwd>(sx, sd, out + (long)i * W, i, valid && i < i1, first_col, true, own0, K, S0, S1);
}

// ---------------------------------------------------------------------------------
// direct kernel
// ---------------------------------------------------------------------------------
template <typename T, bool kBwd, bool kRes, bool kAff, class KT>
__global__ void __launch_bounds__(256)
fgelu3_direct_kernel(const T *__restrict__ x, const T *__restrict__ res, const T *__restrict__ dy,
                     const float *__restrict__ scale, const float *__restrict__ shift,
                     T *__restrict__ out, long planes, int H, int W, int strips, int nseg, int R,
                     const __grid_constant__ KT K)
{
    // 32-bit index arithmetic (the launcher guarantees planes * strips * nseg < 2^31)
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned per_plane = (unsigned)(strips * nseg);
    const bool valid = idx < (unsigned)planes * per_plane;
    if (!valid) idx = 0;                           // idle lanes shadow thread 0, stores off
    const unsigned pu = idx / per_plane, rem = idx - pu * per_plane;
    const int seg = (int)(rem / (unsigned)strips);
    const int s = (int)(rem - (unsigned)seg * (unsigned)strips);
    const long p = pu;
    const int j = 4 * s, i0 = seg * R, i1 = min(H, i0 + R);
    const long base = p * (long)H * W + j;
    GlobalRows<T, kRes, kAff> sx{x + base, kRes ? res + base : nullptr, H, W, j,
                                 make_affine(kAff ? scale : nullptr, shift, p, true, j, W)};
    GlobalRows<T, false> sd{kBwd ? dy + base : nullptr, nullptr, H, W, j, Affine{0.f, 0.f, 0.f, 0.f}};
    const bool own0 = (s > 0) && ((threadIdx.x & 31) == 0);
    // all rows this thread will walk are requested up front (L1 prefetch), so the row-by-row loads
    // of the step loop do not each expose a full memory latency
    for (int r = max(i0 - 1, 0); r <= min(i0 + R, H - 1); ++r) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(x + base + (long)r * W));
        if (kRes) asm volatile("prefetch.global.L1 [%0];" ::"l"(res + base + (long)r * W));
        if (kBwd) asm volatile("prefetch.global.L1 [%0];" ::"l"(dy + base + (long)r * W));
    }
    strip_core<kBwd, KT>(sx, sd, out + base, W, i0, i1, R, valid, j == 0, own0, K);
}

// ---------------------------------------------------------------------------------
// 4x4 and 8x8 planes (the innermost UNet levels): whole planes in registers
// ---------------------------------------------------------------------------------
// Everything the general direct kernel decides at run time is a constant here: a plane is one
// (4x4) or two (8x8) strips wide, there are no row segments, the steps are fully unrolled, and all
// rows of every input are requested up front (one 128-bit load per row and tensor, plus the one
// neighbour column of the other strip for 8x8), so a thread exposes one memory latency per plane.
// 8x8: the two strips of a plane sit in an even/odd lane pair, column 0 of the odd strip comes
// from the even lane by shuffle.  (The 8x8 adjoint reads its rows on demand instead: two tensors
// of nine rows do not fit a reasonable register budget.)
template <int HH>
struct RegRows {
    float v[HH + 1][6];                             // rows 0..HH-1 and the zero row HH; columns -1 .. 4
    __device__ __forceinline__ void load(int row, float (&o)[6]) const
    {
#pragma unroll
        for (int c = 0; c < 6; ++c) o[c] = v[row][c];
    }
};

template <typename T, bool kRes, int HH>
__device__ __forceinline__ void load_plane_rows(const T *__restrict__ xp, const T *__restrict__ rp, int s,
                                                float a, float b, bool affine, RegRows<HH> &R)
{
    constexpr int WW = HH;                          // square planes
#pragma unroll
    for (int r = 0; r < HH; ++r) {
        const T *row = xp + r * WW + 4 * s;
        float4 c4 = ld4(row);
        float nb = 0.f;                             // 8x8: the column next to the strip inside the plane
        if (WW == 8) nb = ld1(s == 0 ? row + 4 : row - 1);
        if (kRes) {
            const T *rrow = rp + r * WW + 4 * s;
            const float4 q4 = ld4(rrow);
            if (affine) { c4.x = fmaf(c4.x, a, b); c4.y = fmaf(c4.y, a, b); c4.z = fmaf(c4.z, a, b); c4.w = fmaf(c4.w, a, b); nb = fmaf(nb, a, b); }
            c4.x += q4.x; c4.y += q4.y; c4.z += q4.z; c4.w += q4.w;
            if (WW == 8) nb += ld1(s == 0 ? rrow + 4 : rrow - 1);
        } else if (affine) {
            c4.x = fmaf(c4.x, a, b); c4.y = fmaf(c4.y, a, b); c4.z = fmaf(c4.z, a, b); c4.w = fmaf(c4.w, a, b); nb = fmaf(nb, a, b);
        }
        R.v[r][1] = c4.x; R.v[r][2] = c4.y; R.v[r][3] = c4.z; R.v[r][4] = c4.w;
        R.v[r][0] = (WW == 8 && s == 1) ? nb : 0.f;
        R.v[r][5] = (WW == 8 && s == 0) ? nb : 0.f;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) R.v[HH][c] = 0.f;
}

// The same rows read when a step asks for them (the 8x8 adjoint: two tensors of nine rows are too many
// registers).  Row indices are compile-time constants after unrolling, so there are still no bounds
// checks; the rows were requested with an L1 prefetch up front.
template <typename T, bool kRes, int HH>
struct LazyRows {
    const T *xp, *rp;
    int s;
    float a, b;                                       // x * a + b on load (a = 1, b = 0 without the affine)
    bool affine;
    __device__ __forceinline__ void load(int row, float (&o)[6]) const
    {
        if (row >= HH) {
#pragma unroll
            for (int c = 0; c < 6; ++c) o[c] = 0.f;
            return;
        }
        const T *p = xp + row * HH + 4 * s;
        float4 c4 = ld4(p);
        float nb = (HH == 8) ? ld1(s == 0 ? p + 4 : p - 1) : 0.f;
        if (affine) {
            c4.x = fmaf(c4.x, a, b); c4.y = fmaf(c4.y, a, b); c4.z = fmaf(c4.z, a, b); c4.w = fmaf(c4.w, a, b);
            nb = fmaf(nb, a, b);
        }
        if (kRes) {
            const T *q = rp + row * HH + 4 * s;
            const float4 q4 = ld4(q);
            c4.x += q4.x; c4.y += q4.y; c4.z += q4.z; c4.w += q4.w;
            if (HH == 8) nb += ld1(s == 0 ? q + 4 : q - 1);
        }
        o[1] = c4.x; o[2] = c4.y; o[3] = c4.z; o[4] = c4.w;
        o[0] = (HH == 8 && s == 1) ? nb : 0.f;
        o[5] = (HH == 8 && s == 0) ? nb : 0.f;
    }
};

template <bool kBwd, class KT, int HH, int I = 0>
struct PlaneSteps {
    template <class SX, class SD, typename T>
    static __device__ __forceinline__ void run(const SX &sx, const SD &sd, T *op, bool valid,
                                               bool first_col, const KT &K, RowSet &A, RowSet &B)
    {
        strip_step<kBwd>(sx, sd, op + I * HH, I, valid, first_col, false, false, K, A, B);
        PlaneSteps<kBwd, KT, HH, I + 1>::run(sx, sd, op, valid, first_col, K, B, A);
    }
};
template <bool kBwd, class KT, int HH>
struct PlaneSteps<kBwd, KT, HH, HH> {
    template <class SX, class SD, typename T>
    static __device__ __forceinline__ void run(const SX &, const SD &, T *, bool, bool, const KT &, RowSet &, RowSet &) {}
};

template <typename T, bool kBwd, bool kRes, bool kAff, class KT, int HH>
__global__ void __launch_bounds__(128)
fgelu3_plane_kernel(const T *__restrict__ x, const T *__restrict__ res, const T *__restrict__ dy,
                    const float *__restrict__ scale, const float *__restrict__ shift,
                    T *__restrict__ out, long planes, const __grid_constant__ KT K)
{
    constexpr int STRIPS = HH / 4, HW = HH * HH;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long p = idx / STRIPS;
    const int s = (int)(idx % STRIPS);
    const bool valid = p < planes;
    if (STRIPS == 1 && !valid) return;              // 4x4 has no shuffles: idle lanes may leave
    if (!valid) p = 0;                              // 8x8: idle lanes shadow plane 0, stores off
    const T *xp = x + p * HW, *rp = kRes ? res + p * HW : nullptr, *dp = kBwd ? dy + p * HW : nullptr;
    RowSet S0, S1;
#pragma unroll
    for (int c = 0; c < 6; ++c) { S0.d[c] = 0.f; S1.d[c] = 0.f; }
    clear_carry(S0);
    if constexpr (kBwd && HH == 8) {                // rows on demand (see LazyRows)
#pragma unroll
        for (int r = 0; r < HH; ++r) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + r * HH + 4 * s));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dp + r * HH + 4 * s));
            if (kRes) asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + r * HH + 4 * s));
        }
        float a = 1.f, b = 0.f;
        if (kAff) { a = __ldg(scale + p); b = __ldg(shift + p); }
        const LazyRows<T, kRes, HH> sx{xp, rp, s, a, b, kAff};
        const LazyRows<T, false, HH> sd{dp, nullptr, s, 1.f, 0.f, false};
        strip_begin<kBwd, KT>(sx, sd, 0, S0);
        PlaneSteps<kBwd, KT, HH>::run(sx, sd, out + p * HW + 4 * s, valid, s == 0, K, S0, S1);
    } else {
        RegRows<HH> sx, sd;
        float a = 1.f, b = 0.f;
        if (kAff) { a = __ldg(scale + p); b = __ldg(shift + p); }
        load_plane_rows<T, kRes, HH>(xp, rp, s, a, b, kAff, sx);
        if (kBwd) load_plane_rows<T, false, HH>(dp, nullptr, s, 1.f, 0.f, false, sd);
        strip_begin<kBwd, KT>(sx, sd, 0, S0);
        PlaneSteps<kBwd, KT, HH>::run(sx, sd, out + p * HW + 4 * s, valid, s == 0, K, S0, S1);
    }
}

// ---------------------------------------------------------------------------------
// TMA-staged kernel
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Column halo of the staged tile, in elements: 16 bytes on each side, so that the box start
// (j0 - halo) stays 16-byte aligned in global memory for fp32 (4) and bf16 (8) alike.
template <typename T> struct Halo { static constexpr int value = 16 / (int)sizeof(T); };

// Row-streaming tile configuration: a CTA owns P planes x one column tile of Tw outputs and
// walks down ALL rows of those planes in chunks of R rows; chunk k is the box
// [Tw + 2*halo, R + 1 rows (kR .. kR+R), P planes] and lands in stage k & 1.  Small batches of
// large planes are additionally split into `nsegs` row segments (blockIdx.y) to fill the GPU; a
// segment that does not start at row 0 begins one row early with its store switched off, which
// builds the carried mid row.
struct TileCfg {
    int Tw, R, P;              // tile width (outputs), rows per chunk, planes per CTA
    int strips;                // Tw / 4 : threads per plane row
    int tiles_x;               // column tiles per plane
    int ghost;                 // planes wider than one tile: leading strips of every tile that only
                               // feed the column-0 shuffle of their right neighbour (never stored)
    int own0_any;              // some lane's left neighbour strip is not lane-1 (W > 128 without ghosts, or a strip
                               // count that does not divide a warp); 0 lets every step skip the recompute branch
    int nsegs, Hs;             // row segments per plane (blockIdx.y) and their height
    int tile_bytes;            // bytes of one staged box, rounded up to 128
};

// Each thread keeps its 4-column strip of one plane for the whole plane height: the carried
// mid row and input row simply stay in registers across chunks, so there is no per-segment
// start-up work at all, and the next two chunks are always in flight (2-stage TMA/mbarrier
// ring) while the current one is being computed.
#ifndef AFR_TMA_MINB
#define AFR_TMA_MINB 1
#endif
template <typename T, bool kBwd, bool kRes, bool kAff, class KT>
__global__ void __launch_bounds__(128, AFR_TMA_MINB)
fgelu3_tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mres,
                  const __grid_constant__ CUtensorMap mdy, const float *__restrict__ scale,
                  const float *__restrict__ shift, T *__restrict__ out, long planes, int H,
                  int W, const __grid_constant__ TileCfg cfg, const __grid_constant__ KT K)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) uint64_t full[2];
    constexpr int HALO = Halo<T>::value;
    constexpr int NIN = 1 + (kRes ? 1 : 0) + (kBwd ? 1 : 0);

    const int tx = (int)(blockIdx.x % cfg.tiles_x);
    const long p0 = (long)(blockIdx.x / cfg.tiles_x) * cfg.P;
    const int j0 = tx * (cfg.Tw - 4 * cfg.ghost) - 4 * cfg.ghost;   // first column of strip 0 (may be < 0)
    const int pitch = cfg.Tw + 2 * HALO, rows = cfg.R + 1;
    const int stage_bytes = NIN * cfg.tile_bytes;
    const uint32_t box_bytes = (uint32_t)(pitch * rows * cfg.P * sizeof(T));

    const int seg_lo = (int)blockIdx.y * cfg.Hs;                 // first / one-past-last output row
    const int seg_hi = min(H, seg_lo + cfg.Hs);
    const int istart = seg_lo > 0 ? seg_lo - 1 : 0;              // first step (a dry one if seg_lo > 0)
    const int nchunks = (seg_hi - istart + cfg.R - 1) / cfg.R;

    auto issue = [&](int k) {          // one thread: arm the stage's barrier, launch its boxes
        unsigned char *base = tile_smem + (k & 1) * stage_bytes;
        uint64_t *bar = &full[k & 1];
        const int row = istart + k * cfg.R;
        mbar_expect_tx(bar, box_bytes * NIN);
        tma_load_3d(base, &mx, bar, j0 - HALO, row, (int)p0);
        if (kRes) tma_load_3d(base + cfg.tile_bytes, &mres, bar, j0 - HALO, row, (int)p0);
        if (kBwd) tma_load_3d(base + (kRes ? 2 : 1) * cfg.tile_bytes, &mdy, bar, j0 - HALO, row, (int)p0);
    };

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

    const int s = threadIdx.x % cfg.strips;
    int pl = threadIdx.x / cfg.strips;
    const int j = j0 + 4 * s;
    const bool valid = (pl < cfg.P) && (p0 + pl < planes) && (s >= cfg.ghost) && (j < W);
    if (pl >= cfg.P) pl = 0;                        // idle lanes shadow plane 0 of the tile, stores off
    const bool first_col = (j == 0);
    // with ghost strips every real strip has its left neighbour in lane-1; otherwise lane 0 of a
    // warp that does not start a plane row recomputes column 0 itself
    const bool any0 = cfg.own0_any != 0;
    const bool own0 = any0 && (cfg.ghost == 0) && (j > 0) && (s == 0 || (threadIdx.x & 31) == 0);
    // halo columns: from the adjacent lane unless that lane holds another plane row / tile / warp
    const bool lds_l = (s == 0) || ((threadIdx.x & 31) == 0);
    const bool lds_r = (s == cfg.strips - 1) || ((threadIdx.x & 31) == 31);
    const int toff = pl * rows * pitch + HALO + 4 * s;
    T *dst = out + (p0 + pl) * (long)H * W + (valid ? j : 0);
    const Affine aff = make_affine(kAff ? scale : nullptr, shift, p0 + pl, p0 + pl < planes, j, W);

    RowSet S0, S1;
#pragma unroll
    for (int c = 0; c < 6; ++c) { S0.d[c] = 0.f; S1.d[c] = 0.f; }
    clear_carry(S0);                                // mid row -1 lies outside the 2x grid

    for (int k = 0; k < nchunks; ++k) {
        const unsigned char *base = tile_smem + (k & 1) * stage_bytes;
        const T *xs = reinterpret_cast<const T *>(base);
        const T *rs = reinterpret_cast<const T *>(base + (kRes ? cfg.tile_bytes : 0));
        const T *ds = reinterpret_cast<const T *>(base + (kRes ? 2 : 1) * cfg.tile_bytes);
        const int r0 = istart + k * cfg.R;
        mbar_wait(&full[k & 1], (k >> 1) & 1);
        if (k == 0) {
            TileRows<T, kRes, kAff> sx{xs + toff, rs + toff, pitch, r0, H, aff, lds_l, lds_r};
            TileRows<T, false> sd{ds + toff, nullptr, pitch, r0, H, Affine{0.f, 0.f, 0.f, 0.f}, lds_l, lds_r};
            strip_begin<kBwd, KT>(sx, sd, r0, S0);
        }
        // R is even: register roles return to S0.  The last chunk stops after the segment's last row
        // (rounded up to a pair of steps) instead of walking rows nobody needs.  Row pointers (staged
        // rows i+1 / i+2 and the output row) are carried instead of being rebuilt from the row index.
        const int rend = min(r0 + cfg.R, seg_hi + ((seg_hi - r0) & 1));
        const T *xr = xs + toff + pitch, *rr = rs + toff + pitch, *dr = ds + toff + pitch;
        T *orow = dst + (long)r0 * W;
        // chunks that lie completely inside the segment (all but the first of a later segment and the
        // last) store under the per-thread predicate alone: no row comparisons in the hot loop
        auto walk = [&](auto inside) {
            constexpr bool kInside = decltype(inside)::value;
            for (int i = r0; i < rend; i += 2) {
                {
                    TileRows<T, kRes, kAff> sx{xr, rr, pitch, i + 1, H, aff, lds_l, lds_r};
                    TileRows<T, false> sd{dr, nullptr, pitch, i + 1, H, Affine{0.f, 0.f, 0.f, 0.f}, lds_l, lds_r};
                    strip_step<kBwd>(sx, sd, orow, i, kInside ? valid : (valid && i >= seg_lo && i < seg_hi), first_col,
                                     any0, own0, K, S0, S1);
                }
                {
                    TileRows<T, kRes, kAff> sx{xr + pitch, rr + pitch, pitch, i + 2, H, aff, lds_l, lds_r};
                    TileRows<T, false> sd{dr + pitch, nullptr, pitch, i + 2, H, Affine{0.f, 0.f, 0.f, 0.f}, lds_l, lds_r};
                    strip_step<kBwd>(sx, sd, orow + W, i + 1,
                                     kInside ? valid : (valid && i + 1 >= seg_lo && i + 1 < seg_hi), first_col, any0,
                                     own0, K, S1, S0);
                }
                xr += 2 * pitch; rr += 2 * pitch; dr += 2 * pitch;
                orow += 2 * W;
            }
        };
        if (r0 >= seg_lo && r0 + cfg.R <= seg_hi) walk(std::true_type{});
        else walk(std::false_type{});
        __syncthreads();                            // stage k & 1 fully consumed
        if (threadIdx.x == 0 && k + 2 < nchunks) issue(k + 2);
    }
}

// ---------------------------------------------------------------------------------
// NHWC (channels-last) flavour of the row-streaming kernel
// ---------------------------------------------------------------------------------
// Same strip core, different geometry.  In a channels-last tensor [B, H, W, C] the contiguous axis is C, so a
// warp's 32 lanes are 32 CHANNELS of one 4-column strip (every load / store of the warp is one 128-byte line) and
// the strips of a tile are the CTA's warps.  The TMA box is [32 channels, 4*strips + 2 columns, R + 1 rows,
// P images] of the rank-4 map {C, W, H, B}; out-of-bounds columns / rows are zero-filled exactly as in the NCHW
// kernel.  The left neighbour strip now lives in another warp, so every strip computes column 0 of its mid rows
// itself (18 GELUs per 4 outputs instead of 16, ~11 % more instructions) -- the price of not transposing the tensor,
// which is what PyTorch / cuDNN otherwise do around every channels-last convolution.  Symmetric taps only (every
// filter the reference designs); other filters take the NCHW kernels through a contiguous copy.
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct NhwcCfg {
    int strips, P, R;          // strips (warps) per column tile, images per CTA (strips * P == 4), rows per chunk
    int tiles_x, cgroups;      // column tiles per image row, 32-channel groups
    int nsegs, Hs;             // row segments (blockIdx.y) and their height
    int tile_bytes;            // bytes of one staged box, rounded up to 128
};

template <typename T, bool kRes>
struct NhwcRows {
    const T *x, *r;            // staged tile at (tile row 0, halo column of this strip, this lane's channel)
    int pitch, row0, H;        // elements per staged row; global row of tile row 0; plane height
    float a, bl, bc, br;       // x * a + b folded into the load (b only for samples inside the plane)
    bool aff;
    __device__ __forceinline__ void load(int row, float (&v)[6]) const
    {
        const int off = (row - row0) * pitch;
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = lds1(x + off + 32 * k);
        if (aff) {
            const bool in = (unsigned)row < (unsigned)H;       // staged zeros (OOB fill) times a stay zero
            v[0] = fmaf(v[0], a, in ? bl : 0.f);
#pragma unroll
            for (int k = 1; k < 5; ++k) v[k] = fmaf(v[k], a, in ? bc : 0.f);
            v[5] = fmaf(v[5], a, in ? br : 0.f);
        }
        if (kRes) {
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] += lds1(r + off + 32 * k);
        }
    }
};

template <typename T, bool kBwd, bool kRes>
__global__ void __launch_bounds__(128)
fgelu3_nhwc_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mres,
                   const __grid_constant__ CUtensorMap mdy, const float *__restrict__ scale,
                   const float *__restrict__ shift, T *__restrict__ out, int B, int C, int H, int W,
                   const __grid_constant__ NhwcCfg cfg, const __grid_constant__ SymK K)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) uint64_t full[2];
    constexpr int NIN = 1 + (kRes ? 1 : 0) + (kBwd ? 1 : 0);

    const int cg = (int)(blockIdx.x % cfg.cgroups);
    const int rest = (int)(blockIdx.x / cfg.cgroups);
    const int tx = rest % cfg.tiles_x;
    const int b0 = (rest / cfg.tiles_x) * cfg.P;
    const int Tw = 4 * cfg.strips, j0 = tx * Tw, cols = Tw + 2, rows = cfg.R + 1;
    const int pitch = cols * 32, img_elems = rows * pitch;
    const int stage_bytes = NIN * cfg.tile_bytes;
    const uint32_t box_bytes = (uint32_t)(img_elems * cfg.P * sizeof(T));

    const int seg_lo = (int)blockIdx.y * cfg.Hs;
    const int seg_hi = min(H, seg_lo + cfg.Hs);
    const int istart = seg_lo > 0 ? seg_lo - 1 : 0;
    const int nchunks = (seg_hi - istart + cfg.R - 1) / cfg.R;

    auto issue = [&](int k) {
        unsigned char *base = tile_smem + (k & 1) * stage_bytes;
        uint64_t *bar = &full[k & 1];
        const int row = istart + k * cfg.R;
        mbar_expect_tx(bar, box_bytes * NIN);
        tma_load_4d(base, &mx, bar, cg * 32, j0 - 1, row, b0);
        if (kRes) tma_load_4d(base + cfg.tile_bytes, &mres, bar, cg * 32, j0 - 1, row, b0);
        if (kBwd) tma_load_4d(base + (kRes ? 2 : 1) * cfg.tile_bytes, &mdy, bar, cg * 32, j0 - 1, row, b0);
    };

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = warp % cfg.strips, pl = warp / cfg.strips;
    const int j = j0 + 4 * s, c = cg * 32 + lane;
    const bool valid = (b0 + pl < B) && (j < W);
    const int b = (b0 + pl < B) ? b0 + pl : b0;               // idle warps shadow image b0, stores off
    const int spl = (b0 + pl < B) ? pl : 0;
    const bool first_col = (j == 0);
    const bool own0 = (j > 0);                                  // no left neighbour in this warp: always recompute
    const int toff = spl * img_elems + (4 * s) * 32 + lane;     // box column 0 is global column j0 - 1
    const long rowstride = (long)W * C;
    const NhwcOut<T> dst{out + ((long)b * H * W + (valid ? j : 0)) * C + c, C};
    float a = 1.f, bsh = 0.f;
    const bool aff = scale != nullptr;
    if (aff) { a = __ldg(scale + (long)b * C + c); bsh = __ldg(shift + (long)b * C + c); }
    const float bl = (j > 0 && j <= W) ? bsh : 0.f, bc = (j < W) ? bsh : 0.f, br = (j + 4 < W) ? bsh : 0.f;

    RowSet S0, S1;
#pragma unroll
    for (int q = 0; q < 6; ++q) { S0.d[q] = 0.f; S1.d[q] = 0.f; }
    clear_carry(S0);

    for (int k = 0; k < nchunks; ++k) {
        const unsigned char *base = tile_smem + (k & 1) * stage_bytes;
        const T *xs = reinterpret_cast<const T *>(base);
        const T *rs = reinterpret_cast<const T *>(base + (kRes ? cfg.tile_bytes : 0));
        const T *ds = reinterpret_cast<const T *>(base + (kRes ? 2 : 1) * cfg.tile_bytes);
        const int r0 = istart + k * cfg.R;
        mbar_wait(&full[k & 1], (k >> 1) & 1);
        if (k == 0) {
            NhwcRows<T, kRes> sx{xs + toff, rs + toff, pitch, r0, H, a, bl, bc, br, aff};
            NhwcRows<T, false> sd{ds + toff, nullptr, pitch, r0, H, 1.f, 0.f, 0.f, 0.f, false};
            strip_begin<kBwd, SymK>(sx, sd, r0, S0);
        }
        const int rend = min(r0 + cfg.R, seg_hi + ((seg_hi - r0) & 1));
        const T *xr = xs + toff + pitch, *rr = rs + toff + pitch, *dr = ds + toff + pitch;
        NhwcOut<T> orow = dst + (long)r0 * rowstride;
        for (int i = r0; i < rend; i += 2) {
            {
                NhwcRows<T, kRes> sx{xr, rr, pitch, i + 1, H, a, bl, bc, br, aff};
                NhwcRows<T, false> sd{dr, nullptr, pitch, i + 1, H, 1.f, 0.f, 0.f, 0.f, false};
                strip_step<kBwd>(sx, sd, orow, i, valid && i >= seg_lo && i < seg_hi, first_col, true, own0, K, S0, S1);
            }
            {
                NhwcRows<T, kRes> sx{xr + pitch, rr + pitch, pitch, i + 2, H, a, bl, bc, br, aff};
                NhwcRows<T, false> sd{dr + pitch, nullptr, pitch, i + 2, H, 1.f, 0.f, 0.f, 0.f, false};
                strip_step<kBwd>(sx, sd, orow + rowstride, i + 1, valid && i + 1 >= seg_lo && i + 1 < seg_hi, first_col,
                                 true, own0, K, S1, S0);
            }
            xr += 2 * pitch; rr += 2 * pitch; dr += 2 * pitch;
            orow = orow + 2 * rowstride;
        }
        __syncthreads();
        if (threadIdx.x == 0 && k + 2 < nchunks) issue(k + 2);
    }
}

// ---------------------------------------------------------------------------------
// standalone N == 3 resamplers (HBM-bound)
// ---------------------------------------------------------------------------------
// up-like: thread = 4 input columns -> 8 output columns, two output rows per input row
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
up3_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int H, int W, int strips,
           int nseg, int R, int C, long out_bstride, const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = 4 * s, i0 = seg * R, i1 = min(H, i0 + R);
    const TI *src = in + p * (long)H * W + j;
    TO *dst = out + strided_base((unsigned)p, C, out_bstride, 4L * H * W) + 2 * j;   // C == 0: dense output
    const int W2 = 2 * W;
    const bool has_r = (j + 4 < W);

    float xa[5], xb[5];
    {
        float4 c = ld4(src + (long)i0 * W);
        xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
        xb[4] = has_r ? ld1(src + (long)i0 * W + 4) : 0.f;
    }
    for (int i = i0; i < i1; ++i) {
#pragma unroll
        for (int c = 0; c < 5; ++c) xa[c] = xb[c];
        // bf16 input: few bytes in flight per thread; pull the row after next into L2 (+19 % on 64x64 planes)
        if (sizeof(TI) == 2 && i + 3 < H) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (long)(i + 3) * W));
        if (i + 1 < H) {
            float4 c = ld4(src + (long)(i + 1) * W);
            xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
            xb[4] = has_r ? ld1(src + (long)(i + 1) * W + 4) : 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < 5; ++c) xb[c] = 0.f;
        }
        float e[8], o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            e[2 * c] = k.k[1][1] * xa[c];
            e[2 * c + 1] = fmaf(k.k[1][2], xa[c + 1], k.k[1][0] * xa[c]);
            o[2 * c] = fmaf(k.k[2][1], xb[c], k.k[0][1] * xa[c]);
            float t = k.k[0][0] * xa[c];
            t = fmaf(k.k[0][2], xa[c + 1], t);
            t = fmaf(k.k[2][0], xb[c], t);
            o[2 * c + 1] = fmaf(k.k[2][2], xb[c + 1], t);
        }
        TO *r0 = dst + (long)(2 * i) * W2;
        st8(r0, e);                                  // 8 outputs: 2x128-bit (fp32) or 1x128-bit (bf16)
        st8(r0 + W2, o);
    }
}

// Short planes (H = 4 or 8, the innermost UNet levels): the same arithmetic with the plane height a
// compile-time constant -- every row of the strip is requested before the first one is used (one
// exposed memory latency per strip instead of one per row) and the row loop is fully unrolled.
template <typename TI, typename TO, int HH>
__global__ void __launch_bounds__(256)
up3_short_kernel(const TI *__restrict__ in, TO *__restrict__ out, long planes, int W, int strips,
                 const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= planes * strips) return;
    const int s = (int)(idx % strips);
    const long p = idx / strips;
    const int j = 4 * s, W2 = 2 * W;
    const TI *src = in + p * (long)HH * W + j;
    TO *dst = out + p * 4L * HH * W + 2 * j;
    const bool has_r = (j + 4 < W);
    float x[HH + 1][5];
#pragma unroll
    for (int r = 0; r < HH; ++r) {
        const float4 c = ld4(src + r * W);
        x[r][0] = c.x; x[r][1] = c.y; x[r][2] = c.z; x[r][3] = c.w;
        x[r][4] = has_r ? ld1(src + r * W + 4) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) x[HH][c] = 0.f;
#pragma unroll
    for (int i = 0; i < HH; ++i) {
        float e[8], o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            e[2 * c] = k.k[1][1] * x[i][c];
            e[2 * c + 1] = fmaf(k.k[1][2], x[i][c + 1], k.k[1][0] * x[i][c]);
            o[2 * c] = fmaf(k.k[2][1], x[i + 1][c], k.k[0][1] * x[i][c]);
            float t = k.k[0][0] * x[i][c];
            t = fmaf(k.k[0][2], x[i][c + 1], t);
            t = fmaf(k.k[2][0], x[i + 1][c], t);
            o[2 * c + 1] = fmaf(k.k[2][2], x[i + 1][c + 1], t);
        }
        TO *r0 = dst + (long)(2 * i) * W2;
        st8(r0, e);
        st8(r0 + W2, o);
    }
}

// down-like: thread = V output columns (4, or 8 for bf16 so that a thread still moves 128-bit
// vectors both ways); input rows 2i-1, 2i, 2i+1, columns 2j-1 .. 2j+2V-1
template <typename T, int V>
__device__ __forceinline__ void load_down_row(const T *plane, int row, int H, int W, int c0, bool has_l,
                                              float (&v)[2 * V + 1])
{
    if (row < 0 || row >= H) {
#pragma unroll
        for (int c = 0; c < 2 * V + 1; ++c) v[c] = 0.f;
        return;
    }
    const T *p = plane + (long)row * W + c0;
#pragma unroll
    for (int h = 0; h < V / 4; ++h) {          // the wide loads go first (issue order = arrival order)
        float w[8];
        ld8(p + 8 * h, w);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[8 * h + c + 1] = w[c];
    }
    v[0] = has_l ? ld1(p - 1) : 0.f;
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
down3_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int H, int W, int Ho,
             int Wo, int strips, int nseg, int R, int C, long in_bstride, const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_plane = (long)strips * nseg;
    if (idx >= planes * per_plane) return;
    const int s = (int)(idx % strips);
    const int seg = (int)((idx / strips) % nseg);
    const long p = idx / per_plane;
    const int j = V * s, i0 = seg * R, i1 = min(Ho, i0 + R);
    const T *plane = in + strided_base((unsigned)p, C, in_bstride, (long)H * W);      // C == 0: dense input
    T *dst = out + p * (long)Ho * Wo + j;
    const bool has_l = (j > 0);

    float vp[2 * V + 1], ve[2 * V + 1], vo[2 * V + 1];
    load_down_row<T, V>(plane, 2 * i0 - 1, H, W, 2 * j, has_l, vp);
    for (int i = i0; i < i1; ++i) {
        // bf16: the loop exposes one memory latency per output row with few bytes in flight per thread
        // (latency-bound at 24 warps/SM, ncu: long_scoreboard): pull the rows of the iteration after
        // next into L2 (+20 %).  The fp32 kernel has twice the bytes in flight and loses 7 % with it.
        if (sizeof(T) == 2 && 2 * i + 5 < H) {
            const T *pf = plane + (long)(2 * i + 4) * W + 2 * j;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + W));
        }
        load_down_row<T, V>(plane, 2 * i, H, W, 2 * j, has_l, ve);
        load_down_row<T, V>(plane, 2 * i + 1, H, W, 2 * j, has_l, vo);
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float acc = k.k[0][0] * vp[2 * q];
            acc = fmaf(k.k[0][1], vp[2 * q + 1], acc);
            acc = fmaf(k.k[0][2], vp[2 * q + 2], acc);
            acc = fmaf(k.k[1][0], ve[2 * q], acc);
            acc = fmaf(k.k[1][1], ve[2 * q + 1], acc);
            acc = fmaf(k.k[1][2], ve[2 * q + 2], acc);
            acc = fmaf(k.k[2][0], vo[2 * q], acc);
            acc = fmaf(k.k[2][1], vo[2 * q + 1], acc);
            acc = fmaf(k.k[2][2], vo[2 * q + 2], acc);
            o[q] = acc;
        }
        if (V == 8) {
            float o8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o8[q] = o[q % V];
            st8(dst + (long)i * Wo, o8);
        } else {
            st4(dst + (long)i * Wo, make_float4(o[0], o[1], o[2], o[3]));
        }
#pragma unroll
        for (int c = 0; c < 2 * V + 1; ++c) vp[c] = vo[c];
    }
}

// H = 8 input rows -> 4 output rows
template <typename T, int HH>
__global__ void __launch_bounds__(256)
down3_short_kernel(const T *__restrict__ in, T *__restrict__ out, long planes, int W, int Wo, int strips,
                   const __grid_constant__ Taps3 k)
{
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= planes * strips) return;
    const int s = (int)(idx % strips);
    const long p = idx / strips;
    const int j = 4 * s;
    const T *plane = in + p * (long)HH * W;
    T *dst = out + p * (long)(HH / 2) * Wo + j;
    const bool has_l = (j > 0);
    float v[HH + 1][9];                             // v[r + 1] = input row r; v[0] = row -1 (zero padding)
#pragma unroll
    for (int c = 0; c < 9; ++c) v[0][c] = 0.f;
#pragma unroll
    for (int r = 0; r < HH; ++r) load_down_row<T, 4>(plane, r, HH, W, 2 * j, has_l, v[r + 1]);
#pragma unroll
    for (int i = 0; i < HH / 2; ++i) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float acc = k.k[0][0] * v[2 * i][2 * q];
            acc = fmaf(k.k[0][1], v[2 * i][2 * q + 1], acc);
            acc = fmaf(k.k[0][2], v[2 * i][2 * q + 2], acc);
            acc = fmaf(k.k[1][0], v[2 * i + 1][2 * q], acc);
            acc = fmaf(k.k[1][1], v[2 * i + 1][2 * q + 1], acc);
            acc = fmaf(k.k[1][2], v[2 * i + 1][2 * q + 2], acc);
            acc = fmaf(k.k[2][0], v[2 * i + 2][2 * q], acc);
            acc = fmaf(k.k[2][1], v[2 * i + 2][2 * q + 1], acc);
            acc = fmaf(k.k[2][2], v[2 * i + 2][2 * q + 2], acc);
            o[q] = acc;
        }
        st4(dst + (long)i * Wo, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline size_t esize(int dtype) { return dtype == AFR_F32 ? 4 : 2; }

bool n3_fgelu_supported(int H, int W, const void *const *ptrs, int nptrs, int dtype)
{
    if (H < 1 || W < 4 || (W % 4) != 0) return false;
    for (int i = 0; i < nptrs; ++i)
        if (ptrs[i] && !aligned_to(ptrs[i], 4 * esize(dtype))) return false;
    return true;
}

static bool pick_tile(long planes, int H, int W, int dtype, int nin, int *threads, struct TileCfg *cfg);

// tuning knobs, read once per process (0 = unset): CTA size and row segments of the TMA kernel
static int env_cta_threads()
{
    static const int v = []() { const char *e = getenv("AFR_CTA_THREADS"); return e ? atoi(e) : 0; }();
    return v;
}
static int env_nsegs()
{
    static const int v = []() { const char *e = getenv("AFR_NSEGS"); return e ? atoi(e) : 0; }();
    return v;
}

static int tma_min_width()
{
    static int v = []() { const char *e = getenv("AFR_TMA_MIN_W"); int w = e ? atoi(e) : 8; return w < 4 ? 4 : w; }();
    return v;
}

static bool plane_kernels_disabled()   // AFR_NO_PLANE=1: A/B runs of the general kernels on 4x4 / 8x8 planes
{
    static const bool v = []() { const char *e = getenv("AFR_NO_PLANE"); return e && atoi(e) != 0; }();
    return v;
}

// AUTO sends these shapes to the whole-plane register kernels (launch_direct) even though TMA could stage them
bool n3_prefers_plane_kernel(int H, int W, bool bwd)
{
    if (plane_kernels_disabled()) return false;
    (void)bwd;
    return (H == 4 && W == 4) || (H == 8 && W == 8);
}

bool n3_fgelu_tma_supported(long planes, int H, int W, const void *const *ptrs, int nptrs, int dtype,
                            int n_inputs)
{
    if (!n3_fgelu_supported(H, W, ptrs, nptrs, dtype)) return false;
    if (H < 2 || W < tma_min_width()) return false;           // narrow planes: direct path
    if ((W * esize(dtype)) % 16 != 0) return false;           // TMA global strides
    for (int i = 0; i < nptrs - 1; ++i)                       // inputs only (last ptr = output)
        if (ptrs[i] && !aligned_to(ptrs[i], 16)) return false;
    int threads; TileCfg cfg;
    return pick_tile(planes, H, W, dtype, n_inputs, &threads, &cfg);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
                cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(sym);
    }();
    return fn;
}

// A tensor map encodes nothing but (base address, shape, strides, box, element type): the same key always
// yields the same 128 bytes, whatever has happened to the allocation in between.  PyTorch's caching allocator
// hands the same addresses back step after step, so the descriptors of a UNet's ~50 fused launches are
// encoded once and then found in this per-thread, direct-mapped cache (no locks; the library keeps no other
// state -- SURVEY.md section 8b).  AFR_NO_DESC_CACHE=1 disables it (A/B timing).
struct MapKey {
    const void *base;
    long planes;               // planes (rank-3 NCHW map) or images (rank-4 NHWC map)
    int H, W, dtype, bw, bh, bp;
    int C;                     // 0: rank-3 map over [planes, H, W]; > 0: rank-4 map over [B, H, W, C]
    bool operator==(const MapKey &o) const
    {
        return base == o.base && planes == o.planes && H == o.H && W == o.W && dtype == o.dtype && bw == o.bw &&
               bh == o.bh && bp == o.bp && C == o.C;
    }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool valid;
};
static constexpr int kMapSlots = 512;

static bool desc_cache_disabled()
{
    static const bool v = []() { const char *e = getenv("AFR_NO_DESC_CACHE"); return e && atoi(e) != 0; }();
    return v;
}

static bool encode_plane_map(CUtensorMap *m, const void *base, long planes, int H, int W, int dtype,
                             int box_w, int box_h, int box_p)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_detail("cuTensorMapEncodeTiled entry point not found"); return false; }
    const size_t es = esize(dtype);
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * es, (cuuint64_t)W * H * es};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = CUDA_SUCCESS;
    for (int attempt = 0; attempt < 2; ++attempt) {
        r = enc(m, dtype == AFR_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_ERROR_INVALID_CONTEXT) break;
        // A thread that has made no runtime call yet (e.g. PyTorch's autograd worker on its first
        // backward) has no context bound for this driver-API call: bind the primary context.
        cudaFree(0);
    }
    if (r != CUDA_SUCCESS)
        set_detail("cuTensorMapEncodeTiled failed (CUresult %d) base=%p dims=[%d,%d,%ld] box=[%d,%d,%d]", (int)r,
                   base, W, H, planes, box_w, box_h, box_p);
    return r == CUDA_SUCCESS;
}

// rank-4 map over a channels-last tensor: dims {C, W, H, B}, box {32 channels, box_w columns, box_h rows, box_p images}
static bool encode_nhwc_map(CUtensorMap *m, const void *base, long B, int C, int H, int W, int dtype, int box_w, int box_h,
                            int box_p)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_detail("cuTensorMapEncodeTiled entry point not found"); return false; }
    const size_t es = esize(dtype);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {32u, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = CUDA_SUCCESS;
    for (int attempt = 0; attempt < 2; ++attempt) {
        r = enc(m, dtype == AFR_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_ERROR_INVALID_CONTEXT) break;
        cudaFree(0);                               // bind the primary context on a fresh thread (see encode_plane_map)
    }
    if (r != CUDA_SUCCESS)
        set_detail("cuTensorMapEncodeTiled (NHWC) failed (CUresult %d) base=%p dims=[%d,%d,%d,%ld] box=[32,%d,%d,%d]", (int)r,
                   base, C, W, H, B, box_w, box_h, box_p);
    return r == CUDA_SUCCESS;
}

static bool make_plane_map(CUtensorMap *m, const void *base, long planes, int H, int W, int dtype,
                           int box_w, int box_h, int box_p, int C = 0)
{
    if (desc_cache_disabled())
        return C > 0 ? encode_nhwc_map(m, base, planes, C, H, W, dtype, box_w, box_h, box_p)
                     : encode_plane_map(m, base, planes, H, W, dtype, box_w, box_h, box_p);
    // allocated on a thread's first TMA launch, freed with the thread (over-aligned type: aligned operator new)
    static thread_local std::unique_ptr<MapSlot[]> slots;
    if (!slots) slots.reset(new (std::nothrow) MapSlot[kMapSlots]());
    if (!slots)
        return C > 0 ? encode_nhwc_map(m, base, planes, C, H, W, dtype, box_w, box_h, box_p)
                     : encode_plane_map(m, base, planes, H, W, dtype, box_w, box_h, box_p);
    const MapKey key = {base, planes, H, W, dtype, box_w, box_h, box_p, C};
    uint64_t h = (uint64_t)reinterpret_cast<uintptr_t>(base) >> 8;     // allocations are 512-byte aligned
    h ^= (uint64_t)planes * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)H << 40) ^ ((uint64_t)W << 20) ^ ((uint64_t)box_h << 8) ^ (uint64_t)box_p ^ ((uint64_t)dtype << 60) ^
         ((uint64_t)C << 30);
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    MapSlot &sl = slots[h % kMapSlots];
    if (sl.valid && sl.key == key) { *m = sl.map; return true; }
    if (!(C > 0 ? encode_nhwc_map(m, base, planes, C, H, W, dtype, box_w, box_h, box_p)
                : encode_plane_map(m, base, planes, H, W, dtype, box_w, box_h, box_p)))
        return false;
    sl.key = key; sl.map = *m; sl.valid = true;
    return true;
}

static inline int pick_rows(int H) { return H < 8 ? H : 8; }


template <typename T, bool kBwd, bool kRes, bool kAff, class KT>
static cudaError_t launch_direct(const void *x, const void *res, const void *dy, const float *scale,
                                 const float *shift, void *out, long planes, int H, int W,
                                 const KT &K, cudaStream_t s)
{
    const int block = 128;
    if (H == 4 && W == 4 && !plane_kernels_disabled()) {          // one thread per plane, everything unrolled
        const long g4 = (planes + block - 1) / block;
        if (g4 > 0x7fffffffL) { set_detail("too many planes"); return cudaErrorInvalidConfiguration; }
        fgelu3_plane_kernel<T, kBwd, kRes, kAff, KT, 4><<<(unsigned)g4, block, 0, s>>>(
            (const T *)x, (const T *)res, (const T *)dy, scale, shift, (T *)out, planes, K);
        return cudaGetLastError();
    }
    if (H == 8 && W == 8 && !plane_kernels_disabled()) {          // a lane pair per plane
        const long g8 = (2 * planes + block - 1) / block;
        if (g8 > 0x7fffffffL) { set_detail("too many planes"); return cudaErrorInvalidConfiguration; }
        fgelu3_plane_kernel<T, kBwd, kRes, kAff, KT, 8><<<(unsigned)g8, block, 0, s>>>(
            (const T *)x, (const T *)res, (const T *)dy, scale, shift, (T *)out, planes, K);
        return cudaGetLastError();
    }
    const int strips = W / 4, R = pick_rows(H), nseg = (H + R - 1) / R;
    const long total = planes * (long)strips * nseg;
    const long grid = (total + block - 1) / block;
    if (total >= 0x7fffffffL) { set_detail("tensor too large for the direct kernel's 32-bit indexing"); return cudaErrorInvalidConfiguration; }
    fgelu3_direct_kernel<T, kBwd, kRes, kAff, KT><<<(unsigned)grid, block, 0, s>>>(
        (const T *)x, (const T *)res, (const T *)dy, scale, shift, (T *)out, planes, H, W, strips, nseg, R, K);
    return cudaGetLastError();
}

// Shared-memory budget of one CTA's two-stage ring.  ~52 KB leaves room for four CTAs per SM
// (16 warps); AFR_RING_KB overrides it for tuning runs.
static size_t ring_budget_bytes()
{
    static size_t v = []() -> size_t {
        const char *e = getenv("AFR_RING_KB");
        int kb = e ? atoi(e) : 52;
        if (kb < 8) kb = 8;
        if (kb > 96) kb = 96;
        return (size_t)kb * 1024;
    }();
    return v;
}

static bool pick_tile(long planes, int H, int W, int dtype, int nin, int *threads, TileCfg *cfg)
{
    const size_t es = esize(dtype);
    TileCfg c;
    c.Tw = W <= 128 ? W : 128;
    c.strips = c.Tw / 4;
    // wide planes: tiles of 32 strips whose first 16 bytes' worth of strips are ghosts (1 for fp32,
    // 2 for bf16 -- keeps every box start 16-byte aligned), see TileCfg::ghost.  Only used when it
    // does not cost an extra, mostly idle tile (W = 256, 512 split evenly into 128-wide tiles and
    // keep the recompute branch instead).
    c.ghost = 0;
    c.tiles_x = (W + c.Tw - 1) / c.Tw;
    if (W > 128) {
        const int g = (int)(16 / (4 * es)), stride = c.Tw - 4 * g;
        if ((W + stride - 1) / stride == c.tiles_x) c.ghost = g;
    }
    // 128-thread CTAs unless that leaves fewer than ~6 CTAs per SM: long-lived streaming CTAs
    // need a few waves to balance, so small batches of large planes use smaller CTAs
    int th = 128;
    if (env_cta_threads() > 0) th = env_cta_threads();   // tuning runs
    while (th > 32 && th > c.strips && (planes * c.tiles_x * c.strips) / th < 6 * 148) th /= 2;
    if (th < c.strips) th = c.strips;
    if (th > 128) return false;
    // chunk height 8 rows (4 if the ring of all staged inputs would not fit the budget); if even
    // that does not fit, stage fewer planes per CTA; 2-row chunks are the last resort
    for (int min_r = 4; min_r >= 2; min_r /= 2) {
        for (int t = th; t >= c.strips && t >= 32; t /= 2) {      // whole warps only (full-mask shuffles)
            c.P = t / c.strips;
            for (c.R = 8; c.R >= min_r; c.R /= 2) {
                const size_t bytes = (size_t)(c.Tw + 2 * (16 / es)) * (c.R + 1) * c.P * es;
                c.tile_bytes = (int)((bytes + 127) / 128 * 128);
                if ((size_t)c.tile_bytes * nin * 2 <= ring_budget_bytes()) {
                    // row segments: when whole-plane streaming gives fewer than ~2.5 waves of CTAs the SMs
                    // finish unevenly (long-lived CTAs, no second wave to even things out); each extra
                    // segment costs one dry step, so segments stay at least 2R rows tall
                    const long ctas = (planes + c.P - 1) / c.P * c.tiles_x;
                    long resident = (long)(220 * 1024) / ((long)c.tile_bytes * nin * 2);
                    resident = resident < 1 ? 1 : resident;
                    if (resident > 65536 / (100 * t)) resident = 65536 / (100 * t);
                    if (resident > 32) resident = 32;
                    c.nsegs = 1;
                    while (2 * ctas * c.nsegs < 5 * 148 * resident && H / (c.nsegs * 2) >= 2 * c.R && c.nsegs < 16) c.nsegs *= 2;
                    // kernels that stage two or three tensors hold fewer CTAs per SM, which makes the rule above stop
                    // early; tall planes still gain from the CTA count a one-input launch would get as long as a
                    // segment keeps >= 64 rows (one dry step per segment): [16,64,256,256] adjoint 0.56 -> 0.60
                    if (nin > 1) {
                        long res1 = (long)(220 * 1024) / ((long)c.tile_bytes * 2);
                        if (res1 > 65536 / (100 * t)) res1 = 65536 / (100 * t);
                        if (res1 > 32) res1 = 32;
                        while (2 * ctas * c.nsegs < 5 * 148 * res1 && H / (c.nsegs * 2) >= 64 && c.nsegs < 16) c.nsegs *= 2;
                    }
                    if (env_nsegs() > 0) c.nsegs = env_nsegs();   // tuning runs
                    c.Hs = ((H + c.nsegs - 1) / c.nsegs + c.R - 1) / c.R * c.R;
                    c.nsegs = (H + c.Hs - 1) / c.Hs;
                    c.own0_any = (c.ghost == 0) && (c.tiles_x > 1 || (32 % c.strips) != 0);
                    *threads = t; *cfg = c;
                    return true;
                }
            }
        }
    }
    return false;
}

template <typename T, bool kBwd, bool kRes, bool kAff, class KT>
static cudaError_t launch_tma(const void *x, const void *res, const void *dy, const float *scale,
                              const float *shift, void *out, long planes, int H, int W, int dtype,
                              const KT &K, cudaStream_t s)
{
    const int nin = 1 + (kRes ? 1 : 0) + (kBwd ? 1 : 0);
    int threads; TileCfg cfg;
    if (!pick_tile(planes, H, W, dtype, nin, &threads, &cfg)) { set_detail("no tile configuration"); return cudaErrorInvalidConfiguration; }
    CUtensorMap mx, mres, mdy;
    const int bw = cfg.Tw + 2 * Halo<T>::value, bh = cfg.R + 1;
    if (!make_plane_map(&mx, x, planes, H, W, dtype, bw, bh, cfg.P)) return cudaErrorInvalidValue;
    mres = mx; mdy = mx;
    if (kRes && !make_plane_map(&mres, res, planes, H, W, dtype, bw, bh, cfg.P)) return cudaErrorInvalidValue;
    if (kBwd && !make_plane_map(&mdy, dy, planes, H, W, dtype, bw, bh, cfg.P)) return cudaErrorInvalidValue;
    const long pgroups = (planes + cfg.P - 1) / cfg.P;
    const long grid = pgroups * cfg.tiles_x;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const dim3 grid3((unsigned)grid, (unsigned)cfg.nsegs);
    const size_t smem = (size_t)cfg.tile_bytes * nin * 2;
    auto kern = fgelu3_tma_kernel<T, kBwd, kRes, kAff, KT>;
    // once per instantiation and device
    static std::atomic<unsigned long long> attr_done{0};
    if (cudaError_t ae = ensure_dyn_smem(kern, attr_done, 100 * 1024)) { set_detail("cudaFuncSetAttribute(max dynamic smem) failed"); return ae; }
    kern<<<grid3, threads, smem, s>>>(mx, mres, mdy, scale, shift, (T *)out, planes, H, W, cfg, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        set_detail("launch grid=%ld block=%d smem=%zu tile Tw=%d R=%d P=%d", grid, threads, smem, cfg.Tw, cfg.R, cfg.P);
    return e;
}

// ---- symmetric-tap fast path: eligibility and host-side coefficient folding ---------------------
static bool d4_symmetric(const Taps3 &t)
{
    return t.k[0][0] == t.k[0][2] && t.k[0][0] == t.k[2][0] && t.k[0][0] == t.k[2][2] &&
           t.k[0][1] == t.k[2][1] && t.k[1][0] == t.k[1][2];
}

static void phase_taps(const Taps3 &t, double (&s)[4])
{
    s[PH_EE] = t.k[1][1]; s[PH_EO] = t.k[1][0]; s[PH_OE] = t.k[0][1]; s[PH_OO] = t.k[0][0];
}

// false: the taps do not qualify (asymmetric, or a negative / non-finite up tap in the forward,
// where gelu(s xi) = s ghat(xi) needs s >= 0) and the general kernels run instead
static bool make_sym(const Taps3 &kU, const Taps3 &kG, const Taps3 &kB, bool bwd, bool low, SymK *K)
{
    if (!d4_symmetric(kU) || !d4_symmetric(kB) || (bwd && !d4_symmetric(kG))) return false;
    double su[4], sb[4], sg[4];
    phase_taps(kU, su); phase_taps(kB, sb); phase_taps(kG, sg);
    static const double P5[5] = {-(double)AFR_P5, (double)AFR_P4, -(double)AFR_P3, (double)AFR_P2, -(double)AFR_P1};
    static const double P4[5] = {0.0, (double)AFR_Q4, -(double)AFR_Q3, (double)AFR_Q2, -(double)AFR_Q1};
    const double *P = low ? P4 : P5;      // bf16 tensors: the degree-4 form, |xi| clamped to T / s_ph
    for (int ph = 0; ph < 4; ++ph) {
        if (!bwd && !(su[ph] >= 0.0 && su[ph] < 1e30)) return false;
        double pw = su[ph] * su[ph] * su[ph] * su[ph] * su[ph];          // s^5 .. s^1
        for (int k = 0; k < 5; ++k) {
            K->p[ph][k] = (float)(P[k] * pw);
            pw = su[ph] != 0.0 ? pw / su[ph] : 0.0;
        }
        K->cl[ph] = su[ph] > 0.0 ? (float)(-(double)AFR_LOW_T / su[ph]) : -3.0e38f;
        K->sw[ph] = (float)su[ph];                                       // adjoint: kU arrives kappa-scaled
        K->dn[ph] = (float)(sb[ph] * (bwd ? sg[ph] : su[ph]));
    }
    return true;
}

cudaError_t n3_fgelu(const void *x, const void *res, const void *dy, const float *scale, const float *shift,
                     void *out, long planes, int H, int W, const Taps3 &kU, const Taps3 &kG, const Taps3 &kB,
                     bool bwd, int dtype, bool use_tma, bool allow_sym, cudaStream_t s, const char **kernel_name)
{
#define AFR_GO(T, B, R, A)                                                                                          \
    do {                                                                                                            \
        if (sym) return use_tma ? launch_tma<T, B, R, A, SymK>(x, res, dy, scale, shift, out, planes, H, W, dtype, KS, s) \
                                : launch_direct<T, B, R, A, SymK>(x, res, dy, scale, shift, out, planes, H, W, KS, s);   \
        return use_tma ? launch_tma<T, B, R, A, StepK>(x, res, dy, scale, shift, out, planes, H, W, dtype, KG, s)   \
                       : launch_direct<T, B, R, A, StepK>(x, res, dy, scale, shift, out, planes, H, W, KG, s);      \
    } while (0)
#define AFR_GO_T(T)                                                                                          \
    if (bwd && scale) { if (res) { AFR_GO(T, true, true, true); } else { AFR_GO(T, true, false, true); } }   \
    else if (bwd) { if (res) { AFR_GO(T, true, true, false); } else { AFR_GO(T, true, false, false); } }     \
    else if (scale) { if (res) { AFR_GO(T, false, true, true); } else { AFR_GO(T, false, false, true); } }   \
    else { if (res) { AFR_GO(T, false, true, false); } else { AFR_GO(T, false, false, false); } }
    SymK KS;
    const StepK KG = {kU, kG, kB};
    const bool sym = allow_sym && make_sym(kU, kG, kB, bwd, dtype == AFR_BF16 && AFR_BF16_LOW_FWD, &KS);
    if (kernel_name)
        *kernel_name = use_tma ? (sym ? "fgelu3_tma_kernel<sym>" : "fgelu3_tma_kernel")
                               : (sym ? "fgelu3_direct_kernel<sym>" : "fgelu3_direct_kernel");
    if (dtype == AFR_F32) { AFR_GO_T(float) }
    AFR_GO_T(bf16)
#undef AFR_GO_T
#undef AFR_GO
}

bool n3_up_supported(int H, int W, const void *in, const void *out, int in_dtype, int out_dtype)
{
    return H >= 1 && W >= 4 && (W % 4) == 0 && aligned_to(in, 4 * esize(in_dtype)) &&
           aligned_to(out, 8 * esize(out_dtype));      // 8 outputs per thread row: 128-bit stores
}

cudaError_t n3_up_like(const void *in, void *out, long planes, int H, int W, const Taps3 &k,
                       int in_dtype, int out_dtype, cudaStream_t s, int C, long out_bstride)
{
    if (planes > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (flat_up_wanted(H, W, in_dtype, out_dtype)) {
        const cudaError_t e = flat_up_like(in, out, planes, C, out_bstride, H, W, k, in_dtype, out_dtype, s);
        if (e != cudaErrorInvalidConfiguration) return e;        // more than 2^32 threads: the strips below take it
    }
    const int strips = W / 4, R = pick_rows(H), nseg = (H + R - 1) / R;
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    // bf16 input only: with fp32 input the up-front loads are 15-60 % slower than the row loop (measured)
    const bool short_plane = (H == 4 || H == 8) && in_dtype == AFR_BF16 && !plane_kernels_disabled() && C == 0;
    const long sgrid = (planes * strips + 255) / 256;
#define AFR_UP(TI, TO)                                                                                       \
    do {                                                                                                     \
        if (short_plane && H == 4)                                                                           \
            up3_short_kernel<TI, TO, 4><<<(unsigned)sgrid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, W, strips, k); \
        else if (short_plane)                                                                                \
            up3_short_kernel<TI, TO, 8><<<(unsigned)sgrid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, W, strips, k); \
        else                                                                                                 \
            up3_kernel<TI, TO><<<(unsigned)grid, 256, 0, s>>>((const TI *)in, (TO *)out, planes, H, W,        \
                                                              strips, nseg, R, C, out_bstride, k);           \
    } while (0)
    if (in_dtype == AFR_F32 && out_dtype == AFR_F32) AFR_UP(float, float);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_BF16) AFR_UP(bf16, bf16);
    else if (in_dtype == AFR_BF16 && out_dtype == AFR_F32) AFR_UP(bf16, float);
    else AFR_UP(float, bf16);
#undef AFR_UP
    return cudaGetLastError();
}

bool n3_down_supported(int H, int W, const void *in, const void *out, int dtype)
{
    return H >= 1 && W >= 8 && (W % 8) == 0 && aligned_to(in, 8 * esize(dtype)) &&     // 8-element loads
           aligned_to(out, 4 * esize(dtype));
}

cudaError_t n3_down_like(const void *in, void *out, long planes, int H, int W, const Taps3 &k,
                         int dtype, cudaStream_t s, int C, long in_bstride)
{
    if (planes > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (flat_down_wanted(H, W, dtype)) return flat_down_like(in, out, planes, C, in_bstride, H, W, k, dtype, s);
    const int Ho = (H + 1) / 2, Wo = W / 2;
    // bf16: 8 outputs per thread when the row width and the bases allow 128-bit vectors both ways
    const bool wide = dtype == AFR_BF16 && (Wo % 8) == 0 && aligned_to(in, 16) && aligned_to(out, 16);
    const int V = wide ? 8 : 4;
    const int strips = Wo / V, R = pick_rows(Ho), nseg = (Ho + R - 1) / R;
    const long total = planes * (long)strips * nseg;
    const long grid = (total + 255) / 256;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    if (H == 8 && dtype == AFR_BF16 && !plane_kernels_disabled() && C == 0) {   // short bf16 planes: all rows up front, unrolled (+16 %; fp32: no gain)
        const int sstrips = Wo / 4;
        const long sgrid = (planes * sstrips + 255) / 256;
        if (sgrid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
        down3_short_kernel<bf16, 8><<<(unsigned)sgrid, 256, 0, s>>>((const bf16 *)in, (bf16 *)out, planes, W, Wo, sstrips, k);
        return cudaGetLastError();
    }
    if (dtype == AFR_F32)
        down3_kernel<float, 4><<<(unsigned)grid, 256, 0, s>>>((const float *)in, (float *)out, planes,
                                                              H, W, Ho, Wo, strips, nseg, R, C, in_bstride, k);
    else if (wide)
        down3_kernel<bf16, 8><<<(unsigned)grid, 256, 0, s>>>((const bf16 *)in, (bf16 *)out, planes, H,
                                                             W, Ho, Wo, strips, nseg, R, C, in_bstride, k);
    else
        down3_kernel<bf16, 4><<<(unsigned)grid, 256, 0, s>>>((const bf16 *)in, (bf16 *)out, planes, H,
                                                             W, Ho, Wo, strips, nseg, R, C, in_bstride, k);
    return cudaGetLastError();
}

// ---- NHWC host side ------------------------------------------------------------------------------------
bool nhwc_fgelu_supported(int C, int H, int W, const void *const *ptrs, int nptrs)
{
    if (C < 32 || (C % 32) != 0 || H < 2 || W < 4 || (W % 4) != 0) return false;
    for (int i = 0; i < nptrs; ++i)
        if (ptrs[i] && !aligned_to(ptrs[i], 16)) return false;
    return true;
}

static bool pick_nhwc(long B, int C, int H, int W, int dtype, int nin, NhwcCfg *cfg)
{
    const size_t es = esize(dtype);
    NhwcCfg c;
    c.strips = W >= 12 ? 4 : (W >= 8 ? 2 : 1);
    c.P = 4 / c.strips;
    c.tiles_x = (W + 4 * c.strips - 1) / (4 * c.strips);
    c.cgroups = C / 32;
    for (c.R = 8; c.R >= 2; c.R /= 2) {
        const size_t bytes = (size_t)32 * (4 * c.strips + 2) * (c.R + 1) * c.P * es;
        c.tile_bytes = (int)((bytes + 127) / 128 * 128);
        if ((size_t)c.tile_bytes * nin * 2 <= ring_budget_bytes() || c.R == 2) break;
    }
    if ((size_t)c.tile_bytes * nin * 2 > 96 * 1024) return false;
    const long ctas = (B + c.P - 1) / c.P * c.tiles_x * c.cgroups;
    long resident = (long)(220 * 1024) / ((long)c.tile_bytes * nin * 2);
    resident = resident < 1 ? 1 : (resident > 5 ? 5 : resident);
    c.nsegs = 1;
    while (2 * ctas * c.nsegs < 5 * 148 * resident && H / (c.nsegs * 2) >= 2 * c.R && c.nsegs < 16) c.nsegs *= 2;
    c.Hs = ((H + c.nsegs - 1) / c.nsegs + c.R - 1) / c.R * c.R;
    c.nsegs = (H + c.Hs - 1) / c.Hs;
    *cfg = c;
    return true;
}

template <typename T, bool kBwd, bool kRes>
static cudaError_t launch_nhwc(const void *x, const void *res, const void *dy, const float *scale, const float *shift,
                               void *out, long B, int C, int H, int W, int dtype, const SymK &K, cudaStream_t s)
{
    const int nin = 1 + (kRes ? 1 : 0) + (kBwd ? 1 : 0);
    NhwcCfg cfg;
    if (!pick_nhwc(B, C, H, W, dtype, nin, &cfg)) { set_detail("no NHWC tile configuration"); return cudaErrorInvalidConfiguration; }
    CUtensorMap mx, mres, mdy;
    const int bw = 4 * cfg.strips + 2, bh = cfg.R + 1;
    if (!make_plane_map(&mx, x, B, H, W, dtype, bw, bh, cfg.P, C)) return cudaErrorInvalidValue;
    mres = mx; mdy = mx;
    if (kRes && !make_plane_map(&mres, res, B, H, W, dtype, bw, bh, cfg.P, C)) return cudaErrorInvalidValue;
    if (kBwd && !make_plane_map(&mdy, dy, B, H, W, dtype, bw, bh, cfg.P, C)) return cudaErrorInvalidValue;
    const long grid = (B + cfg.P - 1) / cfg.P * cfg.tiles_x * cfg.cgroups;
    if (grid > 0x7fffffffL) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)cfg.tile_bytes * nin * 2;
    auto kern = fgelu3_nhwc_kernel<T, kBwd, kRes>;
    static std::atomic<unsigned long long> attr_done{0};
    if (cudaError_t ae = ensure_dyn_smem(kern, attr_done, 100 * 1024)) { set_detail("cudaFuncSetAttribute(max dynamic smem) failed"); return ae; }
    kern<<<dim3((unsigned)grid, (unsigned)cfg.nsegs), 128, smem, s>>>(mx, mres, mdy, scale, shift, (T *)out, (int)B, C, H, W, cfg, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_detail("NHWC launch grid=%ld smem=%zu strips=%d R=%d P=%d", grid, smem, cfg.strips, cfg.R, cfg.P);
    return e;
}

// channels-last fused forward / adjoint.  Returns cudaErrorNotSupported when the taps are not D4-symmetric (or, in
// the forward, the up filter has a negative tap): the caller then takes the NCHW kernels.
cudaError_t nhwc_fgelu(const void *x, const void *res, const void *dy, const float *scale, const float *shift, void *out,
                       long B, int C, int H, int W, const Taps3 &kU, const Taps3 &kG, const Taps3 &kB, bool bwd, int dtype,
                       cudaStream_t s)
{
    SymK KS;
    if (!make_sym(kU, kG, kB, bwd, dtype == AFR_BF16 && AFR_BF16_LOW_FWD, &KS)) return cudaErrorNotSupported;
    if (B > 0x7fffffffL) return cudaErrorInvalidConfiguration;
#define AFR_NH(T)                                                                                                   \
    do {                                                                                                            \
        if (bwd) return res ? launch_nhwc<T, true, true>(x, res, dy, scale, shift, out, B, C, H, W, dtype, KS, s)     \
                            : launch_nhwc<T, true, false>(x, res, dy, scale, shift, out, B, C, H, W, dtype, KS, s);   \
        return res ? launch_nhwc<T, false, true>(x, res, dy, scale, shift, out, B, C, H, W, dtype, KS, s)            \
                   : launch_nhwc<T, false, false>(x, res, dy, scale, shift, out, B, C, H, W, dtype, KS, s);          \
    } while (0)
    if (dtype == AFR_F32) AFR_NH(float);
    AFR_NH(bf16);
#undef AFR_NH
}

}  // namespace afr
