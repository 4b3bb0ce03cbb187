/*
 * afr_oracle.c -- CPU restatement of the alias-free resampling path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA product in
 * aliasfree-diffusion-models-pytorch_b200/csrc; nothing in the product path may
 * link or call it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  The reference holds no golden vectors or tests of its
 * own, so the oracle is pinned against outputs of the
 * reference itself, generated in the build container by
 * tests/golden/make_golden.py (which imports /root/reference) and committed
 * under tests/golden/.  tests/test_oracle.py compares every function
 * below with those fixtures.
 *
 * What each function follows (paths are into the reference repository):
 *   afr_oracle_lowpass_taps   modules/filtrs.py:20-37  (circularLowpassKernel)
 *   afr_oracle_up2x           modules/filtrs.py:79-94  (custom_upsample)
 *   afr_oracle_down2x         modules/filtrs.py:71-77  (custom_downsample)
 *   afr_oracle_filtered_gelu  modules/ddpm_utils.py:123-125,129-131,137-139
 *   *_bwd                     the autograd adjoints of the three above
 *   afr_oracle_rotate         modules/ddpm_models.py:421-429 (scipy.ndimage.rotate,
 *                             order 3, mode='grid-wrap', prefilter on; SciPy is a
 *                             third-party dependency not vendored in the reference,
 *                             unpinned there; the image has SciPy 1.18.1)
 *
 * Conventions: tensors are dense [planes][H][W] float32 (planes = B*C, i.e. NCHW
 * with the two leading dims folded).  Accumulation is in double and rounded once,
 * which is at least as accurate as the reference's fp32 conv.  Loops over planes
 * are OpenMP-parallel so the reference arm of bench.py can use every host core.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define AFR_MAX_TAPS 32

/* ---- low-pass filter design (filtrs.py:20-37) ------------------------------ */

/* Modified Bessel I0 by its power series; beta <= ~50 converges well in double. */
static double bessel_i0(double x)
{
    double q = 0.25 * x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

/* numpy.kaiser(M, beta): I0(beta*sqrt(1-((n-a)/a)^2))/I0(beta), a=(M-1)/2; M==1 -> 1 */
static void kaiser_window(int M, double beta, double *w)
{
    if (M == 1) { w[0] = 1.0; return; }
    double a = 0.5 * (M - 1), d = bessel_i0(beta);
    for (int n = 0; n < M; ++n) {
        double r = (n - a) / a, s = 1.0 - r * r;
        if (s < 0.0) s = 0.0;
        w[n] = bessel_i0(beta * sqrt(s)) / d;
    }
}

/* Radial jinc low-pass sampled on an N x N grid centred at (N-1)/2, optional
 * separable Kaiser window, normalised to unit sum, rounded to float32. */
int afr_oracle_lowpass_taps(double omega_c, int N, int has_beta, double beta, float *out)
{
    if (N < 1 || N > AFR_MAX_TAPS) return 1;
    double k[AFR_MAX_TAPS * AFR_MAX_TAPS], w[AFR_MAX_TAPS];
    double c = 0.5 * (N - 1), sum = 0.0;
    for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) {
            double r = sqrt((a - c) * (a - c) + (b - c) * (b - c));
            k[a * N + b] = (r == 0.0) ? omega_c * omega_c / (4.0 * M_PI)
                                      : omega_c * j1(omega_c * r) / (2.0 * M_PI * r);
        }
    if (has_beta) {
        kaiser_window(N, beta, w);
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) k[a * N + b] *= w[a] * w[b];
    }
    for (int i = 0; i < N * N; ++i) sum += k[i];
    for (int i = 0; i < N * N; ++i) out[i] = (float)(k[i] / sum);
    return 0;
}

/* ---- building blocks in double -------------------------------------------- */

static inline double gelu_erf(double v) { return 0.5 * v * (1.0 + erf(v * M_SQRT1_2)); }
static inline double gelu_erf_grad(double v)
{
    return 0.5 * (1.0 + erf(v * M_SQRT1_2)) + v * exp(-0.5 * v * v) * 0.3989422804014327;
}

/* 'same' padding split used by F.conv2d: low = (N-1)/2, high = N-1-low. */
static inline int pad_lo(int N) { return (N - 1) / 2; }

/* u[Y][X] = sum_{a,b} k[a][b] * z[Y+a-pl][X+b-pl], z[2i][2j]=x[i][j], 0 elsewhere
 * and 0 outside [0,2H)x[0,2W).  No gain (filtrs.py:85-93). */
static void up_plane(const float *x, double *u, int H, int W, const float *k, int N)
{
    int pl = pad_lo(N), H2 = 2 * H, W2 = 2 * W;
    for (int Y = 0; Y < H2; ++Y)
        for (int X = 0; X < W2; ++X) {
            double acc = 0.0;
            for (int a = 0; a < N; ++a) {
                int zy = Y + a - pl;
                if (zy < 0 || zy >= H2 || (zy & 1)) continue;
                for (int b = 0; b < N; ++b) {
                    int zx = X + b - pl;
                    if (zx < 0 || zx >= W2 || (zx & 1)) continue;
                    acc += (double)k[a * N + b] * (double)x[(zy >> 1) * W + (zx >> 1)];
                }
            }
            u[Y * W2 + X] = acc;
        }
}

/* y[i][j] = sum k[a][b] * v[2i+a-pl][2j+b-pl], v = 0 outside; Ho=ceil(H/2). */
static void down_plane(const double *v, double *y, int H, int W, const float *k, int N)
{
    int pl = pad_lo(N), Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    for (int i = 0; i < Ho; ++i)
        for (int j = 0; j < Wo; ++j) {
            double acc = 0.0;
            for (int a = 0; a < N; ++a) {
                int vy = 2 * i + a - pl;
                if (vy < 0 || vy >= H) continue;
                for (int b = 0; b < N; ++b) {
                    int vx = 2 * j + b - pl;
                    if (vx < 0 || vx >= W) continue;
                    acc += (double)k[a * N + b] * v[vy * W + vx];
                }
            }
            y[i * Wo + j] = acc;
        }
}

/* adjoint of up_plane: dx[i][j] = sum k[a][b] * du[2i-a+pl][2j-b+pl] */
static void up_plane_adj(const double *du, double *dx, int H, int W, const float *k, int N)
{
    int pl = pad_lo(N), H2 = 2 * H, W2 = 2 * W;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            double acc = 0.0;
            for (int a = 0; a < N; ++a) {
                int Y = 2 * i - a + pl;
                if (Y < 0 || Y >= H2) continue;
                for (int b = 0; b < N; ++b) {
                    int X = 2 * j - b + pl;
                    if (X < 0 || X >= W2) continue;
                    acc += (double)k[a * N + b] * du[Y * W2 + X];
                }
            }
            dx[i * W + j] = acc;
        }
}

/* adjoint of down_plane: dv[Y][X] = sum over (i,a): 2i+a-pl == Y of k[a][b]*dy[i][j] */
static void down_plane_adj(const double *dy, double *dv, int H, int W, const float *k, int N)
{
    int pl = pad_lo(N), Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    for (int Y = 0; Y < H; ++Y)
        for (int X = 0; X < W; ++X) {
            double acc = 0.0;
            for (int a = 0; a < N; ++a) {
                int t = Y - a + pl;
                if (t < 0 || (t & 1) || (t >> 1) >= Ho) continue;
                for (int b = 0; b < N; ++b) {
                    int s = X - b + pl;
                    if (s < 0 || (s & 1) || (s >> 1) >= Wo) continue;
                    acc += (double)k[a * N + b] * dy[(t >> 1) * Wo + (s >> 1)];
                }
            }
            dv[Y * W + X] = acc;
        }
}

/* ---- public entry points ---------------------------------------------------- */

void afr_oracle_up2x(const float *x, float *u, long planes, int H, int W,
                     const float *taps, int N)
{
    long n_in = (long)H * W, n_out = 4 * n_in;
#pragma omp parallel
    {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)n_out);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            up_plane(x + p * n_in, tmp, H, W, taps, N);
            for (long e = 0; e < n_out; ++e) u[p * n_out + e] = (float)tmp[e];
        }
        free(tmp);
    }
}

void afr_oracle_down2x(const float *v, float *y, long planes, int H, int W,
                       const float *taps, int N)
{
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    long n_in = (long)H * W, n_out = (long)Ho * Wo;
#pragma omp parallel
    {
        double *vin = (double *)malloc(sizeof(double) * (size_t)n_in);
        double *tmp = (double *)malloc(sizeof(double) * (size_t)n_out);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            for (long e = 0; e < n_in; ++e) vin[e] = v[p * n_in + e];
            down_plane(vin, tmp, H, W, taps, N);
            for (long e = 0; e < n_out; ++e) y[p * n_out + e] = (float)tmp[e];
        }
        free(vin); free(tmp);
    }
}

/* y = down(gelu(up(x))); the gelu output is zero-padded, i.e. u outside
 * [0,2H)x[0,2W) never exists (ddpm_utils.py:123-125 feeds a 2Hx2W tensor to conv). */
void afr_oracle_filtered_gelu(const float *x, float *y, long planes, int H, int W,
                              const float *taps_up, int N_up,
                              const float *taps_dn, int N_dn)
{
    long n = (long)H * W, n4 = 4 * n;
#pragma omp parallel
    {
        double *g = (double *)malloc(sizeof(double) * (size_t)n4);
        double *o = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            up_plane(x + p * n, g, H, W, taps_up, N_up);
            /* reference rounds u to fp32 before GELU; keep that rounding step */
            for (long e = 0; e < n4; ++e) g[e] = (double)(float)gelu_erf((double)(float)g[e]);
            down_plane(g, o, 2 * H, 2 * W, taps_dn, N_dn);
            for (long e = 0; e < n; ++e) y[p * n + e] = (float)o[e];
        }
        free(g); free(o);
    }
}

void afr_oracle_up2x_bwd(const float *du, float *dx, long planes, int H, int W,
                         const float *taps, int N)
{
    long n = (long)H * W, n4 = 4 * n;
#pragma omp parallel
    {
        double *d = (double *)malloc(sizeof(double) * (size_t)n4);
        double *o = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            for (long e = 0; e < n4; ++e) d[e] = du[p * n4 + e];
            up_plane_adj(d, o, H, W, taps, N);
            for (long e = 0; e < n; ++e) dx[p * n + e] = (float)o[e];
        }
        free(d); free(o);
    }
}

/* dy is [planes][ceil(H/2)][ceil(W/2)], dv is [planes][H][W] */
void afr_oracle_down2x_bwd(const float *dy, float *dv, long planes, int H, int W,
                           const float *taps, int N)
{
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    long n = (long)H * W, no = (long)Ho * Wo;
#pragma omp parallel
    {
        double *d = (double *)malloc(sizeof(double) * (size_t)no);
        double *o = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            for (long e = 0; e < no; ++e) d[e] = dy[p * no + e];
            down_plane_adj(d, o, H, W, taps, N);
            for (long e = 0; e < n; ++e) dv[p * n + e] = (float)o[e];
        }
        free(d); free(o);
    }
}

/* dx = up^T( gelu'(up(x)) * down^T(dy) ) */
void afr_oracle_filtered_gelu_bwd(const float *x, const float *dy, float *dx,
                                  long planes, int H, int W,
                                  const float *taps_up, int N_up,
                                  const float *taps_dn, int N_dn)
{
    long n = (long)H * W, n4 = 4 * n;
#pragma omp parallel
    {
        double *u = (double *)malloc(sizeof(double) * (size_t)n4);
        double *dg = (double *)malloc(sizeof(double) * (size_t)n4);
        double *d = (double *)malloc(sizeof(double) * (size_t)n);
        double *o = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            up_plane(x + p * n, u, H, W, taps_up, N_up);
            for (long e = 0; e < n; ++e) d[e] = dy[p * n + e];
            down_plane_adj(d, dg, 2 * H, 2 * W, taps_dn, N_dn);
            for (long e = 0; e < n4; ++e) dg[e] *= gelu_erf_grad((double)(float)u[e]);
            up_plane_adj(dg, o, H, W, taps_up, N_up);
            for (long e = 0; e < n; ++e) dx[p * n + e] = (float)o[e];
        }
        free(u); free(dg); free(d); free(o);
    }
}

/* ---- Config-E rotation (ddpm_models.py:421-429) ----------------------------
 * scipy.ndimage.rotate(x, angle, axes=(2,3), reshape=False, order=3,
 * mode='grid-wrap', prefilter=True): per plane, (1) cubic B-spline prefilter
 * along both axes with periodic boundary in double, (2) for every output pixel
 * o the input coordinate is  R o + (c - R c),  R = [[cos, sin], [-sin, cos]],
 * c = (n-1)/2; the value is the 4x4 cubic B-spline sum with indices wrapped
 * modulo n; (3) round to float32. */

static void bspline3_prefilter_periodic(double *c, int n, int stride)
{
    const double z = sqrt(3.0) - 2.0;
    if (n < 2) return;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    for (int i = 0; i < n; ++i) c[i * stride] *= gain;
    /* causal initial value: sum over one full period, closed geometrically */
    double zi = z, s = c[0];
    for (int i = 1; i < n; ++i) { s += zi * c[(n - i) * stride]; zi *= z; }
    c[0] = s / (1.0 - zi);
    for (int i = 1; i < n; ++i) c[i * stride] += z * c[(i - 1) * stride];
    /* anti-causal initial value */
    zi = z; s = c[(n - 1) * stride];
    for (int i = 0; i < n - 1; ++i) { s += zi * c[i * stride]; zi *= z; }
    c[(n - 1) * stride] = s * z / (zi - 1.0);
    for (int i = n - 2; i >= 0; --i) c[i * stride] = z * (c[(i + 1) * stride] - c[i * stride]);
}

static inline void bspline3_weights(double t, double w[4])
{
    double t1 = 1.0 - t;
    w[0] = t1 * t1 * t1 / 6.0;
    w[1] = 2.0 / 3.0 - 0.5 * t * t * (2.0 - t);
    w[2] = 2.0 / 3.0 - 0.5 * t1 * t1 * (2.0 - t1);
    w[3] = t * t * t / 6.0;
}

static inline int wrap_idx(long i, int n) { long m = i % n; return (int)(m < 0 ? m + n : m); }

void afr_oracle_rotate(const float *x, float *y, long planes, int H, int W, double degrees)
{
    double th = degrees * M_PI / 180.0, cs = cos(th), sn = sin(th);
    double ch = 0.5 * (H - 1), cw = 0.5 * (W - 1);
    double off_r = ch - (cs * ch + sn * cw), off_c = cw - (-sn * ch + cs * cw);
    long n = (long)H * W;
#pragma omp parallel
    {
        double *c = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (long p = 0; p < planes; ++p) {
            for (long e = 0; e < n; ++e) c[e] = x[p * n + e];
            for (int j = 0; j < W; ++j) bspline3_prefilter_periodic(c + j, H, W);
            for (int i = 0; i < H; ++i) bspline3_prefilter_periodic(c + (long)i * W, W, 1);
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    double rr = cs * i + sn * j + off_r, cc = -sn * i + cs * j + off_c;
                    double fr = floor(rr), fc = floor(cc), wr[4], wc[4], acc = 0.0;
                    bspline3_weights(rr - fr, wr);
                    bspline3_weights(cc - fc, wc);
                    for (int a = 0; a < 4; ++a) {
                        int ri = wrap_idx((long)fr - 1 + a, H);
                        double row = 0.0;
                        for (int b = 0; b < 4; ++b)
                            row += wc[b] * c[(long)ri * W + wrap_idx((long)fc - 1 + b, W)];
                        acc += wr[a] * row;
                    }
                    y[p * n + (long)i * W + j] = (float)acc;
                }
        }
        free(c);
    }
}

int afr_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the OpenMP regions above.  torchrun exports OMP_NUM_THREADS=1 and libgomp reads the
 * environment once, when it is first loaded (by `import torch`, long before bench.py's CPU arm runs),
 * so the arm sets the count explicitly.  Returns the count now in force. */
int afr_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
