// extern "C" surface of libafr_b200.so (declared in include/afr.h): argument checking,
// taps arrangement for forward/adjoint stencils, and kernel selection.
//
// Stencil table (pl = (N-1)/2 and ph = N-1-pl are F.conv2d's 'same' padding split):
//   up2x   fwd  up-like  (x,  k,        pad = pl)   -> [2H, 2W]
//   up2x   bwd  down-like(du, flip(k),  pad = ph)   -> [H, W]
//   down2x fwd  down-like(v,  k,        pad = pl)   -> [ceil(H/2), ceil(W/2)]
//   down2x bwd  up-like  (dy, flip(k),  pad = ph)   -> [H, W]
//   fused  fwd  u = up-like(x, k_up, pl_u);  y  = down-like(gelu(u),       k_dn,       pl_d)
//   fused  bwd  dg = up-like(dy, flip(k_dn), ph_d);
//               dx = down-like(gelu'(u) * dg, flip(k_up), ph_u)
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "afr_common.cuh"
#include "afr_kernels.h"

using namespace afr;

namespace {

thread_local char g_err[512] = "";
thread_local char g_detail[320] = "";
thread_local const char *g_last_kernel = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_path{-1};

int current_path()
{
    int p = g_path.load();
    if (p < 0) {
        const char *e = getenv("AFR_PATH");
        p = AFR_PATH_AUTO;
        if (e) {
            if (!strcmp(e, "direct")) p = AFR_PATH_DIRECT;
            else if (!strcmp(e, "tma")) p = AFR_PATH_TMA;
            else if (!strcmp(e, "generic")) p = AFR_PATH_GENERIC;
            else if (!strcmp(e, "direct_general")) p = AFR_PATH_DIRECT_GENERAL;
            else if (!strcmp(e, "tma_general")) p = AFR_PATH_TMA_GENERAL;
        }
        g_path.store(p);
    }
    return p;
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_status(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return AFR_OK;
    }
    return fail(AFR_ERR_CUDA, "%s: %s%s%s", what, cudaGetErrorString(e), g_detail[0] ? " -- " : "", g_detail);
}

// The runtime is shared with PyTorch (one libcudart.so.12): drop any stale non-sticky error a
// previous, unrelated runtime call left behind so that it is not attributed to our launch.
void begin_call()
{
    g_detail[0] = 0;
    (void)cudaGetLastError();
}

bool dtype_ok(int d) { return d == AFR_F32 || d == AFR_BF16; }
size_t esz(int d) { return d == AFR_F32 ? 4 : 2; }
bool elem_aligned(const void *p, int d) { return (reinterpret_cast<uintptr_t>(p) % esz(d)) == 0; }

int check_common(int B, int C, int H, int W, const float *taps, int N)
{
    if (B < 0 || C < 0 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    if ((long)H * W > (1L << 30)) return fail(AFR_ERR_BAD_SHAPE, "plane too large");
    if (!taps) return fail(AFR_ERR_BAD_TAPS, "taps is NULL");
    if (N < 1 || N > AFR_MAX_TAPS) return fail(AFR_ERR_BAD_TAPS, "N=%d outside [1,%d]", N, AFR_MAX_TAPS);
    return AFR_OK;
}

void set_taps(TapsG &t, const float *k, int N, bool flip)
{
    t.n = N;
    const int pl = (N - 1) / 2;
    t.pad = flip ? (N - 1 - pl) : pl;
    for (int i = 0; i < N * N; ++i) t.k[i] = flip ? k[N * N - 1 - i] : k[i];
    for (int i = N * N; i < AFR_MAX_TAPS * AFR_MAX_TAPS; ++i) t.k[i] = 0.f;
}

void set_taps3(Taps3 &t, const float *k, bool flip)
{
    for (int i = 0; i < 9; ++i) t.k[i / 3][i % 3] = flip ? k[8 - i] : k[i];
}

}  // namespace

namespace afr {
void set_detail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
    va_end(ap);
}
}  // namespace afr

extern "C" {

int afr_version(void) { return AFR_VERSION; }
const char *afr_last_error(void) { return g_err; }
const char *afr_last_kernel(void) { return g_last_kernel; }
uint64_t afr_launch_count(void) { return g_launches.load(); }

const char *afr_status_string(int s)
{
    switch (s) {
    case AFR_OK: return "ok";
    case AFR_ERR_BAD_SHAPE: return "bad shape";
    case AFR_ERR_BAD_TAPS: return "bad taps";
    case AFR_ERR_BAD_DTYPE: return "bad dtype";
    case AFR_ERR_NULL_POINTER: return "null pointer";
    case AFR_ERR_MISALIGNED: return "misaligned pointer";
    case AFR_ERR_CUDA: return "CUDA error";
    case AFR_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

int afr_set_path(int path)
{
    int old = current_path();
    if (path >= AFR_PATH_AUTO && path <= AFR_PATH_TMA_GENERAL) g_path.store(path);
    return old;
}

int afr_up2x_fwd(const void *x, void *u, int B, int C, int H, int W, const float *taps, int N,
                 int in_dtype, int out_dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!x || !u) return fail(AFR_ERR_NULL_POINTER, "x or u is NULL");
    if (!elem_aligned(x, in_dtype) || !elem_aligned(u, out_dtype)) return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    begin_call();
    const int path = current_path();
    const bool small_up = N == 3 && path == AFR_PATH_AUTO && small_up_supported(H, W, x, u, 4L * H * W, out_dtype);
    // fp32 outputs: the loop-free kernel with one float4 store per lane beats the warp-shuffle planes too
    const bool flat_first = n3_up_supported(H, W, x, u, in_dtype, out_dtype) && flat_up_wanted(H, W, in_dtype, out_dtype);
    if (small_up && !flat_first && (warp_up_shape(H, W) || !n3_up_supported(H, W, x, u, in_dtype, out_dtype))) {
        Taps3 k; set_taps3(k, taps, false);
        g_last_kernel = warp_up_shape(H, W) ? "up3_warp_kernel" : "up3_group_kernel";
        return cuda_status(small_up_like(x, u, planes, 1, 4L * H * W, H, W, k, in_dtype, out_dtype, s), g_last_kernel);
    }
    if (N == 3 && path != AFR_PATH_GENERIC && n3_up_supported(H, W, x, u, in_dtype, out_dtype)) {
        Taps3 k; set_taps3(k, taps, false);
        g_last_kernel = flat_up_wanted(H, W, in_dtype, out_dtype) ? "up3_flat_kernel" : "up3_kernel";
        return cuda_status(n3_up_like(x, u, planes, H, W, k, in_dtype, out_dtype, s), g_last_kernel);
    }
    TapsG t; set_taps(t, taps, N, false);
    if (path == AFR_PATH_AUTO && stripn_resample_supported(N) && n3_up_supported(H, W, x, u, in_dtype, out_dtype)) {
        g_last_kernel = "upn_kernel";
        return cuda_status(stripn_up_like(x, u, planes, H, W, t, in_dtype, out_dtype, s), "upn_kernel");
    }
    g_last_kernel = "up_like_generic_kernel";
    return cuda_status(generic_up_like(x, u, planes, H, W, 2 * H, 2 * W, t, in_dtype, out_dtype, s),
                       "up_like_generic_kernel");
}

int afr_up2x_bwd(const void *du, void *dx, int B, int C, int H, int W, const float *taps, int N,
                 int du_dtype, int dx_dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(du_dtype) || !dtype_ok(dx_dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!du || !dx) return fail(AFR_ERR_NULL_POINTER, "du or dx is NULL");
    if (!elem_aligned(du, du_dtype) || !elem_aligned(dx, dx_dtype)) return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    begin_call();
    const int path = current_path();
    if (du_dtype != dx_dtype)
        return fail(AFR_ERR_UNSUPPORTED, "up2x_bwd needs du and dx of one dtype (cast du on the host side)");
    const bool small_dn = N == 3 && path == AFR_PATH_AUTO && small_down_supported(2 * H, 2 * W, du, dx, 4L * H * W, du_dtype);
    if (small_dn && (warp_down_shape(2 * H, 2 * W) || !n3_down_supported(2 * H, 2 * W, du, dx, du_dtype))) {
        Taps3 k; set_taps3(k, taps, true);
        g_last_kernel = warp_down_shape(2 * H, 2 * W) ? "down3_warp_kernel" : "down3_group_kernel";
        return cuda_status(small_down_like(du, dx, planes, 1, 4L * H * W, 2 * H, 2 * W, k, du_dtype, s), g_last_kernel);
    }
    if (N == 3 && path != AFR_PATH_GENERIC && n3_down_supported(2 * H, 2 * W, du, dx, du_dtype)) {
        Taps3 k; set_taps3(k, taps, true);
        g_last_kernel = flat_down_wanted(2 * H, 2 * W, du_dtype) ? "down3_flat_kernel" : "down3_kernel";
        return cuda_status(n3_down_like(du, dx, planes, 2 * H, 2 * W, k, du_dtype, s), g_last_kernel);
    }
    TapsG t; set_taps(t, taps, N, true);
    if (path == AFR_PATH_AUTO && stripn_resample_supported(N) && n3_down_supported(2 * H, 2 * W, du, dx, du_dtype)) {
        g_last_kernel = "downn_kernel";
        return cuda_status(stripn_down_like(du, dx, planes, 2 * H, 2 * W, t, du_dtype, s), "downn_kernel");
    }
    g_last_kernel = "down_like_generic_kernel";
    return cuda_status(generic_down_like(du, dx, planes, 2 * H, 2 * W, H, W, t, du_dtype, s),
                       "down_like_generic_kernel");
}

int afr_up2x_fwd_strided(const void *x, void *u, int B, int C, int H, int W, int64_t out_batch_stride,
                         const float *taps, int N, int in_dtype, int out_dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (out_batch_stride < 4L * C * H * W) return fail(AFR_ERR_BAD_SHAPE, "out_batch_stride smaller than C*2H*2W");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!x || !u) return fail(AFR_ERR_NULL_POINTER, "x or u is NULL");
    begin_call();
    if (N != 3 || current_path() != AFR_PATH_AUTO)
        return fail(AFR_ERR_UNSUPPORTED, "strided up2x needs N == 3 and the AUTO kernel path");
    Taps3 k; set_taps3(k, taps, false);
    const bool strip_ok = n3_up_supported(H, W, x, u, in_dtype, out_dtype) && (out_batch_stride % 8) == 0;
    const bool flat_first = strip_ok && flat_up_wanted(H, W, in_dtype, out_dtype);
    if (!flat_first && small_up_supported(H, W, x, u, (long)out_batch_stride, out_dtype) && (warp_up_shape(H, W) || !strip_ok)) {
        g_last_kernel = warp_up_shape(H, W) ? "up3_warp_kernel" : "up3_group_kernel";
        return cuda_status(small_up_like(x, u, planes, C, (long)out_batch_stride, H, W, k, in_dtype, out_dtype,
                                         (cudaStream_t)stream), g_last_kernel);
    }
    if (!strip_ok)
        return fail(AFR_ERR_UNSUPPORTED, "strided up2x needs W %% 4 == 0, 32-byte aligned slice base and batch stride (H=%d W=%d)", H, W);
    g_last_kernel = flat_up_wanted(H, W, in_dtype, out_dtype) ? "up3_flat_kernel" : "up3_kernel";
    return cuda_status(n3_up_like(x, u, planes, H, W, k, in_dtype, out_dtype, (cudaStream_t)stream, C, (long)out_batch_stride),
                       g_last_kernel);
}

int afr_up2x_bwd_strided(const void *du, void *dx, int B, int C, int H, int W, int64_t du_batch_stride,
                         const float *taps, int N, int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (du_batch_stride < 4L * C * H * W) return fail(AFR_ERR_BAD_SHAPE, "du_batch_stride smaller than C*2H*2W");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!du || !dx) return fail(AFR_ERR_NULL_POINTER, "du or dx is NULL");
    begin_call();
    if (N != 3 || current_path() != AFR_PATH_AUTO)
        return fail(AFR_ERR_UNSUPPORTED, "strided up2x adjoint needs N == 3 and the AUTO kernel path");
    Taps3 k; set_taps3(k, taps, true);
    const bool strip_ok = n3_down_supported(2 * H, 2 * W, du, dx, dtype) && (du_batch_stride % 8) == 0;
    if (small_down_supported(2 * H, 2 * W, du, dx, (long)du_batch_stride, dtype) && (warp_down_shape(2 * H, 2 * W) || !strip_ok)) {
        g_last_kernel = warp_down_shape(2 * H, 2 * W) ? "down3_warp_kernel" : "down3_group_kernel";
        return cuda_status(small_down_like(du, dx, planes, C, (long)du_batch_stride, 2 * H, 2 * W, k, dtype,
                                           (cudaStream_t)stream), g_last_kernel);
    }
    if (!strip_ok)
        return fail(AFR_ERR_UNSUPPORTED, "strided up2x adjoint needs W %% 4 == 0 and 16-byte aligned slice base / batch stride (H=%d W=%d)", H, W);
    g_last_kernel = flat_down_wanted(2 * H, 2 * W, dtype) ? "down3_flat_kernel" : "down3_kernel";
    return cuda_status(n3_down_like(du, dx, planes, 2 * H, 2 * W, k, dtype, (cudaStream_t)stream, C, (long)du_batch_stride),
                       g_last_kernel);
}

int afr_down2x_fwd(const void *v, void *y, int B, int C, int H, int W, const float *taps, int N,
                   int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!v || !y) return fail(AFR_ERR_NULL_POINTER, "v or y is NULL");
    if (!elem_aligned(v, dtype) || !elem_aligned(y, dtype)) return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    begin_call();
    const int path = current_path();
    const bool small_dn = N == 3 && path == AFR_PATH_AUTO && small_down_supported(H, W, v, y, (long)H * W, dtype);
    if (small_dn && (warp_down_shape(H, W) || !n3_down_supported(H, W, v, y, dtype))) {
        Taps3 k; set_taps3(k, taps, false);
        g_last_kernel = warp_down_shape(H, W) ? "down3_warp_kernel" : "down3_group_kernel";
        return cuda_status(small_down_like(v, y, planes, 1, (long)H * W, H, W, k, dtype, s), g_last_kernel);
    }
    if (N == 3 && path != AFR_PATH_GENERIC && n3_down_supported(H, W, v, y, dtype)) {
        Taps3 k; set_taps3(k, taps, false);
        g_last_kernel = flat_down_wanted(H, W, dtype) ? "down3_flat_kernel" : "down3_kernel";
        return cuda_status(n3_down_like(v, y, planes, H, W, k, dtype, s), g_last_kernel);
    }
    TapsG t; set_taps(t, taps, N, false);
    if (path == AFR_PATH_AUTO && stripn_resample_supported(N) && n3_down_supported(H, W, v, y, dtype)) {
        g_last_kernel = "downn_kernel";
        return cuda_status(stripn_down_like(v, y, planes, H, W, t, dtype, s), "downn_kernel");
    }
    g_last_kernel = "down_like_generic_kernel";
    return cuda_status(generic_down_like(v, y, planes, H, W, (H + 1) / 2, (W + 1) / 2, t, dtype, s),
                       "down_like_generic_kernel");
}

int afr_down2x_bwd(const void *dy, void *dv, int B, int C, int H, int W, const float *taps, int N,
                   int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!dy || !dv) return fail(AFR_ERR_NULL_POINTER, "dy or dv is NULL");
    if (!elem_aligned(dy, dtype) || !elem_aligned(dv, dtype)) return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    begin_call();
    const int path = current_path();
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const bool small_up = N == 3 && path == AFR_PATH_AUTO && (H % 2) == 0 && (W % 2) == 0 &&
                          small_up_supported(Ho, Wo, dy, dv, (long)H * W, dtype);
    const bool flat_first = n3_up_supported(Ho, Wo, dy, dv, dtype, dtype) && flat_up_wanted(Ho, Wo, dtype, dtype);
    if (small_up && !flat_first && (warp_up_shape(Ho, Wo) || !n3_up_supported(Ho, Wo, dy, dv, dtype, dtype))) {
        Taps3 k; set_taps3(k, taps, true);
        g_last_kernel = warp_up_shape(Ho, Wo) ? "up3_warp_kernel" : "up3_group_kernel";
        return cuda_status(small_up_like(dy, dv, planes, 1, (long)H * W, Ho, Wo, k, dtype, dtype, s), g_last_kernel);
    }
    if (N == 3 && path != AFR_PATH_GENERIC && (H % 2) == 0 && (W % 2) == 0 &&
        n3_up_supported(Ho, Wo, dy, dv, dtype, dtype)) {
        Taps3 k; set_taps3(k, taps, true);
        g_last_kernel = flat_up_wanted(Ho, Wo, dtype, dtype) ? "up3_flat_kernel" : "up3_kernel";
        return cuda_status(n3_up_like(dy, dv, planes, Ho, Wo, k, dtype, dtype, s), g_last_kernel);
    }
    TapsG t; set_taps(t, taps, N, true);
    if (path == AFR_PATH_AUTO && stripn_resample_supported(N) && (H % 2) == 0 && (W % 2) == 0 &&
        n3_up_supported(Ho, Wo, dy, dv, dtype, dtype)) {
        g_last_kernel = "upn_kernel";
        return cuda_status(stripn_up_like(dy, dv, planes, Ho, Wo, t, dtype, dtype, s), "upn_kernel");
    }
    g_last_kernel = "up_like_generic_kernel";
    return cuda_status(generic_up_like(dy, dv, planes, Ho, Wo, H, W, t, dtype, dtype, s),
                       "up_like_generic_kernel");
}

static int fused_common(const void *x, const void *residual, const void *dy, void *out, int B, int C,
                        int H, int W, const float *taps_up, int N_up, const float *taps_down,
                        int N_down, int dtype, void *stream, bool bwd, const float *scale = nullptr,
                        const float *shift = nullptr)
{
    if (int rc = check_common(B, C, H, W, taps_up, N_up)) return rc;
    if (int rc = check_common(B, C, H, W, taps_down, N_down)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!x || !out || (bwd && !dy)) return fail(AFR_ERR_NULL_POINTER, "NULL tensor pointer");
    if (!elem_aligned(x, dtype) || !elem_aligned(out, dtype) || (residual && !elem_aligned(residual, dtype)) ||
        (bwd && !elem_aligned(dy, dtype)))
        return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    begin_call();
    int path = current_path();
    // *_GENERAL: same kernel family with the symmetric-tap fast path switched off (tests, A/B timing)
    const bool allow_sym = (path != AFR_PATH_DIRECT_GENERAL && path != AFR_PATH_TMA_GENERAL);
    if (path == AFR_PATH_DIRECT_GENERAL) path = AFR_PATH_DIRECT;
    if (path == AFR_PATH_TMA_GENERAL) path = AFR_PATH_TMA;
    const void *ptrs[4] = {x, residual, bwd ? dy : nullptr, out};
    if (N_up == 3 && N_down == 3 && path != AFR_PATH_GENERIC && n3_fgelu_supported(H, W, ptrs, 4, dtype)) {
        const bool tma_ok = n3_fgelu_tma_supported(planes, H, W, ptrs, 4, dtype, 1 + (residual ? 1 : 0) + (bwd ? 1 : 0));
        if (path == AFR_PATH_TMA && !tma_ok)
            return fail(AFR_ERR_UNSUPPORTED, "TMA path forced but shape/alignment not eligible (H=%d W=%d)", H, W);
        const bool use_tma = (path == AFR_PATH_TMA) || (path == AFR_PATH_AUTO && tma_ok && !n3_prefers_plane_kernel(H, W, bwd));
        Taps3 kU, kG, kB;
        set_taps3(kU, taps_up, false);
        if (bwd)                       // the adjoint kernels evaluate gelu' in w = kappa * u
            for (int i = 0; i < 9; ++i) kU.k[i / 3][i % 3] *= AFR_KAPPA;
        set_taps3(kG, taps_down, true);
        set_taps3(kB, bwd ? taps_up : taps_down, bwd);
        return cuda_status(n3_fgelu(x, residual, dy, scale, shift, out, planes, H, W, kU, kG, kB, bwd, dtype,
                                    use_tma, allow_sym, s, &g_last_kernel),
                           "fgelu3 kernel");
    }
    if (scale)
        return fail(AFR_ERR_UNSUPPORTED, "affine fusion needs the N==3 kernels (N_up=%d N_down=%d H=%d W=%d)", N_up,
                    N_down, H, W);
    if (path == AFR_PATH_TMA || path == AFR_PATH_DIRECT)
        return fail(AFR_ERR_UNSUPPORTED, "N==3 path forced but N_up=%d N_down=%d H=%d W=%d not eligible", N_up, N_down, H, W);
    TapsG tU, tG, tB;
    set_taps(tU, taps_up, N_up, false);
    set_taps(tG, taps_down, N_down, true);
    if (bwd) set_taps(tB, taps_up, N_up, true); else set_taps(tB, taps_down, N_down, false);
    if (path == AFR_PATH_AUTO && N_up == N_down && stripn_supported(N_up, H, W, x, residual, bwd ? dy : nullptr, out, dtype)) {
        g_last_kernel = "fgelu_strip_kernel";
        return cuda_status(stripn_fgelu(x, residual, dy, out, planes, H, W, tU, tG, tB, bwd, dtype, s),
                           "fgelu_strip_kernel");
    }
    g_last_kernel = "fgelu_generic_kernel";
    return cuda_status(generic_fgelu(x, residual, dy, out, planes, H, W, tU, tG, tB, bwd, dtype, s),
                       "fgelu_generic_kernel");
}

int afr_filtered_gelu_fwd(const void *x, const void *residual, void *y, int B, int C, int H, int W,
                          const float *taps_up, int N_up, const float *taps_down, int N_down,
                          int dtype, void *stream)
{
    return fused_common(x, residual, nullptr, y, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype,
                        stream, false);
}

int afr_filtered_gelu_affine_fwd(const void *x, const void *residual, const float *scale_dev,
                                 const float *shift_dev, void *y, int B, int C, int H, int W,
                                 const float *taps_up, int N_up, const float *taps_down, int N_down,
                                 int dtype, void *stream)
{
    if (!scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "scale or shift is NULL");
    return fused_common(x, residual, nullptr, y, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype,
                        stream, false, scale_dev, shift_dev);
}

int afr_filtered_gelu_bwd(const void *x, const void *residual, const void *dy, void *dx, int B, int C,
                          int H, int W, const float *taps_up, int N_up, const float *taps_down,
                          int N_down, int dtype, void *stream)
{
    return fused_common(x, residual, dy, dx, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype,
                        stream, true);
}

int afr_filtered_gelu_affine_bwd(const void *x, const void *residual, const float *scale_dev, const float *shift_dev,
                                 const void *dy, void *dz, int B, int C, int H, int W, const float *taps_up, int N_up,
                                 const float *taps_down, int N_down, int dtype, void *stream)
{
    if (!scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "scale or shift is NULL");
    return fused_common(x, residual, dy, dz, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype, stream, true,
                        scale_dev, shift_dev);
}

static int fused_nhwc(const void *x, const void *residual, const float *scale, const float *shift, const void *dy,
                      void *out, int B, int C, int H, int W, const float *taps_up, int N_up, const float *taps_down,
                      int N_down, int dtype, void *stream, bool bwd)
{
    if (int rc = check_common(B, C, H, W, taps_up, N_up)) return rc;
    if (int rc = check_common(B, C, H, W, taps_down, N_down)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if ((scale == nullptr) != (shift == nullptr)) return fail(AFR_ERR_NULL_POINTER, "scale and shift go together");
    if ((long)B * C == 0) return AFR_OK;
    if (!x || !out || (bwd && !dy)) return fail(AFR_ERR_NULL_POINTER, "NULL tensor pointer");
    begin_call();
    const void *ptrs[4] = {x, residual, bwd ? dy : nullptr, out};
    if (N_up != 3 || N_down != 3 || current_path() == AFR_PATH_GENERIC || !nhwc_fgelu_supported(C, H, W, ptrs, 4))
        return fail(AFR_ERR_UNSUPPORTED, "channels-last fused kernel needs 3x3 filters, C %% 32 == 0, W %% 4 == 0, H >= 2 and "
                                         "16-byte aligned buffers (C=%d H=%d W=%d)", C, H, W);
    Taps3 kU, kG, kB;
    set_taps3(kU, taps_up, false);
    if (bwd)
        for (int i = 0; i < 9; ++i) kU.k[i / 3][i % 3] *= AFR_KAPPA;
    set_taps3(kG, taps_down, true);
    set_taps3(kB, bwd ? taps_up : taps_down, bwd);
    const cudaError_t e = nhwc_fgelu(x, residual, dy, scale, shift, out, B, C, H, W, kU, kG, kB, bwd, dtype, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        return fail(AFR_ERR_UNSUPPORTED, "channels-last fused kernel needs D4-symmetric taps (non-negative up taps in the forward)");
    g_last_kernel = "fgelu3_nhwc_kernel<sym>";
    return cuda_status(e, "fgelu3_nhwc_kernel");
}

int afr_filtered_gelu_nhwc_fwd(const void *x, const void *residual, const float *scale_dev, const float *shift_dev, void *y,
                               int B, int C, int H, int W, const float *taps_up, int N_up, const float *taps_down,
                               int N_down, int dtype, void *stream)
{
    return fused_nhwc(x, residual, scale_dev, shift_dev, nullptr, y, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype,
                      stream, false);
}

int afr_filtered_gelu_nhwc_bwd(const void *x, const void *residual, const float *scale_dev, const float *shift_dev,
                               const void *dy, void *dz, int B, int C, int H, int W, const float *taps_up, int N_up,
                               const float *taps_down, int N_down, int dtype, void *stream)
{
    return fused_nhwc(x, residual, scale_dev, shift_dev, dy, dz, B, C, H, W, taps_up, N_up, taps_down, N_down, dtype,
                      stream, true);
}

int afr_affine_apply_nhwc(const void *x, const float *scale_dev, const float *shift_dev, void *y, int B, int C, int H, int W,
                          int dtype, void *stream)
{
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (B == 0) return AFR_OK;
    if (!x || !y || !scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    if (C % 4 != 0 || (reinterpret_cast<uintptr_t>(x) % (4 * esz(dtype))) != 0 || (reinterpret_cast<uintptr_t>(y) % (4 * esz(dtype))) != 0 ||
        (reinterpret_cast<uintptr_t>(scale_dev) % 16) != 0 || (reinterpret_cast<uintptr_t>(shift_dev) % 16) != 0)
        return fail(AFR_ERR_UNSUPPORTED, "C must be a multiple of 4 and the buffers aligned to 4 elements");
    begin_call();
    g_last_kernel = "affine_apply_nhwc_kernel";
    return cuda_status(affine_apply_nhwc(x, scale_dev, shift_dev, y, B, C, (long)H * W, dtype, (cudaStream_t)stream),
                       "affine_apply_nhwc_kernel");
}

int afr_up2x_nhwc(const void *x, void *u, int B, int C, int H, int W, int64_t x_pixel_stride, int64_t u_pixel_stride,
                  const float *taps, int N, int adjoint_of_down, int in_dtype, int out_dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if ((long)B * C == 0) return AFR_OK;
    if (!x || !u) return fail(AFR_ERR_NULL_POINTER, "x or u is NULL");
    begin_call();
    if (N != 3 || current_path() == AFR_PATH_GENERIC ||
        !nhwc_resample_supported(C, x, u, (long)x_pixel_stride, (long)u_pixel_stride, in_dtype, out_dtype))
        return fail(AFR_ERR_UNSUPPORTED, "channels-last up2x needs a 3x3 filter, C %% 4 == 0 and 4-element aligned bases / pixel strides");
    Taps3 k; set_taps3(k, taps, adjoint_of_down != 0);
    g_last_kernel = "up3_nhwc_kernel";
    return cuda_status(nhwc_up_like(x, u, B, C, H, W, (long)x_pixel_stride, (long)u_pixel_stride, k, in_dtype, out_dtype,
                                    (cudaStream_t)stream), "up3_nhwc_kernel");
}

int afr_down2x_nhwc(const void *v, void *y, int B, int C, int H, int W, int64_t v_pixel_stride, int64_t y_pixel_stride,
                    const float *taps, int N, int adjoint_of_up, int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if ((long)B * C == 0) return AFR_OK;
    if (!v || !y) return fail(AFR_ERR_NULL_POINTER, "v or y is NULL");
    begin_call();
    if (N != 3 || current_path() == AFR_PATH_GENERIC ||
        !nhwc_resample_supported(C, v, y, (long)v_pixel_stride, (long)y_pixel_stride, dtype, dtype))
        return fail(AFR_ERR_UNSUPPORTED, "channels-last down2x needs a 3x3 filter, C %% 4 == 0 and 4-element aligned bases / pixel strides");
    Taps3 k; set_taps3(k, taps, adjoint_of_up != 0);
    g_last_kernel = "down3_nhwc_kernel";
    return cuda_status(nhwc_down_like(v, y, B, C, H, W, (long)v_pixel_stride, (long)y_pixel_stride, k, dtype, (cudaStream_t)stream),
                       "down3_nhwc_kernel");
}

int afr_groupnorm1_bwd(const void *x, const void *dz, const float *gamma_dev, const float *mean_dev, const float *rstd_dev,
                       void *dx, float *dgamma_dev, float *dbeta_dev, float *workspace_dev, int B, int C, int H, int W,
                       int dtype, int channels_last, void *stream)
{
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (B == 0) return AFR_OK;
    if (!x || !dz || !gamma_dev || !mean_dev || !rstd_dev || !dx || !dgamma_dev || !dbeta_dev || !workspace_dev)
        return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    const long hw = (long)H * W;
    const bool ok_shape = channels_last ? (C % 32 == 0) : (hw % 4 == 0);
    if (!ok_shape || (reinterpret_cast<uintptr_t>(x) % 16) || (reinterpret_cast<uintptr_t>(dz) % 16) ||
        (reinterpret_cast<uintptr_t>(dx) % 16) || (reinterpret_cast<uintptr_t>(gamma_dev) % 16))
        return fail(AFR_ERR_UNSUPPORTED, "GroupNorm backward needs H*W %% 4 == 0 (NCHW) or C %% 32 == 0 (channels-last) and "
                                         "16-byte aligned buffers");
    begin_call();
    g_last_kernel = "gn_bwd_apply_kernel";
    const int rc = cuda_status(groupnorm1_bwd(x, dz, gamma_dev, mean_dev, rstd_dev, dx, dgamma_dev, dbeta_dev, workspace_dev, B, C,
                                              hw, dtype, channels_last != 0, (cudaStream_t)stream), "gn_bwd kernels");
    if (rc == AFR_OK) g_launches.fetch_add(2, std::memory_order_relaxed);      // three kernels per call
    return rc;
}

int afr_groupnorm1_stats(const void *x, const float *gamma_dev, const float *beta_dev, float eps, const float *add_dev,
                         float *scale_dev, float *shift_dev, float *mean_dev, float *rstd_dev, int B, int C, int H, int W,
                         int dtype, void *stream)
{
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (B == 0) return AFR_OK;
    if (!x || !gamma_dev || !beta_dev || !scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    const long hw = (long)H * W;
    if (((long)C * hw) % 4 != 0 || (reinterpret_cast<uintptr_t>(x) % (4 * esz(dtype))) != 0)
        return fail(AFR_ERR_UNSUPPORTED, "C*H*W must be a multiple of 4 and x aligned to 4 elements");
    begin_call();
    g_last_kernel = "groupnorm1_affine_kernel";
    return cuda_status(groupnorm1_affine(x, gamma_dev, beta_dev, eps, scale_dev, shift_dev, B, C, hw, dtype,
                                         (cudaStream_t)stream, add_dev, mean_dev, rstd_dev),
                       "groupnorm1_affine_kernel");
}

int afr_affine_apply(const void *x, const float *scale_dev, const float *shift_dev, void *y, int B, int C, int H, int W,
                     int dtype, void *stream)
{
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (B == 0) return AFR_OK;
    if (!x || !y || !scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    const long hw = (long)H * W;
    if (hw % 4 != 0 || (reinterpret_cast<uintptr_t>(x) % (4 * esz(dtype))) != 0 || (reinterpret_cast<uintptr_t>(y) % (4 * esz(dtype))) != 0)
        return fail(AFR_ERR_UNSUPPORTED, "H*W must be a multiple of 4 and x, y aligned to 4 elements");
    begin_call();
    g_last_kernel = "affine_apply_kernel";
    return cuda_status(affine_apply(x, scale_dev, shift_dev, y, (long)B * C, hw, dtype, (cudaStream_t)stream), "affine_apply_kernel");
}

int afr_gelu_down2x_fwd(const void *v, const float *scale_dev, const float *shift_dev, void *y, int B, int C, int H,
                        int W, const float *taps, int N, int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if ((scale_dev == nullptr) != (shift_dev == nullptr)) return fail(AFR_ERR_NULL_POINTER, "scale and shift go together");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!v || !y) return fail(AFR_ERR_NULL_POINTER, "v or y is NULL");
    begin_call();
    if (N != 3 || current_path() == AFR_PATH_GENERIC || !actdown_supported(H, W, v, y, dtype))
        return fail(AFR_ERR_UNSUPPORTED, "gelu_down2x needs N == 3, even H, W %% 8 == 0 and aligned buffers (N=%d H=%d W=%d)", N, H, W);
    Taps3 k; set_taps3(k, taps, false);
    g_last_kernel = "gelu_down3_kernel";
    return cuda_status(actdown_fwd(v, scale_dev, shift_dev, y, planes, H, W, k, dtype, (cudaStream_t)stream), "gelu_down3_kernel");
}

int afr_gelu_down2x_bwd(const void *v, const void *dy, void *dv, int B, int C, int H, int W, const float *taps, int N,
                        int dtype, void *stream)
{
    if (int rc = check_common(B, C, H, W, taps, N)) return rc;
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!v || !dy || !dv) return fail(AFR_ERR_NULL_POINTER, "NULL tensor pointer");
    begin_call();
    if (N != 3 || current_path() == AFR_PATH_GENERIC || !actdown_supported(H, W, v, dy, dtype) ||
        (reinterpret_cast<uintptr_t>(dv) % (8 * esz(dtype))) != 0)
        return fail(AFR_ERR_UNSUPPORTED, "gelu_down2x adjoint needs N == 3, even H, W %% 8 == 0 and aligned buffers (N=%d H=%d W=%d)", N, H, W);
    Taps3 k; set_taps3(k, taps, true);
    g_last_kernel = "gelu_up3_bwd_kernel";
    return cuda_status(actdown_bwd(v, dy, dv, planes, H, W, k, dtype, (cudaStream_t)stream), "gelu_up3_bwd_kernel");
}

int afr_rotate_periodic_cubic(const void *x, void *y, int B, int C, int H, int W, double degrees,
                              int dtype, void *stream)
{
    if (B < 0 || C < 0 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (dtype != AFR_F32) return fail(AFR_ERR_BAD_DTYPE, "rotate supports fp32 only");
    if ((long)H * (W | 1) > 16384 + 128) return fail(AFR_ERR_UNSUPPORTED, "plane larger than 16384 px");
    const long planes = (long)B * C;
    if (planes == 0) return AFR_OK;
    if (!x || !y) return fail(AFR_ERR_NULL_POINTER, "x or y is NULL");
    if (x == y) return fail(AFR_ERR_UNSUPPORTED, "in-place rotate is not supported");
    if (!elem_aligned(x, dtype) || !elem_aligned(y, dtype)) return fail(AFR_ERR_MISALIGNED, "misaligned buffer");
    begin_call();
    g_last_kernel = "rotate_kernel";
    return cuda_status(rotate_periodic_cubic((const float *)x, (float *)y, planes, H, W, degrees,
                                             (cudaStream_t)stream),
                       "rotate_kernel");
}

int afr_groupnorm1_affine(const void *x, const float *gamma_dev, const float *beta_dev, float eps,
                          float *scale_dev, float *shift_dev, int B, int C, int H, int W, int dtype,
                          void *stream)
{
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(AFR_ERR_BAD_SHAPE, "bad shape");
    if (!dtype_ok(dtype)) return fail(AFR_ERR_BAD_DTYPE, "bad dtype");
    if (B == 0) return AFR_OK;
    if (!x || !gamma_dev || !beta_dev || !scale_dev || !shift_dev) return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    const long hw = (long)H * W;
    if (((long)C * hw) % 4 != 0 || (reinterpret_cast<uintptr_t>(x) % (4 * esz(dtype))) != 0)
        return fail(AFR_ERR_UNSUPPORTED, "C*H*W must be a multiple of 4 and x aligned to 4 elements");
    begin_call();
    g_last_kernel = "groupnorm1_affine_kernel";
    return cuda_status(groupnorm1_affine(x, gamma_dev, beta_dev, eps, scale_dev, shift_dev, B, C, hw, dtype,
                                         (cudaStream_t)stream),
                       "groupnorm1_affine_kernel");
}

int afr_ddpm_update(void *x, const void *eps, const void *noise, int64_t n, float ca, float cb,
                    float cc, void *stream)
{
    if (n < 0) return fail(AFR_ERR_BAD_SHAPE, "n < 0");
    if (n == 0) return AFR_OK;
    if (!x || !eps) return fail(AFR_ERR_NULL_POINTER, "x or eps is NULL");
    begin_call();
    g_last_kernel = "ddpm_update_kernel";
    return cuda_status(ddpm_update((float *)x, (const float *)eps, (const float *)noise, (long)n, ca, cb,
                                   cc, nullptr, nullptr, (cudaStream_t)stream),
                       "ddpm_update_kernel");
}

int afr_ddpm_update_table(void *x, const void *eps, const void *noise, int64_t n, const float *table_dev,
                          const int *step_dev, void *stream)
{
    if (n < 0) return fail(AFR_ERR_BAD_SHAPE, "n < 0");
    if (n == 0) return AFR_OK;
    if (!x || !eps || !table_dev || !step_dev) return fail(AFR_ERR_NULL_POINTER, "NULL pointer");
    begin_call();
    g_last_kernel = "ddpm_update_kernel";
    return cuda_status(ddpm_update((float *)x, (const float *)eps, (const float *)noise, (long)n, 0.f, 0.f,
                                   0.f, table_dev, step_dev, (cudaStream_t)stream),
                       "ddpm_update_kernel");
}

}  // extern "C"
