"""PyTorch-facing operators of the alias-free resampling path.

Thin host code: argument checks, output allocation and ``torch.autograd.Function``
wrappers around the C ABI of ``libafr_b200.so`` (include/afr.h).  All compute happens in
the CUDA kernels; a non-CUDA tensor is an error (there is no CPU fallback).

Names mirror the reference: ``custom_upsample`` / ``custom_downsample`` have the
signatures of modules/filtrs.py:79 and :71.
"""
import ctypes
import os

import torch

from . import _native

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _check(rc):
    if rc != 0:
        L = _native.lib()
        raise RuntimeError("afr: %s: %s" % (L.afr_status_string(rc).decode(), L.afr_last_error().decode()))


try:                                           # raw handle of torch's current stream without building a Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                         # older torch
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def _stream(t):
    return ctypes.c_void_p(_raw_stream(t.device.index))


class _on_device:
    """``with torch.cuda.device(t.device)`` that costs nothing when that device is already current (the common
    case: one process per GPU).  A launch from Python is ~10 us of host time; this and the raw stream handle take
    ~4 us off it, which is what the eager small-batch UNet feels (63 of our launches per reverse step)."""
    __slots__ = ("idx", "guard")

    def __init__(self, device):
        self.idx, self.guard = device.index, None

    def __enter__(self):
        if torch.cuda.current_device() != self.idx:
            self.guard = torch.cuda.device(self.idx)
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _is_cl(x):
    """Dense channels-last memory ([B, H, W, C]) that is not also dense NCHW (C == 1 or H == W == 1 are both)."""
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()


def _nhwc_eligible(x, ku, kd):
    """Shapes the channels-last fused kernel takes (csrc/afr_n3.cu: fgelu3_nhwc_kernel)."""
    return (_is_cl(x) and ku.n == 3 and kd.n == 3 and x.shape[1] % 32 == 0 and x.shape[3] % 4 == 0 and x.shape[2] >= 2
            and x.dtype in _DT)


def _require(x, name="x", keep_cl=False):
    if not isinstance(x, torch.Tensor):
        raise TypeError("afr: %s must be a torch.Tensor" % name)
    if not x.is_cuda:
        raise RuntimeError("afr: %s must be a CUDA tensor (this build has no CPU fallback)" % name)
    if x.dim() != 4:
        raise ValueError("afr: %s must be [B, C, H, W], got %s" % (name, tuple(x.shape)))
    if x.dtype not in _DT:
        raise TypeError("afr: %s dtype %s unsupported (float32 / bfloat16)" % (name, x.dtype))
    if keep_cl and _is_cl(x):
        return x                      # channels-last stays channels-last where a kernel can take it as it is
    return x.contiguous()


class Taps:
    """Host copy of an N x N filter.  Only the tensor is state: the pointer handed to the C ABI is taken
    from it at launch time, so a ``Taps`` (and any module caching one) can be deep-copied and pickled --
    the reference's ``copy.deepcopy(model)`` EMA pattern and whole-model ``torch.save`` keep working."""
    __slots__ = ("t", "n")

    def __init__(self, filt):
        if isinstance(filt, Taps):
            self.t, self.n = filt.t, filt.n
            return
        t = torch.as_tensor(filt).detach().to("cpu", torch.float32).contiguous()
        if t.dim() != 2 or t.shape[0] != t.shape[1]:
            raise ValueError("afr: filter must be N x N, got %s" % (tuple(t.shape),))
        self.t, self.n = t, int(t.shape[0])

    @property
    def ptr(self):
        return ctypes.c_void_p(self.t.data_ptr())


def _taps(filt):
    return filt if isinstance(filt, Taps) else Taps(filt)


# ---- raw launches (no autograd) ---------------------------------------------------
_CL = torch.channels_last


def _cl_resample_ok(x, k):
    """Channels-last tensors the NHWC resamplers take as they are (csrc/afr_nhwc_resample.cu)."""
    return _is_cl(x) and k.n == 3 and x.shape[1] % 4 == 0 and x.dtype in _DT


def _up_nhwc(x, u, k, adjoint, B, C, H, W, xps, ups):
    with _on_device(x.device):
        _check(_native.lib().afr_up2x_nhwc(x.data_ptr(), u.data_ptr(), B, C, H, W, xps, ups, k.ptr, k.n, 1 if adjoint else 0,
                                           _DT[x.dtype], _DT[u.dtype], _stream(x)))


def _down_nhwc(v, y, k, adjoint, B, C, H, W, vps, yps):
    with _on_device(v.device):
        _check(_native.lib().afr_down2x_nhwc(v.data_ptr(), y.data_ptr(), B, C, H, W, vps, yps, k.ptr, k.n, 1 if adjoint else 0,
                                             _DT[v.dtype], _stream(v)))


def _up_fwd(x, k, out_dtype):
    B, C, H, W = x.shape
    if _cl_resample_ok(x, k):
        u = torch.empty((B, C, 2 * H, 2 * W), dtype=out_dtype, device=x.device, memory_format=_CL)
        _up_nhwc(x, u, k, False, B, C, H, W, C, C)
        return u
    x = x.contiguous()
    u = torch.empty((B, C, 2 * H, 2 * W), dtype=out_dtype, device=x.device)
    with _on_device(x.device):
        _check(_native.lib().afr_up2x_fwd(x.data_ptr(), u.data_ptr(), B, C, H, W, k.ptr, k.n,
                                          _DT[x.dtype], _DT[out_dtype], _stream(x)))
    return u


def _up_bwd(du, k, H, W):
    B, C = du.shape[:2]
    if _cl_resample_ok(du, k):
        dx = torch.empty((B, C, H, W), dtype=du.dtype, device=du.device, memory_format=_CL)
        _down_nhwc(du, dx, k, True, B, C, 2 * H, 2 * W, C, C)
        return dx
    du = du.contiguous()
    dx = torch.empty((B, C, H, W), dtype=du.dtype, device=du.device)
    with _on_device(du.device):
        _check(_native.lib().afr_up2x_bwd(du.data_ptr(), dx.data_ptr(), B, C, H, W, k.ptr, k.n,
                                          _DT[du.dtype], _DT[dx.dtype], _stream(du)))
    return dx


def _down_fwd(v, k):
    B, C, H, W = v.shape
    if _cl_resample_ok(v, k):
        y = torch.empty((B, C, (H + 1) // 2, (W + 1) // 2), dtype=v.dtype, device=v.device, memory_format=_CL)
        _down_nhwc(v, y, k, False, B, C, H, W, C, C)
        return y
    v = v.contiguous()
    y = torch.empty((B, C, (H + 1) // 2, (W + 1) // 2), dtype=v.dtype, device=v.device)
    with _on_device(v.device):
        _check(_native.lib().afr_down2x_fwd(v.data_ptr(), y.data_ptr(), B, C, H, W, k.ptr, k.n,
                                            _DT[v.dtype], _stream(v)))
    return y


def _down_bwd(dy, k, H, W):
    B, C = dy.shape[:2]
    if _cl_resample_ok(dy, k) and H % 2 == 0 and W % 2 == 0:
        dv = torch.empty((B, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=_CL)
        _up_nhwc(dy, dv, k, True, B, C, H // 2, W // 2, C, C)
        return dv
    dy = dy.contiguous()
    dv = torch.empty((B, C, H, W), dtype=dy.dtype, device=dy.device)
    with _on_device(dy.device):
        _check(_native.lib().afr_down2x_bwd(dy.data_ptr(), dv.data_ptr(), B, C, H, W, k.ptr, k.n,
                                            _DT[dy.dtype], _stream(dy)))
    return dv


def _fgelu_nhwc(x, res, scale, shift, dy, ku, kd):
    """Channels-last launch (forward, or adjoint when `dy` is given; optional GroupNorm affine).  Returns the
    channels-last result, or None when the library declines (filters that are not D4-symmetric): the caller then
    goes through NCHW-contiguous copies."""
    B, C, H, W = x.shape
    cl = torch.channels_last
    if res is not None:
        res = res.contiguous(memory_format=cl)
    out = torch.empty_like(x, memory_format=cl)
    L = _native.lib()
    with _on_device(x.device):
        if dy is None:
            rc = L.afr_filtered_gelu_nhwc_fwd(x.data_ptr(), None if res is None else res.data_ptr(),
                                              None if scale is None else scale.data_ptr(),
                                              None if shift is None else shift.data_ptr(), out.data_ptr(), B, C, H, W,
                                              ku.ptr, ku.n, kd.ptr, kd.n, _DT[x.dtype], _stream(x))
        else:
            dy = dy.contiguous(memory_format=cl).to(x.dtype)
            rc = L.afr_filtered_gelu_nhwc_bwd(x.data_ptr(), None if res is None else res.data_ptr(),
                                              None if scale is None else scale.data_ptr(),
                                              None if shift is None else shift.data_ptr(), dy.data_ptr(), out.data_ptr(),
                                              B, C, H, W, ku.ptr, ku.n, kd.ptr, kd.n, _DT[x.dtype], _stream(x))
    if rc == 7:                                   # AFR_ERR_UNSUPPORTED
        return None
    _check(rc)
    return out


def _fgelu_fwd(x, res, ku, kd, out=None):
    if out is None and _nhwc_eligible(x, ku, kd) and (res is None or res.dtype == x.dtype):
        y = _fgelu_nhwc(x, res, None, None, None, ku, kd)
        if y is not None:
            return y
    x = x.contiguous()
    res = None if res is None else res.contiguous()
    B, C, H, W = x.shape
    y = torch.empty_like(x) if out is None else out
    with _on_device(x.device):
        _check(_native.lib().afr_filtered_gelu_fwd(
            x.data_ptr(), None if res is None else res.data_ptr(), y.data_ptr(), B, C, H, W,
            ku.ptr, ku.n, kd.ptr, kd.n, _DT[x.dtype], _stream(x)))
    return y


def _fgelu_bwd(x, res, dy, ku, kd):
    if _nhwc_eligible(x, ku, kd) and (res is None or res.dtype == x.dtype):
        dx = _fgelu_nhwc(x, res, None, None, dy, ku, kd)
        if dx is not None:
            return dx
    x, dy = x.contiguous(), dy.contiguous()
    res = None if res is None else res.contiguous()
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    with _on_device(x.device):
        _check(_native.lib().afr_filtered_gelu_bwd(
            x.data_ptr(), None if res is None else res.data_ptr(), dy.data_ptr(), dx.data_ptr(),
            B, C, H, W, ku.ptr, ku.n, kd.ptr, kd.n, _DT[x.dtype], _stream(x)))
    return dx


def _needs_grad(*tensors):
    """Autograd bookkeeping (``Function.apply`` costs ~5 us of host time per call) is only paid when a gradient can
    actually flow: the no_grad sampler calls the raw launchers directly."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ---- autograd -------------------------------------------------------------------------
class _Up2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, out_dtype):
        ctx.k, ctx.hw, ctx.in_dtype = k, tuple(x.shape[-2:]), x.dtype
        return _up_fwd(x, k, out_dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, du):
        dx = _up_bwd(du, ctx.k, *ctx.hw)
        return dx.to(ctx.in_dtype), None, None


class _Down2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, k):
        ctx.k, ctx.hw = k, tuple(v.shape[-2:])
        return _down_fwd(v, k)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        return _down_bwd(dy, ctx.k, *ctx.hw), None


class _FilteredGelu(torch.autograd.Function):
    """Saves only x (and the residual): u, gelu(u) and their 4x-sized gradients are
    recomputed inside the backward kernel."""

    @staticmethod
    def forward(ctx, x, res, ku, kd):
        ctx.ku, ctx.kd = ku, kd
        ctx.has_res = res is not None
        if ctx.has_res:
            ctx.save_for_backward(x, res)
        else:
            ctx.save_for_backward(x)
        return _fgelu_fwd(x, res, ku, kd)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x, res = saved[0], (saved[1] if ctx.has_res else None)
        dx = _fgelu_bwd(x, res, dy.to(x.dtype), ctx.ku, ctx.kd)
        return dx, (dx if ctx.has_res else None), None, None


class _Up2xCat(torch.autograd.Function):
    """``torch.cat([skip, up2x(x)], dim=1)`` with the upsampler writing its result straight into the
    channel slice of the concatenated buffer (and its adjoint reading the gradient slice in place)."""

    @staticmethod
    def run(skip, x, k):
        B, C, H, W = x.shape
        Cs = skip.shape[1]
        out = torch.empty((B, Cs + C, 2 * H, 2 * W), dtype=skip.dtype, device=skip.device)
        out[:, :Cs].copy_(skip)
        with _on_device(x.device):
            _check(_native.lib().afr_up2x_fwd_strided(x.data_ptr(), out[:, Cs:].data_ptr(), B, C, H, W, out.stride(0),
                                                      k.ptr, k.n, _DT[x.dtype], _DT[out.dtype], _stream(x)))
        return out

    @staticmethod
    def forward(ctx, skip, x, k):
        ctx.k, ctx.shape, ctx.cs, ctx.in_dtype = k, tuple(x.shape), skip.shape[1], x.dtype
        return _Up2xCat.run(skip, x, k)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous()
        dx = torch.empty((B, C, H, W), dtype=g.dtype, device=g.device)
        with _on_device(g.device):
            _check(_native.lib().afr_up2x_bwd_strided(g[:, ctx.cs:].data_ptr(), dx.data_ptr(), B, C, H, W, g.stride(0),
                                                      ctx.k.ptr, ctx.k.n, _DT[g.dtype], _stream(g)))
        return g[:, :ctx.cs], dx.to(ctx.in_dtype), None


class _Up2xCatCL(torch.autograd.Function):
    """The same on channels-last memory: the upsampler's channel slice of the concatenated tensor has pixel stride
    Cs + C, which is all the NHWC kernels need."""

    @staticmethod
    def run(skip, x, k):
        B, C, H, W = x.shape
        Cs = skip.shape[1]
        out = torch.empty((B, Cs + C, 2 * H, 2 * W), dtype=skip.dtype, device=skip.device, memory_format=_CL)
        out[:, :Cs].copy_(skip)
        _up_nhwc(x, out[:, Cs:], k, False, B, C, H, W, C, Cs + C)
        return out

    @staticmethod
    def forward(ctx, skip, x, k):
        ctx.k, ctx.shape, ctx.cs, ctx.in_dtype = k, tuple(x.shape), skip.shape[1], x.dtype
        return _Up2xCatCL.run(skip, x, k)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous(memory_format=_CL)
        dx = torch.empty((B, C, H, W), dtype=g.dtype, device=g.device, memory_format=_CL)
        _down_nhwc(g[:, ctx.cs:], dx, ctx.k, True, B, C, 2 * H, 2 * W, ctx.cs + C, C)
        return g[:, :ctx.cs], dx.to(ctx.in_dtype), None


def up2x_cat(skip, x, filt):
    """``torch.cat([skip, custom_upsample(x, filt)], dim=1)`` (modules/ddpm_utils.py:344-345, 413-414) without
    materialising the upsampled tensor separately: saves one read and one write of it.  Shapes the strided
    kernels do not cover (N != 3, W % 4 != 0, a forced kernel path) take the two-step form."""
    x, skip = _require(x, keep_cl=True), _require(skip, "skip_x", keep_cl=True)
    k = _taps(filt)
    B, C, H, W = x.shape
    if (_is_cl(x) or _is_cl(skip)) and k.n == 3 and C % 4 == 0 and skip.shape[1] % 4 == 0 and skip.shape[0] == B \
            and tuple(skip.shape[2:]) == (2 * H, 2 * W) and _native.lib().afr_set_path(-1) != _native.PATHS["generic"]:
        x, skip = x.contiguous(memory_format=_CL), skip.contiguous(memory_format=_CL)
        return _Up2xCatCL.apply(skip, x, k) if _needs_grad(skip, x) else _Up2xCatCL.run(skip, x, k)
    x, skip = x.contiguous(), skip.contiguous()
    if (k.n == 3 and W % 4 == 0 and skip.shape[0] == B and tuple(skip.shape[2:]) == (2 * H, 2 * W)
            and (skip.shape[1] * 4 * H * W * skip.element_size()) % 32 == 0 and _native.PATHS_AUTO()):
        return _Up2xCat.apply(skip, x, k) if _needs_grad(skip, x) else _Up2xCat.run(skip, x, k)
    return torch.cat([skip, up2x(x, k, skip.dtype)], dim=1)


def _match_residual(x, residual):
    """``x + residual`` follows PyTorch type promotion (e.g. bf16 conv output + fp32 skip under
    autocast -> fp32), like the reference's ``x = x + residual`` (modules/ddpm_utils.py:128)."""
    residual = _require(residual, "residual", keep_cl=True)
    if residual.shape != x.shape:
        raise ValueError("afr: residual must match x in shape")
    if residual.dtype != x.dtype:
        dt = torch.promote_types(x.dtype, residual.dtype)
        x, residual = x.to(dt), residual.to(dt)
    return x, residual


# ---- public functions -----------------------------------------------------------------
def up2x(x, filt, out_dtype=None):
    """Zero-stuff x2 + depthwise N x N low-pass, 'same' zero padding, no gain."""
    x = _require(x, keep_cl=True)
    if not _needs_grad(x):
        return _up_fwd(x, _taps(filt), out_dtype or x.dtype)
    return _Up2x.apply(x, _taps(filt), out_dtype or x.dtype)


def down2x(x, filt):
    """Depthwise N x N low-pass then keep every 2nd row/column; output is dense (channels-last if x is)."""
    x = _require(x, keep_cl=True)
    return _Down2x.apply(x, _taps(filt)) if _needs_grad(x) else _down_fwd(x, _taps(filt))


def filtered_gelu(x, filt_up, filt_down, residual=None):
    """down2x(gelu(up2x(x + residual, filt_up)), filt_down) in ONE kernel (exact erf GELU).
    Replaces modules/ddpm_utils.py:123-125 / 128-131 / 137-139."""
    x = _require(x, keep_cl=True)
    if residual is not None:
        x, residual = _match_residual(x, residual)
    if not _needs_grad(x, residual):
        return _fgelu_fwd(x, residual, _taps(filt_up), _taps(filt_down))
    return _FilteredGelu.apply(x, residual, _taps(filt_up), _taps(filt_down))


def filtered_gelu_affine(x, scale, shift, filt_up, filt_down, residual=None):
    """Inference-only: ``filtered_gelu(x * scale[:, :, None, None] + shift[:, :, None, None] + residual)``
    in one kernel -- the normalise + affine step of the GroupNorm that precedes the activation
    (modules/ddpm_utils.py:122-125, 127-131) is folded into the kernel's load.  ``scale`` / ``shift``:
    float32 CUDA tensors [B, C].  No autograd; raises ``NotImplementedError`` for shapes / filter sizes
    the fused kernels do not cover (callers fall back to GroupNorm + ``filtered_gelu``)."""
    x = _require(x)
    if torch.is_grad_enabled() and (x.requires_grad or scale.requires_grad or shift.requires_grad):
        raise RuntimeError("afr: filtered_gelu_affine is forward-only (use GroupNorm + filtered_gelu to train)")
    B, C, H, W = x.shape
    ku, kd = _taps(filt_up), _taps(filt_down)
    if ku.n != 3 or kd.n != 3 or W % 4 != 0:
        raise NotImplementedError("afr: affine fusion needs N == 3 filters and W % 4 == 0")
    scale = scale.detach().to(torch.float32).contiguous()
    shift = shift.detach().to(torch.float32).contiguous()
    if tuple(scale.shape) != (B, C) or tuple(shift.shape) != (B, C) or not scale.is_cuda or not shift.is_cuda:
        raise ValueError("afr: scale and shift must be CUDA tensors of shape [B, C]")
    if residual is not None:
        x, residual = _match_residual(x, residual)
    y = torch.empty_like(x)
    with _on_device(x.device):
        _check(_native.lib().afr_filtered_gelu_affine_fwd(
            x.data_ptr(), None if residual is None else residual.data_ptr(), scale.data_ptr(), shift.data_ptr(),
            y.data_ptr(), B, C, H, W, ku.ptr, ku.n, kd.ptr, kd.n, _DT[x.dtype], _stream(x)))
    return y


def _actdown_ok(v, k):
    H, W = v.shape[-2:]
    return k.n == 3 and H % 2 == 0 and W % 8 == 0 and _native.lib().afr_set_path(-1) != _native.PATHS["generic"]


class _GeluDown2x(torch.autograd.Function):
    """down2x(gelu(v)) in one kernel; saves v only and recomputes gelu'(v) in the adjoint kernel."""

    @staticmethod
    def forward(ctx, v, k):
        B, C, H, W = v.shape
        y = torch.empty((B, C, H // 2, W // 2), dtype=v.dtype, device=v.device)
        with _on_device(v.device):
            _check(_native.lib().afr_gelu_down2x_fwd(v.data_ptr(), None, None, y.data_ptr(), B, C, H, W, k.ptr, k.n,
                                                     _DT[v.dtype], _stream(v)))
        ctx.k = k
        ctx.save_for_backward(v)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        (v,) = ctx.saved_tensors
        B, C, H, W = v.shape
        dv = torch.empty_like(v)
        dy = dy.contiguous().to(v.dtype)
        with _on_device(v.device):
            _check(_native.lib().afr_gelu_down2x_bwd(v.data_ptr(), dy.data_ptr(), dv.data_ptr(), B, C, H, W, ctx.k.ptr,
                                                     ctx.k.n, _DT[v.dtype], _stream(v)))
        return dv, None


def gelu_down2x(v, filt):
    """``custom_downsample(gelu(v), filt)`` (modules/ddpm_utils.py:172-173, 182-183: the second half of variant
    4's activation, after the GroupNorm on the 2x grid) in ONE kernel with autograd; gelu(v), a 4x-sized tensor,
    is never written.  Shapes the fused kernel does not cover run as F.gelu + down2x."""
    v = _require(v, "v")
    k = _taps(filt)
    if _actdown_ok(v, k):
        return _GeluDown2x.apply(v, k)
    # the composite keeps gelu(v) in fp32 so that bf16 tensors see ONE rounding, like the fused kernel
    return _Down2x.apply(torch.nn.functional.gelu(v.float()), k).to(v.dtype)


def gelu_down2x_affine(v, scale, shift, filt):
    """Inference-only: ``custom_downsample(gelu(v * scale[:, :, None, None] + shift[:, :, None, None]), filt)`` --
    GroupNorm's normalise + affine (statistics from ``groupnorm1_affine``), GELU, low-pass and decimation in one
    kernel (modules/ddpm_utils.py:171-173).  Raises ``NotImplementedError`` for unsupported shapes."""
    v = _require(v, "v")
    k = _taps(filt)
    if torch.is_grad_enabled() and (v.requires_grad or scale.requires_grad or shift.requires_grad):
        raise RuntimeError("afr: gelu_down2x_affine is forward-only (use GroupNorm + gelu_down2x to train)")
    if not _actdown_ok(v, k):
        raise NotImplementedError("afr: gelu_down2x_affine needs N == 3, even H and W % 8 == 0")
    B, C, H, W = v.shape
    scale = scale.detach().to(torch.float32).contiguous()
    shift = shift.detach().to(torch.float32).contiguous()
    if tuple(scale.shape) != (B, C) or tuple(shift.shape) != (B, C) or not scale.is_cuda or not shift.is_cuda:
        raise ValueError("afr: scale and shift must be CUDA tensors of shape [B, C]")
    y = torch.empty((B, C, H // 2, W // 2), dtype=v.dtype, device=v.device)
    with _on_device(v.device):
        _check(_native.lib().afr_gelu_down2x_fwd(v.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), B, C, H, W,
                                                 k.ptr, k.n, _DT[v.dtype], _stream(v)))
    return y


def groupnorm1_affine(x, weight, bias, eps):
    """Statistics of ``GroupNorm(1, C)`` as the [B, C] (scale, shift) pair that
    ``filtered_gelu_affine`` consumes, in ONE kernel.  Forward only."""
    x = _require(x)
    B, C, H, W = x.shape
    w = weight.detach().to(torch.float32).contiguous()
    b = bias.detach().to(torch.float32).contiguous()
    scale = torch.empty((B, C), dtype=torch.float32, device=x.device)
    shift = torch.empty((B, C), dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _check(_native.lib().afr_groupnorm1_affine(x.data_ptr(), w.data_ptr(), b.data_ptr(), float(eps),
                                                   scale.data_ptr(), shift.data_ptr(), B, C, H, W,
                                                   _DT[x.dtype], _stream(x)))
    return scale, shift


# ---- GroupNorm(1, C) folded into its neighbours, with autograd -------------------
def _gn_stats(h, weight, bias, eps, add=None):
    """One reduction kernel: per-(sample, channel) scale / shift of GroupNorm(1, C) (+ `add` [B, C] folded into
    shift) and the per-sample mean / rstd the backward needs."""
    B, C, H, W = h.shape
    w = weight.detach().to(torch.float32).contiguous()
    b = bias.detach().to(torch.float32).contiguous()
    scale = torch.empty((B, C), dtype=torch.float32, device=h.device)
    shift = torch.empty((B, C), dtype=torch.float32, device=h.device)
    mean = torch.empty((B, 1), dtype=torch.float32, device=h.device)
    rstd = torch.empty((B, 1), dtype=torch.float32, device=h.device)
    a = None if add is None else add.detach().to(torch.float32).contiguous()
    with _on_device(h.device):
        _check(_native.lib().afr_groupnorm1_stats(h.data_ptr(), w.data_ptr(), b.data_ptr(), float(eps),
                                                  None if a is None else a.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                  mean.data_ptr(), rstd.data_ptr(), B, C, H, W, _DT[h.dtype], _stream(h)))
    return scale, shift, mean, rstd


# Our three-launch GroupNorm backward (csrc/afr_norm.cu) instead of ATen's; AFR_GN_BWD=aten switches back (A/B).
USE_OWN_GN_BACKWARD = os.environ.get("AFR_GN_BWD", "afr") != "aten"


def _gn_backward(dz, h, mean, rstd, weight):
    """GroupNorm(1, C) backward (dz = gradient of the normalised + affine output) -> (dh, dgamma, dbeta)."""
    B, C, H, W = h.shape
    cl = _is_cl(h)
    ok = (USE_OWN_GN_BACKWARD and h.dtype in _DT and weight.dtype == torch.float32 and weight.is_contiguous()
          and ((cl and C % 32 == 0) or (not cl and h.is_contiguous() and (H * W) % 4 == 0)))
    if ok:
        dz = dz.to(h.dtype)
        dz = dz.contiguous(memory_format=torch.channels_last) if cl else dz.contiguous()
        dh = torch.empty_like(h)
        dw = torch.empty(C, dtype=torch.float32, device=h.device)
        db = torch.empty(C, dtype=torch.float32, device=h.device)
        work = torch.empty(2 * B * C + 2 * B, dtype=torch.float32, device=h.device)
        with _on_device(h.device):
            _check(_native.lib().afr_groupnorm1_bwd(h.data_ptr(), dz.data_ptr(), weight.data_ptr(), mean.data_ptr(),
                                                    rstd.data_ptr(), dh.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                                    work.data_ptr(), B, C, H, W, _DT[h.dtype], 1 if cl else 0, _stream(h)))
        return dh, dw, db
    # ATen's own kernels.  Called directly, the op reads its tensors as dense NCHW whatever their strides (measured:
    # channels-last inputs give dgamma / dbeta that are off by O(1)), so they are made NCHW-contiguous first.
    return torch.ops.aten.native_group_norm_backward(dz.contiguous(), h.contiguous(), mean, rstd, weight, B, C, H * W, 1,
                                                     [True, True, True])


class _NormFilteredGelu(torch.autograd.Function):
    """filtered_gelu(GroupNorm(1, C)(h) + residual): statistics in one reduction kernel, normalise + affine inside
    the fused activation kernel's load -- the normalised tensor is never written.  Backward: the adjoint kernel
    recomputes it the same way and yields the gradient of the normalised tensor; ATen finishes the GroupNorm."""

    @staticmethod
    def run(h, weight, bias, eps, res, ku, kd):
        B, C, H, W = h.shape
        if _is_cl(h) and not _nhwc_eligible(h, ku, kd):
            h = h.contiguous()
        scale, shift, mean, rstd = _gn_stats(h, weight, bias, eps)        # layout-agnostic: a sample is one dense block
        if _nhwc_eligible(h, ku, kd):
            y = _fgelu_nhwc(h, res, scale, shift, None, ku, kd)
            if y is not None:
                return y, scale, shift, mean, rstd
            h = h.contiguous()
            scale, shift, mean, rstd = _gn_stats(h, weight, bias, eps)
        res = None if res is None else res.contiguous()
        y = torch.empty_like(h)
        with _on_device(h.device):
            _check(_native.lib().afr_filtered_gelu_affine_fwd(
                h.data_ptr(), None if res is None else res.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(),
                B, C, H, W, ku.ptr, ku.n, kd.ptr, kd.n, _DT[h.dtype], _stream(h)))
        return y, scale, shift, mean, rstd

    @staticmethod
    def forward(ctx, h, weight, bias, eps, res, ku, kd):
        y, scale, shift, mean, rstd = _NormFilteredGelu.run(h, weight, bias, eps, res, ku, kd)
        ctx.ku, ctx.kd, ctx.has_res = ku, kd, res is not None
        ctx.save_for_backward(h, weight, scale, shift, mean, rstd, *(() if res is None else (res,)))
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        h, weight, scale, shift, mean, rstd = ctx.saved_tensors[:6]
        res = ctx.saved_tensors[6] if ctx.has_res else None
        B, C, H, W = h.shape
        dz = None
        if _nhwc_eligible(h, ctx.ku, ctx.kd):
            dz = _fgelu_nhwc(h, res, scale, shift, dy, ctx.ku, ctx.kd)
        if dz is not None:
            dh, dw, db = _gn_backward(dz, h, mean, rstd, weight)
            return dh, dw, db, None, (dz if ctx.has_res else None), None, None
        h = h.contiguous()
        res = None if res is None else res.contiguous()
        dz = torch.empty_like(h)
        dy = dy.contiguous().to(h.dtype)
        with _on_device(h.device):
            _check(_native.lib().afr_filtered_gelu_affine_bwd(
                h.data_ptr(), None if res is None else res.data_ptr(), scale.data_ptr(), shift.data_ptr(), dy.data_ptr(),
                dz.data_ptr(), B, C, H, W, ctx.ku.ptr, ctx.ku.n, ctx.kd.ptr, ctx.kd.n, _DT[h.dtype], _stream(h)))
        dh, dw, db = _gn_backward(dz, h, mean, rstd, weight)
        return dh, dw, db, None, (dz if ctx.has_res else None), None, None


class _NormAddEmb(torch.autograd.Function):
    """GroupNorm(1, C)(h) + emb[:, :, None, None] as statistics kernel + ONE apply pass (emb folded into the shift;
    emb may be None: a plain GroupNorm in the tensor's own memory format with the three-launch backward)."""

    @staticmethod
    def run(h, weight, bias, eps, emb):
        B, C, H, W = h.shape
        scale, shift, mean, rstd = _gn_stats(h, weight, bias, eps, add=emb)
        y = torch.empty_like(h)                                        # keeps h's memory format
        apply = _native.lib().afr_affine_apply_nhwc if (_is_cl(h) and C % 4 == 0) else _native.lib().afr_affine_apply
        if _is_cl(h) and C % 4 != 0:
            h = h.contiguous()
            y = torch.empty_like(h)
        with _on_device(h.device):
            _check(apply(h.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), B, C, H, W, _DT[h.dtype], _stream(h)))
        return y, mean, rstd

    @staticmethod
    def forward(ctx, h, weight, bias, eps, emb):
        y, mean, rstd = _NormAddEmb.run(h, weight, bias, eps, emb)
        ctx.save_for_backward(h, weight, mean, rstd)
        ctx.emb_dtype = None if emb is None else emb.dtype
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        h, weight, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        dh, dw, db = _gn_backward(dy, h, mean, rstd, weight)
        return dh, dw, db, None, (None if ctx.emb_dtype is None else dy.sum(dim=(2, 3)).to(ctx.emb_dtype))


def norm_fusable(h, norm, n_up=3, n_down=3):
    """Can GroupNorm `norm` applied to `h` be folded into the kernels that follow it?  (N == 3 kernels, rows of
    4-element groups, GroupNorm(1, C) with affine; with autograd only in fp32 -- ATen's backward takes our fp32
    statistics.)"""
    if not (isinstance(h, torch.Tensor) and h.is_cuda and h.dim() == 4 and h.dtype in _DT):
        return False
    if norm.num_groups != 1 or not norm.affine or n_up != 3 or n_down != 3 or h.shape[-1] % 4 != 0:
        return False
    if torch.is_grad_enabled() and (h.requires_grad or norm.weight.requires_grad):
        return h.dtype == torch.float32 and norm.weight.dtype == torch.float32 and not torch.is_autocast_enabled()
    return True


def norm_filtered_gelu(h, norm, filt_up, filt_down, residual=None):
    """``filtered_gelu(norm(h) + residual)`` for a ``GroupNorm(1, C)`` module `norm` (modules/ddpm_utils.py:122-125,
    127-131) without writing the normalised tensor; differentiable.  Check ``norm_fusable`` first."""
    h = _require(h, "h", keep_cl=True)
    if residual is not None:
        residual = _require(residual, "residual", keep_cl=True)
        if residual.dtype != h.dtype or residual.shape != h.shape:
            raise ValueError("afr: residual must match h in shape and dtype")
    if not _needs_grad(h, residual, norm.weight, norm.bias):
        return _NormFilteredGelu.run(h, norm.weight, norm.bias, norm.eps, residual, _taps(filt_up), _taps(filt_down))[0]
    return _NormFilteredGelu.apply(h, norm.weight, norm.bias, norm.eps, residual, _taps(filt_up), _taps(filt_down))


def norm_add_emb(h, norm, emb):
    """``norm(h) + emb[:, :, None, None]`` (the tail of every Down / Up stage, modules/ddpm_utils.py:385-387,
    415-417: GroupNorm, ``.repeat`` of the time embedding, add) in one statistics kernel + one apply pass."""
    h = _require(h, "h", keep_cl=True)
    if emb.dim() != 2 or tuple(emb.shape) != tuple(h.shape[:2]):
        raise ValueError("afr: emb must be [B, C]")
    if not _needs_grad(h, emb, norm.weight, norm.bias):
        return _NormAddEmb.run(h, norm.weight, norm.bias, norm.eps, emb)[0]
    return _NormAddEmb.apply(h, norm.weight, norm.bias, norm.eps, emb)


def groupnorm1(h, norm):
    """``norm(h)`` for a ``GroupNorm(1, C)`` module through the statistics + apply kernels and the three-launch
    backward: same values as ``nn.GroupNorm``, but it keeps a channels-last tensor channels-last (ATen's GroupNorm
    returns NCHW) and costs fewer launches.  Check ``norm_fusable`` first."""
    h = _require(h, "h", keep_cl=True)
    if not _needs_grad(h, norm.weight, norm.bias):
        return _NormAddEmb.run(h, norm.weight, norm.bias, norm.eps, None)[0]
    return _NormAddEmb.apply(h, norm.weight, norm.bias, norm.eps, None)


def custom_upsample(x, sinc_filter, factor=2):
    """Signature of modules/filtrs.py:79.  Like the reference, the result is float32
    whatever the input dtype (filtrs.py:85 allocates an fp32 buffer)."""
    if factor != 2:
        raise NotImplementedError("afr: only factor=2 is implemented (every reference call site uses 2)")
    return up2x(x, sinc_filter, out_dtype=torch.float32)


def custom_downsample(x, jinc_filter, factor=2):
    """Signature of modules/filtrs.py:71."""
    if factor != 2:
        raise NotImplementedError("afr: only factor=2 is implemented (every reference call site uses 2)")
    return down2x(x, jinc_filter)


def rotate(x, degrees):
    """scipy.ndimage.rotate(x, degrees, axes=(2,3), reshape=False, mode='grid-wrap') on device
    (modules/ddpm_models.py:421-429).  No autograd (the sampler runs under no_grad)."""
    x = _require(x)
    if x.dtype != torch.float32:
        raise TypeError("afr: rotate needs float32")
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    with _on_device(x.device):
        _check(_native.lib().afr_rotate_periodic_cubic(x.data_ptr(), y.data_ptr(), B, C, H, W,
                                                       float(degrees), 0, _stream(x)))
    return y


def ddpm_update_(x, eps, noise, ca, cb, cc):
    """In place  x <- ca * (x - cb * eps) + cc * noise  (modules/ddpm_models.py:374)."""
    for t in (x, eps) + (() if noise is None else (noise,)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("afr: ddpm_update_ needs contiguous float32 CUDA tensors")
    with _on_device(x.device):
        _check(_native.lib().afr_ddpm_update(x.data_ptr(), eps.data_ptr(),
                                             None if noise is None else noise.data_ptr(),
                                             x.numel(), float(ca), float(cb), float(cc), _stream(x)))
    return x


def ddpm_update_table_(x, eps, noise, table, step):
    """Like ``ddpm_update_`` but (ca, cb, cc) are read on the device from ``table[step]``
    (``table``: float32 CUDA [T, 3]; ``step``: int32 CUDA scalar tensor) -- graph-capturable."""
    for t in (x, eps, table) + (() if noise is None else (noise,)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("afr: ddpm_update_table_ needs contiguous float32 CUDA tensors")
    if not (step.is_cuda and step.dtype == torch.int32):
        raise RuntimeError("afr: step must be an int32 CUDA tensor")
    with _on_device(x.device):
        _check(_native.lib().afr_ddpm_update_table(x.data_ptr(), eps.data_ptr(),
                                                   None if noise is None else noise.data_ptr(),
                                                   x.numel(), table.data_ptr(), step.data_ptr(), _stream(x)))
    return x
