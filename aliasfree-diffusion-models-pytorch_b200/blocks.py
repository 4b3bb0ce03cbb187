"""Drop-in ``nn.Module`` blocks of the alias-free UNet (Configs B-D).

Same class names, constructor arguments, forward signatures and ``state_dict`` keys as the
reference blocks in modules/ddpm_utils.py (DoubleConv_F :97-143, Down_F :253, Up_F :276,
Down_FF :301, Up_FF :330, Down_FFF :360, Up_FFF :389, plus the baseline DoubleConv :76,
Down :199, Up :222, SelfAttention :54 that variants 0/1 need), so reference checkpoints
load with ``strict=True``.  The filter taps are plain attributes (``jinc_filter``,
``sinc_filter``), not buffers -- exactly like the reference, they are absent from the
state_dict.

What differs is the execution: every ``up -> GELU -> down`` triple is ONE launch of the
fused CUDA kernel (ops.filtered_gelu), the residual add of the second activation is folded
into that launch, and the standalone resamplers are single kernels as well; the 3x3
convolutions, GroupNorm and attention stay on cuDNN / cuBLAS.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .filters import taps_from_settings


# DoubleConv_F folds each GroupNorm's normalise + affine into the kernel that follows it -- the fused
# activation kernel, or the stage's embedding add -- in inference AND training;
# set to False to always run nn.GroupNorm separately.  (FUSE_GROUPNORM_INFERENCE: variant 4's
# inference-only fold.)
FUSE_GROUPNORM = os.environ.get("AFR_FUSE_GROUPNORM", "1") != "0"
FUSE_GROUPNORM_INFERENCE = FUSE_GROUPNORM


def _groupnorm1_affine(h, norm):
    """GroupNorm(1, C) as a per-(sample, channel) affine: returns (scale, shift), both [B, C] fp32,
    with  norm(h) == h * scale[:, :, None, None] + shift[:, :, None, None]  (one reduction kernel)."""
    return ops.groupnorm1_affine(h, norm.weight, norm.bias, norm.eps)


def _cached_taps(module, name):
    """``ops.Taps`` of the filter attribute `name`, rebuilt only when a user swaps the tensor (the conversion to a
    pinned-down fp32 host copy costs ~5 us per call otherwise)."""
    filt = getattr(module, name)
    cache = module.__dict__.setdefault("_taps_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] is not filt:
        hit = cache[name] = (filt, ops.Taps(filt))
    return hit[1]


def _conv3(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False)


class SelfAttention(nn.Module):
    """modules/ddpm_utils.py:54-74 (unchanged maths; stays on cuBLAS)."""

    def __init__(self, channels, size):
        super().__init__()
        self.channels, self.size = channels, size
        self.mha = nn.MultiheadAttention(channels, 4, batch_first=True)
        self.ln = nn.LayerNorm([channels])
        self.ff_self = nn.Sequential(nn.LayerNorm([channels]), nn.Linear(channels, channels),
                                     nn.GELU(), nn.Linear(channels, channels))

    def forward(self, x):
        tokens = x.flatten(2).transpose(1, 2)                    # [B, HW, C]
        q = self.ln(tokens)
        att = self.mha(q, q, q, need_weights=False)[0] + tokens   # same values, no [B,L,L] weights tensor
        att = self.ff_self(att) + att
        return att.transpose(1, 2).reshape(-1, self.channels, self.size, self.size)


class DoubleConv(nn.Module):
    """Baseline (unfiltered) block, modules/ddpm_utils.py:76-95."""

    def __init__(self, in_channels, out_channels, mid_channels=None, residual=False):
        super().__init__()
        mid = mid_channels or out_channels
        self.residual = residual
        self.double_conv = nn.Sequential(_conv3(in_channels, mid), nn.GroupNorm(1, mid), nn.GELU(),
                                         _conv3(mid, out_channels), nn.GroupNorm(1, out_channels))

    def forward(self, x):
        y = self.double_conv(x)
        return F.gelu(x + y) if self.residual else y


class DoubleConv_F(nn.Module):
    """conv3x3 -> GroupNorm -> [up2x -> GELU -> low-pass -> down2x] -> conv3x3 -> GroupNorm,
    and for ``residual=True`` a second filtered GELU applied to (x + that)."""

    def __init__(self, in_channels, out_channels, mid_channels=None, residual=False, f_settings=None):
        super().__init__()
        self.residual = residual
        self.f_settings = f_settings
        self.jinc_filter = taps_from_settings(f_settings, "down")
        self.sinc_filter = taps_from_settings(f_settings, "up")
        mid = mid_channels or out_channels
        self.conv1 = _conv3(in_channels, mid)
        self.norm1 = nn.GroupNorm(1, mid)
        self.gelu = nn.GELU()          # parameter-free; kept so the module tree matches the reference
        self.conv2 = _conv3(mid, out_channels)
        self.norm2 = nn.GroupNorm(1, out_channels)
        self._taps = None

    def _filters(self):
        key = (id(self.sinc_filter), id(self.jinc_filter))
        if self._taps is None or self._taps[0] != key:      # rebuilt if a user swaps the filters
            self._taps = (key, ops.Taps(self.sinc_filter), ops.Taps(self.jinc_filter))
        return self._taps[1], self._taps[2]

    def forward(self, x, emb=None):
        """``emb`` ([B, out_channels], optional, non-residual blocks only): the stage's time embedding, added to
        the output (``+ emb[:, :, None, None]``) inside the last GroupNorm's apply pass."""
        up, dn = self._filters()
        nu, nd = self.sinc_filter.shape[-1], self.jinc_filter.shape[-1]
        h = self.conv1(x)
        if FUSE_GROUPNORM and ops.norm_fusable(h, self.norm1, nu, nd):
            h = ops.norm_filtered_gelu(h, self.norm1, up, dn)      # norm1 folded into the activation kernel
        else:
            h = ops.filtered_gelu(self.norm1(h), up, dn)
        h = self.conv2(h)
        if self.residual:
            if FUSE_GROUPNORM and ops.norm_fusable(h, self.norm2, nu, nd) and x.dtype == h.dtype:
                return ops.norm_filtered_gelu(h, self.norm2, up, dn, residual=x)
            return ops.filtered_gelu(self.norm2(h), up, dn, residual=x)     # gelu-filter(x + h), add fused
        if emb is not None:
            if FUSE_GROUPNORM and ops.norm_fusable(h, self.norm2) and emb.dtype == h.dtype:
                return ops.norm_add_emb(h, self.norm2, emb)
            return self.norm2(h) + emb[:, :, None, None]
        if FUSE_GROUPNORM and ops.norm_fusable(h, self.norm2):
            return ops.groupnorm1(h, self.norm2)                  # same values; keeps h's memory format
        return self.norm2(h)


def up_is_n3(block):
    return block.sinc_filter.shape[-1] == 3 and block.jinc_filter.shape[-1] == 3


class _TimeConditioned(nn.Module):
    """Shared tail of every Down*/Up* stage: ``x + Linear(SiLU(t))`` broadcast over H, W."""

    def _make_emb(self, emb_dim, out_channels):
        self.emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(emb_dim, out_channels))

    def _add_emb(self, x, t):
        return x + self.emb_layer(t)[:, :, None, None]

    def _blocks_then_emb(self, first, last, x, t):
        """The stage's two DoubleConv blocks and the embedding add; filtered blocks take the embedding into their
        last GroupNorm's apply pass (no broadcast tensor, no separate add)."""
        h = first(x)
        if type(last) is DoubleConv_F:
            return last(h, emb=self.emb_layer(t))
        return self._add_emb(last(h), t)


def _pair(block, cin, cout, mid=None, **kw):
    return nn.Sequential(block(cin, cin, residual=True, **kw), block(cin, cout, mid, **kw))


class Down(_TimeConditioned):
    """modules/ddpm_utils.py:199-220 (MaxPool, baseline blocks)."""

    def __init__(self, in_channels, out_channels, emb_dim=256):
        super().__init__()
        blocks = _pair(DoubleConv, in_channels, out_channels)
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), blocks[0], blocks[1])
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, t):
        return self._add_emb(self.maxpool_conv(x), t)


class Up(_TimeConditioned):
    """modules/ddpm_utils.py:222-245 (bilinear upsample, baseline blocks)."""

    def __init__(self, in_channels, out_channels, emb_dim=256):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = _pair(DoubleConv, in_channels, out_channels, in_channels // 2)
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, skip_x, t):
        return self._add_emb(self.conv(torch.cat([skip_x, self.up(x)], dim=1)), t)


class Down_F(_TimeConditioned):
    """Config C down stage: MaxPool kept, filtered activations inside (ddpm_utils.py:253-274)."""

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__()
        self.f_settings = f_settings
        blocks = _pair(DoubleConv_F, in_channels, out_channels, f_settings=f_settings)
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), blocks[0], blocks[1])
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, t):
        return self._blocks_then_emb(self.maxpool_conv[1], self.maxpool_conv[2], self.maxpool_conv[0](x), t)


class Up_F(_TimeConditioned):
    """Config C up stage (ddpm_utils.py:276-299)."""

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__()
        self.f_settings = f_settings
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = _pair(DoubleConv_F, in_channels, out_channels, in_channels // 2, f_settings=f_settings)
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, skip_x, t):
        return self._blocks_then_emb(self.conv[0], self.conv[1], torch.cat([skip_x, self.up(x)], dim=1), t)


class _FilteredDown(_TimeConditioned):
    _block = None

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__()
        self.f_settings = f_settings
        self.jinc_filter = taps_from_settings(f_settings, "down")
        kw = {"f_settings": f_settings} if issubclass(self._block, DoubleConv_F) else {}
        self.conv = _pair(self._block, in_channels, out_channels, **kw)
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, t):
        return self._blocks_then_emb(self.conv[0], self.conv[1], ops.custom_downsample(x, _cached_taps(self, "jinc_filter")), t)


class _FilteredUp(_TimeConditioned):
    _block = None

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__()
        self.f_settings = f_settings
        self.sinc_filter = taps_from_settings(f_settings, "up")
        kw = {"f_settings": f_settings} if issubclass(self._block, DoubleConv_F) else {}
        self.conv = _pair(self._block, in_channels, out_channels, in_channels // 2, **kw)
        self._make_emb(emb_dim, out_channels)

    def forward(self, x, skip_x, t):
        # the upsampler writes straight into its half of the concatenated buffer
        return self._blocks_then_emb(self.conv[0], self.conv[1], ops.up2x_cat(skip_x, x, _cached_taps(self, "sinc_filter")), t)


class Down_FF(_FilteredDown):
    """Config B: low-pass + decimate instead of MaxPool, baseline blocks (ddpm_utils.py:301-328)."""
    _block = DoubleConv


class Up_FF(_FilteredUp):
    """Config B: zero-stuff + low-pass instead of bilinear (ddpm_utils.py:330-358)."""
    _block = DoubleConv


class Down_FFF(_FilteredDown):
    """Config D: filtered resampling and filtered activations (ddpm_utils.py:360-387)."""
    _block = DoubleConv_F


class Up_FFF(_FilteredUp):
    """Config D (ddpm_utils.py:389-417)."""
    _block = DoubleConv_F


# ---- variant 4 (GroupNorm moved to the 2x grid, between upsample and GELU) ------------------------
# modules/ddpm_utils.py:145-197, 419-480.  The norm sits between the two resamplers, so the one-kernel
# up -> GELU -> down fusion does not apply; what fuses is the second half.  Training: up2x kernel ->
# nn.GroupNorm -> ONE kernel for GELU + low-pass + decimation (ops.gelu_down2x, with its own adjoint).
# Inference: the GroupNorm's normalise + affine moves into that kernel as well (statistics from one
# reduction kernel), so neither norm(u) nor gelu(norm(u)) -- each 4x the activation -- is written.
class DoubleConv_F4(DoubleConv_F):
    def _act(self, h, norm):
        u = ops.up2x(h, self.sinc_filter)
        if (FUSE_GROUPNORM_INFERENCE and not torch.is_grad_enabled() and u.is_cuda and self.jinc_filter.shape[-1] == 3
                and u.shape[-1] % 8 == 0 and u.shape[-2] % 2 == 0 and norm.num_groups == 1 and norm.affine):
            try:
                return ops.gelu_down2x_affine(u, *_groupnorm1_affine(u, norm), self.jinc_filter)
            except NotImplementedError:
                pass
        return ops.gelu_down2x(norm(u), self.jinc_filter)

    def forward(self, x):
        h = self._act(self.conv1(x), self.norm1)
        h = self.norm2(self.conv2(h))
        if self.residual:
            h = self._act(h + x, self.norm2)            # the reference re-uses norm2 here (:180)
        return h


class Down_F4(_FilteredDown):
    _block = DoubleConv_F4

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__(in_channels, out_channels, emb_dim, f_settings)
        self.norm1 = nn.GroupNorm(1, in_channels)       # present (unused) upstream: keeps state_dict keys


class Up_F4(_FilteredUp):
    _block = DoubleConv_F4

    def __init__(self, in_channels, out_channels, emb_dim=256, f_settings=None):
        super().__init__(in_channels, out_channels, emb_dim, f_settings)
        self.norm1 = nn.GroupNorm(1, in_channels // 2)
