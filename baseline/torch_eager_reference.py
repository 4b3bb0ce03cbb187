"""The reference's resampling path restated with the same PyTorch calls it makes
(modules/filtrs.py:71-94, modules/ddpm_utils.py:123-125): zero-stuff + depthwise F.conv2d
('same' padding, groups=C) + strided slice + F.gelu.  BASELINE / TEST INFRASTRUCTURE ONLY (never imported by the product package): it is
what the upstream code executes when run in eager mode on a GPU, i.e. the "kernel to beat" on the
B200 (the reference itself cannot travel to the GPU box; it has no setup.py, so there is nothing to
pip-install into baseline/_ref).  Pinned by tests/test_oracle.py against the same golden fixtures
as the C oracle."""
import torch
import torch.nn.functional as F


def custom_downsample(x, jinc_filter, factor=2):
    k = jinc_filter[None, None].to(x.device).repeat(x.size(1), 1, 1, 1)      # per-call upload + repeat, as upstream
    return F.conv2d(x, k, padding="same", groups=x.size(1))[:, :, ::factor, ::factor]


def custom_upsample(x, sinc_filter, factor=2):
    b, c, h, w = x.shape
    up = torch.zeros(b, c, h * factor, w * factor, device=x.device)
    up[:, :, ::factor, ::factor] = x
    k = sinc_filter[None, None].to(x.device).repeat(c, 1, 1, 1)
    return F.conv2d(up, k, padding="same", groups=c)


def filtered_gelu(x, sinc_filter, jinc_filter):
    return custom_downsample(F.gelu(custom_upsample(x, sinc_filter)), jinc_filter)
