"""aliasfree_b200 -- B200-native alias-free resampling path.

The package directory is ``aliasfree-diffusion-models-pytorch_b200/``; import it as
``aliasfree_b200`` (see the alias package at the repository root).
"""
from . import _native
from .filters import circularLowpassKernel, taps_from_settings
from .ops import (Taps, custom_downsample, custom_upsample, ddpm_update_, down2x, filtered_gelu, gelu_down2x,
                  rotate, up2x, up2x_cat)
from .blocks import (DoubleConv, DoubleConv_F, DoubleConv_F4, Down, Down_F, Down_F4, Down_FF, Down_FFF,
                     SelfAttention, Up, Up_F, Up_F4, Up_FF, Up_FFF)
from .unet import UNet
from .diffusion import Diffusion
from .patch import patch, unpatch
from . import parallel
from .tasks import rotation_results, shift_results
from .hostio import HostPipeline

build = _native.build
set_path = _native.set_path
last_kernel = _native.last_kernel
launch_count = _native.launch_count

__all__ = ["circularLowpassKernel", "taps_from_settings", "Taps", "custom_upsample",
           "custom_downsample", "up2x", "down2x", "up2x_cat", "filtered_gelu", "gelu_down2x", "rotate", "ddpm_update_",
           "DoubleConv", "DoubleConv_F", "DoubleConv_F4", "Down", "Down_F", "Down_F4", "Down_FF", "Down_FFF",
           "Up", "Up_F", "Up_F4",
           "Up_FF", "Up_FFF", "SelfAttention", "UNet", "Diffusion", "patch", "unpatch", "parallel", "rotation_results", "shift_results", "HostPipeline",
           "build", "set_path", "last_kernel", "launch_count"]
