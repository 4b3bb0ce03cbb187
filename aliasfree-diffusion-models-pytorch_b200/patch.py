"""The reference's only plugin surface is Python name binding:
``custom_upsample`` / ``custom_downsample`` are star-imported globals of
``modules.ddpm_utils`` and the block classes are globals of ``modules.ddpm_models``.
``patch()`` rebinds those names to this package so that the reference's own ``UNet``,
``train`` and ``Diffusion.sample`` run unchanged on the CUDA kernels; ``unpatch()`` restores
them.  The reference checkout must be importable (its root on ``sys.path``)."""
import importlib

from . import blocks, ops
from .diffusion import Diffusion

_FUNCS = ("custom_upsample", "custom_downsample")
_CLASSES = ("DoubleConv_F", "Down_F", "Up_F", "Down_FF", "Up_FF", "Down_FFF", "Up_FFF",
            "DoubleConv_F4", "Down_F4", "Up_F4")
_saved = {}


def patch(ddpm_utils=None, ddpm_models=None, filtrs=None, rotate_on_device=True):
    """Rebind the hot-path names inside the reference modules. Returns the list of patched names."""
    mods = {
        "filtrs": filtrs or importlib.import_module("modules.filtrs"),
        "ddpm_utils": ddpm_utils or importlib.import_module("modules.ddpm_utils"),
        "ddpm_models": ddpm_models or importlib.import_module("modules.ddpm_models"),
    }
    done = []
    for mname, mod in mods.items():
        for name in _FUNCS + (_CLASSES if mname != "filtrs" else ()):
            if hasattr(mod, name):
                _saved.setdefault((mod, name), getattr(mod, name))
                setattr(mod, name, getattr(ops if name in _FUNCS else blocks, name))
                done.append(f"{mod.__name__}.{name}")
    if rotate_on_device and hasattr(mods["ddpm_models"], "Diffusion"):
        D = mods["ddpm_models"].Diffusion
        _saved.setdefault((D, "rotate_2d_matrix"), D.__dict__["rotate_2d_matrix"])
        D.rotate_2d_matrix = staticmethod(Diffusion.rotate_2d_matrix)
        done.append(f"{D.__module__}.Diffusion.rotate_2d_matrix")
    return done


def unpatch():
    for (owner, name), val in list(_saved.items()):
        setattr(owner, name, val)
        del _saved[(owner, name)]
