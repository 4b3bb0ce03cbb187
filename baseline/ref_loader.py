"""Import the UNMODIFIED reference (``modules.filtrs``, ``modules.ddpm_utils``, ``modules.ddpm_models``).

Search order: a live checkout at ``/root/reference`` (the build container), then the byte-for-byte
install under ``baseline/_ref`` that ``tools/install_ref.py`` makes (git-ignored; it travels to the GPU
box inside the gpurun snapshot).  The two plotting packages the reference imports at module top
(matplotlib, imageio -- absent from this image) are replaced by empty stand-ins; nothing else is
touched.  Test / baseline infrastructure only: the product package never imports this file.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))


def reference_root():
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "modules", "filtrs.py")):
            return root
    return None


def available():
    return reference_root() is not None


def load():
    """Returns ``(filtrs, ddpm_utils, ddpm_models)`` of the reference, or raises ImportError."""
    root = reference_root()
    if root is None:
        raise ImportError("no reference checkout: neither /root/reference nor baseline/_ref "
                          "(run tools/install_ref.py in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    return tuple(importlib.import_module("modules." + m) for m in ("filtrs", "ddpm_utils", "ddpm_models"))
