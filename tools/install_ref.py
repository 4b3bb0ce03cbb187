#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the
GPU box with the gpurun snapshot (the box has no ``/root/reference``).

  python tools/install_ref.py [--src /root/reference]

The reference has no setup.py / pyproject.toml (nothing to ``pip install``): its importable unit is
the ``modules/`` package at the repository root, imported as ``modules.filtrs`` etc. with that root on
``sys.path``.  This copies exactly that package, byte for byte, and records a SHA-256 manifest next
to it; nothing from it is ever committed (``.gitignore``: ``baseline/_ref/``).  ``baseline/ref_loader``
imports it (preferring a live ``/root/reference`` checkout) for the tests that run the reference's own
UNet / sampler / train loop on our kernels under ``patch()`` and for bench.py's eager-GPU arms.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def install(src="/root/reference", verbose=False):
    """Returns the install directory, or None when there is no reference checkout to copy from."""
    pkg = os.path.join(src, "modules")
    if not os.path.isdir(pkg):
        return DST if os.path.isdir(os.path.join(DST, "modules")) else None
    out = os.path.join(DST, "modules")
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(pkg)):
        if not name.endswith(".py"):
            continue
        s, d = os.path.join(pkg, name), os.path.join(out, name)
        data = open(s, "rb").read()
        manifest["modules/" + name] = hashlib.sha256(data).hexdigest()
        if not os.path.exists(d) or open(d, "rb").read() != data:
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
    for extra in ("LICENSE",):
        s = os.path.join(src, extra)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(DST, extra))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("installed %d reference modules into %s" % (len(manifest), DST))
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if install(a.src, verbose=True) else 1)
