"""Deterministic, RNG-stream-independent parameter fill shared by the golden generator
(which fills the *reference* modules) and the parity tests (which fill ours).

Every tensor of ``module.state_dict()`` is drawn from a numpy Generator seeded by the
CRC32 of its key, so two modules with the same keys/shapes get bit-identical
parameters no matter in which order they were constructed.
"""
import zlib

import numpy as np
import torch


def fill_params_(module, salt=""):
    sd = module.state_dict()
    for key in sorted(sd):
        t = sd[key]
        if not t.dtype.is_floating_point:
            continue
        rng = np.random.default_rng(zlib.crc32((salt + key).encode()))
        v = rng.standard_normal(tuple(t.shape)).astype(np.float32)
        if t.ndim >= 2:
            fan_in = int(np.prod(t.shape[1:]))
            v *= 1.0 / np.sqrt(fan_in)
        elif key.endswith("weight"):          # norm scales
            v = 1.0 + 0.1 * v
        else:                                 # biases
            v *= 0.1
        with torch.no_grad():
            t.copy_(torch.from_numpy(v))
    return module


def seeded_array(tag, shape, scale=1.0):
    rng = np.random.default_rng(zlib.crc32(tag.encode()))
    return (scale * rng.standard_normal(shape)).astype(np.float32)
