"""Drivers around the sampler that the reference keeps in modules/ddpm_tasks.py, reduced to the
parts that exercise the accelerated path (no checkpoint / plotting / disk glue):

  rotation_results  Config-E sweep, modules/ddpm_tasks.py:346-369 (+ Results.ipynb:534, 640):
                    for every angle theta re-seed and run the full sampler with a per-step
                    rotation of theta/T degrees; frames are independent, so with several ranks they
                    are dealt round-robin and nothing is communicated.
  shift_results     the translation twin, modules/ddpm_tasks.py:371-392.
"""
import numpy as np
import torch

from .parallel import world


def _seed_all(seed):
    """What the reference's set_seed (modules/utils.py:98-105) does for the generators we use."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def rotation_results(model, diffusion, thetas, n=4, image_channels=3, seed=42, cuda_graph=False):
    """Returns ``(x_all, results_all)``: lists indexed like ``thetas``; entries of frames owned by
    other ranks are None.  Same seed for every frame, so frames differ only by the accumulated
    rotation (the reference's 'per-frame rotation sweep')."""
    rank, ws = world()
    x_all, results_all = [None] * len(thetas), [None] * len(thetas)
    for idx, th in enumerate(thetas):
        if idx % ws != rank:
            continue
        _seed_all(seed)
        x, results = diffusion.sample(model, n=n, image_channels=image_channels, theta=float(th),
                                      cuda_graph=cuda_graph)
        x_all[idx], results_all[idx] = x, results
    return x_all, results_all


def shift_results(model, diffusion, shifts, n=4, image_channels=3, seed=42):
    rank, ws = world()
    x_all = [None] * len(shifts)
    for idx, sh in enumerate(shifts):
        if idx % ws != rank:
            continue
        _seed_all(seed)
        x_all[idx] = diffusion.sample_shift(model, n=n, image_channels=image_channels, shift=sh)
    return x_all
