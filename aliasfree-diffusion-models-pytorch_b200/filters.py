"""Low-pass filter design on the host (runs once per block at construction).

``circularLowpassKernel`` mirrors the reference function of the same name
(modules/filtrs.py:20-37): a radial jinc  w*J1(w r)/(2 pi r)  sampled on an N x N grid
centred at (N-1)/2, optionally multiplied by the outer product of a Kaiser window,
normalised to unit sum and returned as a float32 CPU tensor.  The taps never become a
device tensor: the kernels receive them by value in parameter space.
"""
import math

import numpy as np
import torch
from scipy.special import j1


def circularLowpassKernel(omega_c=np.pi, N=6, beta=None):
    N = int(N)
    if N < 1:
        raise ValueError("N must be >= 1")
    centre = 0.5 * (N - 1)
    d = np.arange(N, dtype=np.float64) - centre
    r = np.hypot(d[:, None], d[None, :])
    at_zero = r == 0.0
    r_safe = np.where(at_zero, 1.0, r)
    taps = np.where(at_zero, omega_c * omega_c / (4.0 * math.pi),
                    omega_c * j1(omega_c * r_safe) / (2.0 * math.pi * r_safe))
    if beta is not None:
        win = np.kaiser(N, beta)
        taps = taps * win[:, None] * win[None, :]
    taps = taps / taps.sum()
    return torch.from_numpy(taps.astype(np.float32))


def taps_from_settings(f_settings, which):
    """``which`` is 'up' or 'down'; ``f_settings`` is the reference's dict with keys
    kernel_size, kaiser_beta, omega_c_down, omega_c_up (modules/ddpm_tasks.py:47-51)."""
    if f_settings is None:
        raise ValueError("f_settings is empty")
    return circularLowpassKernel(omega_c=f_settings["omega_c_" + which],
                                 N=f_settings["kernel_size"], beta=f_settings["kaiser_beta"])
