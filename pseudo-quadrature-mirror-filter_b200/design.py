"""Host-side PQMF bank design (one-off, ~15 ms): same public helper names as the reference's pqmf.py.

Stays numpy/scipy on the host exactly like the reference (SURVEY.md layer L0b): the only requirement is
that `h` comes out bit-identical and `hk` within cos-ulp noise of what the reference registers, because the
CUDA kernels consume these buffers as given.  Reference lines: kaiser_filter pqmf.py:66-85, loss_wc :88-95,
get_prototype :98-112, get_qmf_bank :44-63, center_pad_next_pow_2 :26-32, make_odd :35-41.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def kaiser_filter(wc, atten, N=None):
    """Kaiser-windowed low-pass with cutoff `wc` (rad/sample) and stop-band attenuation `atten` dB.
    N (odd) overrides the minimum length that kaiserord proposes."""
    from scipy.signal import firwin, kaiserord

    n_est, beta = kaiserord(atten, wc / np.pi)
    n_est = 2 * (n_est // 2) + 1
    taps = n_est if N is None else N
    return firwin(taps, wc, window=("kaiser", beta), scale=False, fs=2 * np.pi)


def loss_wc(wc, atten, M, N):
    """Objective of Creusere & Mitra: worst autocorrelation sample of the prototype at non-zero multiples of 2M."""
    proto = kaiser_filter(wc, atten, N)
    auto = np.convolve(proto, proto[::-1], "full")
    lags = np.abs(auto[auto.shape[-1] // 2 :: 2 * M][1:])
    return np.max(lags)


def get_prototype(atten, M, N=None):
    """Prototype low-pass for an M-band bank: Nelder-Mead on the cutoff, started at 1/M."""
    from scipy.optimize import fmin

    best = fmin(lambda w: loss_wc(w, atten, M, N), 1 / M, disp=0)[0]
    return kaiser_filter(best, atten, N)


def get_qmf_bank(h: torch.Tensor, n_band: int) -> torch.Tensor:
    """Cosine-modulate the prototype into n_band band-pass filters, in float32 like the reference does
    (integer index grids times Python floats), so the registered `hk` carries the same rounding."""
    taps = h.shape[-1]
    band = torch.arange(n_band).reshape(-1, 1)
    pos = torch.arange(-(taps // 2), taps // 2 + 1)
    quarter = (-1) ** band * math.pi / 4
    carrier = torch.cos((2 * band + 1) * math.pi / (2 * n_band) * pos + quarter)
    return 2 * h * carrier


def center_pad_next_pow_2(x: torch.Tensor) -> torch.Tensor:
    """Zero-pad the last axis symmetrically (extra sample on the right) up to the next power of two."""
    target = 2 ** math.ceil(math.log2(x.shape[-1]))
    extra = target - x.shape[-1]
    return F.pad(x, (extra // 2, extra // 2 + int(extra % 2)))


def make_odd(x: torch.Tensor) -> torch.Tensor:
    """Append one zero when the last axis has even length."""
    return x if x.shape[-1] % 2 else F.pad(x, (0, 1))
