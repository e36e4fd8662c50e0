"""torch-CPU port of the reference's convolution formulation -- TEST INFRASTRUCTURE.

Second half of the oracle (see pqmf_oracle.py for the rules on who may import
`oracle/`).  Where pqmf_oracle.py evaluates the closed form in float64, this file
follows the reference's own operator sequence (de-interleave -> conv1d -> crop ->
sign mask, and the transposed chain) so that

  * on the same box with the same torch build it reproduces the reference's fp32
    results bit for bit (pinned in tests/test_oracle.py against tests/golden/), and
  * timed on the GPU box's host cores it is the `cpu_baseline` / `--impl reference`
    number of bench.py (kind "port": /root/reference does not travel to that box,
    and the reference is Python, so there is nothing to compile into oracle/_ref).

Reference lines followed: pqmf.py:13-22 (mask), :115-130 (polyphase analysis),
:133-157 (polyphase synthesis), :160-199 (classic pair), :306-354 (cached variant,
with the non-cached `cached_conv.Conv1d` body that is baked into
PitchShifterPvoc/torchscript/pqmfpvoc.ts: F.pad(x, _pad) + conv1d).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def flip_sign_odd_bands_even_frames(y: torch.Tensor) -> torch.Tensor:
    # pqmf.py:19-22 -- built the same way (ones, strided fill, multiply) so rounding/-0.0 match
    mask = torch.ones_like(y)
    mask[..., 1::2, ::2] = -1
    return y * mask


def analysis_polyphase(x: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """[B,1,T] -> [B,M,T/M]; pqmf.py:126-129 then :257."""
    m, length = hk.shape
    b, c, t = x.shape
    if t % m:
        raise ValueError("polyphase analysis needs T % n_band == 0")
    phases = x.reshape(b, c, t // m, m).permute(0, 1, 3, 2).reshape(b, c * m, t // m)
    w = hk.reshape(m, length // m, m).permute(0, 2, 1)  # [band, phase, tap]
    y = F.conv1d(phases, w, padding=w.shape[-1] // 2)[..., :-1]
    return flip_sign_odd_bands_even_frames(y)


def analysis_classic(x: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """[B,1,T] -> [B,M,floor(T/M)]; pqmf.py:171-176 then :257."""
    y = F.conv1d(x, hk.unsqueeze(1), stride=hk.shape[0], padding=hk.shape[-1] // 2)[..., :-1]
    return flip_sign_odd_bands_even_frames(y)


def synthesis_polyphase(s: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """[B,M,F] -> [B,1,M*F]; pqmf.py:283 then :145-156."""
    m, length = hk.shape
    s = flip_sign_odd_bands_even_frames(s)
    w = hk.flip(-1).reshape(m, length // m, m).permute(2, 0, 1)  # [phase, band, tap]
    y = F.conv1d(s, w, padding=w.shape[-1] // 2 + 1)[..., :-1] * m
    y = y.flip(1)
    b, _, f = y.shape
    y = y.reshape(b, 1, m, f).permute(0, 1, 3, 2).reshape(b, 1, f * m)
    return y[..., 2 * m :]


def synthesis_classic(s: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """[B,M,F] -> [B,1,M*F]; pqmf.py:283 then :190-198 (zero-stuffing)."""
    m, length = hk.shape
    s = flip_sign_odd_bands_even_frames(s)
    up = torch.zeros(*s.shape[:2], m * s.shape[-1]).to(s)
    up[..., ::m] = s * m
    return F.conv1d(up, hk.flip(-1).unsqueeze(0), padding=length // 2)[..., 1:]


def _odd(w: torch.Tensor) -> torch.Tensor:
    return F.pad(w, (0, 1)) if w.shape[-1] % 2 == 0 else w  # pqmf.py:35-41


def analysis_cached_offline(x: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """CachedPQMF.forward in non-cached mode: pqmf.py:310, :316-323, :339-343 + .ts conv body."""
    m, length = hk.shape
    w = _odd(hk).unsqueeze(1)  # [M,1,L+1]
    pad = ((w.shape[-1] - 1 + 1) // 2,) * 2  # cc.get_padding(L+1) = (L/2, L/2) for odd kernels
    y = F.conv1d(F.pad(x, pad), w, stride=m)
    return flip_sign_odd_bands_even_frames(y)


def synthesis_cached_offline(s: torch.Tensor, hk: torch.Tensor) -> torch.Tensor:
    """CachedPQMF.inverse in non-cached mode: pqmf.py:312-314, :326-332, :345-354 + .ts conv body."""
    m, length = hk.shape
    w = _odd(hk.flip(-1).reshape(m, length // m, m).permute(2, 0, 1))  # [M,M,K+1]
    pad = (w.shape[-1] // 2,) * 2
    s = flip_sign_odd_bands_even_frames(s)
    y = F.conv1d(F.pad(s, pad), w) * m
    y = y.flip(1)
    y = y.permute(0, 2, 1)
    y = y.reshape(y.shape[0], y.shape[1], -1, m).permute(0, 2, 1, 3)
    return y.reshape(y.shape[0], y.shape[1], -1)


def time_roundtrip(hk: torch.Tensor, batch: int, n_samples: int, repeats: int = 5, warmup: int = 2, threads: int | None = None,
                   seed: int = 1234):
    """Wall-clock of analysis_polyphase + synthesis_polyphase on all host threads.
    Returns (best_seconds_fwd, best_seconds_inv, threads_used)."""
    import os
    import time

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    x = (0.5 * torch.randn(batch, 1, n_samples, generator=g)).clamp_(-1, 1)
    best_f = best_i = float("inf")
    with torch.no_grad():
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            y = analysis_polyphase(x, hk)
            t1 = time.perf_counter()
            synthesis_polyphase(y, hk)
            t2 = time.perf_counter()
            if it >= warmup:
                best_f = min(best_f, t1 - t0)
                best_i = min(best_i, t2 - t1)
    return best_f, best_i, threads
