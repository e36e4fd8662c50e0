"""CPU oracle for the PQMF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  The product package
(`pqmf_b200`) never imports anything from `oracle/`; it fails loudly when its
CUDA library is missing.

What is restated here (numpy, float64 accumulation unless noted), each function
citing the reference lines it follows (paths relative to the reference repo):

* prototype design + cosine modulation + centre padding   pqmf.py:26-112
* analysis  (polyphase == classic == cached forward)      pqmf.py:115-130, 160-177, 339-343
* sign mask                                               pqmf.py:13-22
* synthesis (polyphase == classic; cached = +1 frame)     pqmf.py:133-157, 180-199, 345-354
* block streaming with carried FIR history                SURVEY.md A.4 (upstream
  `cached_conv` is a third-party package that is NOT vendored in the reference
  and is not installed here: **parity of the streaming mode is unpinned**; the
  offline CachedPQMF path is pinned by the committed TorchScript archive
  PitchShifterPvoc/torchscript/pqmfpvoc.ts, see tests/golden/make_golden.py).

Pinning: tests/test_oracle.py checks every function below against
tests/golden/*.npz, which were produced by importing the reference's own
pqmf.py (and its committed .ts archive) in the dev container.

The arithmetic is the closed form of SURVEY.md A.1 evaluated directly: no
convolution library, no folding trick.  With `dtype=np.float64` and the
reference's fp32 `hk` this is "the truth for the reference's coefficients";
the reference's own fp32 run sits ~1e-6 away from it.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# filter design (host side, one-off)                      reference pqmf.py:26-112
# --------------------------------------------------------------------------


def kaiser_lowpass(wc: float, atten: float, n_taps: int | None = None) -> np.ndarray:
    """Kaiser-window low-pass, cutoff `wc` rad/sample.  Follows pqmf.py:66-85:
    kaiserord on the normalised transition width wc/pi, length forced odd,
    firwin with scale=False and fs=2*pi (third-party: scipy.signal)."""
    from scipy.signal import firwin, kaiserord

    n_min, beta = kaiserord(atten, wc / np.pi)
    n_min = 2 * (n_min // 2) + 1
    n = n_min if n_taps is None else n_taps
    return firwin(n, wc, window=("kaiser", beta), scale=False, fs=2 * np.pi)


def aliasing_objective(wc: float, atten: float, n_band: int, n_taps: int | None) -> float:
    """Creusere-Mitra objective (pqmf.py:88-95): peak of the prototype
    autocorrelation sampled every 2M lags, zero lag excluded."""
    h = kaiser_lowpass(wc, atten, n_taps)
    g = np.convolve(h, h[::-1], "full")
    g = np.abs(g[g.shape[-1] // 2 :: 2 * n_band][1:])
    return float(np.max(g))


def design_prototype(atten: float, n_band: int, n_taps: int | None = None) -> np.ndarray:
    """Nelder-Mead over the cutoff starting at 1/M (pqmf.py:98-112). float64."""
    from scipy.optimize import fmin

    wc = fmin(lambda w: aliasing_objective(w, atten, n_band, n_taps), 1 / n_band, disp=0)[0]
    return kaiser_lowpass(wc, atten, n_taps)


def modulate_bank(h32: np.ndarray, n_band: int) -> np.ndarray:
    """hk[k, t] = 2 h[t] cos((2k+1) pi/(2M) (t - N//2) + (-1)^k pi/4), computed in
    float32 the way pqmf.py:54-61 does it (int64 index x python float -> fp32).
    torch is used for the fp32 cosine so the table matches the reference's
    platform libm path; the values are compared numerically (<=1e-8), not by hash."""
    import torch

    h = torch.from_numpy(np.asarray(h32, np.float32))
    n = h.shape[-1]
    k = torch.arange(n_band).reshape(-1, 1)
    t = torch.arange(-(n // 2), n // 2 + 1)
    phase = (-1) ** k * math.pi / 4
    hk = 2 * h * torch.cos((2 * k + 1) * math.pi / (2 * n_band) * t + phase)
    return hk.numpy()


def pad_pow2_centered(hk: np.ndarray) -> np.ndarray:
    """Centre-pad the last axis to the next power of two (pqmf.py:26-32):
    left pad//2, right pad//2 + pad%2."""
    n = hk.shape[-1]
    length = 2 ** math.ceil(math.log2(n))
    pad = length - n
    width = [(0, 0)] * (hk.ndim - 1) + [(pad // 2, pad // 2 + pad % 2)]
    return np.pad(hk, width)


def design_bank(atten: float, n_band: int):
    """(h fp32 [N], hk fp32 [M, L]) exactly as PQMF.__init__ builds them (pqmf.py:216-231)."""
    h = design_prototype(atten, n_band).astype(np.float32)
    hk = pad_pow2_centered(modulate_bank(h, n_band)).astype(np.float32)
    return h, hk


# --------------------------------------------------------------------------
# sign mask                                                 reference pqmf.py:13-22
# --------------------------------------------------------------------------


def sign_mask(n_band: int, n_frames: int, frame0: int = 0, dtype=np.float64) -> np.ndarray:
    """sigma(k, n) = -1 where the band index is odd AND the (global) frame index
    is even (pqmf.py:19-20: mask[..., 1::2, ::2] = -1), else +1."""
    sig = np.ones((n_band, n_frames), dtype)
    n = np.arange(n_frames) + frame0
    sig[1::2, (n % 2) == 0] = -1
    return sig


# --------------------------------------------------------------------------
# analysis                                   reference pqmf.py:115-130 / 160-177 / 339-343
# --------------------------------------------------------------------------


def _windows(xp: np.ndarray, n_frames: int, hop: int, length: int) -> np.ndarray:
    """[B, n_frames, length] strided view of xp[B, *]: window n starts at n*hop."""
    b, _ = xp.shape
    s0, s1 = xp.strides
    return np.lib.stride_tricks.as_strided(xp, (b, n_frames, length), (s0, s1 * hop, s1), writeable=False)


def analysis(x: np.ndarray, hk: np.ndarray, n_frames: int | None = None, *, history: np.ndarray | None = None,
             frame0: int = 0, dtype=np.float64) -> np.ndarray:
    """y[b,k,n] = sigma(k, n+frame0) * sum_j hk[k,j] * X[b, n*M + j - off].

    Offline (history=None): off = L/2 and X is x zero-extended on both sides --
    the closed form of polyphase_forward / classic_forward / CachedPQMF.forward
    followed by reverse_half.  n_frames defaults to T // M
    (polyphase, classic); CachedPQMF.forward yields ceil(T / M) (conv k=L+1,
    stride M, pad (L/2, L/2), no crop).

    Streaming (history [B, L] = the L samples that preceded x, zeros at stream
    start): off = L, i.e. frame n only sees samples strictly before (n+1)*M - M
    (cached padding puts both pads on the left).
    x: [B, T] -> [B, M, n_frames].
    """
    x = np.asarray(x, dtype)
    hk = np.asarray(hk, dtype)
    m, length = hk.shape
    b, t = x.shape
    if n_frames is None:
        n_frames = t // m
    if history is None:
        left = np.zeros((b, length // 2), dtype)
    else:
        left = np.asarray(history, dtype)
        assert left.shape == (b, length)
    need = (n_frames - 1) * m + length if n_frames > 0 else 0
    right = np.zeros((b, max(0, need - left.shape[1] - t)), dtype)
    xp = np.ascontiguousarray(np.concatenate([left, x, right], axis=1))
    if n_frames == 0:
        return np.zeros((b, m, 0), dtype)
    y = np.empty((b, m, n_frames), dtype)
    hk_t = np.ascontiguousarray(hk.T)
    step = 16384  # frames per chunk: bounds the materialised window matrix
    for bi in range(b):
        win = _windows(xp[bi : bi + 1], n_frames, m, length)[0]  # [F, L] strided view
        for f0 in range(0, n_frames, step):
            y[bi, :, f0 : f0 + step] = (win[f0 : f0 + step] @ hk_t).T
    return y * sign_mask(m, n_frames, frame0, dtype)[None]


# --------------------------------------------------------------------------
# synthesis                                  reference pqmf.py:133-157 / 180-199 / 345-354
# --------------------------------------------------------------------------


def synthesis(s: np.ndarray, hk: np.ndarray, *, delay_frames: int = 0, history: np.ndarray | None = None,
              frame0: int = 0, dtype=np.float64) -> np.ndarray:
    """out[b,tau] = M * sum_k sum_n sigma(k,n) S[b,k,n] * hk[k, tau - n*M + off2].

    Offline PQMF.inverse (delay_frames=0): off2 = L/2 (pqmf.py:148-156: conv with the
    flipped polyphase bank, pad K/2+1, drop 2 frames).  Offline CachedPQMF.inverse
    (delay_frames=1): the same sum one frame late, off2 = L/2 - M (pqmf.py:345-354 with
    the k=K+1, pad (K/2, K/2) conv baked into pqmfpvoc.ts).

    Streaming (history [B, M, K] = the K sub-band frames that preceded s, zeros at
    stream start; frame0 = global index of s[..., 0], only its parity matters):
    off2 = -M, output frame f uses frames f-K .. f-1.
    s: [B, M, F] -> [B, M*F].
    """
    s = np.asarray(s, dtype)
    hk = np.asarray(hk, dtype)
    m, length = hk.shape
    k_taps = length // m
    b, m2, f = s.shape
    assert m2 == m and length % m == 0
    if history is None:
        off2 = length // 2 - delay_frames * m
        sp = s * sign_mask(m, f, frame0, dtype)[None]
        n_hist = 0
    else:
        hist = np.asarray(history, dtype)
        assert hist.shape == (b, m, k_taps)
        off2 = -m
        n_hist = k_taps
        sp = np.concatenate([hist, s], axis=2) * sign_mask(m, f + k_taps, frame0 - k_taps, dtype)[None]
    # out[tau] = M * sum_n sum_k sp[k, n] hk[k, tau - (n - n_hist) M + off2]; put frame n's
    # L-sample contribution at position (n - n_hist) M - off2 of a long accumulator.
    ntot = sp.shape[2]
    base = n_hist * m + max(0, off2) + length
    acc = np.zeros((b, base + ntot * m + length), dtype)
    step = 16384
    for bi in range(b):
        for n0 in range(0, ntot, step):
            blk = sp[bi, :, n0 : n0 + step]                    # [M, nb]
            contrib = (blk.T @ hk) * m                          # [nb, L]
            nb = blk.shape[1]
            for q in range(k_taps):  # L = K*M: chunk q of every frame, a contiguous strided add
                chunk = contrib[:, q * m : (q + 1) * m].reshape(nb * m)
                start = base - n_hist * m - off2 + q * m + n0 * m
                acc[bi, start : start + nb * m] += chunk
    return acc[:, base : base + f * m]


# --------------------------------------------------------------------------
# block streaming driver (explicit state), SURVEY.md A.4
# --------------------------------------------------------------------------


class StreamState:
    """Carried FIR history for B streams: last L input samples and last K sub-band frames."""

    def __init__(self, batch: int, n_band: int, length: int, dtype=np.float64):
        self.x_hist = np.zeros((batch, length), dtype)
        self.s_hist = np.zeros((batch, n_band, length // n_band), dtype)
        self.frames_in = 0   # analysis frames produced so far
        self.frames_out = 0  # synthesis frames consumed so far


def stream_analysis(x_block: np.ndarray, hk: np.ndarray, st: StreamState, dtype=np.float64) -> np.ndarray:
    m, length = hk.shape
    t = x_block.shape[1]
    assert t % m == 0
    y = analysis(x_block, hk, t // m, history=st.x_hist, frame0=st.frames_in, dtype=dtype)
    cat = np.concatenate([st.x_hist, np.asarray(x_block, st.x_hist.dtype)], axis=1)
    st.x_hist = cat[:, -length:]
    st.frames_in += t // m
    return y


def stream_synthesis(s_block: np.ndarray, hk: np.ndarray, st: StreamState, dtype=np.float64) -> np.ndarray:
    m, length = hk.shape
    k_taps = length // m
    out = synthesis(s_block, hk, history=st.s_hist, frame0=st.frames_out, dtype=dtype)
    cat = np.concatenate([st.s_hist, np.asarray(s_block, st.s_hist.dtype)], axis=2)
    st.s_hist = cat[:, :, -k_taps:]
    st.frames_out += s_block.shape[2]
    return out


# --------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------


def snr_db(ref: np.ndarray, est: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    est = np.asarray(est, np.float64)
    return float(10 * np.log10(np.sum(ref ** 2) / np.sum((ref - est) ** 2)))


def audio_like(shape, seed: int) -> np.ndarray:
    """Audio-scale synthetic input of SURVEY.md 8d: 0.5*N(0,1) clamped to [-1, 1], fp32."""
    rng = np.random.default_rng(seed)
    return np.clip(0.5 * rng.standard_normal(shape), -1.0, 1.0).astype(np.float32)
