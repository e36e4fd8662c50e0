"""What do trimmed correction K-steps of the Hankel-4 kernels cost in accuracy?  (CPU, numpy; no GPU needed.)

The kernels compute  sum_j x_j c_j  with x = h1 + h2 (fp16 terms), 2^10 c = c1 + c2, as  h1 c1 + h1 c2 + h2 c1  (h2 c2 ~ 2^-22 dropped).
A trimmed K-step keeps only h1 c1.  For every trim this prints, for analysis and synthesis at n_band 16 / attenuation 100:
  * the worst-case bound  2 * 2^-11 * max|input| * sum |bank entries of the trimmed steps|   (what hankel4_pick_trim uses),
  * the error actually made on audio-like noise (0.5 N(0,1) clamped; synthesis fed analysis outputs) -- RMS and max over ~2M outputs,
  * the error on the adversarial signals of tests/test_gpu_parity.py::test_hankel4_worst_case_signals.
Usage: python tools/trim_budget.py [n_band]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pqmf_oracle as O  # noqa: E402  (analysis tool, not product code)


def f16(v):
    return np.asarray(v, np.float32).astype(np.float16).astype(np.float64)


def split(v):
    h1 = f16(v)
    return h1, f16(np.asarray(v, np.float64) - h1)


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    hk = np.load(os.path.join(ROOT, "tests", "golden", f"bank_M{m}.npz"))["hk"].astype(np.float64)
    L = hk.shape[1]
    nz = np.nonzero(np.abs(hk).sum(0))[0]
    al = max(m, 32)
    jlo = (nz[0] // al) * al
    kt = ((nz[-1] + al) // al) * al - jlo
    ks = (kt + 64 - m + 15) // 16
    fr = 64 // m
    print(f"n_band {m}: L {L}, taps [{jlo}, {jlo + kt}), K-steps {ks}")
    rng = np.random.default_rng(1)
    b, t = 8, 16384 * 2
    x = np.clip(0.5 * rng.standard_normal((b, t)), -1, 1).astype(np.float32).astype(np.float64)
    y = O.analysis(x, hk)                      # float64 truth
    s = y.astype(np.float32).astype(np.float64)
    c1a, c2a = split(hk * 1024.0)
    c1s, c2s = split(hk * 1024.0 * m)
    x1, x2 = split(x)
    # which taps does trimming `tr` K-steps per side drop the corrections for?  K index kap = 16 s + e, tap j = kap - M delta (analysis)
    def dropped_taps_analysis(tr, delta):
        kaps = np.r_[np.arange(0, 16 * tr), np.arange(16 * (ks - tr), 16 * ks)]
        j = kaps - m * delta
        return jlo + j[(j >= 0) & (j < kt)]

    print("analysis: trim | worst-case bound | noise rms / max | adversarial max")
    # adversarial: x = sign(hk[k]) pattern (period L), amplitude 1 - 2^-12
    pat_rows = []
    for r in range(m):
        pat = np.sign(hk[r]); pat[pat == 0] = 1.0
        pat_rows.append(np.tile(pat, t // L + 1)[:t] * (1.0 - 2.0 ** -12))
    xa = np.array(pat_rows[:b]).astype(np.float32).astype(np.float64)
    xa1, xa2 = split(xa)
    for tr in range(0, 9):
        if ks - 2 * tr < 1:
            break
        worst, err_all, adv_all = 0.0, [], []
        for delta in range(fr):
            taps = dropped_taps_analysis(tr, delta)
            if taps.size == 0:
                err_all.append(np.zeros(1)); adv_all.append(np.zeros(1)); continue
            worst = max(worst, 2 * 2.0 ** -11 * np.abs(hk[:, taps]).sum(1).max())
            # dropped = sum_j (h1 c2 + h2 c1) / 1024 over the dropped taps, frames n = fr i + delta
            frames = np.arange(L // m + 4, t // m - L // m - 4)
            frames = frames[frames % fr == delta][:4000]
            for xx1, xx2, sink in ((x1, x2, err_all), (xa1, xa2, adv_all)):
                idx = frames[:, None] * m + taps[None, :] - L // 2
                d = (np.einsum("bnj,kj->bkn", xx1[:, idx], c2a[:, taps]) + np.einsum("bnj,kj->bkn", xx2[:, idx], c1a[:, taps])) / 1024.0
                sink.append(d.ravel())
        e, a = np.concatenate(err_all), np.concatenate(adv_all)
        print(f"  {tr} | {worst:.2e} | {np.sqrt((e ** 2).mean()):.2e} / {np.abs(e).max():.2e} | {np.abs(a).max():.2e}")

    # synthesis: out[16 f + p] = sum_lag sum_k sigma s[k, f + o - lag] * M hk[k, M lag + p]; K index kap = M e + kb, lag = delta + ehi - e
    elo, ehi = jlo // m, (jlo + kt) // m - 1
    print("synthesis: trim | worst-case bound (|s| <= 1, all bands adversarial) | audio sub-bands rms / max | full-scale random-sign max")
    f = s.shape[-1]
    sa = ((rng.integers(0, 2, (b, m, f // 32 + 1)) * 2 - 1).repeat(32, axis=2)[..., :f] * (1.0 - 2.0 ** -12)).astype(np.float32).astype(np.float64)
    mask = np.ones((m, f)); mask[1::2, ::2] = -1
    for tr in range(0, 9):
        if ks - 2 * tr < 1:
            break
        worst, err_all, adv_all = 0.0, [], []
        for delta in range(fr):
            kaps = np.r_[np.arange(0, 16 * tr), np.arange(16 * (ks - tr), 16 * ks)]
            e_idx, kb = kaps // m, kaps % m
            lag = delta + ehi - e_idx
            ok = (lag >= elo) & (lag <= ehi)
            e_idx, kb, lag = e_idx[ok], kb[ok], lag[ok]
            if lag.size == 0:
                err_all.append(np.zeros(1)); adv_all.append(np.zeros(1)); continue
            for p in range(m):
                worst = max(worst, 2 * 2.0 ** -11 * (m * np.abs(hk[kb, m * lag + p])).sum())
            frames = np.arange(L // m + 4, f - L // m - 4)
            frames = frames[frames % fr == delta][:1500]
            o = L // (2 * m)
            for ss, sink in ((s, err_all), (sa, adv_all)):
                sm = ss * mask[None]
                s1, s2 = split(sm)
                n_idx = frames[:, None] + o - lag[None, :]          # [frames, terms]
                g1 = s1[:, kb[None, :], n_idx]                      # [b, frames, terms]
                g2 = s2[:, kb[None, :], n_idx]
                for p in range(0, m, max(1, m // 4)):
                    w1, w2 = c1s[kb, m * lag + p], c2s[kb, m * lag + p]
                    sink.append(((g1 * w2).sum(-1) + (g2 * w1).sum(-1)).ravel() / 1024.0)
        e, a = np.concatenate(err_all), np.concatenate(adv_all)
        print(f"  {tr} | {worst:.2e} | {np.sqrt((e ** 2).mean()):.2e} / {np.abs(e).max():.2e} | {np.abs(a).max():.2e}")


if __name__ == "__main__":
    main()
