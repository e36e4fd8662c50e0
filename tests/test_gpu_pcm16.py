"""int16 PCM edge: analysis straight from interleaved 16-bit WAV frames, synthesis straight to them.
The load side must be BIT-identical to the float path on pcm / 32768 (what torchaudio.load gives the reference's scripts,
PQMFWrapper.py:113; the down-mix of 2-TestBlocks.py:26-30 is mean(dim=0)); the store side to clamp(round(v * 32768))."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pq():
    import pqmf_b200

    return pqmf_b200


def _pcm(b, t, c, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-32768, 32768, (b, t, c), generator=g, dtype=torch.int32).to(torch.int16).cuda()


def _on_hankel(rows, samples_per_row):
    """>= 96 tiles of 8192 samples: the dispatcher runs the tensor-core Hankel kernels (float and PCM instantiations of ONE kernel,
    hence bit-identical); below that the PCM entry points use the fp32 direct form, which is what ``fp32=True`` selects for floats."""
    return rows * -(-samples_per_row // 8192) >= 96


def _as_float_rows(pcm):
    """torchaudio.load's view of the same WAV frames: [clips, channels, time] float32 = int16 / 32768."""
    return (pcm.to(torch.float32) / 32768.0).transpose(1, 2).contiguous()


# small shapes run the direct form, >= 96 tiles of 8192 samples the tensor-core Hankel kernels
@pytest.mark.parametrize("m,b,t,c", ((16, 2, 4096, 1), (16, 3, 2048, 2), (8, 2, 4000, 3), (16, 24, 32768, 1), (16, 12, 32768, 2), (8, 24, 32768, 1),
                                     (32, 12, 65536, 2), (64, 24, 32768, 1), (4, 25, 32768, 1), (16, 4, 32768, 6),
                                     # stereo on the Hankel kernels (a CTA loads a tile's frames once and visits both channels): several
                                     # visits per CTA, odd clip counts, partial last tiles
                                     (16, 101, 40960, 2), (8, 37, 40000, 2), (32, 40, 40960, 2), (4, 33, 24576, 2), (16, 7, 1048576, 2)))
def test_analysis_from_pcm_is_bit_identical_to_the_float_path(pq, m, b, t, c):
    mod = pq.PQMF(100, m).cuda()
    plain = pq.PQMF(100, m, fp32=True).cuda()
    pcm = _pcm(b, t, c, 100 * m + c)
    y = mod.forward_pcm16(pcm)
    assert y.shape == (b, c * m, t // m)
    same = mod if _on_hankel(b * c, t) else plain
    assert torch.equal(y, same(_as_float_rows(pcm)))
    assert (y - mod(_as_float_rows(pcm))).abs().max().item() <= 3e-6
    if c > 1:
        yd = mod.forward_pcm16(pcm, downmix=True)
        assert yd.shape == (b, m, t // m)
        mono = _as_float_rows(pcm).mean(dim=1, keepdim=True)  # 2-TestBlocks.py:26-30
        same = mod if _on_hankel(b, t) else plain
        if c == 2:
            assert torch.equal(yd, same(mono))
        else:  # torch's reduction order over more than two channels is its own business: same values to fp32 rounding
            assert (yd - same(mono)).abs().max().item() <= 1e-6


@pytest.mark.parametrize("m,b,f,c", ((16, 2, 256, 1), (16, 3, 128, 2), (16, 24, 2048, 1), (16, 12, 2048, 2), (8, 24, 4096, 1), (32, 24, 1024, 1),
                                     (64, 24, 512, 1), (16, 2, 2048, 5),
                                     # stereo on the Hankel kernels: several (left, right) visits per CTA, odd clip counts, partial last tiles
                                     (16, 101, 2560, 2), (8, 37, 5000, 2), (32, 40, 1280, 2), (4, 33, 6144, 2), (16, 7, 65536, 2)))
def test_synthesis_to_pcm_quantises_like_torch(pq, m, b, f, c):
    torch.manual_seed(m + f + c)
    mod = pq.CachedPQMF(100, m).cuda()
    x = (0.5 * torch.randn(b, c, m * f, device="cuda")).clamp_(-1, 1)
    y = mod(x) * 1.7   # loud enough to clip some samples: the store saturates
    # the float kernel that computes the same values: the Hankel kernels for big batches (not n_band 64: its second tap range
    # accumulates into the output, which int16 cannot do), else the fp32 direct form
    same = mod if (_on_hankel(b * c, m * f) and m != 64) else pq.CachedPQMF(100, m, fp32=True).cuda()
    ref = torch.clamp(torch.round(same.inverse(y) * 32768.0), -32768, 32767).to(torch.int16).transpose(1, 2).contiguous()  # [b, time, c]
    out = mod.inverse_pcm16(y)
    assert out.dtype == torch.int16 and out.shape == (b, m * f, c)
    assert (out == 32767).any() or (out == -32768).any()
    assert torch.equal(out, ref)


def test_flutemulti_stereo_wav_frames(golden, pq):
    """The reference's stereo fixture (audio/flutemulti.wav, int16): interleave it the way the WAV file stores it."""
    g = golden("flutemulti_excerpt.npz")
    pcm_ct = torch.from_numpy(g["pcm"].astype(np.int16))          # [2, T] as torchaudio.load lays it out
    t = (pcm_ct.shape[1] // 16) * 16
    frames = pcm_ct[:, :t].t().contiguous()[None].cuda()          # [1, T, 2] interleaved
    mod = pq.PQMF(100, 16).cuda()
    y = mod.forward_pcm16(frames)
    xf = (pcm_ct[:, :t].to(torch.float32) / 32768.0)[None].cuda()
    assert torch.equal(y, pq.PQMF(100, 16, fp32=True).cuda()(xf))    # one short clip: the PCM entry runs the fp32 direct form
    assert (y - mod(xf)).abs().max().item() <= 3e-6                  # ... and the default float path (fold kernels) agrees
    back = mod.inverse_pcm16(y)
    # near-perfect reconstruction survives the 16-bit quantisation: interior within 2 LSB of the input (65 dB SNR on this material)
    err = (back.to(torch.int32) - frames.to(torch.int32))[:, 1024:-1024].abs()
    assert err.float().mean().item() < 8.0


def test_cabi_pcm_argument_checks(pq):
    from pqmf_b200 import _lib

    c = _lib.cabi
    assert c.pqmf_analysis_pcm16(None, None, None, None, 1, 64, 0, 0, 4, 16, 512, 0, None) == -1      # zero channels
    assert c.pqmf_analysis_pcm16(None, None, None, None, 1, 64, 2, 0, 4, 16, 512, 0, None) == -1      # null pointers
    assert c.pqmf_analysis_pcm16(None, None, None, None, 0, 64, 2, 0, 4, 16, 512, 0, None) == 0       # empty batch
    assert c.pqmf_synthesis_pcm16(None, None, None, None, 1, 2, 4, 16, 512, 2, 0, None) == -1         # bad delay
    with pytest.raises(RuntimeError):
        pq.PQMF(100, 16).cuda().forward_pcm16(torch.zeros(1, 64, 1, device="cuda"))                   # float frames: wrong dtype
    with pytest.raises(RuntimeError):
        pq.PQMF(100, 16).cuda().forward_pcm16(torch.zeros(1, 64, 1, dtype=torch.int16))               # CPU tensor: no fallback


def test_host_entry_pcm16_matches_the_device_path(pq):
    """pqmf_roundtrip_host_pcm16: int16 WAV frames in host memory in, int16 WAV frames out (2 B/sample/channel over the link)."""
    from pqmf_b200 import _lib

    mod = pq.PQMF(100, 16).cuda()
    b, t, c = 40, 1 << 16, 2
    pcm = _pcm(b, t, c, 5)
    want = mod.inverse_pcm16(mod.forward_pcm16(pcm)).cpu()
    hp = pcm.cpu().pin_memory()
    ho = torch.empty_like(hp).pin_memory()
    hy = torch.empty(b * c, 16, t // 16).pin_memory()
    hk_h, tab_h = mod.hk.cpu().contiguous(), mod._tables.cpu().contiguous()
    rc = _lib.cabi.pqmf_roundtrip_host_pcm16(hp.data_ptr(), hy.data_ptr(), ho.data_ptr(), hk_h.data_ptr(), tab_h.data_ptr(), b, t, c, 16, 512, 0,
                                             int(mod._flags), 0)
    assert rc == 0, _lib.strerror(rc)
    assert torch.equal(ho, want)
    assert torch.equal(hy.reshape(b, c * 16, -1), mod.forward_pcm16(pcm).cpu())
    _lib.cabi.pqmf_host_release()
