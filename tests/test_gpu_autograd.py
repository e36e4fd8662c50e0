"""Gradients of the two ops: the reference path is differentiable (conv1d), so the CUDA ops register autograd
kernels -- analysis-backward is the synthesis kernel, synthesis-backward the analysis kernel.  Checked against autograd through
the oracle's torch-CPU port of the reference's op sequence (float64), on every kernel family."""
import numpy as np
import pytest
import torch

from oracle import pqmf_oracle as O
from oracle import pqmf_port_torch as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pq():
    import pqmf_b200

    return pqmf_b200


def _ref_grads(fwd, inv, x, wy, wo, hk):
    """d/dx <fwd(x), wy> and d/ds <inv(s), wo> at s = fwd(x), through the CPU port in float64."""
    hk64 = torch.from_numpy(hk).double()
    x64 = torch.from_numpy(x).double().requires_grad_(True)
    y = fwd(x64, hk64)
    (y * torch.from_numpy(wy).double()).sum().backward()
    s64 = y.detach().clone().requires_grad_(True)
    o = inv(s64, hk64)
    (o * torch.from_numpy(wo).double()).sum().backward()
    return y.detach().numpy(), x64.grad.numpy(), s64.grad.numpy()


# (module ctor kwargs, cached?, n_band, batch, samples): fold kernels, Hankel-4 pair kernels, direct form, ragged cached
CASES = (
    ({}, False, 16, 3, 16 * 300),
    ({}, False, 16, 24, 16 * 2048),
    ({"exact": True}, False, 16, 2, 16 * 200),
    ({}, False, 8, 2, 8 * 500),
    ({}, True, 16, 2, 16 * 300 - 7),
    ({}, True, 16, 26, 16 * 2048),
)


@pytest.mark.parametrize("kwargs,cached,m,b,t", CASES)
def test_gradients_match_the_reference_path(golden, pq, kwargs, cached, m, b, t):
    hk = golden(f"bank_M{m}.npz")["hk"]
    rng = np.random.default_rng(b * 1000 + t)
    x = O.audio_like((b, 1, t), 31 + t)
    frames = -(-t // m) if cached else t // m
    wy = rng.standard_normal((b, m, frames)).astype(np.float32)
    wo = rng.standard_normal((b, 1, m * frames)).astype(np.float32)
    fwd, inv = (P.analysis_cached_offline, P.synthesis_cached_offline) if cached else (P.analysis_polyphase, P.synthesis_polyphase)
    y_ref, gx_ref, gs_ref = _ref_grads(fwd, inv, x, wy, wo, hk)

    mod = (pq.CachedPQMF if cached else pq.PQMF)(100, m, **kwargs).cuda()
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    y = mod(xg)
    assert y.requires_grad
    (y * torch.from_numpy(wy).cuda()).sum().backward()
    scale = max(1.0, float(np.abs(gx_ref).max()))
    assert np.abs(xg.grad.cpu().numpy() - gx_ref).max() <= 2e-5 * scale
    sg = torch.from_numpy(y_ref.astype(np.float32)).cuda().requires_grad_(True)
    out = mod.inverse(sg)
    (out * torch.from_numpy(wo).cuda()).sum().backward()
    scale = max(1.0, float(np.abs(gs_ref).max()))
    assert np.abs(sg.grad.cpu().numpy() - gs_ref).max() <= 2e-5 * scale


def test_no_grad_and_scripted_paths_are_unchanged(pq):
    mod = pq.CachedPQMF(100, 16).cuda()
    x = torch.randn(2, 1, 4096, device="cuda").clamp_(-1, 1)
    with torch.no_grad():
        y0 = mod(x)
    y1 = mod(x.clone().requires_grad_(True))
    assert torch.equal(y0, y1.detach()) and not y0.requires_grad
    scripted = torch.jit.script(mod)
    assert torch.equal(scripted(x), y0)
    xs = x.clone().requires_grad_(True)
    scripted.inverse(scripted(xs)).square().sum().backward()
    xe = x.clone().requires_grad_(True)
    mod.inverse(mod(xe)).square().sum().backward()
    assert torch.allclose(xs.grad, xe.grad)


@pytest.mark.parametrize("scale", (1e-7, 1e-3, 1.0, 3e5))
def test_gradients_of_any_magnitude(pq, scale):
    """gradients have no natural scale (a mean-reduced loss gives 1e-7 per sample; an int-scaled signal 1e5), and
    the tensor-core kernels carry values as fp16 pairs.  The backward passes normalise by an exact power of two on the device, so
    relative accuracy must not depend on the magnitude -- on the Hankel path (>= 96 tiles) and on the small-batch paths alike."""
    from oracle import pqmf_port_torch as P

    for b, t in ((24, 32768), (2, 4096)):
        torch.manual_seed(b)
        mod = pq.PQMF(100, 16).cuda()
        hk64 = mod.hk.double().cpu()
        x = (0.5 * torch.randn(b, 1, t)).clamp_(-1, 1)
        w = torch.randn(b, 16, t // 16) * scale            # upstream gradient of the analysis output
        xg = x.cuda().requires_grad_(True)
        (mod(xg) * w.cuda()).sum().backward()
        xr = x.double().requires_grad_(True)
        (P.analysis_polyphase(xr, hk64) * w.double()).sum().backward()
        assert torch.isfinite(xg.grad).all()
        assert (xg.grad.cpu().double() - xr.grad).abs().max().item() <= 2e-5 * scale
        s = (0.3 * torch.randn(b, 16, t // 16))
        v = torch.randn(b, 1, t) * scale
        sg = s.cuda().requires_grad_(True)
        (mod.inverse(sg) * v.cuda()).sum().backward()
        sr = s.double().requires_grad_(True)
        (P.synthesis_polyphase(sr, hk64) * v.double()).sum().backward()
        assert torch.isfinite(sg.grad).all()
        assert (sg.grad.cpu().double() - sr.grad).abs().max().item() <= 3e-4 * scale   # gain 16 x 16 bands
