"""Streaming (cached) mode on the GPU: block-by-block output equals the offline result with the fixed
latency of SURVEY.md A.4, for the config-3 shape (many streams x block 2048) and for awkward block sizes."""
import numpy as np
import pytest
import torch

from oracle import pqmf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def pq():
    import pqmf_b200

    return pqmf_b200


def _run_stream(mod, x, block):
    ys, outs = [], []
    for i in range(0, x.shape[-1], block):
        yb = mod.forward_stream(x[..., i : i + block].contiguous())
        ys.append(yb)
        outs.append(mod.inverse_stream(yb))
    return torch.cat(ys, dim=-1), torch.cat(outs, dim=-1)


@pytest.mark.parametrize("m,block,streams,exact", ((16, 2048, 7, False), (16, 2048, 7, True), (16, 512, 3, False), (16, 48, 2, False),
                                                 (8, 256, 3, False), (64, 4096, 2, False), (16, 4096 + 32, 2, False)))
def test_streamed_equals_offline(golden, pq, m, block, streams, exact):
    hk = golden(f"bank_M{m}.npz")["hk"]
    length = hk.shape[1]
    k_taps = length // m
    n_blocks = 5
    t = n_blocks * block
    x = O.audio_like((streams, 1, t), 99 + block)
    mod = pq.CachedPQMF(100, m, exact=exact).cuda()
    xd = torch.from_numpy(x).cuda()
    y_s, out_s = _run_stream(mod, xd, block)
    # invariant 1: stream_analysis(x) == offline_cached_forward(cat[zeros(L/2), x])[..., :T/M]
    xz = torch.cat([torch.zeros(streams, 1, length // 2, device="cuda"), xd], dim=-1)
    y_off = mod.forward(xz)[..., : t // m]
    assert (y_s - y_off).abs().max().item() <= 2e-6
    # invariant 2: stream_synthesis(s) == offline_cached_inverse(cat[zeros(K/2 frames), s])[..., :T]
    sz = torch.cat([torch.zeros(streams, m, k_taps // 2, device="cuda"), y_s], dim=-1)
    out_off = mod.inverse(sz)[..., :t]
    assert (out_s - out_off).abs().max().item() <= 4e-6
    # and both against the float64 oracle run as one long causal stream
    st = O.StreamState(streams, m, length)
    y64 = O.stream_analysis(x[:, 0], hk, st)
    o64 = O.stream_synthesis(y_s.cpu().numpy(), hk, st)
    assert np.abs(y_s.cpu().numpy() - y64).max() <= TOL / 2
    assert np.abs(out_s.cpu().numpy()[:, 0] - o64).max() <= TOL / 2
    # reconstruction after the fixed latency
    lat = mod.cumulative_delay
    assert lat == length // 2 + (k_taps // 2) * m + m
    if t > 4 * lat:
        assert O.snr_db(x[..., : t - lat], out_s.cpu().numpy()[..., lat:]) > 30.0


def test_streaming_flag_routes_forward_and_reset(golden, pq):
    mod = pq.CachedPQMF(100, 16, streaming=True).cuda()
    x = torch.from_numpy(O.audio_like((4, 1, 4096), 3)).cuda()
    a = torch.cat([mod(x[..., :2048].contiguous()), mod(x[..., 2048:].contiguous())], dim=-1)
    mod.reset_stream()
    b = torch.cat([mod.forward_stream(x[..., :2048].contiguous()), mod.forward_stream(x[..., 2048:].contiguous())], dim=-1)
    assert torch.equal(a, b)
    mod.reset_stream()
    c = mod(x[..., :2048].contiguous())
    assert torch.equal(c, a[..., :128])


def test_config3_many_streams(pq):
    """BASELINE config 3 at reduced stream count for the parity check (4096 streams is the bench shape):
    512 streams x block 2048 x 8 blocks, state carried across blocks."""
    torch.manual_seed(1)
    streams, block, n_blocks = 512, 2048, 8
    mod = pq.CachedPQMF(100, 16).cuda()
    x = (0.5 * torch.randn(streams, 1, block * n_blocks, device="cuda")).clamp_(-1, 1)
    y_s, out_s = _run_stream(mod, x, block)
    xz = torch.cat([torch.zeros(streams, 1, 256, device="cuda"), x], dim=-1)
    assert (y_s - mod.forward(xz)[..., : y_s.shape[-1]]).abs().max().item() <= 5e-6  # streamed (fold kernels) vs offline (Hankel GEMM)
    lat = mod.cumulative_delay
    err = out_s[..., lat:] - x[..., :-lat]
    snr = 10 * torch.log10((x[..., :-lat] ** 2).sum() / (err ** 2).sum())
    assert snr.item() > 30.0


@pytest.mark.parametrize("streams,block,exact_bank", ((300, 2048, False), (1000, 512, False), (290, 4096, False), (97, 7680, False), (299, 1024, False),
                                                      (300, 2048, True)))
def test_many_streams_on_the_hankel_kernels(golden, pq, streams, block, exact_bank):
    """>= 96 tiles of streams: the streaming blocks run on the tensor-core Hankel kernels (several streams per MMA tile, history rows
    in front of every block, state rolled by the loader threads).  Checked against the float64 streaming oracle, against the fold
    kernels (PQMF_FLAG_FOLD) and against the offline result."""
    from pqmf_b200 import _lib

    att = 120 if exact_bank else 100  # attenuation 120: no zero taps -> 8 history rows instead of 7
    mod = pq.CachedPQMF(att, 16).cuda()
    hk = mod.hk.cpu().numpy()
    n_blocks = 3
    t = n_blocks * block
    x = O.audio_like((streams, 1, t), 5 + block)
    xd = torch.from_numpy(x).cuda()
    y_s, out_s = _run_stream(mod, xd, block)
    st = O.StreamState(streams, 16, 512)
    y64 = O.stream_analysis(x[:, 0], hk, st)
    o64 = O.stream_synthesis(y_s.cpu().numpy(), hk, st)
    assert np.abs(y_s.cpu().numpy() - y64).max() <= TOL / 2
    assert np.abs(out_s.cpu().numpy()[:, 0] - o64).max() <= TOL / 2
    # the fold kernels on the same blocks
    fold = pq.CachedPQMF(att, 16).cuda()
    fold._flags |= _lib.PQMF_FLAG_FOLD
    y_f, out_f = _run_stream(fold, xd, block)
    assert (y_f - y_s).abs().max().item() <= 3e-6 and (out_f - out_s).abs().max().item() <= 8e-6
    # state really is carried: the streamed sub-bands equal the offline ones of the zero-prefixed signal
    xz = torch.cat([torch.zeros(streams, 1, 256, device="cuda"), xd], dim=-1)
    assert (y_s - mod.forward(xz)[..., : y_s.shape[-1]]).abs().max().item() <= 3e-6


def test_config3_sixty_four_steps_equal_offline(pq):
    """64 consecutive 2048-sample blocks with carried state equal the offline result on the concatenated
    131 072-sample signal (sub-bands of the zero-prefixed input; reconstruction after the fixed 528-sample latency)."""
    torch.manual_seed(64)
    streams, block, n_blocks = 300, 2048, 64
    mod = pq.CachedPQMF(100, 16).cuda()
    x = (0.5 * torch.randn(streams, 1, block * n_blocks, device="cuda")).clamp_(-1, 1)
    y_s, out_s = _run_stream(mod, x, block)
    xz = torch.cat([torch.zeros(streams, 1, 256, device="cuda"), x], dim=-1)
    assert (y_s - mod.forward(xz)[..., : y_s.shape[-1]]).abs().max().item() <= 2e-6
    sz = torch.cat([torch.zeros(streams, 16, 16, device="cuda"), y_s], dim=-1)
    assert (out_s - mod.inverse(sz)[..., : out_s.shape[-1]]).abs().max().item() <= 4e-6
    lat = mod.cumulative_delay
    assert lat == 528
    err = out_s[..., lat:] - x[..., :-lat]
    assert (10 * torch.log10((x[..., :-lat] ** 2).sum() / (err ** 2).sum())).item() > 55.0


@pytest.mark.parametrize("streams,block", ((1, 512), (1, 4096), (300, 2048), (5, 16384)))
def test_stream_graph_replay_equals_eager_streaming(pq, streams, block):
    """StreamGraph: the block step (forward_stream + inverse_stream) as two alternating CUDA graphs over fixed buffers -- the same
    bits as the eager calls, block after block, and reset() starts a new stream without re-capturing."""
    torch.manual_seed(streams + block)
    n_blocks = 7
    x = (0.5 * torch.randn(streams, 1, block * n_blocks, device="cuda")).clamp_(-1, 1)
    eager = pq.CachedPQMF(100, 16).cuda()
    y_e, out_e = _run_stream(eager, x, block)
    graph = pq.StreamGraph(pq.CachedPQMF(100, 16).cuda(), streams, block)
    for rep in range(2):
        ys, outs = [], []
        for i in range(n_blocks):
            y, out = graph.step(x[..., i * block : (i + 1) * block])
            ys.append(y.clone())
            outs.append(out.clone())
        assert torch.equal(torch.cat(ys, -1), y_e) and torch.equal(torch.cat(outs, -1), out_e)
        graph.reset()
    with pytest.raises(ValueError):
        pq.StreamGraph(pq.CachedPQMF(100, 16).cuda(), 1, 48)  # three frames per block: the frame parity would have to alternate


def test_stream_ops_refuse_to_be_differentiated(pq):
    mod = pq.CachedPQMF(100, 16).cuda()
    x = torch.randn(2, 1, 2048, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError, match="not differentiable"):
        mod.forward_stream(x)
    with torch.no_grad():
        y = mod.forward_stream(x)
    with pytest.raises(RuntimeError, match="not differentiable"):
        mod.inverse_stream(y.requires_grad_(True))
    with pytest.raises(RuntimeError, match="not differentiable"):
        mod.process_stream(x)


@pytest.mark.parametrize("streams,block", ((300, 2048), (4096, 2048), (1000, 512), (97, 7680), (7, 2048), (1, 512)))
def test_fused_block_step_equals_the_two_calls(pq, streams, block):
    """process_stream(x) = (inverse_stream(forward_stream(x)), forward_stream(x)) as ONE op call (pqmf_stream_step_f32): the same bits
    and the same carried state as the two calls, block after block, on every streaming kernel family."""
    torch.manual_seed(streams + block)
    n_blocks = 5
    x = (0.5 * torch.randn(streams, 1, block * n_blocks, device="cuda")).clamp_(-1, 1)
    two = pq.CachedPQMF(100, 16).cuda()
    y_e, out_e = _run_stream(two, x, block)
    one = pq.CachedPQMF(100, 16).cuda()
    scripted = torch.jit.script(pq.CachedPQMF(100, 16).cuda())
    for mod in (one, scripted):
        ys, outs = [], []
        for i in range(n_blocks):
            out, y = mod.process_stream(x[..., i * block : (i + 1) * block].contiguous())
            ys.append(y)
            outs.append(out)
        assert torch.equal(torch.cat(ys, -1), y_e)
        assert torch.equal(torch.cat(outs, -1), out_e)
    assert torch.equal(one._x_state[one._x_slot], two._x_state[two._x_slot]) and torch.equal(one._s_state[one._s_slot], two._s_state[two._s_slot])
    # mixing the fused step with the separate calls keeps one consistent stream
    one.reset_stream()
    o0, y0 = one.process_stream(x[..., :block].contiguous())
    y1 = one.forward_stream(x[..., block : 2 * block].contiguous())
    o1 = one.inverse_stream(y1)
    assert torch.equal(torch.cat([y0, y1], -1), y_e[..., : 2 * block // 16]) and torch.equal(torch.cat([o0, o1], -1), out_e[..., : 2 * block])


@pytest.mark.parametrize("m,streams,block", ((8, 300, 2048), (32, 300, 2048), (8, 1000, 512), (32, 290, 4096), (8, 7, 2048), (32, 5, 1024)))
def test_streaming_other_band_counts(golden, pq, m, streams, block):
    """n_band 8 and 32: many streams run the streaming Hankel kernels (tiles of several streams, as n_band 16), few streams the fp32
    direct form with the history roll fused into the kernel.  Against the float64 streaming oracle, the offline kernels and the
    plain-fp32 module."""
    hk = golden(f"bank_M{m}.npz")["hk"]
    length = hk.shape[1]
    n_blocks = 3
    x = O.audio_like((streams, 1, n_blocks * block), 11 * m + block)
    xd = torch.from_numpy(x).cuda()
    mod = pq.CachedPQMF(100, m).cuda()
    y_s, out_s = _run_stream(mod, xd, block)
    st = O.StreamState(streams, m, length)
    y64 = O.stream_analysis(x[:, 0], hk, st)
    o64 = O.stream_synthesis(y_s.cpu().numpy(), hk, st)
    assert np.abs(y_s.cpu().numpy() - y64).max() <= TOL / 2
    assert np.abs(out_s.cpu().numpy()[:, 0] - o64).max() <= TOL / 2
    xz = torch.cat([torch.zeros(streams, 1, length // 2, device="cuda"), xd], dim=-1)
    assert (y_s - mod.forward(xz)[..., : y_s.shape[-1]]).abs().max().item() <= 3e-6
    plain = pq.CachedPQMF(100, m, fp32=True).cuda()
    y_p, out_p = _run_stream(plain, xd, block)
    assert (y_p - y_s).abs().max().item() <= 3e-6 and (out_p - out_s).abs().max().item() <= 8e-6
    lat = mod.cumulative_delay
    if out_s.shape[-1] > 4 * lat:
        err = out_s[..., lat:] - xd[..., :-lat]
        assert (10 * torch.log10((xd[..., :-lat] ** 2).sum() / (err ** 2).sum())).item() > 30.0
