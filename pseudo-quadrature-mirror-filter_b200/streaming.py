"""CUDA-graph replay of the real-time block step (forward_stream + inverse_stream of one block of every stream).

The reference's real-time use is one stream and blocks of 512 ... 16384 samples (PQMFWrapper.py:40-41: m_buffer_size / max_buffer_size,
driven block by block from a Pure Data external, README.md:16): there the two kernel launches take a few microseconds and the host
side of the call (op dispatch, output allocation, launch) dominates.  A captured graph removes it.  The streaming ops read their FIR
history from one buffer and write the new history to another (ping-pong), so a single graph would replay stale pointers; this helper
captures TWO graphs -- even and odd steps -- over fixed input / state / output buffers and alternates between them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class StreamGraph:
    """``g = StreamGraph(cached_pqmf, streams, block)``; then per block ``y, out = g.step(x_block)``.

    ``x_block`` ``[streams, 1, block]`` is copied into the graph's static input (or write into ``g.x`` yourself and call ``g.step()``);
    the returned sub-bands ``[streams, n_band, block / n_band]`` and signal ``[streams, 1, block]`` are the graph's static outputs
    of this step's parity: valid until the step after next.  ``block`` must hold an even number of frames (the sign mask follows the
    global frame parity, which a captured graph cannot advance)."""

    def __init__(self, mod, streams: int, block: int, device: Optional[torch.device] = None):
        if block % (2 * mod.n_band) != 0:
            raise ValueError(f"StreamGraph needs an even number of frames per block: block={block} is not a multiple of {2 * mod.n_band}")
        dev = torch.device(device) if device is not None else mod.hk.device
        if dev.type != "cuda":
            raise RuntimeError("StreamGraph needs the module on a CUDA device")
        self.mod = mod
        self.x = torch.zeros(streams, 1, block, device=dev)
        self._k = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            mod.reset_stream()
            for _ in range(2):  # allocates the ping-pong state, opts the kernels into their shared memory: none of that may happen in a capture
                mod.process_stream(self.x)
            side.synchronize()
            assert mod._x_slot == 0 and mod._s_slot == 0
            self._graphs, self._outs = [], []
            for _ in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    out, y = mod.process_stream(self.x)
                self._graphs.append(g)
                self._outs.append((y, out))
        torch.cuda.current_stream(dev).wait_stream(side)
        self.reset()

    def reset(self) -> None:
        """Forget the carried history (zeroes the state buffers IN PLACE: the captured pointers stay valid)."""
        self.mod._x_state.zero_()
        self.mod._s_state.zero_()
        self._k = 0

    def step(self, x: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self._graphs[self._k & 1].replay()
        out = self._outs[self._k & 1]
        self._k += 1
        return out
