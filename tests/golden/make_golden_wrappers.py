"""Golden outputs of the reference's three WRAPPERS, produced by running them -- unmodified, on top of the reference's own pqmf.py --
in the dev container:

    python tests/golden/make_golden_wrappers.py      # needs /root/reference

tests/test_gpu_reference_wrappers.py runs the same wrapper files on CUDA on top of the DROP-IN (dropin/pqmf.py, dropin/PQMF/) and
compares with these.  Written: tests/golden/wrappers.npz."""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (cached_conv stand-in, REF)

sys.path.insert(0, ROOT)
from oracle import fetch_ref_wrappers  # noqa: E402

WRAPPERS = tempfile.mkdtemp(prefix="ref_wrappers_")


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    assert fetch_ref_wrappers.fetch() and fetch_ref_wrappers.unpack(WRAPPERS), "needs /root/reference"
    G.install_cached_conv_stand_in()
    sys.path.insert(0, G.REF)                      # `from pqmf import CachedPQMF`  -> the reference's pqmf.py
    pkg = types.ModuleType("PQMF")                 # `from PQMF.pqmf import ...`     -> the same file, as the wrappers' package name
    pkg.__path__ = [G.REF, os.path.join(WRAPPERS, "PQMF")]
    sys.modules["PQMF"] = pkg
    import pqmf as ref_pqmf  # noqa: F401

    pcm = np.load(os.path.join(HERE, "flute_C1.npz"))["pcm"]
    x = torch.from_numpy(pcm[100000 : 100000 + 16384].astype(np.float32) / 32768.0)[None]  # [1, 16384]
    out = {"x": x.numpy()}
    with torch.no_grad():
        w = load_module("PQMFWrapper", os.path.join(WRAPPERS, "PQMFWrapper.py")).PQMFWrapper(100, 16, 8192).eval()
        recon, sub = w.process(x)
        out["process_recon"], out["process_sub"] = recon.numpy(), sub.numpy()

        pv = load_module("pvoc_wrapper", os.path.join(WRAPPERS, "PQMF", "PitchShifterPvoc", "1-PitchShifterWrapper.py"))
        wp = pv.PQMFPitchShiftWrapper(100, 16, 8192, 44100).eval()
        out["pvoc_forward"] = wp.forward(x[:, :8192]).numpy()
        out["pvoc_pitch_block0"] = wp.pitchshift(x[:, :8192]).numpy()
        out["pvoc_pitch_block1"] = wp.pitchshift(x[:, 8192:]).numpy()     # second block: the prev_tail cross-fade is live

        ps = load_module("ps_wrapper", os.path.join(WRAPPERS, "PQMF", "PitchShifterTorchaudio", "PQMFPsWrapper.py"))
        wt = ps.PQMFPitchShiftWrapper(100, 16, 512, 44100).eval()
        sub_t = wt.forward(x[:, :8192])
        out["ps_forward"] = sub_t.numpy()
        out["ps_inverse"] = wt.inverse(sub_t).numpy()
        out["ps_pitch"] = wt.pitchshifter(x[:, :8192]).numpy()
    np.savez_compressed(os.path.join(HERE, "wrappers.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
