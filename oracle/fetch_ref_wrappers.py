"""Puts the reference's three wrapper scripts where the GPU tests can run them UNCHANGED against the drop-in -- TEST INFRASTRUCTURE.

    python oracle/fetch_ref_wrappers.py        # needs /root/reference (read-only mount); __graft_entry__.build() calls it

north_star: the drop-in must be "usable unchanged by PQMFWrapper.py and both pitch-shifter wrappers".  Those files live in
/root/reference, which does not exist on the GPU box.  They are copied VERBATIM into oracle/_ref/wrappers/ -- a directory that is
git-ignored (the reference's sources never enter this repository's history) but not gpurun-ignored, so it travels to the GPU box
exactly like the built .so files do.  Layout: the pitch-shifter wrappers import `PQMF.pqmf` and
`PQMF.PitchShifterPvoc.VocoderPitchShifter` (1-PitchShifterWrapper.py:12-14), so their directory sits under a `PQMF/` folder that the
test appends to the drop-in package's __path__.  tests/test_gpu_reference_wrappers.py skips when the directory is absent.
"""
import os
import shutil
import sys

REF = os.environ.get("PQMF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "wrappers")
FILES = {
    "PQMFWrapper.py": "PQMFWrapper.py",
    "PitchShifterPvoc/1-PitchShifterWrapper.py": "PQMF/PitchShifterPvoc/1-PitchShifterWrapper.py",
    "PitchShifterPvoc/VocoderPitchShifter.py": "PQMF/PitchShifterPvoc/VocoderPitchShifter.py",
    "PitchShifterTorchaudio/PQMFPsWrapper.py": "PQMF/PitchShifterTorchaudio/PQMFPsWrapper.py",
}


def fetch() -> bool:
    if not os.path.isdir(REF):
        return False
    for src, dst in FILES.items():
        out = os.path.join(DST, dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), out)
    return True


if __name__ == "__main__":
    ok = fetch()
    print("copied the reference wrappers to", DST if ok else "(nothing: %s not found)" % REF)
    sys.exit(0)
