"""Packs the reference's three wrapper scripts so that the GPU tests can run them UNCHANGED against the drop-in -- TEST INFRASTRUCTURE.

    python oracle/fetch_ref_wrappers.py        # needs /root/reference (read-only mount); __graft_entry__.build() calls it

north_star: the drop-in must be "usable unchanged by PQMFWrapper.py and both pitch-shifter wrappers".  Those files live in
/root/reference, which does not exist on the GPU box.  They are packed VERBATIM into ONE archive, oracle/_ref/wrappers.tar -- a
directory that is git-ignored (the reference's sources never enter this repository's history or its source tree) but not
gpurun-ignored, so the archive travels to the GPU box exactly like the built .so files do.  Layout inside the archive: the
pitch-shifter wrappers import `PQMF.pqmf` and `PQMF.PitchShifterPvoc.VocoderPitchShifter` (1-PitchShifterWrapper.py:12-14), so their
directories sit under a `PQMF/` folder that the test appends to the drop-in package's __path__ after unpacking the archive into a
temporary directory.  tests/test_gpu_reference_wrappers.py skips when the archive is absent.
"""
import io
import os
import sys
import tarfile

REF = os.environ.get("PQMF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "wrappers.tar")
FILES = {
    "PQMFWrapper.py": "PQMFWrapper.py",
    "PitchShifterPvoc/1-PitchShifterWrapper.py": "PQMF/PitchShifterPvoc/1-PitchShifterWrapper.py",
    "PitchShifterPvoc/VocoderPitchShifter.py": "PQMF/PitchShifterPvoc/VocoderPitchShifter.py",
    "PitchShifterTorchaudio/PQMFPsWrapper.py": "PQMF/PitchShifterTorchaudio/PQMFPsWrapper.py",
}


def fetch() -> bool:
    """(Re)builds the archive from the mounted reference; False when the reference is not there."""
    if not os.path.isdir(REF):
        return False
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    with tarfile.open(ARCHIVE, "w") as tar:
        for src, dst in FILES.items():
            data = open(os.path.join(REF, src), "rb").read()
            info = tarfile.TarInfo(dst)
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    return True


def unpack(dest: str) -> bool:
    """Extracts the archive into `dest` (a scratch directory); False when there is no archive."""
    if not os.path.isfile(ARCHIVE):
        return False
    with tarfile.open(ARCHIVE) as tar:
        tar.extractall(dest)
    return True


if __name__ == "__main__":
    ok = fetch()
    print("packed the reference wrappers into", ARCHIVE if ok else "(nothing: %s not found)" % REF)
    sys.exit(0)
