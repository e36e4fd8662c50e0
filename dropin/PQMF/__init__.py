"""Drop-in for the package name the pitch-shifter wrappers import (``from PQMF.pqmf import CachedPQMF``,
PitchShifterPvoc/1-PitchShifterWrapper.py:12, PitchShifterTorchaudio/PQMFPsWrapper.py:12)."""
