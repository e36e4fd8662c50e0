"""Drop-in for the reference's top-level ``pqmf`` module (``from pqmf import CachedPQMF``, PQMFWrapper.py:5).
Put this directory (and the repo root) on sys.path ahead of the reference's own pqmf.py."""
from pqmf_b200._api import *  # noqa: F401,F403
from pqmf_b200._api import __all__  # noqa: F401
