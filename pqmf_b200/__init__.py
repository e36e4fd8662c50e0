"""Importable name of the package whose sources live in ``pseudo-quadrature-mirror-filter_b200/``.

The task fixes the on-disk package directory name (it contains hyphens, so Python cannot import
it by that name); this stub makes the same files importable as ``pqmf_b200`` -- a clean dotted
name is also what TorchScript needs for the qualified names of scripted modules.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pseudo-quadrature-mirror-filter_b200")
if not _os.path.isdir(_impl):  # pragma: no cover
    raise ImportError(f"pqmf_b200: implementation directory not found: {_impl}")
__path__.append(_impl)

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: F401,E402
