"""Sources of the ``pqmf_b200`` package (import it as ``pqmf_b200``; see ../pqmf_b200/__init__.py).

Layout: csrc/ (sm_100a kernels, the C ABI of include/pqmf_b200.h and the torch op registration),
_lib.py (loads the two in-tree shared objects, fails loudly when they are missing), design.py (host-side
filter design, numpy/scipy like the reference), pqmf.py (drop-in mirror of the reference's pqmf.py API).
"""
