"""sphexa_b200 — B200-native SPH-VE hydro step behind SPH-EXA's propagator API.

The product is libsphx.so (CUDA sm_100a kernels + C ABI, include/sphx.h). This package is the thin Python plumbing
used by tests and bench.py: ctypes bindings, torch tensors as device memory, torch.distributed for multi-GPU launch.
"""
from . import _cabi, dist, host  # noqa: F401

try:  # torch is plumbing only; the host helpers work without it
    from . import sim  # noqa: F401
except ImportError:  # pragma: no cover
    sim = None
from ._cabi import SphxError, load  # noqa: F401
