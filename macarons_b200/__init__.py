"""macarons_b200 -- Blackwell (sm_100a) implementation of the SCONE / MACARONS next-best-view
scoring path behind the reference's `macarons.networks` / `macarons.utility` API.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/macarons_b200.h)
  _lib.py    ctypes binding (no fallback: raises if the library is missing)
  ops.py     tensor-level wrappers (checks, output/workspace allocation, current stream)
  networks/  SconeVis, SconeOcc, Macarons ... mirrors of reference macarons/networks
  utility/   spherical_harmonics, CustomGeometry, scone_utils mirrors (hot-path subset)
  parallel.py  candidate-camera partition + score all-gather across ranks
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def library_path():
    return _lib.lib_path()
