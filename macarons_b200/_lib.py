"""ctypes binding of libmacarons_b200.so (the C ABI declared in include/macarons_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised.  The library is built in-tree by `macarons_b200.build.build()` (nvcc, sm_100a).
"""
import ctypes
import os

from . import build as _build

_LIB = None

_c_float_p = ctypes.c_void_p  # device or host address passed as an integer

# name -> (restype, argtypes); must list every symbol of include/macarons_b200.h
SYMBOLS = {
    "mac_version": (ctypes.c_int, []),
    "mac_last_error": (ctypes.c_char_p, []),
    "mac_built_for_sm": (ctypes.c_int, []),
    "mac_launch_count": (ctypes.c_ulonglong, []),
    "mac_host_release": (None, []),
    "mac_covgain_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "mac_covgain_f32": (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mac_visibility_f32": (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p]),
    "mac_covgain_host": (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int]),
}


SYMBOLS["mac_covgain_backward_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_void_p])

MAX_PEERS = 16


class PeerBoard(ctypes.Structure):
    """mac_peer_board_t of include/macarons_b200.h"""
    _fields_ = [("world", ctypes.c_int), ("rank", ctypes.c_int), ("epoch", ctypes.c_uint),
                ("scores", ctypes.c_void_p * MAX_PEERS), ("flags", ctypes.c_void_p * MAX_PEERS),
                ("partials", ctypes.c_void_p * MAX_PEERS)]


SYMBOLS["mac_covgain_push_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(PeerBoard), ctypes.c_void_p])
SYMBOLS["mac_covgain_push_argmax_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p,
                                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                         ctypes.POINTER(PeerBoard), ctypes.c_void_p, ctypes.c_void_p,
                                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p])
SYMBOLS["mac_covgain_accumulate_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_int,
                                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                        ctypes.c_void_p])
SYMBOLS["mac_covgain_partial_region_bytes"] = (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int])
SYMBOLS["mac_covgain_push_partial_argmax_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p,
                                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                                 ctypes.POINTER(PeerBoard), ctypes.c_void_p, ctypes.c_void_p,
                                                                 ctypes.c_void_p])
SYMBOLS["mac_gather_wait_argmax"] = (ctypes.c_int, [_c_float_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p])


SYMBOLS["mac_linear_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_int, _c_float_p,
                                            _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p, ctypes.c_int,
                                            _c_float_p, _c_float_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p])


SYMBOLS["mac_linear_lnio_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_int, _c_float_p,
                                                 _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                                 _c_float_p, ctypes.c_float, ctypes.c_void_p])


SYMBOLS["mac_knn16_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_void_p])
SYMBOLS["mac_sconevis_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int])
SYMBOLS["mac_sconevis_forward_f32"] = (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])
SYMBOLS["mac_sconevis_forward_ragged_f32"] = (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                             ctypes.c_void_p])
SYMBOLS["mac_sconeocc_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
SYMBOLS["mac_sconeocc_forward_f32"] = (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_void_p,
                                                      ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                      ctypes.c_void_p])


SYMBOLS["mac_sconeocc_cells_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong])
SYMBOLS["mac_sconeocc_forward_cells_f32"] = (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_float_p, ctypes.c_int,
                                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_float_p,
                                                            _c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                            _c_float_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                                            ctypes.c_size_t, ctypes.c_void_p])


SYMBOLS["mac_view_state_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p])
SYMBOLS["mac_view_harmonics_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p])
SYMBOLS["mac_gather_bins_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_void_p])
SYMBOLS["mac_viewstate_harm_f32"] = (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_void_p])
SYMBOLS["mac_sample_proxy_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_int])
SYMBOLS["mac_sample_proxy_points_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_float, _c_float_p, _c_float_p,
                                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                         ctypes.c_size_t, ctypes.c_void_p])
SYMBOLS["mac_fov_sample_proxy_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int])
SYMBOLS["mac_fov_sample_proxy_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p,
                                                      ctypes.c_float, ctypes.c_float, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p,
                                                      _c_float_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])


SYMBOLS["mac_points_in_fov_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int,
                                                   ctypes.c_void_p, ctypes.c_void_p])


SYMBOLS["mac_cell_min_dist_f64"] = (ctypes.c_int, [_c_float_p, ctypes.c_void_p, _c_float_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_int, ctypes.c_void_p])
SYMBOLS["mac_unproject_depth_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_void_p])
SYMBOLS["mac_signed_distance_f32"] = (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p])
SYMBOLS["mac_manydepth_workspace_bytes"] = (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_int])
SYMBOLS["mac_manydepth_forward_f32"] = (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p,
                                                       _c_float_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                       ctypes.c_void_p])


class MacaronsB200Error(RuntimeError):
    pass


def lib_path():
    return _build.lib_path()


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MacaronsB200Error(
            "%s is missing: run `python -m macarons_b200.build` (needs nvcc). "
            "macarons_b200 has no CPU or PyTorch fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc):
    """Turn a negative return code into an exception carrying mac_last_error()."""
    if rc != 0:
        msg = load().mac_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("macarons_b200: " + msg)
        raise MacaronsB200Error("macarons_b200 (code %d): %s" % (rc, msg))
