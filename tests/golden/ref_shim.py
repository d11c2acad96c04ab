"""Make the upstream reference importable in the build container (NOT on the GPU box).

The reference (`/root/reference`, Anttwo/MACARONS) imports pytorch3d and matplotlib at module
scope (networks/Attention.py:5, utility/utils.py:11-40); neither is installed here.  The NBV
scoring path only ever calls one pytorch3d symbol, `pytorch3d.ops.knn_gather`
(utility/utils.py:1509), which is a plain gather.  This module registers placeholder modules so
the reference's own files import unmodified, and supplies that single gather.

Used only by `tests/golden/make_golden.py` (fixture generation) and by the optional
`-m "not gpu"` tests that re-check the oracle against the live reference when it is present.
Nothing in the product path, `bench.py` or the `-m gpu` tests imports this file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MACARONS_REFERENCE_ROOT", "/root/reference")


class _Placeholder:
    """Stands in for any pytorch3d / matplotlib symbol that is imported but never executed."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


def _placeholder_module(name):
    mod = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):  # keep inspect.getmodule() & friends working
            raise AttributeError(attr)
        return _Placeholder

    mod.__getattr__ = _getattr
    sys.modules[name] = mod
    return mod


def _knn_gather(x, idx):
    """pytorch3d.ops.knn_gather: x (N,M,U), idx (N,L,K) -> (N,L,K,U)."""
    n, m, u = x.shape
    _, l, k = idx.shape
    return x[:, :, None].expand(n, m, k, u).gather(1, idx[..., None].expand(n, l, k, u))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "macarons", "networks"))


_installed = False


def install():
    """Idempotently register the placeholders and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torchvision  # noqa: F401  (must be imported before the placeholders exist)

    if "pytorch3d" not in sys.modules:
        _placeholder_module("pytorch3d")
        for sub in ("ops", "io", "structures", "datasets", "renderer", "renderer.mesh",
                    "renderer.mesh.shading", "renderer.mesh.rasterizer", "renderer.mesh.renderer",
                    "renderer.cameras", "renderer.lighting", "loss", "transforms"):
            _placeholder_module("pytorch3d." + sub)
        sys.modules["pytorch3d.ops"].knn_gather = _knn_gather
        # the depth path executes pytorch3d cameras / rotation conversions: restated stand-ins (oracle/cameras.py)
        root = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle import cameras as _cams
        sys.modules["pytorch3d.renderer.cameras"].FoVPerspectiveCameras = _cams.FoVPerspectiveCameras
        sys.modules["pytorch3d.renderer"].FoVPerspectiveCameras = _cams.FoVPerspectiveCameras
        for _name in ("axis_angle_to_matrix", "matrix_to_quaternion", "quaternion_apply", "quaternion_to_matrix"):
            setattr(sys.modules["pytorch3d.transforms"], _name, getattr(_cams, _name))
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = _placeholder_module("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
