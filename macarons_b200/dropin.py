"""Switch a reference checkout (Anttwo/MACARONS, package `macarons`) to the sm_100a implementation of the NBV scoring
path, without editing its files:

    import macarons_b200.dropin
    macarons_b200.dropin.install()          # before importing macarons.testers / macarons.trainers
    from macarons.testers.shapenet import *  # the reference's own driver code, unchanged

The reference binds names at import time (`from ..networks.SconeVis import SconeVis`, `from ..utility.scone_utils import *`,
reference utility/macarons_utils.py:26-72, testers/shapenet.py:1), so replacing whole modules would drop the many symbols of
those files that are not on the hot path (optimisers, loaders, losses, the Camera / Scene / Memory classes).  `install()`
therefore imports the reference's own modules and replaces, in EVERY loaded `macarons.*` namespace, exactly the objects
that have a counterpart here -- classes SconeOcc / SconeVis / Macarons / ManyDepth (+ their sub-modules) and the hot
functions of scone_utils / macarons_utils / spherical_harmonics / CustomGeometry -- and binds the depth-side helpers
as methods of the reference's `Camera`.  Everything else stays the reference's.
"""
import importlib
import sys

# (reference module, our module, names); a name is replaced only if both modules define it
_TABLE = (
    ("networks.Attention", "macarons_b200.networks.Attention", ("Embedding", "MultiHeadSelfAttention", "FeedForward", "Encoder")),
    ("networks.SconeOcc", "macarons_b200.networks.SconeOcc", ("XEmbedding", "PCTransformer", "SconeOcc")),
    ("networks.SconeVis", "macarons_b200.networks.SconeVis", ("SconeVis",)),
    ("networks.ManyDepth", "macarons_b200.networks.ManyDepth", ("FeatureExtractor", "CostVolumeBuilder", "ExpansionLayer",
                                                                "DisparityLayer", "DepthDecoder", "ManyDepth")),
    ("networks.Macarons", "macarons_b200.networks.Macarons", ("Macarons",)),
    ("utility.spherical_harmonics", "macarons_b200.utility.spherical_harmonics", ("get_spherical_harmonics",
                                                                                   "clear_spherical_harmonics_cache")),
    ("utility.CustomGeometry", "macarons_b200.utility.CustomGeometry", ("get_spherical_coords", "get_cartesian_coords")),
    ("utility.scone_utils", "macarons_b200.utility.scone_utils", ("get_all_harmonics_under_degree", "compute_view_state",
                                                                  "move_view_state_to_view_space", "compute_view_harmonics",
                                                                  "compute_occupancy_probability", "sample_proxy_points")),
    ("utility.macarons_utils", "macarons_b200.utility.macarons_utils", ("compute_occupancy_probability",
                                                                        "compute_scene_occupancy_probability_field",
                                                                        "load_images_for_depth_model",
                                                                        "predict_coverage_gain_for_single_camera",
                                                                        "get_distance_factor", "get_distance_factor_threshold",
                                                                        "get_distance_factor_smooth")),
)
_CAMERA_METHODS = ("project_depth_in_3D", "compute_partial_point_cloud", "get_signed_distance_to_depth_maps",
                   "get_points_in_fov")


def install(package="macarons", depth=True):
    """Patch the loaded reference package in place.  Returns {"replaced": [(namespace, name), ...], "camera_methods": [...]}.
    `depth=False` leaves the reference's ManyDepth classes alone (e.g. to keep training the depth module with autograd)."""
    replacements = {}
    for ref_suffix, ours_name, names in _TABLE:
        if not depth and ref_suffix == "networks.ManyDepth":
            continue
        ref_mod = importlib.import_module(package + "." + ref_suffix)
        ours = importlib.import_module(ours_name)
        for n in names:
            if hasattr(ref_mod, n) and hasattr(ours, n):
                replacements[id(vars(ref_mod)[n])] = (getattr(ours, n), vars(ref_mod)[n])
    report = {"replaced": [], "camera_methods": []}
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == package or mod_name.startswith(package + ".")):
            continue
        for attr, val in list(vars(mod).items()):
            hit = replacements.get(id(val))
            if hit is not None and hit[1] is val:
                setattr(mod, attr, hit[0])
                report["replaced"].append((mod_name, attr))
    ref_mu = sys.modules.get(package + ".utility.macarons_utils")
    camera_cls = getattr(ref_mu, "Camera", None) if ref_mu is not None else None
    if camera_cls is not None:
        ours = importlib.import_module("macarons_b200.utility.macarons_utils")
        for n in _CAMERA_METHODS:
            setattr(camera_cls, n, getattr(ours, n))
            report["camera_methods"].append(n)
        # the batched scoring entry point next to the single-camera one, for callers that want all candidates at once
        ref_mu.predict_coverage_gains_for_cameras = ours.predict_coverage_gains_for_cameras
    return report
