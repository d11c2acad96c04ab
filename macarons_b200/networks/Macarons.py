"""Macarons wrapper (reference macarons/networks/Macarons.py:91-178): dispatches `forward(mode=...)`
to the depth / occupancy / visibility modules and exposes `compute_visibility_gains`, which runs the
sm_100a coverage-gain kernel in per-point mode."""
from torch import nn

from .. import ops
from ..utility.spherical_harmonics import clear_spherical_harmonics_cache


class Macarons(nn.Module):
    def __init__(self, depth_model, occupancy_model, visibility_model):
        super().__init__()
        self.depth = depth_model
        self.occupancy = occupancy_model
        self.visibility = visibility_model
        if depth_model is not None:
            self.image_height = depth_model.input_height
            self.image_width = depth_model.input_width

    def forward(self, mode, x=None, x_alpha=None, R=None, T=None, zfar=None, device=None, gt_pose=None,
                partial_point_cloud=None, proxy_points=None, view_harmonics=None):
        def missing(*args):
            return any(a is None for a in args)

        if mode == 'depth':
            if missing(x, x_alpha, R, T, zfar, device):
                raise NameError("For 'depth' mode, you should provide the following args:"
                                "x, x_alpha, R, T, zfar, device")
            return self.depth(x=x, x_alpha=x_alpha, R=R, T=T, zfar=zfar, device=device, gt_pose=gt_pose)
        if mode == 'occupancy':
            if missing(partial_point_cloud, proxy_points, view_harmonics):
                raise NameError("For 'occupancy' mode, you should provide the following args:"
                                "partial_point_cloud, proxy_points, view_harmonics")
            return self.occupancy(pc=partial_point_cloud, x=proxy_points, view_harmonics=view_harmonics)
        if mode == 'visibility':
            if missing(proxy_points, view_harmonics):
                raise NameError("For 'visibility' mode, you should provide the following args:"
                                "proxy_points, view_harmonics")
            return self.visibility(proxy_points, view_harmonics=view_harmonics)
        raise NameError("Invalid mode. Please select a mode between 'depth', 'occupancy' and 'visibility'.")

    def compute_visibility_gains(self, pts, harmonics, X_cam):
        """(B,P,pts_dim), (B,P,64), (B,C,3) -> (B,C,P) per-point visibility gains
        [reference Macarons.py:138-178; like the reference, the ReLU variant is refused]."""
        clear_spherical_harmonics_cache()
        if not self.visibility.use_sigmoid:
            raise NameError("WARNING! ReLU has been used in visibility model.")
        return ops.sh_integration(pts, harmonics, X_cam, use_sigmoid=True, per_point=True)
