"""Oracle for rows a10-a12 of SURVEY.md section 8: view-state spherical-histogram binning and its
projection onto the SH basis.  TEST INFRASTRUCTURE (see oracle/__init__.py); torch CPU fp32 in the
reference's operation order.

Reference (paths relative to /root/reference/macarons):
  utility/utils.py:113-117            floor_divide  (x - x % d) / d
  utility/scone_utils.py:714-738      get_all_harmonics_under_degree
  utility/scone_utils.py:799-860      compute_view_state
  utility/scone_utils.py:863-930      move_view_state_to_view_space
  utility/scone_utils.py:934-960      compute_view_harmonics
  utility/CustomGeometry.py:5-24      get_cartesian_coords
"""
import numpy as np
import torch

from .sh_cov import real_sh_basis, spherical_coords


def float_floor_divide(x, d):
    """utils.py:113-117 -- float floor division built on torch.remainder."""
    return (x - x % d) / d


def bin_centre_harmonics(degree, n_elev, n_azim):
    """scone_utils.py:714-738 -> (base (degree^2, n_elev*n_azim), h_polar, h_azim).
    Bin (i, j) (elevation-major) sits at elev = -pi/2 + (i+1) pi/(n_elev+1), azim = 2 pi j / n_azim."""
    h_elev = torch.Tensor([-np.pi / 2 + (i + 1) / (n_elev + 1) * np.pi
                           for i in range(n_elev) for _ in range(n_azim)])
    h_polar = -h_elev + np.pi / 2
    h_azim = torch.Tensor([2 * np.pi * j / n_azim for _ in range(n_elev) for j in range(n_azim)])
    base = real_sh_basis(h_polar, h_azim, n_degree=degree)
    return base.transpose(0, 1), h_polar, h_azim


def view_state_bins(pts, X_view, n_elev, n_azim):
    """Bin index (B,P,V) int64 of every visited camera in every point's 7x14 spherical histogram
    (scone_utils.py:815-849, incl. the Python floor divisions -n_elev // 2 and -n_azim // 2 and the
    final wrap modulo n_elev*n_azim)."""
    B, P = pts.shape[0], pts.shape[1]
    V = len(X_view)
    elev_step = np.pi / (n_elev + 1)
    azim_step = 2 * np.pi / n_azim
    X_pts = pts[..., :3]
    rays = X_view.view(1, 1, V, 3).expand(B, P, -1, -1) - X_pts.view(B, P, 1, 3).expand(-1, -1, V, -1)
    _, elev, azim = spherical_coords(rays.reshape(-1, 3))
    elev, azim = elev.view(B, P, V), azim.view(B, P, V)
    ie = float_floor_divide(elev, elev_step)
    ia = float_floor_divide(azim, azim_step)
    ie = torch.where(elev % elev_step > elev_step / 2., ie + 1, ie)
    ia = torch.where(azim % azim_step > azim_step / 2., ia + 1, ia)
    ie = torch.where(ie >= n_elev, torch.full_like(ie, n_elev - 1), ie)
    ie = torch.where(ie < -n_elev // 2, torch.full_like(ie, -n_elev // 2), ie)
    ia = torch.where(ia > n_azim // 2, torch.full_like(ia, -n_azim // 2), ia)
    ie = ie + n_elev // 2
    ia = torch.where(ia < 0, ia + n_azim, ia)
    idx = ie.long() * n_azim + ia.long()
    return idx % (n_elev * n_azim)


def view_state(pts, X_view, n_elev, n_azim):
    """scone_utils.py:799-860 -> (B,P,n_elev*n_azim) fp32 in {0,1}: 1 where a visited camera falls."""
    B, P = pts.shape[0], pts.shape[1]
    n_bins = n_elev * n_azim
    idx = view_state_bins(pts, X_view, n_elev, n_azim)
    state = torch.zeros(B, P, n_bins)
    state.scatter_(2, idx, 1.0)
    return state


def view_harmonics(state, base, h_polar, n_elev, n_azim, point_chunk=4096):
    """scone_utils.py:934-960: spherical L2 product of the histogram with each basis function,
    sum_j state_j * base_kj * sin(polar_j) * polar_step * azim_step -> (B,P,n_harmonics).
    Evaluated in slices of `point_chunk` points (the reduction is per point, slicing is bit-neutral)."""
    n_h = base.shape[0]
    B, P, n_bins = state.shape
    polar_step = np.pi / (n_elev + 1)
    azim_step = 2 * np.pi / n_azim
    out = []
    for p0 in range(0, P, point_chunk):
        s = state[:, p0:p0 + point_chunk]
        n = s.shape[1]
        vals = s.view(B, n, 1, n_bins).expand(-1, -1, n_h, -1)
        polar = h_polar.view(1, 1, 1, n_bins).expand(B, n, n_h, -1)
        out.append(torch.sum(vals * base * torch.sin(polar) * polar_step * azim_step, dim=-1))
    return torch.cat(out, dim=1) if len(out) != 1 else out[0]


def cartesian_coords(r, elev, azim, in_degrees=False):
    """CustomGeometry.py:5-24: r, elev, azim (N,1) -> (N,3); x = cos(elev) sin(azim), y = sin(elev), z = cos(elev) cos(azim)."""
    factor = 1
    if in_degrees:
        factor *= np.pi / 180.
    X = torch.stack((torch.cos(factor * elev) * torch.sin(factor * azim),
                     torch.sin(factor * elev),
                     torch.cos(factor * elev) * torch.cos(factor * azim)), dim=2)
    return r * X.view(-1, 3)


def view_space_bin_indices(fov_camera, n_elev, n_azim):
    """scone_utils.py:876-926: for every bin of the n_elev x n_azim sphere grid, the index of the bin its direction
    falls into after mapping through the camera's inverse world-to-view transform minus the camera centre.
    `fov_camera` is any object with pytorch3d's FoVPerspectiveCameras interface (oracle/cameras.py stands in for it,
    pytorch3d itself is not installed: camera convention restated, see that file's header).
    NB the elevation clamps are +-(n_elev // 2) here, unlike compute_view_state (:838-839 vs :914-915)."""
    n_view = n_elev * n_azim
    candidate_dist = torch.Tensor([1. for _ in range(n_view)])
    candidate_elev = torch.Tensor([-90. + (i + 1) / (n_elev + 1) * 180. for i in range(n_elev) for _ in range(n_azim)])
    candidate_azim = torch.Tensor([360. * j / n_azim for _ in range(n_elev) for j in range(n_azim)])
    X_cam_ref = cartesian_coords(r=candidate_dist.view(-1, 1), elev=candidate_elev.view(-1, 1),
                                 azim=candidate_azim.view(-1, 1), in_degrees=True)
    X_cam_inv = fov_camera.get_world_to_view_transform().inverse().transform_points(X_cam_ref) \
        - fov_camera.get_camera_center()
    elev_step = np.pi / (n_elev + 1)
    azim_step = 2 * np.pi / n_azim
    _, ray_elev, ray_azim = spherical_coords(X_cam_inv.view(-1, 3))
    ray_elev, ray_azim = ray_elev.view(n_view), ray_azim.view(n_view)
    idx_elev = float_floor_divide(ray_elev, elev_step)
    idx_azim = float_floor_divide(ray_azim, azim_step)
    idx_elev = torch.where(ray_elev % elev_step > elev_step / 2., idx_elev + 1, idx_elev)
    idx_azim = torch.where(ray_azim % azim_step > azim_step / 2., idx_azim + 1, idx_azim)
    idx_elev = torch.where(idx_elev > n_elev // 2, torch.full_like(idx_elev, n_elev // 2), idx_elev)
    idx_elev = torch.where(idx_elev < -(n_elev // 2), torch.full_like(idx_elev, -(n_elev // 2)), idx_elev)
    idx_azim = torch.where(idx_azim > n_azim // 2, torch.full_like(idx_azim, -(n_azim // 2)), idx_azim)
    idx_elev = idx_elev + n_elev // 2
    idx_azim = torch.where(idx_azim < 0, idx_azim + n_azim, idx_azim)
    return idx_elev.long() * n_azim + idx_azim.long()


def move_view_state_to_view_space(state, fov_camera, n_elev, n_azim):
    """scone_utils.py:863-930: gather the bins of every point's view state with view_space_bin_indices."""
    n_clouds, seq_len = state.shape[0], state.shape[1]
    indices = view_space_bin_indices(fov_camera, n_elev, n_azim)
    return torch.gather(input=state, dim=2, index=indices.view(1, 1, -1).expand(n_clouds, seq_len, -1))
