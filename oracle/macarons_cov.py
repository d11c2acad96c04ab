"""Oracle for the MACARONS per-candidate coverage-gain prediction (FoV / occupancy
masking + proxy sampling on either side of `SconeVis.forward`, and the weighted integration after it).
TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: /root/reference/macarons/utility/macarons_utils.py
    Camera.get_points_in_fov                    :2400-2435
    predict_coverage_gain_for_single_camera     :1580-1738
    get_distance_factor / _threshold / _smooth  :1741-1788
and utility/scone_utils.py:788-796 (normalize_points_in_prediction_box).

The functions are restated over plain tensors + a `state_dict` of SconeVis (oracle/scone_nets.py) so that they run
without the reference tree; `tests/golden/make_golden.py` asserts bit-equality with the reference's own functions
(its `Camera`, `Macarons` and `SconeVis` classes) on seeded scenes.  The cameras are pytorch3d
`FoVPerspectiveCameras`-like objects (oracle/cameras.py restates them: that part is "parity unpinned").
"""
import numpy as np
import torch

from . import sampling, scone_nets, sh_cov


def points_in_fov(pts, fov_camera, ndc_bounds, fov_range=None):
    """pts (N,3) -> boolean mask (N,) of the points that project inside the image and lie in front of the camera,
    closer than `fov_range` to its centre [macarons_utils.py:2411-2430].  ndc_bounds = (min_x, max_x, min_y, max_y)."""
    min_x, max_x, min_y, max_y = ndc_bounds
    camera_center = fov_camera.get_camera_center()
    proj = fov_camera.get_full_projection_transform().transform_points(pts)
    view = fov_camera.get_world_to_view_transform().transform_points(pts)
    mask = (proj[:, 0] >= min_x) * (proj[:, 0] <= max_x) * (proj[:, 1] >= min_y) * (proj[:, 1] <= max_y) * (view[:, 2] > 0.)
    if fov_range is not None:
        mask = (torch.linalg.norm(pts - camera_center, dim=-1) < fov_range) * mask
    return mask


def distance_factor(pts, X_cam, fov_deg, image_height, image_width, cell_resolution):
    """[macarons_utils.py:1741-1765] 1 up to the distance at which a surface cell covers one pixel, then ~ 1/d^2."""
    focal_length = 1. / torch.tan(np.pi / 180. * fov_deg / 2.)
    pixel_size = 2. / min(image_height, image_width)
    epsilon = np.sqrt(np.pi) / 2. * cell_resolution
    distance_th = focal_length * epsilon / pixel_size
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    dist_mask = dists > distance_th
    res = torch.ones(pts.shape[0], 1)
    res[dist_mask] = epsilon ** 2 * (focal_length / pixel_size / dists[dist_mask]) ** 2
    return res


def distance_factor_threshold(pts, X_cam, distance_th=17.):
    """[macarons_utils.py:1768-1776]"""
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    dist_mask = dists > distance_th
    res = torch.ones(pts.shape[0], 1)
    res[dist_mask] *= distance_th ** 2 / (dists[dist_mask]) ** 2
    return res


def distance_factor_smooth(pts, X_cam, fov_deg, image_height, image_width, cell_resolution):
    """[macarons_utils.py:1779-1788]"""
    focal_length = 1. / torch.tan(np.pi / 180. * fov_deg / 2.)
    pixel_size = 2. / min(image_height, image_width)
    epsilon = np.sqrt(np.pi) / 2. * cell_resolution
    distance_th = focal_length * epsilon / pixel_size
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    return 1. / (1. + (dists / distance_th) ** 2)


def predict_coverage_gain_for_single_camera(vis_sd, X_world, proxy_view_harmonics, occ_probs, X_cam_world, fov_camera,
                                            prediction_camera, ndc_bounds, prediction_box_diag, sensor_range=70.,
                                            min_occ=0.1, seq_len=2048, distance_factor_th=17., image_height=256,
                                            image_width=456, cell_resolution=None, u=None, return_stages=False):
    """X_world (N,3), proxy_view_harmonics (N,64), occ_probs (N,1), X_cam_world (1,3)
    -> (proxy_points_world (1,seq_len,4), view_harmonics (1,seq_len,64), visibility_gains (1,1,seq_len),
        coverage_gain (1,1))  [macarons_utils.py:1600-1738, use_occ_to_sample_proxy_points=True].
    `u` injects the uniforms of the proxy sampling (the reference draws torch.rand(seq_len, 1))."""
    fov_mask = points_in_fov(X_world, fov_camera, ndc_bounds, sensor_range)                      # :1603-1605
    fov_X, fov_vh, fov_occ = X_world[fov_mask], proxy_view_harmonics[fov_mask], occ_probs[fov_mask]
    occ_mask = fov_occ[..., 0] > min_occ                                                          # :1610-1613
    fov_X, fov_vh, fov_occ = fov_X[occ_mask], fov_vh[occ_mask], fov_occ[occ_mask]
    if len(fov_X) == 0:
        # empty field of view (:1704-1736): a dummy forward pass, coverage gain 0; with all-zero inputs the
        # prediction of the model is irrelevant for the returned gain
        k = 16
        dummy_pts, dummy_vh = torch.zeros(1, k, 4), torch.zeros(1, k, 64)
        harm = scone_nets.scone_vis_forward(vis_sd, dummy_pts, dummy_vh)
        vis = sh_cov.visibility_gains(dummy_pts, harm, X_cam_world.view(1, -1, 3))
        return dummy_pts, dummy_vh, vis, (torch.mean(vis, dim=-1) * 0.).view(-1, 1)

    volume = fov_occ.sum()                                                                        # :1621
    proxy, vh, sample_idx = sampling.sample_proxy_points(fov_X, fov_occ, fov_vh, seq_len, min_occ, u=u)   # :1624-1628
    proxy_world = 0. + proxy
    center = (proxy[..., :3].max(dim=0, keepdim=True)[0] + proxy[..., :3].min(dim=0, keepdim=True)[0]).view(1, 3) / 2.
    view_transform = prediction_camera.get_world_to_view_transform()
    box_center = view_transform.transform_points(center)
    proxy[..., :3] = view_transform.transform_points(proxy[..., :3])
    proxy[..., :3] = (proxy[..., :3] - box_center) / prediction_box_diag                          # scone_utils.py:796
    proxy, vh = proxy.unsqueeze(0), vh.unsqueeze(0)
    X_cam = ((view_transform.transform_points(X_cam_world) - box_center) / prediction_box_diag).unsqueeze(0)
    harm = scone_nets.scone_vis_forward(vis_sd, proxy, vh)                                        # :1663
    proxy_s = proxy[0][sample_idx].unsqueeze(0)                                                   # :1669-1672
    proxy_world_s = proxy_world[sample_idx].unsqueeze(0)
    harm_s = harm[0][sample_idx].unsqueeze(0)
    vh_s = vh[0][sample_idx].unsqueeze(0)
    vis = sh_cov.visibility_gains(proxy_s, harm_s, X_cam)                                         # :1676-1683
    world_xyz = proxy_world_s[..., :3].view(-1, 3)
    if distance_factor_th is None:                                                                # :1685-1701
        factor = distance_factor(world_xyz, X_cam_world, fov_camera.fov, image_height, image_width, cell_resolution)
    elif distance_factor_th == 'smooth':
        factor = distance_factor_smooth(world_xyz, X_cam_world, fov_camera.fov, image_height, image_width, cell_resolution)
    else:
        factor = distance_factor_threshold(world_xyz, X_cam_world, distance_th=distance_factor_th)
    vis = vis * factor.view(1, 1, -1)
    coverage = torch.mean(vis, dim=-1) * volume                                                   # :1703
    if return_stages:
        return proxy_world_s, vh_s, vis, coverage.view(-1, 1), {"proxy": proxy, "sample_idx": sample_idx, "harm": harm,
                                                                "X_cam": X_cam, "n_unique": proxy.shape[1], "volume": volume}
    return proxy_world_s, vh_s, vis, coverage.view(-1, 1)
