"""Oracle for the depth-side helpers on either side of the depth network.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: /root/reference/macarons/utility/macarons_utils.py
    Camera.__init__ (NDC tables)                 :1929-1938
    Camera.project_depth_in_3D                   :2339-2360
    Camera.compute_partial_point_cloud           :2362-2398
    Camera.get_points_zbuf                       :2437-2449
    Camera.get_signed_distance_to_depth_maps     :2451-2500
`tests/golden/make_golden.py` asserts bit-equality with the reference's own methods (called unbound on a stand-in for
`self`); the pytorch3d camera (`unproject_points`, transforms) is the restated one of oracle/cameras.py ("parity
unpinned" for that part).
"""
import torch


def ndc_tables(image_height, image_width):
    """(ndc_x_tab, ndc_y_tab), each (H, W): NDC coordinates of the pixel centres as Camera.__init__ builds them."""
    H, W = image_height, image_width
    x_tab = torch.Tensor([[i for j in range(W)] for i in range(H)])
    y_tab = torch.Tensor([[j for j in range(W)] for i in range(H)])
    ndc_x_tab = W / min(W, H) - (y_tab / (min(W, H) - 1)) * 2
    ndc_y_tab = H / min(W, H) - (x_tab / (min(W, H) - 1)) * 2
    return ndc_x_tab, ndc_y_tab


def project_depth_in_3D(depth, fov_cameras, image_height, image_width):
    """depth (B, H, W, 1) -> world points (B, H*W, 3)   [:2339-2360]"""
    batch_size = depth.shape[0]
    ndc_x_tab, ndc_y_tab = ndc_tables(image_height, image_width)
    ndc_points = torch.cat((ndc_x_tab.view(1, -1, 1).expand(batch_size, -1, -1),
                            ndc_y_tab.view(1, -1, 1).expand(batch_size, -1, -1),
                            depth.view(batch_size, -1, 1)), dim=-1).view(batch_size, image_height * image_width, 3)
    return fov_cameras.unproject_points(ndc_points, scaled_depth_input=False)


def compute_partial_point_cloud(depth, mask, fov_cameras, image_height, image_width, gathering_factor, images=None,
                                fov_range=None, perm=None):
    """depth, mask (1, H, W, 1) -> the masked world points, a random fraction `gathering_factor` of them  [:2362-2398].
    `perm` injects the permutation the reference draws with torch.randperm (CPU generator)."""
    points_mask = mask.view(1, -1) if fov_range is None else mask.view(1, -1) * (depth < fov_range).view(1, -1)
    world_points = project_depth_in_3D(depth, fov_cameras, image_height, image_width)[points_mask]
    n_points = int(len(world_points) * gathering_factor)
    if perm is None:
        perm = torch.randperm(len(world_points))
    idx = perm[:n_points]
    if images is None:
        return world_points[idx]
    return world_points[idx], (0. + images.view(1, -1, 3))[points_mask][idx]


def signed_distance_to_depth_maps(pts, depth_maps, mask, fov_camera, image_height, image_width, zfar):
    """pts (P, 3), depth_maps / mask (n, H, W, 1) -> (n, P, 1): z of the point in the camera minus the depth map sampled
    bilinearly at its projection (masked pixels count as 1.1 zfar); positive = behind the surface  [:2451-2500]."""
    n_depths = depth_maps.shape[0]
    pts_zbuf = fov_camera.get_world_to_view_transform().transform_points(pts)[..., 2:]          # get_points_zbuf
    if pts_zbuf.dim() == 2:
        pts_zbuf = pts_zbuf[None]
    depths = 0. + depth_maps
    depths[~mask.view(n_depths, image_height, image_width)] = 1.1 * zfar
    depths = (0. + torch.transpose(depths, -1, -2)).transpose(-2, -3)                           # (n, 1, H, W)
    proj = fov_camera.get_full_projection_transform().transform_points(pts)
    if proj.dim() == 2:
        proj = proj[None]
    factor = -1 * min(image_height, image_width)
    proj[..., 0] = factor / image_width * proj[..., 0]
    proj[..., 1] = factor / image_height * proj[..., 1]
    grid = proj[..., :2].view(n_depths, -1, 1, 2)
    map_zbuf = torch.nn.functional.grid_sample(input=depths, grid=grid, mode='bilinear', padding_mode='border')
    map_zbuf = (0. + torch.transpose(map_zbuf, -3, -2)).transpose(-2, -1).view(n_depths, -1, 1)
    return pts_zbuf - map_zbuf
