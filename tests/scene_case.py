"""Shared set-up of the scene-occupancy-field tests: a synthetic MACARONS scene state after
three frames -- a height-field surface seen in three overlapping patches, proxy points uniform in the scene box, their
view states / pseudo ground-truth occupancy / out-of-field flags updated per frame through the `Scene` methods -- built
through whichever `Scene` class is handed in (the reference's in tests/golden/make_golden.py, the product's on the GPU,
the product's on the CPU as the container of the oracle run)."""
import types

import torch

import synth
from oracle import cameras as o_cams

X_MIN = (-4.0, -1.5, -4.5)
X_MAX = (4.5, 3.0, 4.0)
GRID = (3, 2, 3)


def params():
    return types.SimpleNamespace(harmonic_degree=8, view_state_n_elev=7, view_state_n_azim=14, n_view_state_cameras=98,
                                 n_harmonics=64, k_for_knn=16, prediction_neighborhood_size=3, jz=False, ddp=False)


def height(x, z):
    return 0.6 * torch.sin(0.9 * x) * torch.cos(0.7 * z) + 0.15 * x - 0.3


def prediction_camera(device="cpu"):
    R, T = synth.look_at_RT(torch.tensor([[3.0, 2.5, -5.0]]), torch.zeros(1, 3))
    return o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=1000., device=device)


def build(scene_cls, device, seed, n_proxy=6000, n_surface=9000):
    """-> (surface_scene, proxy_scene).  Consumes the global CPU generator (Cell.fill's randperm) after seeding it."""
    gen = torch.Generator().manual_seed(int(seed))
    x_min, x_max = torch.tensor(X_MIN), torch.tensor(X_MAX)
    common = dict(x_min=x_min.to(device), x_max=x_max.to(device), grid_l=GRID[0], grid_w=GRID[1], grid_h=GRID[2],
                  n_proxy_points=n_proxy, device=device, view_state_n_elev=7, view_state_n_azim=14)
    surface_scene = scene_cls(cell_capacity=300, cell_resolution=None, feature_dim=1, **common)
    proxy_scene = scene_cls(cell_capacity=100000, cell_resolution=0.001, feature_dim=1, score_threshold=0.95, **common)
    proxy_scene.initialize_proxy_points()
    proxy_scene.proxy_points = (x_min + (x_max - x_min) * torch.rand(n_proxy, 3, generator=gen)).to(device)
    xz = torch.rand(n_surface, 2, generator=gen) * torch.tensor([8.5, 8.5]) + torch.tensor([-4.0, -4.5])
    surface = torch.stack((xz[:, 0], height(xz[:, 0], xz[:, 1]), xz[:, 1]), dim=-1)
    frames = torch.tensor([[-2.5, 2.2, -2.5], [0.5, 2.6, 0.0], [2.5, 2.0, 2.0]])
    torch.manual_seed(int(seed))
    for f in range(frames.shape[0]):
        X_cam = frames[f:f + 1]
        seen = (surface - X_cam).norm(dim=-1) < 4.2
        part_pc = surface[seen][torch.randperm(int(seen.sum()), generator=gen)[:2500]].to(device)
        surface_scene.fill_cells(part_pc, features=torch.zeros(len(part_pc), 1, device=device))
        pp = proxy_scene.proxy_points
        pp_cpu = pp.cpu()
        fov_mask_cpu = (pp_cpu - X_cam).norm(dim=-1) < 4.6
        fov_mask = fov_mask_cpu.to(device)
        fov_pts = pp[fov_mask]
        proxy_scene.fill_cells(fov_pts, features=proxy_scene.get_proxy_indices_from_mask(fov_mask).view(-1, 1))
        # signed distance to the surface along -y: positive below (behind) the surface (evaluated on the host so that
        # every implementation sees identical values)
        fp = pp_cpu[fov_mask_cpu]
        sgn = (height(fp[:, 0], fp[:, 2]) - fp[:, 1]).view(-1, 1).to(device)
        camera = types.SimpleNamespace(X_cam=X_cam.to(device))
        proxy_scene.update_proxy_view_states(camera, fov_mask, signed_distances=sgn, distance_to_surface=None, X_cam=None)
        proxy_scene.update_proxy_supervision_occ(fov_mask, sgn, tol=0.3)
        proxy_scene.update_proxy_out_of_field(fov_mask)
        surface_scene.set_all_features_to_value(value=1.)
    return surface_scene, proxy_scene


def scene_digest(surface_scene, proxy_scene):
    """Summary of the bookkeeping state (cell contents, proxy state tensors) for cross-implementation comparison."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    counts = []
    for scene in (surface_scene, proxy_scene):
        for key in sorted(scene.cells):
            cell = scene.cells[key]
            counts.append(int(cell.cell_pts.shape[0]))
            h.update(np.ascontiguousarray(cell.cell_pts.cpu().numpy()).tobytes())
            if cell.use_feature:
                h.update(np.ascontiguousarray(cell.cell_features.cpu().numpy()).tobytes())
    for t in (proxy_scene.view_states, proxy_scene.proxy_supervision_occ, proxy_scene.out_of_field,
              proxy_scene.proxy_n_inside_fov, proxy_scene.proxy_n_behind_depth):
        h.update(np.ascontiguousarray(t.cpu().numpy()).tobytes())
    return h.hexdigest()[:16], counts
