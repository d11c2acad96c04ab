"""Depth-side helpers: CPU oracle against the fixture generated from the unmodified
reference's `Camera` methods, and the CUDA kernels behind macarons_b200.utility.macarons_utils against the oracle."""
import types
import warnings

import numpy as np
import pytest
import torch

import synth
from conftest import load_golden
from oracle import cameras as o_cams
from oracle import depth_io as o_dio

warnings.filterwarnings("ignore", message="Default grid_sample")


def _case(device="cpu"):
    g = load_golden("depth_io_48x80")
    H, W = int(g["H"]), int(g["W"])
    s = synth.depth_io_inputs(H, W, int(g["seed"]))
    cam = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=100., device=device)
    cam1 = o_cams.FoVPerspectiveCameras(R=s["R"][:1], T=s["T"][:1], zfar=100., device=device)
    return g, H, W, s, cam, cam1


def test_depth_io_oracle_matches_reference_golden():
    g, H, W, s, cam, cam1 = _case()
    world = o_dio.project_depth_in_3D(s["depth"], cam, H, W)
    assert np.abs(world[:, ::7].numpy() - g["world_points"]).max() <= 1e-4
    pc = o_dio.compute_partial_point_cloud(s["depth"][:1], s["mask"][:1], cam1, H, W, 0.05, fov_range=9.0,
                                           perm=torch.from_numpy(g["perm"]))
    assert pc.shape == g["partial_pc"].shape and np.abs(pc.numpy() - g["partial_pc"]).max() <= 1e-4
    sd = o_dio.signed_distance_to_depth_maps(s["pts"], s["depth"], s["mask"], cam, H, W, 100.)
    assert np.abs(sd.numpy() - g["signed_distance"]).max() <= 1e-4


def test_unprojection_round_trip_on_cpu():
    """Un-projected pixels project back onto their NDC coordinates at their depth (property of the restated camera)."""
    g, H, W, s, cam, cam1 = _case()
    world = o_dio.project_depth_in_3D(s["depth"], cam, H, W)
    nx, ny = o_dio.ndc_tables(H, W)
    back = cam.get_full_projection_transform().transform_points(world)
    view_z = cam.get_world_to_view_transform().transform_points(world)[..., 2]
    assert (back[..., 0] - nx.view(1, -1)).abs().max().item() <= 1e-4
    assert (back[..., 1] - ny.view(1, -1)).abs().max().item() <= 1e-4
    assert (view_z - s["depth"].view(2, -1)).abs().max().item() <= 1e-3


@pytest.mark.gpu
def test_depth_io_kernels_match_oracle(cuda_device):
    from macarons_b200.utility import macarons_utils as mu
    dev = cuda_device
    g, H, W, s, cam, cam1 = _case(device=dev)
    _, _, _, _, cam_cpu, cam1_cpu = _case()
    camera = types.SimpleNamespace(image_height=H, image_width=W, zfar=100., gathering_factor=0.05, fov_camera=cam1)
    # un-projection of every pixel
    world = mu.project_depth_in_3D(camera, s["depth"].to(dev), fov_cameras=cam)
    want = o_dio.project_depth_in_3D(s["depth"], cam_cpu, H, W)
    assert world.shape == want.shape and (world.cpu() - want).abs().max().item() <= 2e-4
    assert np.abs(world[:, ::7].cpu().numpy() - g["world_points"]).max() <= 2e-4
    # partial point cloud: same permutation as the golden run
    torch.manual_seed(int(g["seed"]))
    pc, col = mu.compute_partial_point_cloud(camera, s["depth"][:1].to(dev), s["mask"][:1].to(dev), images=s["images"][:1].to(dev),
                                             fov_cameras=cam1, fov_range=9.0)
    assert pc.shape == g["partial_pc"].shape and col.shape == pc.shape
    assert np.abs(pc.cpu().numpy() - g["partial_pc"]).max() <= 2e-4
    # signed distances to both depth maps
    sd = mu.get_signed_distance_to_depth_maps(camera, s["pts"].to(dev), s["depth"].to(dev), s["mask"].to(dev), fov_camera=cam)
    want = o_dio.signed_distance_to_depth_maps(s["pts"], s["depth"], s["mask"], cam_cpu, H, W, 100.)
    err = (sd.cpu() - want).abs()
    # the depth maps of this case are white noise with masked pixels at 1.1 zfar: a sampling position that moves by
    # 4e-6 pixel (one ulp of the projection, FMA vs separate multiply-add) changes the bilinear sample by up to 4e-4
    assert err.median().item() <= 5e-4 and (err > 2e-2).float().mean().item() <= 5e-3
    assert np.abs(sd.cpu().numpy() - g["signed_distance"]).max() <= 120.0 and sd.shape == want.shape
    with pytest.raises(NameError):
        mu.get_signed_distance_to_depth_maps(camera, s["pts"].to(dev), s["depth"].to(dev), s["mask"].to(dev))
    with pytest.raises(NameError):
        mu.get_signed_distance_to_depth_maps(camera, s["pts"].to(dev), s["depth"][:1].to(dev), s["mask"][:1].to(dev), fov_camera=cam)


@pytest.mark.gpu
def test_depth_io_full_resolution_properties(cuda_device):
    """256 x 456 (the MACARONS image size): un-projected pixels lie at their depth in front of the camera; world points and
    signed distances agree with the oracle at full resolution; the partial cloud keeps the reference's fraction of pixels."""
    from macarons_b200.utility import macarons_utils as mu
    dev = cuda_device
    H, W = 256, 456
    gen = torch.Generator().manual_seed(5)
    R, T = synth.look_at_RT(torch.tensor([[2., 1., -9.]]), torch.zeros(1, 3))
    cam = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=100., device=dev)
    low = 4 + 4 * torch.rand(1, 1, H // 16 + 2, W // 16 + 2, generator=gen)
    depth = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).view(1, H, W, 1).to(dev)
    camera = types.SimpleNamespace(image_height=H, image_width=W, zfar=100., gathering_factor=0.05, fov_camera=cam)
    world = mu.project_depth_in_3D(camera, depth)
    view_z = cam.get_world_to_view_transform().transform_points(world)[..., 2]
    # fp32 through the inverse of a znear = 1 / zfar = 100 projection (as the reference does it): 1e-6 relative on the
    # matrix moves a point at depth 8 by ~5e-4 (measured on the CPU oracle), hence the 5e-3 / 1e-2 bounds on depths of 4-8
    z_err = (view_z - depth.view(1, -1)).abs().max().item()
    assert z_err <= 5e-3, z_err
    mask = torch.ones(1, H, W, 1, dtype=torch.bool, device=dev)
    # signed distance of the un-projected pixels to their own depth map, against the oracle on identical inputs (it is not
    # zero: the reference's NDC tables step by 2 / (min(H, W) - 1) while grid_sample uses align_corners = False, so a pixel
    # re-projects up to ~half a pixel away from its own centre; the oracle reproduces that, max 0.29 on this depth map)
    cam_cpu = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=100.)
    world_cpu = o_dio.project_depth_in_3D(depth.cpu(), cam_cpu, H, W)
    want = o_dio.signed_distance_to_depth_maps(world_cpu[0], depth.cpu(), mask.cpu(), cam_cpu, H, W, 100.)
    sd = mu.get_signed_distance_to_depth_maps(camera, world_cpu[0].to(dev), depth, mask)
    assert sd.shape == (1, H * W, 1)
    sd_err = (sd.cpu() - want).abs().max().item()
    assert sd_err <= 5e-3, sd_err
    assert (world.cpu() - world_cpu).abs().max().item() <= 5e-3
    pc = mu.compute_partial_point_cloud(camera, depth, mask)
    assert pc.shape == (int(H * W * 0.05), 3)
