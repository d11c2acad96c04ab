"""GPU: one full MACARONS NBV step (BASELINE.json configs[2]: depth 256 x 256 -> partial cloud -> scene update -> occupancy
field -> 128 candidate poses -> argmax) through macarons_b200.nbv.macarons_nbv_step, against the CHAINED oracle.

The chain is checked stage by stage: every oracle stage is fed the product's previous-stage output, so a legitimate
rounding difference in an early stage (e.g. one depth pixel 1e-4 off) cannot change the random sub-sample sizes of a later
one (Cell.fill's randperm length) and turn the comparison into noise.  The final NBV index must be the oracle's."""
import contextlib
import io
import types

import numpy as np
import pytest
import torch

import scene_case
import synth
from macarons_b200 import nbv, ops
from macarons_b200.networks import ManyDepth as MD
from macarons_b200.networks.Macarons import Macarons
from macarons_b200.networks.SconeOcc import SconeOcc
from macarons_b200.networks.SconeVis import SconeVis
from macarons_b200.utility import scene
from oracle import cameras as o_cams
from oracle import depth as o_depth
from oracle import depth_io as o_dio
from oracle import macarons_cov as o_mcov
from oracle import scene as o_scene
from oracle import view_state as o_vs

pytestmark = pytest.mark.gpu

H = W = 256
N_CAND = 128


def _params():
    p = scene_case.params()
    for k, v in dict(znear=0.5, zfar=750., gathering_factor=0.05, sensor_range=6.0, carving_tolerance=0.3, seq_len=512,
                     min_occ_for_proxy_points=0.1, use_occ_to_sample_proxy_points=True, distance_factor_th=17.0,
                     image_height=H, image_width=W).items():
        setattr(p, k, v)
    return p


def _models(dev):
    with contextlib.redirect_stdout(io.StringIO()):
        resnet = MD.ResNet18Trunk()
        depth = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet, input_height=H, input_width=W), None)
        occ, vis = SconeOcc(), SconeVis()
    sds = [synth.seeded_state_dict(m.state_dict(), 5) for m in (depth, occ, vis)]
    for m, sd in zip((depth, occ, vis), sds):
        m.load_state_dict(sd)
    return Macarons(depth, occ, vis).to(dev).eval(), sds


def _scenes(scene_cls, dev, gen_seed):
    gen = torch.Generator().manual_seed(gen_seed)
    x_min, x_max = torch.tensor([-3.0, -2.0, -4.0]), torch.tensor([3.0, 2.5, 2.5])
    common = dict(x_min=x_min.to(dev), x_max=x_max.to(dev), grid_l=3, grid_w=2, grid_h=3, n_proxy_points=8000, device=dev,
                  view_state_n_elev=7, view_state_n_azim=14)
    surface_scene = scene_cls(cell_capacity=400, cell_resolution=None, feature_dim=1, **common)
    proxy_scene = scene_cls(cell_capacity=100000, cell_resolution=0.001, feature_dim=1, score_threshold=0.95, **common)
    proxy_scene.initialize_proxy_points()
    proxy_scene.proxy_points = (x_min + (x_max - x_min) * torch.rand(8000, 3, generator=gen)).to(dev)
    return surface_scene, proxy_scene


def _camera(dev):
    eye = torch.tensor([[0.2, 0.4, -3.2]])
    R, T = synth.look_at_RT(eye, torch.zeros(1, 3))
    cam = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=750., device=dev)
    nb = synth.ndc_bounds(H, W)
    return types.SimpleNamespace(image_height=H, image_width=W, zfar=750., gathering_factor=0.05, fov_camera=cam,
                                 fov_camera_0=cam, X_cam=eye.to(dev), min_ndc_x=nb[0], max_ndc_x=nb[1], min_ndc_y=nb[2],
                                 max_ndc_y=nb[3]), nb


def _candidates(dev):
    gen = torch.Generator().manual_seed(99)
    eye = (torch.rand(N_CAND, 3, generator=gen) - 0.5) * torch.tensor([5.0, 3.0, 5.5]) + torch.tensor([0.0, 0.5, -0.8])
    at = (torch.rand(N_CAND, 3, generator=gen) - 0.5) * 2.0
    R, T = synth.look_at_RT(eye, at)
    return eye, R, T, o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=750., device=dev)


def test_full_macarons_nbv_step_against_chained_oracle(cuda_device, monkeypatch):
    dev = cuda_device
    params = _params()
    macarons, (depth_sd, occ_sd, vis_sd) = _models(dev)
    x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(1, H, W, 11)
    frames = {k: v.to(dev) for k, v in dict(x=x, x_alpha=x_alpha, R=R, T=T, zfar=zfar, gt_pose=gt_pose).items()}
    camera, nb = _camera(dev)
    camera_cpu, _ = _camera("cpu")
    eye, cR, cT, cand = _candidates(dev)
    _, _, _, cand_cpu = _candidates("cpu")
    surface_scene, proxy_scene = _scenes(scene.Scene, dev, 7)
    u = torch.rand(N_CAND, params.seq_len, generator=torch.Generator().manual_seed(3))

    torch.manual_seed(123)
    n0 = ops.launch_count()
    cov, best, st = nbv.macarons_nbv_step(params, macarons, camera, surface_scene, proxy_scene, frames, eye.to(dev), cand,
                                          samples=u.to(dev))
    torch.cuda.synchronize()
    print("full MACARONS NBV step: %d kernel launches of this library" % (ops.launch_count() - n0))
    assert cov.shape == (N_CAND, 1) and torch.isfinite(cov).all()

    # ---- stage 1: depth ----
    with torch.no_grad():
        want = o_depth.many_depth_forward(depth_sd, x, x_alpha, R, T, zfar, gt_pose)
    assert (st["disp1"].cpu() - want[1]).abs().max().item() <= 5e-4
    depth_cpu = st["depth"].cpu()
    assert (depth_cpu - o_depth.depth_from_disparity(st["disp1"].cpu()).permute(0, 2, 3, 1)).abs().max().item() <= 1e-5 * 750

    # ---- stage 2: partial cloud from the product's depth (same permutation: same generator state) ----
    mask = torch.ones(1, H, W, 1, dtype=torch.bool)
    torch.manual_seed(123)
    pc_want = o_dio.compute_partial_point_cloud(depth_cpu, mask, camera_cpu.fov_camera, H, W, 0.05, fov_range=params.sensor_range)
    part_pc = st["part_pc"].cpu()
    assert part_pc.shape == pc_want.shape and part_pc.shape[0] > 1000
    assert (part_pc - pc_want).abs().max().item() <= 5e-4

    # ---- stage 3: field of view + signed distances of the proxy points ----
    pp = proxy_scene.proxy_points.cpu()
    fov_want = o_mcov.points_in_fov(pp, camera_cpu.fov_camera, nb, params.sensor_range)
    fov_got = st["fov_proxy_mask"].cpu()
    assert (fov_want != fov_got).float().mean().item() <= 1e-3 and fov_got.sum().item() > 500
    sgn_want = o_dio.signed_distance_to_depth_maps(pp[fov_got], depth_cpu, mask, camera_cpu.fov_camera, H, W, 750.)
    sgn_err = (st["sgn_dists"].cpu() - sgn_want).abs()
    assert sgn_err.median().item() <= 5e-4 and (sgn_err > 2e-2).float().mean().item() <= 5e-3

    # ---- stage 4: the same scene update on a CPU container scene, fed the product's stage outputs ----
    monkeypatch.setattr(scene, "compute_view_state", o_vs.view_state)
    o_surface, o_proxy = _scenes(scene.Scene, "cpu", 7)
    torch.manual_seed(123)
    torch.randperm(int((depth_cpu < params.sensor_range).sum()))   # the partial cloud's permutation, as the step drew it
    o_surface.fill_cells(part_pc, features=torch.zeros(len(part_pc), 1))
    o_proxy.fill_cells(pp[fov_got], features=o_proxy.get_proxy_indices_from_mask(fov_got).view(-1, 1))
    sgn = st["sgn_dists"].cpu()
    o_proxy.update_proxy_view_states(camera_cpu, fov_got, signed_distances=sgn)
    o_proxy.update_proxy_supervision_occ(fov_got, sgn, tol=params.carving_tolerance)
    o_proxy.update_proxy_out_of_field(fov_got)
    o_surface.set_all_features_to_value(value=1.)
    monkeypatch.undo()
    for key in o_proxy.cells:
        assert torch.equal(o_proxy.cells[key].cell_features, proxy_scene.cells[key].cell_features.cpu()), key
    for key in o_surface.cells:
        assert torch.equal(o_surface.cells[key].cell_pts, surface_scene.cells[key].cell_pts.cpu()), key
    vs_diff = (o_proxy.view_states != proxy_scene.view_states.cpu()).any(-1).float().mean().item()
    assert vs_diff <= 1e-3
    o_proxy.view_states = proxy_scene.view_states.cpu().clone()       # re-synchronise (bin-boundary rays)
    assert torch.equal(o_proxy.proxy_supervision_occ, proxy_scene.proxy_supervision_occ.cpu())
    assert torch.equal(o_proxy.out_of_field, proxy_scene.out_of_field.cpu())

    # ---- stage 5: occupancy field (the product ran it inside the step with the generator state that follows) ----
    # the CPU generator now stands where the step's field computation found it: 1 randperm (partial cloud) + the
    # Cell.fill permutations of both scenes have been replayed above
    with torch.no_grad():
        oX, ovh, oprobs = o_scene.scene_occupancy_field(params, occ_sd, o_surface, o_proxy, camera_cpu.fov_camera_0)
    assert torch.equal(oX, st["X_world"].cpu())
    assert (ovh - st["view_harmonics"].cpu()).abs().max().item() <= 2e-6
    err = (oprobs - st["occ_probs"].cpu()).abs()[:, 0]
    scale = max(1.0, oprobs.abs().max().item())
    print("occupancy field: %d points, err median %.2e q98 %.2e max %.2e" % (len(err), err.median(), np.quantile(err.numpy(), 0.98),
                                                                         err.max()))
    assert np.quantile(err.numpy(), 0.98) <= 1e-4 * scale and err.max().item() <= 0.05 * scale

    # ---- stage 6: candidates, on the product's field ----
    Xw, vh, occp = st["X_world"].cpu(), st["view_harmonics"].cpu(), st["occ_probs"].cpu()
    diag = torch.linalg.norm(proxy_scene.x_max - proxy_scene.x_min).item()
    want_cov = torch.zeros(N_CAND)
    with torch.no_grad():
        for c in range(N_CAND):
            cam_c = o_cams.FoVPerspectiveCameras(R=cR[c:c + 1], T=cT[c:c + 1], zfar=750.)
            want_cov[c] = o_mcov.predict_coverage_gain_for_single_camera(
                vis_sd, Xw, vh, occp, eye[c:c + 1], cam_c, camera_cpu.fov_camera_0, nb, diag, sensor_range=params.sensor_range,
                seq_len=params.seq_len, distance_factor_th=17.0, image_height=H, image_width=W, cell_resolution=None,
                u=u[c].view(-1, 1))[3].item()
    got_cov = cov.view(-1).cpu()
    rel = (got_cov - want_cov).abs().max().item() / max(want_cov.abs().max().item(), 1e-9)
    print("candidates: max |gain - oracle| / max gain = %.2e; NBV %d (oracle %d)" % (rel, int(best), int(want_cov.argmax())))
    assert rel <= 2e-3
    assert int(best) == int(want_cov.argmax())
