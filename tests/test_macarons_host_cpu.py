"""CPU tests of the host-side logic around the MACARONS scoring kernels (macarons_b200/utility/macarons_utils.py): the
distance factors (reference utility/macarons_utils.py:1741-1788) against the oracle restatement, the camera rows handed to
the kernels, and the argument / error behaviour that needs no GPU."""
import types

import numpy as np
import pytest
import torch

import synth
from macarons_b200.utility import macarons_utils as mu
from oracle import cameras as o_cams
from oracle import macarons_cov as o_mcov


def _params(th):
    return types.SimpleNamespace(distance_factor_th=th, image_height=256, image_width=456, sensor_range=70.,
                                 min_occ_for_proxy_points=0.1, seq_len=64, use_occ_to_sample_proxy_points=True, jz=False,
                                 ddp=False, k_for_knn=16, n_harmonics=64)


@pytest.mark.parametrize("th", [17.0, "smooth", None])
def test_distance_factors_match_oracle(th):
    gen = torch.Generator().manual_seed(3)
    pts = (torch.rand(500, 3, generator=gen) - 0.5) * 120
    X_cam = torch.tensor([[3., -2., 5.]])
    R, T = synth.look_at_RT(X_cam, torch.zeros(1, 3))
    cam = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=1000.)
    params = _params(th)
    if th is None:
        got = mu.get_distance_factor(params, pts, X_cam, cam, 0.5)
        want = o_mcov.distance_factor(pts, X_cam, cam.fov, 256, 456, 0.5)
    elif th == "smooth":
        got = mu.get_distance_factor_smooth(params, pts, X_cam, cam, 0.5)
        want = o_mcov.distance_factor_smooth(pts, X_cam, cam.fov, 256, 456, 0.5)
    else:
        got = mu.get_distance_factor_threshold(pts, X_cam, distance_th=th)
        want = o_mcov.distance_factor_threshold(pts, X_cam, distance_th=th)
    assert got.shape == want.shape == (500, 1)
    assert torch.allclose(got, want, rtol=1e-6, atol=0)
    # the batched form used by predict_coverage_gains_for_cameras: C cameras at once == one camera at a time
    eye = torch.tensor([[3., -2., 5.], [-10., 4., 1.], [0.5, 0.5, 30.]])
    R3, T3 = synth.look_at_RT(eye, torch.zeros(3, 3))
    batch = o_cams.FoVPerspectiveCameras(R=R3, T=T3, zfar=1000.)
    world = pts[:64][None].expand(3, -1, -1).contiguous()
    f = mu._distance_factors(params, world, eye, batch, 0.5)
    singles = [o_cams.FoVPerspectiveCameras(R=R3[c:c + 1], T=T3[c:c + 1], zfar=1000.) for c in range(3)]
    f_list = mu._distance_factors(params, world, eye, singles, 0.5)
    assert f.shape == (3, 64) and torch.equal(f, f_list)
    for c in range(3):
        if th is None:
            w = o_mcov.distance_factor(world[c], eye[c:c + 1], singles[c].fov, 256, 456, 0.5)
        elif th == "smooth":
            w = o_mcov.distance_factor_smooth(world[c], eye[c:c + 1], singles[c].fov, 256, 456, 0.5)
        else:
            w = o_mcov.distance_factor_threshold(world[c], eye[c:c + 1], distance_th=th)
        assert torch.allclose(f[c], w.view(-1), rtol=1e-6, atol=0)


def test_camera_rows_batched_equals_list():
    s = synth.macarons_scene(100, 5, 3)
    batch = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=1000.)
    singles = [o_cams.FoVPerspectiveCameras(R=s["R"][c:c + 1], T=s["T"][c:c + 1], zfar=1000.) for c in range(5)]
    rows_b, rows_l = mu._camera_rows(batch, "cpu"), mu._camera_rows(singles, "cpu")
    assert rows_b.shape == (5, 36) and torch.allclose(rows_b, rows_l, rtol=1e-6, atol=1e-6)
    # [full projection | world-to-view | centre | 0]: the centre is where the world-to-view transform maps to the origin
    centre = rows_b[:, 32:35]
    assert np.abs((centre - s["X_cam"]).numpy()).max() <= 1e-4
    view = rows_b[:, 16:32].view(5, 4, 4)
    origin = torch.cat((centre, torch.ones(5, 1)), dim=1)[:, None, :] @ view
    assert origin[:, 0, :3].abs().max().item() <= 1e-4
    inv = mu._unproject_rows(batch, "cpu")
    assert inv.shape == (5, 18)
    full = rows_b[:, :16].view(5, 4, 4)
    eye4 = full @ inv[:, :16].view(5, 4, 4)
    assert (eye4 - torch.eye(4)).abs().max().item() <= 1e-3


def test_argument_errors_without_gpu():
    params = _params(17.0)
    vis = types.SimpleNamespace(use_sigmoid=False)
    macarons = types.SimpleNamespace(visibility=vis)
    s = synth.macarons_scene(50, 2, 4)
    cams = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=1000.)
    args = (s["X_world"], s["vh"], s["occ"])
    with pytest.raises(NameError):       # ReLU visibility model, as Macarons.compute_visibility_gains (reference :176)
        mu.predict_coverage_gains_for_cameras(params, macarons, None, None, *args, None, s["X_cam"], cams, prediction_camera=cams)
    vis.use_sigmoid = True
    with pytest.raises(NameError):       # neither camera nor prediction camera (reference :1639-1641)
        mu.predict_coverage_gains_for_cameras(params, macarons, None, None, *args, None, s["X_cam"], cams)
    params.use_occ_to_sample_proxy_points = False
    with pytest.raises(NotImplementedError):
        mu.predict_coverage_gains_for_cameras(params, macarons, None, None, *args, None, s["X_cam"], cams, prediction_camera=cams)
    camera = types.SimpleNamespace(image_height=4, image_width=4, zfar=10., fov_camera=cams)
    with pytest.raises(NameError):       # several depth maps for the camera's single pose (reference :2466-2470)
        mu.get_signed_distance_to_depth_maps(camera, s["X_world"], torch.zeros(2, 4, 4, 1), torch.ones(2, 4, 4, 1, dtype=torch.bool))
    with pytest.raises(NameError):       # number of cameras != number of depth maps (reference :2472-2473)
        mu.get_signed_distance_to_depth_maps(camera, s["X_world"], torch.zeros(3, 4, 4, 1), torch.ones(3, 4, 4, 1, dtype=torch.bool),
                                             fov_camera=cams)


def test_view_space_bin_permutation_matches_reference_golden():
    """Host side of row a12: the 98 gather indices the product derives on the host must equal, bin for bin, the indices
    the reference function produced for the 24 golden camera poses (scone_utils.py:876-926)."""
    from conftest import load_golden
    from oracle import cameras as o_cams
    from macarons_b200.utility import scone_utils
    g = load_golden("move_view_state")
    aa, T = torch.from_numpy(g["axis_angle"]), torch.from_numpy(g["T"])
    for i in range(aa.shape[0]):
        R = o_cams.axis_angle_to_matrix(aa[i:i + 1])[0]
        assert np.array_equal(scone_utils.view_space_bin_permutation(R, 7, 14, T=T[i]).numpy(), g["indices"][i]), i


def test_frame_files_round_trip_and_match_the_reference_loader(tmp_path):
    """frames written by save_frame are the reference's `<n>.pt` dicts; the loader returns
    what the reference's load_images_for_depth_model returns from the same directory (when the reference tree is here),
    serves repeated loads from the resident cache and still reads frames it has never seen from disk."""
    import types
    from macarons_b200.utility import macarons_utils as mu
    from oracle import cameras as o_cams
    mu.clear_frame_cache()
    gen = torch.Generator().manual_seed(4)
    H, W = 12, 20
    camera = types.SimpleNamespace(image_height=H, image_width=W, zfar=700., device="cpu", n_frames_captured=0,
                                   save_dir_path=str(tmp_path), fov_camera=None)
    frames = []
    for i in range(5):
        R, T = synth.look_at_RT(torch.rand(1, 3, generator=gen) * 4 + 1, torch.zeros(1, 3))
        camera.fov_camera = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=700.)
        rgb = torch.rand(1, H, W, 3, generator=gen)
        zbuf = torch.where(torch.rand(1, H, W, 1, generator=gen) > 0.3, 1 + 5 * torch.rand(1, H, W, 1, generator=gen),
                           torch.full((1, H, W, 1), -1.0))
        path = mu.save_frame(camera, rgb, zbuf)
        assert path.endswith("%d.pt" % i) and camera.n_frames_captured == i + 1
        frames.append((rgb, zbuf, R, T))
    on_disk = torch.load(str(tmp_path / "3.pt"))
    assert set(on_disk) == {"rgb", "zbuf", "mask", "R", "T", "zfar"} and on_disk["zfar"] == 700.
    assert torch.equal(on_disk["mask"], frames[3][1] > -1)
    got = mu.load_images_for_depth_model(camera, n_frames=1, n_alpha=2, return_gt_zbuf=True)
    images, zbuf, mask, R, T, zfar = got
    assert images.shape == (3, H, W, 3) and mask.dtype == torch.bool and zfar.shape == (3,)
    for k, i in enumerate((2, 3, 4)):      # oldest first, ending at the latest frame
        assert torch.equal(images[k:k + 1], frames[i][0]) and torch.equal(zbuf[k:k + 1], frames[i][1])
        assert torch.equal(R[k:k + 1], frames[i][2]) and torch.equal(T[k:k + 1], frames[i][3])
    mu.clear_frame_cache()                   # cold: everything comes from the files
    cold = mu.load_images_for_depth_model(camera, n_frames=1, n_alpha=2, frame_nb=3)
    assert torch.equal(cold[0][2:3], frames[3][0]) and len(cold) == 5
    import os
    if os.path.isdir("/root/reference/macarons"):
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
        import ref_shim
        ref_shim.install()
        from macarons.utility import macarons_utils as ref_mu
        want = ref_mu.load_images_for_depth_model(camera, n_frames=1, n_alpha=2, return_gt_zbuf=True)
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_fov_camera_matches_the_pytorch3d_restatement():
    """macarons_b200.utility.cameras.FoVCamera (the package's own camera for machines without pytorch3d) against
    oracle/cameras.py (the restated pytorch3d conventions the reference goldens were generated with)."""
    from macarons_b200.utility import cameras
    from oracle import cameras as o_cams
    gen = torch.Generator().manual_seed(12)
    eye = (torch.rand(5, 3, generator=gen) - 0.5) * 8
    at = (torch.rand(5, 3, generator=gen) - 0.5) * 2
    R, T = cameras.look_at(eye, at)
    R0, T0 = synth.look_at_RT(eye, at)
    assert torch.allclose(R, R0, atol=1e-6) and torch.allclose(T, T0, atol=1e-5)
    ours = cameras.FoVCamera(R, T, zfar=750.)
    ref = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=750.)
    pts = (torch.rand(5, 40, 3, generator=gen) - 0.5) * 6
    for name in ("get_world_to_view_transform", "get_full_projection_transform", "get_projection_transform"):
        a, b = getattr(ours, name)(), getattr(ref, name)()
        assert torch.allclose(a.get_matrix(), b.get_matrix(), atol=1e-6), name
        assert torch.allclose(a.transform_points(pts), b.transform_points(pts), rtol=1e-5, atol=1e-5), name
    assert torch.allclose(ours.get_camera_center(), ref.get_camera_center(), atol=1e-5)
    inv = ours.get_world_to_view_transform().inverse()
    assert torch.allclose(inv.transform_points(ours.get_world_to_view_transform().transform_points(pts)), pts, atol=1e-4)
    full_inv = ours.get_full_projection_transform().inverse().get_matrix()
    assert torch.allclose(full_inv, ref.get_full_projection_transform().inverse().get_matrix(), rtol=1e-3, atol=1e-3)
    one = ours[2]
    assert len(one) == 1 and torch.equal(one.R, R[2:3]) and one.transform_points if hasattr(one, "transform_points") else True
    assert ours.get_world_to_view_transform().transform_points(pts[0, :, :].reshape(-1, 3)[:7]).shape[-1] == 3
