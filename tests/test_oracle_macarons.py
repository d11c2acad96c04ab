"""CPU: the MACARONS candidate-scoring oracle (oracle/macarons_cov.py) against the fixtures generated from the
unmodified reference (`predict_coverage_gain_for_single_camera`, tests/golden/make_golden.py::macarons_cov_goldens)."""
import numpy as np
import pytest
import torch

import macarons_case
import synth
from conftest import load_golden
from oracle import macarons_cov as o_mcov


def _vis_sd(g):
    from macarons_b200.networks.SconeVis import SconeVis
    sd = synth.seeded_state_dict(SconeVis().state_dict(), int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"]), "seeded weights differ from the golden run"
    return sd


@pytest.mark.parametrize("name", ["macarons_cov_th17", "macarons_cov_smooth", "macarons_cov_pixel"])
def test_macarons_cov_oracle_matches_reference_golden(name):
    g = load_golden(name)
    N, C, seq_len = int(g["N"]), int(g["C"]), int(g["seq_len"])
    th = macarons_case.threshold_from_golden(g)
    s, params, cams, pred, camera, proxy_scene, surface_scene, nb, u = macarons_case.build(N, C, int(g["seed"]), th, seq_len)
    sd = _vis_sd(g)
    with torch.no_grad():
        for c in range(C):
            X_cam = cams[c].get_camera_center()
            pw, vh, vis, cov = o_mcov.predict_coverage_gain_for_single_camera(
                sd, s["X_world"], s["vh"], s["occ"], X_cam, cams[c], pred, nb, s["diag"], seq_len=seq_len,
                distance_factor_th=th, image_height=macarons_case.H, image_width=macarons_case.W, cell_resolution=0.5,
                u=u[c].view(-1, 1))
            assert pw.shape[1] == int(g["n_returned"][c])
            # same arithmetic as the golden run; other machines may reorder fp32 sums inside torch's GEMMs
            assert abs(cov.item() - float(g["coverage"][c])) <= 2e-4 * max(1.0, abs(float(g["coverage"][c])))
            assert abs(vis.double().sum().item() - float(g["visibility_sum"][c])) <= 2e-4 * max(1.0, float(g["visibility_sum"][c]))
    assert float(g["coverage"][-1]) == 0.0     # the empty field-of-view branch


def test_points_in_fov_geometry():
    """A point on the optical axis is inside; behind the camera, beyond the range or outside the image it is not."""
    from oracle import cameras as o_cams
    R, T = synth.look_at_RT(torch.tensor([[0., 0., -10.]]), torch.zeros(1, 3))
    cam = o_cams.FoVPerspectiveCameras(R=R, T=T, zfar=1000.)
    nb = synth.ndc_bounds()
    # depth 10 in front of the camera, fov 60 deg: ndc = offset / (10 tan 30 deg); the image spans |x| <= 1.78, |y| <= 1
    pts = torch.tensor([[0., 0., 0.], [0., 0., -20.], [0., 0., 100.], [0., 5., 0.], [0., 6.5, 0.], [9., 0., 0.], [-9., 0., 0.],
                        [11., 0., 0.]])
    mask = o_mcov.points_in_fov(pts, cam, nb, 70.)
    assert mask.tolist() == [True, False, False, True, False, True, True, False]
