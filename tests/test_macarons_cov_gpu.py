"""GPU parity of the MACARONS candidate-scoring path (macarons_b200/utility/macarons_utils.py over csrc/sampling.cu,
the ragged SconeVis forward and the coverage-gain kernel) against the oracle (oracle/macarons_cov.py) and the fixtures
generated from the unmodified reference (tests/golden/macarons_cov_*.npz)."""
import numpy as np
import pytest
import torch

import macarons_case
import synth
from conftest import load_golden
from oracle import macarons_cov as o_mcov
from oracle import scone_nets as o_nets

pytestmark = pytest.mark.gpu

COV_RTOL = 2e-3    # a proxy point whose projection rounds across an image border, or a uniform that falls within one
                   # fp32 step of a CDF value, changes one draw of 2048: <= 1e-3 of the integral
NET_RTOL = 1e-4    # ragged forward vs separate forwards / oracle (fp32-accurate 3xTF32 tensor-core layers)


def _model(dev):
    from macarons_b200.networks.Macarons import Macarons
    from macarons_b200.networks.SconeVis import SconeVis
    vis = SconeVis()
    sd = synth.seeded_state_dict(vis.state_dict(), 5)
    vis.load_state_dict(sd)
    return Macarons(None, None, vis.to(dev).eval()), sd


def test_ragged_forward_matches_separate_forwards(cuda_device):
    from macarons_b200 import netpack, ops
    dev = cuda_device
    macarons, sd = _model(dev)
    lens = [300, 64, 1, 257, 0, 130]
    B, S = len(lens), 320
    pts, vh = synth.sconevis_inputs(B, S, 77)
    for b, n in enumerate(lens):
        pts[b, n:] = 0
        vh[b, n:] = 0
    w = netpack.pack_sconevis(macarons.visibility)
    with torch.no_grad():
        ragged = ops.sconevis_forward(w, pts.to(dev), vh.to(dev), lens=torch.tensor(lens, dtype=torch.int32, device=dev))
        for b, n in enumerate(lens):
            if n == 0:
                continue
            single = macarons.visibility(pts[b:b + 1, :n].to(dev), view_harmonics=vh[b:b + 1, :n].to(dev))
            want = o_nets.scone_vis_forward(sd, pts[b:b + 1, :n], vh[b:b + 1, :n])
            scale = want.abs().max().item()
            assert (ragged[b, :n] - single[0]).abs().max().item() <= NET_RTOL * scale, "cloud %d vs separate forward" % b
            assert (ragged[b, :n].cpu() - want[0]).abs().max().item() <= NET_RTOL * scale, "cloud %d vs oracle" % b
    assert torch.isfinite(ragged).all()


def test_fov_selection_and_sampling_match_oracle(cuda_device):
    from macarons_b200 import ops
    from macarons_b200.utility import macarons_utils as mu
    from oracle import sampling as o_sampling
    dev = cuda_device
    N, C, S = 20000, 6, 2048
    s, params, cams, pred, camera, _, _, nb, u = macarons_case.build(N, C, 701, 17.0, S)
    res, res_h, inverse, counts, volume = ops.fov_sample_proxy(
        s["X_world"].to(dev), s["occ"].to(dev), s["vh"].to(dev), mu._camera_rows(cams, dev), nb, 70., 0.1, u.to(dev))
    counts, res, res_h, inverse, volume = counts.cpu(), res.cpu(), res_h.cpu(), inverse.cpu(), volume.cpu()
    for c in range(C):
        mask = o_mcov.points_in_fov(s["X_world"], cams[c], nb, 70.) & (s["occ"][:, 0] > 0.1)
        assert abs(int(counts[c, 0]) - int(mask.sum())) <= 1, "candidate %d: points kept" % c
        if int(mask.sum()) == 0:
            assert int(counts[c, 1]) == 0
            continue
        assert abs(volume[c].item() - s["occ"][mask].sum().item()) <= 1e-5 * s["occ"][mask].sum().item() + 1.0
        if int(counts[c, 0]) != int(mask.sum()):
            continue   # a border point flipped: every later index shifts, compared through the coverage gain instead
        want, want_h, want_inv = o_sampling.sample_proxy_points(s["X_world"][mask], s["occ"][mask], s["vh"][mask], S, 0.1,
                                                               u=u[c].view(-1, 1))
        got_draws = res[c][inverse[c]]                      # the 2048 draws with duplicates
        want_draws = want[want_inv]
        differing = (got_draws != want_draws).any(dim=-1).float().mean().item()
        assert differing <= 5e-3, "candidate %d: %.4f of the draws differ" % (c, differing)
        if differing == 0:
            n = int(counts[c, 1])
            assert n == want.shape[0] and torch.equal(res[c, :n], want) and torch.equal(res_h[c, :n], want_h)
            assert torch.equal(inverse[c], want_inv)


@pytest.mark.parametrize("name", ["macarons_cov_th17", "macarons_cov_smooth", "macarons_cov_pixel"])
def test_batched_candidates_match_oracle_and_reference_golden(cuda_device, name):
    from macarons_b200.utility import macarons_utils as mu
    dev = cuda_device
    g = load_golden(name)
    N, C, S = int(g["N"]), int(g["C"]), int(g["seq_len"])
    th = macarons_case.threshold_from_golden(g)
    s, params, cams, pred, camera, proxy_scene, surface_scene, nb, u = macarons_case.build(N, C, int(g["seed"]), th, S, device=dev)
    macarons, sd = _model(dev)
    X_cams = torch.cat([cam.get_camera_center() for cam in cams]).to(dev)
    with torch.no_grad():
        out = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, s["X_world"].to(dev),
                                                    s["vh"].to(dev), s["occ"].to(dev), camera, X_cams, cams,
                                                    prediction_camera=pred, samples=u.to(dev))
    cov = out["coverage_gain"].view(-1).cpu().numpy()
    ref = g["coverage"]
    # the same candidates as ONE batched camera object (three matrix calls instead of 3 C)
    from oracle import cameras as o_cams
    batch_cam = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=1000., device=dev)
    with torch.no_grad():
        out_b = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, s["X_world"].to(dev),
                                                      s["vh"].to(dev), s["occ"].to(dev), camera, batch_cam.get_camera_center(),
                                                      batch_cam, prediction_camera=pred, samples=u.to(dev))
    assert np.abs(out_b["coverage_gain"].view(-1).cpu().numpy() - cov).max() <= COV_RTOL * np.abs(ref).max()
    assert np.abs(cov - ref).max() <= COV_RTOL * np.abs(ref).max(), (cov, ref)
    assert cov[-1] == 0.0 and int(out["n_points_in_fov"][-1]) == 0           # empty field of view
    assert int(np.argmax(cov)) == int(np.argmax(ref))                        # the NBV among the candidates
    vis_sum = out["visibility_gains"].double().sum(dim=(1, 2)).cpu().numpy()
    live = ref > 0
    assert np.abs(vis_sum[live] - g["visibility_sum"][live]).max() <= COV_RTOL * np.abs(g["visibility_sum"]).max()
    # against the oracle, candidate by candidate, with the per-point gains of the draws
    s_cpu, _, cams_cpu, pred_cpu, _, _, _, _, _ = macarons_case.build(N, C, int(g["seed"]), th, S)
    with torch.no_grad():
        for c in range(C - 1):
            pw, vh, vis, want_cov = o_mcov.predict_coverage_gain_for_single_camera(
                sd, s_cpu["X_world"], s_cpu["vh"], s_cpu["occ"], cams_cpu[c].get_camera_center(), cams_cpu[c], pred_cpu, nb,
                s_cpu["diag"], seq_len=S, distance_factor_th=th, image_height=macarons_case.H, image_width=macarons_case.W,
                cell_resolution=0.5, u=u[c].view(-1, 1))
            assert abs(cov[c] - want_cov.item()) <= COV_RTOL * abs(want_cov.item())
            got_pw = out["proxy_points_world"][c].cpu()
            same = (got_pw == pw[0]).all(dim=-1)
            assert same.float().mean().item() >= 1 - 5e-3
            if bool(same.all()):
                err = (out["visibility_gains"][c, 0].cpu() - vis[0, 0]).abs()
                assert err.max().item() <= 5e-3 and err.median().item() <= 2e-5


def test_single_camera_api(cuda_device):
    """Reference signature: 4 return values, (1, seq_len, .) shapes, dummy 16-point pass for an empty field of view."""
    from macarons_b200.utility import macarons_utils as mu
    dev = cuda_device
    N, C, S = 6000, 4, 512
    s, params, cams, pred, camera, proxy_scene, surface_scene, nb, u = macarons_case.build(N, C, 702, "smooth", S, device=dev)
    macarons, _ = _model(dev)
    args = (s["X_world"].to(dev), s["vh"].to(dev), s["occ"].to(dev), camera)
    torch.manual_seed(3)
    with torch.no_grad():
        pw, vh, vis, cov = mu.predict_coverage_gain_for_single_camera(params, macarons, proxy_scene, surface_scene, *args,
                                                                     cams[0].get_camera_center(), cams[0])
        assert pw.shape == (1, S, 4) and vh.shape == (1, S, 64) and vis.shape == (1, 1, S) and cov.shape == (1, 1)
        assert cov.item() > 0
        pw, vh, vis, cov = mu.predict_coverage_gain_for_single_camera(params, macarons, proxy_scene, surface_scene, *args,
                                                                     cams[-1].get_camera_center(), cams[-1],
                                                                     prediction_camera=pred)
        assert pw.shape == (1, 16, 4) and vh.shape == (1, 16, 64) and vis.shape == (1, 1, 16) and cov.item() == 0.0
    with pytest.raises(NameError):
        camera.fov_camera_0 = None
        mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, *args[:3], None,
                                              cams[0].get_camera_center(), [cams[0]])


def test_full_size_scene_properties(cuda_device):
    """BASELINE config 3 shape (128 candidate poses over 100 k proxy points): the batched pass must agree with scoring the
    candidates one at a time through the reference-signature function (same kernels, C = 1), gains are non-negative and
    bounded by the proxy volume in the field of view, and an all-zero occupancy field scores zero everywhere."""
    from macarons_b200.utility import macarons_utils as mu
    from oracle import cameras as o_cams
    dev = cuda_device
    N, C, S = 100000, 128, 2048
    s, params, cams, pred, camera, proxy_scene, surface_scene, nb, _ = macarons_case.build(N, C, 31, 17.0, S, device=dev)
    macarons, _ = _model(dev)
    batch_cam = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=1000., device=dev)
    X_cams = batch_cam.get_camera_center()
    u = torch.rand(C, S, generator=torch.Generator().manual_seed(5)).to(dev)
    args = (s["X_world"].to(dev), s["vh"].to(dev), s["occ"].to(dev), camera)
    with torch.no_grad():
        out = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, *args, X_cams, batch_cam,
                                                    prediction_camera=pred, samples=u)
        cov = out["coverage_gain"].view(-1)
        assert torch.isfinite(cov).all() and (cov >= 0).all()
        assert (cov <= out["fov_proxy_volume"] + 1e-3).all()          # sigmoid gains and distance factors are <= 1
        assert int((out["n_points_in_fov"] > 0).sum()) >= C // 2 and int(out["n_points_in_fov"][-1]) == 0
        for c in (0, 37, 126):
            one = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, *args, X_cams[c:c + 1],
                                                        [cams[c]], prediction_camera=pred, samples=u[c:c + 1])
            assert int(one["n_unique"][0]) == int(out["n_unique"][c])
            assert abs(one["coverage_gain"].item() - cov[c].item()) <= 1e-4 * max(1.0, cov[c].item())
        zero = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, args[0], args[1],
                                                     torch.zeros_like(args[2]), camera, X_cams, batch_cam,
                                                     prediction_camera=pred, samples=u)
        assert float(zero["coverage_gain"].abs().max()) == 0.0
