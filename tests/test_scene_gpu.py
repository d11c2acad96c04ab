"""GPU parity of SURVEY.md section 8f rank 2: the ragged SconeOcc forward over cells, the Scene bookkeeping on the
device, and the batched occupancy probability field of a scene against the outputs of the reference's own per-cell loop
(tests/golden/scene_field_*.npz) and the oracle (oracle/scene.py)."""
import contextlib
import io

import numpy as np
import pytest
import torch

import scene_case
import synth
from conftest import load_golden
from macarons_b200 import ops
from macarons_b200.networks.Macarons import Macarons
from macarons_b200.networks.SconeOcc import SconeOcc
from macarons_b200.utility import macarons_utils, scene
from oracle import scone_nets as o_nets

pytestmark = pytest.mark.gpu

NET_RTOL = 1e-4


def _occ(dev, seed=5):
    with contextlib.redirect_stdout(io.StringIO()):
        occ = SconeOcc()
    sd = synth.seeded_state_dict(occ.state_dict(), seed)
    occ.load_state_dict(sd)
    return occ.to(dev).eval(), sd


def test_forward_cells_equals_per_cell_forwards(cuda_device):
    """One ragged launch sequence over 7 cells of different cloud / query sizes (incl. a cloud above seq_len, one with
    65 points -> down-sampling factor 0 -> 2, and a cell with a single query) == the reference's call pattern, one
    forward per cell, with the same RNG draws."""
    occ, sd = _occ(cuda_device)
    gen = torch.Generator().manual_seed(12)
    sizes = [(700, 150), (65, 40), (2500, 333), (130, 1), (2048, 64), (300, 500), (1000, 17)]
    clouds = [(torch.rand(n, 3, generator=gen) - 0.5).to(cuda_device) for n, _ in sizes]
    queries = [(torch.rand(q, 3, generator=gen) - 0.5).to(cuda_device) for _, q in sizes]
    vhs = [(0.3 * torch.randn(q, 64, generator=gen)).to(cuda_device) for _, q in sizes]
    with torch.no_grad():
        torch.manual_seed(77)
        n0 = ops.launch_count()
        batched = occ.forward_cells(clouds, queries, vhs)
        n_batched = ops.launch_count() - n0
        torch.manual_seed(77)
        n0 = ops.launch_count()
        single = [occ(c[None], q[None], v[None])[0] for c, q, v in zip(clouds, queries, vhs)]
        n_single = ops.launch_count() - n0
    assert n_batched < n_single / 3
    for c, (a, b) in enumerate(zip(batched, single)):
        assert a.shape == b.shape == (sizes[c][1], 1)
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item()), c
    # and against the oracle (reference arithmetic) for two of the cells
    torch.manual_seed(77)
    with torch.no_grad():
        for c in range(2):
            want = o_nets.scone_occ_forward(sd, clouds[c].cpu()[None], queries[c].cpu()[None], vhs[c].cpu()[None])[0]
            err = (batched[c].cpu() - want).abs()
            assert np.quantile(err.numpy(), 0.98) <= NET_RTOL * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("name", ["scene_field_s31", "scene_field_s32"])
def test_scene_occupancy_field_matches_reference_golden(name, cuda_device):
    g = load_golden(name)
    seed = int(g["seed"])
    occ, sd = _occ(cuda_device, int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])
    surface_scene, proxy_scene = scene_case.build(scene.Scene, cuda_device, seed, n_proxy=int(g["n_proxy"]),
                                                  n_surface=int(g["n_surface"]))
    digest, counts = scene_case.scene_digest(surface_scene, proxy_scene)
    assert counts == g["cell_counts"].tolist()       # same cell populations as the reference's Scene on the CPU
    macarons = Macarons(None, occ, None)
    pred = scene_case.prediction_camera(cuda_device)
    torch.manual_seed(seed + 1000)
    n0 = ops.launch_count()
    with torch.no_grad():
        X_world, vh, probs = macarons_utils.compute_scene_occupancy_probability_field(
            scene_case.params(), macarons, None, surface_scene, proxy_scene, cuda_device, prediction_camera=pred)
    launches = ops.launch_count() - n0
    n, n_oof = int(g["n_points"]), int(g["n_out_of_field"])
    assert X_world.shape == (n, 3) and vh.shape == (n, 64) and probs.shape == (n, 1)
    assert np.array_equal(X_world[:64].cpu().numpy(), g["X_world_head"])
    # view harmonics: equal to 2e-6 except where a view-state bin sits on a boundary (<= 1e-3 of the points, see
    # test_viewstate_sampling_gpu.py)
    verr = np.abs(vh[:n - n_oof:8].cpu().numpy() - g["view_harmonics_stride8"]).max(axis=-1)
    assert (verr > 2e-6).mean() <= 2e-3
    scale = max(1.0, float(np.abs(g["occupancy"]).max()))
    err = np.abs(probs[:n - n_oof, 0].cpu().numpy() - g["occupancy"])
    print("%s: %d cells' queries in %d launches; occupancy err median %.2e q98 %.2e max %.2e (scale %.2f)"
          % (name, n - n_oof, launches, np.median(err), np.quantile(err, 0.98), err.max(), scale))
    assert np.quantile(err, 0.98) <= NET_RTOL * scale      # kNN rounding ties excepted (tests/test_nets_gpu.py)
    assert err.max() <= 0.05 * scale
    perr = np.abs(proxy_scene.proxy_proba[:, 0].cpu().numpy() - g["proxy_proba"])
    assert np.quantile(perr, 0.98) <= NET_RTOL * scale and perr.max() <= 0.05 * scale
    assert torch.all(vh[n - n_oof:] == 0) and torch.all(probs[n - n_oof:] == 0.5)
    assert launches < 400        # one ragged forward, not one per cell (the per-cell loop takes > 100 launches per cell)


def test_forward_cells_edge_cases(cuda_device):
    """A cell without queries, a single cell, and the empty list."""
    occ, _ = _occ(cuda_device)
    gen = torch.Generator().manual_seed(3)
    mk = lambda n, d: (torch.rand(n, d, generator=gen) - 0.5).to(cuda_device)
    assert occ.forward_cells([], [], []) == []
    with torch.no_grad():
        torch.manual_seed(5)
        outs = occ.forward_cells([mk(200, 3), mk(90, 3), mk(300, 3)], [mk(50, 3), mk(0, 3), mk(7, 3)],
                                 [mk(50, 64), mk(0, 64), mk(7, 64)])
    assert [tuple(o.shape) for o in outs] == [(50, 1), (0, 1), (7, 1)]
    assert all(torch.isfinite(o).all() for o in outs)
    with torch.no_grad():
        torch.manual_seed(6)
        one = occ.forward_cells([mk(500, 3)], [mk(33, 3)], [mk(33, 64)])
    assert one[0].shape == (33, 1)


def test_scene_field_without_observations(cuda_device):
    """Before any frame: every proxy point is out of field -> the field is the stored default probability with zero
    harmonics, and no network call is made (reference :1520-1538)."""
    occ, _ = _occ(cuda_device)
    x_min, x_max = torch.tensor([-1., -1., -1.]).to(cuda_device), torch.tensor([1., 1., 1.]).to(cuda_device)
    common = dict(x_min=x_min, x_max=x_max, grid_l=2, grid_w=2, grid_h=2, n_proxy_points=500, device=cuda_device)
    surface_scene = scene.Scene(cell_capacity=100, cell_resolution=None, feature_dim=1, **common)
    proxy_scene = scene.Scene(cell_capacity=1000, cell_resolution=0.001, feature_dim=1, **common)
    proxy_scene.initialize_proxy_points()
    n0 = ops.launch_count()
    X_world, vh, probs = macarons_utils.compute_scene_occupancy_probability_field(
        scene_case.params(), Macarons(None, occ, None), None, surface_scene, proxy_scene, cuda_device,
        prediction_camera=scene_case.prediction_camera(cuda_device))
    assert ops.launch_count() == n0
    assert X_world.shape == (500, 3) and torch.equal(X_world, proxy_scene.proxy_points)
    assert torch.all(vh == 0) and torch.all(probs == 0.5)
