"""CPU: the Scene / Cell bookkeeping of macarons_b200.utility.scene against the state the reference's own classes reach
on the same seeded three-frame scenario (tests/golden/scene_field_*.npz, tests/golden/make_golden.py --scene-only), and the
scene-field oracle (oracle/scene.py) against the reference's outputs.  The binning kernel behind
`update_proxy_view_states` needs a GPU, so here the oracle's CPU restatement of compute_view_state stands in for it."""
import numpy as np
import pytest
import torch

import scene_case
import synth
from conftest import load_golden
from oracle import scene as o_scene
from oracle import view_state as o_vs


@pytest.fixture()
def cpu_scene_module(monkeypatch):
    from macarons_b200.utility import scene
    monkeypatch.setattr(scene, "compute_view_state", o_vs.view_state)
    return scene


@pytest.mark.parametrize("name", ["scene_field_s31", "scene_field_s32"])
def test_scene_bookkeeping_and_field_oracle_match_reference_golden(name, cpu_scene_module):
    g = load_golden(name)
    seed = int(g["seed"])
    surface_scene, proxy_scene = scene_case.build(cpu_scene_module.Scene, "cpu", seed, n_proxy=int(g["n_proxy"]),
                                                  n_surface=int(g["n_surface"]))
    digest, counts = scene_case.scene_digest(surface_scene, proxy_scene)
    assert counts == g["cell_counts"].tolist()
    assert digest == str(g["scene_digest"])          # every cell's points / features and every proxy state tensor
    from macarons_b200.networks.SconeOcc import SconeOcc
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        sd = synth.seeded_state_dict(SconeOcc().state_dict(), int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])
    torch.manual_seed(seed + 1000)
    with torch.no_grad():
        X_world, vh, probs = o_scene.scene_occupancy_field(scene_case.params(), sd, surface_scene, proxy_scene,
                                                           scene_case.prediction_camera())
    n, n_oof = int(g["n_points"]), int(g["n_out_of_field"])
    assert X_world.shape == (n, 3) and vh.shape == (n, 64) and probs.shape == (n, 1)
    assert np.array_equal(X_world[:64].numpy(), g["X_world_head"])
    assert np.abs(vh[:n - n_oof:8].numpy() - g["view_harmonics_stride8"]).max() <= 2e-6
    assert np.abs(probs[:n - n_oof, 0].numpy() - g["occupancy"]).max() <= 2e-5
    assert np.abs(proxy_scene.proxy_proba[:, 0].numpy() - g["proxy_proba"]).max() <= 2e-5
    assert torch.all(vh[n - n_oof:] == 0) and torch.all(probs[n - n_oof:] == 0.5)


def test_cell_capacity_resolution_rules(cpu_scene_module):
    Scene = cpu_scene_module.Scene
    x_min, x_max = torch.tensor([0., 0., 0.]), torch.tensor([4., 2., 6.])
    s = Scene(x_min, x_max, 2, 1, 3, cell_capacity=50, cell_resolution=None, n_proxy_points=100, device="cpu")
    assert len(s.cells) == 6 and s.cell_capacity == 50 and s.cell_resolution > 0
    c = s.cells["[1, 0, 2]"]
    assert torch.allclose(c.center, torch.tensor([3., 1., 5.])) and torch.allclose(c.x_min, torch.tensor([[2., 0., 4.]]))
    s2 = Scene(x_min, x_max, 2, 1, 3, cell_capacity=None, cell_resolution=0.25, n_proxy_points=100, device="cpu")
    assert s2.cell_capacity == int((2 * np.sqrt(8.)) // (np.pi * 0.125 ** 2))
    with pytest.raises(NameError):
        Scene(x_min, x_max, 1, 1, 1, cell_capacity=None, cell_resolution=None, n_proxy_points=10, device="cpu")
    # points on the border of the grid are clamped to the last cell; neighbours are clamped and unique
    pts = torch.tensor([[4., 2., 6.], [0., 0., 0.], [2.5, 1.0, 3.9]])
    assert s.get_cells_for_each_pt(pts).tolist() == [[1, 0, 2], [0, 0, 0], [1, 0, 1]]
    assert s.get_neighboring_cells(torch.tensor([0, 0, 0])).tolist() == [[0, 0, 0], [0, 0, 1], [1, 0, 0], [1, 0, 1]]
    # capacity: a cell never holds more than `capacity` points, new points closer than `resolution` are dropped
    torch.manual_seed(0)
    cloud = torch.rand(500, 3) * torch.tensor([2., 2., 2.])
    s.fill_cells(cloud)
    assert s.cells["[0, 0, 0]"].cell_pts.shape[0] == 50
    before = s.cells["[0, 0, 0]"].cell_pts.clone()
    s.fill_cells(before + 1e-6)          # all within the resolution of stored points: only the random re-draw happens
    assert s.cells["[0, 0, 0]"].cell_pts.shape[0] == 50
