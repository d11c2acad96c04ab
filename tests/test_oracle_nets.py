"""CPU: the network oracle (oracle/scone_nets.py) against the fixtures generated from the unmodified reference
(tests/golden/make_golden.py: SconeVis.forward, SconeOcc.forward, get_knn_points, compute_occupancy_probability)."""
import numpy as np
import pytest
import torch

import synth
from conftest import load_golden
from oracle import scone_nets as o_nets


def _sd(kind, g):
    # the template only supplies names and shapes; the values come from synth.seeded_state_dict
    from macarons_b200.networks.SconeOcc import SconeOcc
    from macarons_b200.networks.SconeVis import SconeVis
    module = SconeOcc() if kind == "occ" else SconeVis()
    sd = synth.seeded_state_dict(module.state_dict(), int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"]), "seeded weights differ from the golden run"
    return sd


@pytest.mark.parametrize("name", ["sconevis_small", "sconevis_2048"])
def test_sconevis_oracle_matches_reference_golden(name):
    g = load_golden(name)
    pts, vh = synth.sconevis_inputs(int(g["B"]), int(g["S"]), int(g["seed"]))
    with torch.no_grad():
        out = o_nets.scone_vis_forward(_sd("vis", g), pts, vh)[:, ::int(g["row_stride"])]
    assert out.shape == g["harmonics"].shape
    # same arithmetic; different machines / thread counts may reorder fp32 sums inside torch's GEMMs
    assert np.abs(out.numpy() - g["harmonics"]).max() <= 2e-5 * np.abs(g["harmonics"]).max()


@pytest.mark.parametrize("name", ["sconeocc_small", "sconeocc_cfg1"])
def test_sconeocc_oracle_matches_reference_golden(name):
    g = load_golden(name)
    pc, x, vh = synth.sconeocc_inputs(int(g["B"]), int(g["N"]), int(g["Q"]), int(g["seed"]), grid=bool(g["grid"]))
    sd = _sd("occ", g)                  # (building the template module consumes the global RNG: do it first)
    torch.manual_seed(int(g["seed"]))
    with torch.no_grad():
        out = o_nets.scone_occ_forward(sd, pc, x, vh)
    assert out.shape == g["occupancy"].shape
    assert np.abs(out.numpy() - g["occupancy"]).max() <= 2e-5 * max(1.0, np.abs(g["occupancy"]).max())
    _, dist, idx = o_nets.knn_points(x, pc, 16)
    if int(g["Q"]) > 512:
        dist, idx = dist[:, ::16], idx[:, ::16]
    assert np.abs(dist.numpy() - g["knn_dist"]).max() <= 1e-6
    assert (idx.numpy() != g["knn_idx"]).mean() <= 1e-3   # only exact-tie reorderings may differ


def test_chunked_occupancy_oracle_matches_reference_golden():
    g = load_golden("sconeocc_chunked")
    pc, x, vh = synth.sconeocc_inputs(1, int(g["N"]), int(g["Q"]), int(g["seed"]))
    sd = _sd("occ", g)
    torch.manual_seed(int(g["seed"]))
    with torch.no_grad():
        out = o_nets.compute_occupancy_probability(sd, pc, x, vh,
                                                   max_points_per_pass=int(g["max_points_per_pass"]))
    assert np.abs(out.numpy() - g["occupancy"]).max() <= 2e-5 * max(1.0, np.abs(g["occupancy"]).max())


def test_subsample_helper_draws_like_the_forward():
    torch.manual_seed(7)
    gi, si = o_nets.scone_occ_subsamples(700)
    torch.manual_seed(7)
    assert torch.equal(gi, torch.randperm(700)[:2048])
    assert [len(s) for s in si] == [350, 175]
