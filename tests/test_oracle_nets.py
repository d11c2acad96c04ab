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


# ---- depth (row a14) ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["depth_64x96", "depth_96x160_b2"])
def test_depth_oracle_matches_reference_golden(name):
    from macarons_b200.networks import ManyDepth as MD
    from oracle import depth as o_depth
    g = load_golden(name)
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    resnet = MD.ResNet18Trunk()
    model = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet, input_height=H, input_width=W), None)
    sd = synth.seeded_state_dict(model.state_dict(), int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])      # same keys / shapes as the reference model
    x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(B, H, W, int(g["seed"]))
    with torch.no_grad():
        out = o_depth.many_depth_forward(sd, x, x_alpha, R, T, zfar, gt_pose)
    s = int(g["row_stride"])
    assert np.abs(out[1].numpy()[..., ::s, ::s] - g["disp1"]).max() <= 2e-5
    for k, key in enumerate(("disp2", "disp3", "disp4")):
        assert np.abs(out[2 + k].numpy() - g[key]).max() <= 2e-5


def test_batchnorm_folding_and_transposed_conv_packing_are_exact():
    """netpack: conv + eval BatchNorm == folded conv; ConvTranspose2d(k3,s1,p1) == conv with the flipped kernel, in the
    (ky, kx, c) im2col order the CUDA gather uses."""
    import torch.nn.functional as F
    from macarons_b200 import netpack
    gen = torch.Generator().manual_seed(0)
    conv = torch.nn.Conv2d(5, 7, 3, 2, 1, bias=False)
    bn = torch.nn.BatchNorm2d(7).eval()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen))
        bn.weight.copy_(torch.rand(7, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(7, generator=gen))
        bn.running_mean.copy_(torch.randn(7, generator=gen))
        bn.running_var.copy_(torch.rand(7, generator=gen) + 0.5)
    x = torch.randn(2, 5, 9, 11, generator=gen)
    w, b = netpack._fold_bn(conv.weight, None, bn)
    with torch.no_grad():
        want = bn(conv(x))
        got = F.conv2d(x.double(), w, b, 2, 1)
    assert (got - want.double()).abs().max().item() <= 1e-5

    up = torch.nn.ConvTranspose2d(4, 6, 3, 1, 1)
    dst = netpack.ConvW()
    pk = netpack._Packer()
    netpack._pack_conv(pk, dst, up, None, netpack.ACT_ELU, transposed=True)
    assert (dst.k, dst.stride, dst.pad, dst.reflect, dst.act) == (3, 1, 1, 0, netpack.ACT_ELU)
    assert (dst.lin.N, dst.lin.K) == (6, 36)
    mat = (pk.keep[-1].hi + pk.keep[-1].lo)[:, :36]                     # (Cout, ky, kx, c)
    wconv = mat.view(6, 3, 3, 4).permute(0, 3, 1, 2)
    xin = torch.randn(1, 4, 8, 8, generator=gen)
    with torch.no_grad():
        assert (F.conv2d(xin, wconv, up.bias, 1, 1) - up(xin)).abs().max().item() <= 1e-5
