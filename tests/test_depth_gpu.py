"""GPU parity of the fused ManyDepth forward (row a14): CUDA path through the C ABI vs the oracle on identical seeded
frames / poses / weights, and vs the fixtures generated from the reference's own ManyDepth code.

Stated tolerance: disparities are sigmoid outputs in (0, 1); |cuda - reference| <= DISP_ATOL.  The CUDA path folds the
(eval-mode) BatchNorm into the convolution weights and evaluates the camera chain in closed form instead of through a
numerically inverted 4x4 matrix, so it is an fp32 implementation with a different rounding sequence."""
import numpy as np
import pytest
import torch

import synth
from conftest import load_golden
from macarons_b200 import ops
from macarons_b200.networks import ManyDepth as MD
from macarons_b200.networks.Macarons import Macarons
from oracle import depth as o_depth

pytestmark = pytest.mark.gpu

DISP_ATOL = 5e-4      # observed: max 2.4e-4, median ~1e-5 through ~30 convolution layers with random Kaiming weights
DISP_MEDIAN_ATOL = 5e-5


def _model(H, W, g, dev):
    resnet = MD.ResNet18Trunk()
    model = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet, input_height=H, input_width=W), None)
    sd = synth.seeded_state_dict(model.state_dict(), int(g["weight_seed"]))
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])
    model.load_state_dict(sd)
    return model.to(dev).eval(), sd


@pytest.mark.parametrize("name", ["depth_64x96", "depth_96x160_b2"])
def test_many_depth_forward(name, cuda_device):
    g = load_golden(name)
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    model, sd = _model(H, W, g, cuda_device)
    x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(B, H, W, int(g["seed"]))
    to = lambda t: t.to(cuda_device)
    n0 = ops.launch_count()
    with torch.no_grad():
        pose, d1, d2, d3, d4 = model(to(x), to(x_alpha), to(R), to(T), to(zfar), cuda_device, gt_pose=to(gt_pose))
        want = o_depth.many_depth_forward(sd, x, x_alpha, R, T, zfar, gt_pose)
    assert ops.launch_count() - n0 > 40, "the CUDA kernels did not run"
    assert torch.equal(pose.cpu(), gt_pose)
    for got, ref, key in zip((d1, d2, d3, d4), want[1:], ("disp1", "disp2", "disp3", "disp4")):
        got = got.cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() <= DISP_ATOL, key
        assert (got - ref).abs().median().item() <= DISP_MEDIAN_ATOL, key
        gold = g[key]
        s = int(g["row_stride"]) if key == "disp1" else 1
        assert np.abs(got.numpy()[..., ::s, ::s] - gold).max() <= DISP_ATOL, key
    # depth = 1 / (a disp + b)  (utility/depth_model_utils.py:844-848) keeps the ordering of disparities
    depth = o_depth.depth_from_disparity(d1.cpu())
    assert torch.all(depth >= 0.5 - 1e-4) and torch.all(depth <= 750.0 + 1e-2)


def test_macarons_depth_mode_and_argument_checks(cuda_device):
    g = load_golden("depth_64x96")
    model, _ = _model(64, 96, g, cuda_device)
    mac = Macarons(model, None, None)
    x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(1, 64, 96, int(g["seed"]))
    to = lambda t: t.to(cuda_device)
    with torch.no_grad():
        out = mac(mode='depth', x=to(x), x_alpha=to(x_alpha), R=to(R), T=to(T), zfar=to(zfar), device=cuda_device,
                  gt_pose=to(gt_pose))
    assert np.abs(out[1].cpu().numpy() - g["disp1"]).max() <= DISP_ATOL
    with pytest.raises(NameError):
        mac(mode='depth', x=to(x))
    with pytest.raises(NameError):
        model(to(x), to(x_alpha), to(R), to(T), to(zfar), cuda_device, gt_pose=None)
    with pytest.raises(NameError):
        model(to(x), to(x_alpha), to(R), to(T), to(zfar) * 0 + 100.0, cuda_device, gt_pose=to(gt_pose))


def test_full_resolution_shapes(cuda_device):
    """256 x 456 (reference default): output shapes of the 4 scales, finite values, repeatable."""
    resnet = MD.ResNet18Trunk()
    model = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet), None)
    model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 5))
    model = model.to(cuda_device).eval()
    x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(1, 256, 456, 7)
    to = lambda t: t.to(cuda_device)
    with torch.no_grad():
        a = model(to(x), to(x_alpha), to(R), to(T), to(zfar), cuda_device, gt_pose=to(gt_pose))
        b = model(to(x), to(x_alpha), to(R), to(T), to(zfar), cuda_device, gt_pose=to(gt_pose))
    assert [tuple(t.shape) for t in a[1:]] == [(1, 1, 256, 456), (1, 1, 128, 228), (1, 1, 64, 114), (1, 1, 32, 57)]
    for u, v in zip(a[1:], b[1:]):
        assert torch.isfinite(u).all() and torch.equal(u, v)
