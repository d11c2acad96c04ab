"""GPU: backward of the coverage-gain / visibility-gain integration with respect to the harmonics (SURVEY.md section 8f
rank 3) -- mac_covgain_backward_f32 behind torch.autograd -- against autograd through the oracle's restatement of the
reference arithmetic (oracle/sh_cov.py, torch CPU) and against a float64 analytic gradient.

Tolerance: gradients are sums over C cameras of g * act'(z) * Y_k with |Y_k| <= ~1.5; stated bound
|cuda - float64| <= GRAD_RTOL * max|grad| per tensor (observed ~1e-6)."""
import numpy as np
import pytest
import torch

import synth
from macarons_b200 import ops
from macarons_b200.networks.Macarons import Macarons
from macarons_b200.networks.SconeVis import SconeVis
from oracle import sh_cov

pytestmark = pytest.mark.gpu

GRAD_RTOL = 2e-5


def _f64_grad(pts, harm, cams, g, use_sigmoid, per_point):
    """Analytic float64 gradient: dH[b,p,:] = sum_c w[b,c,p] * Y(cam_c - pt_p)."""
    pts64, H, cam = pts.double().numpy()[..., :3], harm.double().numpy(), cams.double().numpy()
    B, P, C = pts64.shape[0], pts64.shape[1], cam.shape[1]
    out = np.zeros((B, P, 64))
    for c in range(C):
        Y = sh_cov.sh_basis_closed_form_f64(cam[:, c, None, :] - pts64)        # (B,P,64)
        z = (Y * H).sum(-1)
        if use_sigmoid:
            v = 1.0 / (1.0 + np.exp(-z))
            da = v * (1 - v)
        else:
            da = (z > 0).astype(np.float64)
        w = g.double().numpy()[:, c, :] if per_point else g.double().numpy()[:, c, None] / P
        out += (w * da)[..., None] * Y
    return out


@pytest.mark.parametrize("B,P,C,D,sig,per_point", [(2, 300, 7, 4, True, False), (1, 257, 33, 3, False, False),
                                                    (1, 2048, 52, 4, True, False), (2, 130, 5, 4, True, True),
                                                    (1, 77, 600, 4, True, False), (1, 64, 3, 4, False, True)])
def test_backward_matches_oracle_autograd_and_float64(B, P, C, D, sig, per_point, cuda_device):
    pts, harm, cams = synth.covgain_inputs(B, P, C, seed=B * 100 + P + C, pts_dim=D)
    gen = torch.Generator().manual_seed(3)
    g = torch.randn((B, C, P) if per_point else (B, C), generator=gen)
    h_dev = harm.to(cuda_device).requires_grad_(True)
    out = ops.sh_integration(pts.to(cuda_device), h_dev, cams.to(cuda_device), use_sigmoid=sig, per_point=per_point)
    assert out.requires_grad and out.grad_fn is not None
    n0 = ops.launch_count()
    out.backward(g.to(cuda_device))
    assert ops.launch_count() == n0 + 1
    got = h_dev.grad.cpu()
    truth = _f64_grad(pts, harm, cams, g, sig, per_point)
    scale = np.abs(truth).max()
    assert np.abs(got.numpy() - truth).max() <= GRAD_RTOL * scale
    # autograd through the reference's fp32 arithmetic (same torch ops as the reference, CPU)
    h_cpu = harm.clone().requires_grad_(True)
    fn = sh_cov.visibility_gains if per_point else sh_cov.coverage_gain
    fn(pts, h_cpu, cams, use_sigmoid=sig).backward(g)
    ref = h_cpu.grad
    if sig:
        # away from ill-conditioned rays the reference's own fp32 gradient is within 1e-4 of float64
        err = (got - ref).abs()
        assert np.quantile(err.numpy(), 0.999) <= 1e-4 * scale and err.max().item() <= 2e-2 * scale
    else:
        # relu'(z) flips where |z| is at rounding level: compare where the reference's z is clearly non-zero
        z64 = sh_cov.visibility_gains_f64(pts.numpy(), harm.numpy(), cams.numpy(), use_sigmoid=False)
        assert np.quantile((got - ref).abs().numpy(), 0.99) <= 1e-4 * scale


def test_module_methods_are_differentiable(cuda_device):
    """The three SH-integration methods of the mirrored classes give gradients w.r.t. the harmonics; the loss of
    trainers/pretrain_scone_vis.py:186 (MSE on the coverage) back-propagates through compute_coverage_gain."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        vis = SconeVis().to(cuda_device)
    pts, harm, cams = synth.covgain_inputs(1, 512, 20, seed=5)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    h = d[1].clone().requires_grad_(True)
    cov = vis.compute_coverage_gain(d[0], h, d[2])
    loss = torch.nn.functional.mse_loss(cov, torch.full_like(cov, 0.3))
    loss.backward()
    assert h.grad is not None and torch.isfinite(h.grad).all() and h.grad.abs().max() > 0
    # finite-difference check of the loss along a random direction (float32 forward: coarse step, loose tolerance)
    v = torch.randn_like(h)
    eps = 1e-2
    with torch.no_grad():
        lp = torch.nn.functional.mse_loss(vis.compute_coverage_gain(d[0], h + eps * v, d[2]), torch.full_like(cov, 0.3))
        lm = torch.nn.functional.mse_loss(vis.compute_coverage_gain(d[0], h - eps * v, d[2]), torch.full_like(cov, 0.3))
    fd = ((lp - lm) / (2 * eps)).item()
    an = (h.grad * v).sum().item()
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1e-6) + 1e-7
    # per-point methods
    h2 = d[1].clone().requires_grad_(True)
    vis.compute_visibilities(d[0], h2, d[2]).sum().backward()
    h3 = d[1].clone().requires_grad_(True)
    Macarons(None, None, vis).compute_visibility_gains(d[0], h3, d[2]).sum().backward()
    assert torch.equal(h2.grad, h3.grad)
    # sum of the per-point outputs == P * coverage: the two backward modes agree
    h4 = d[1].clone().requires_grad_(True)
    (vis.compute_coverage_gain(d[0], h4, d[2]).sum() * pts.shape[1]).backward()
    assert (h4.grad - h2.grad).abs().max().item() <= 1e-5 * h2.grad.abs().max().item()
    # no-grad calls and frozen inputs keep the plain inference path
    with torch.no_grad():
        assert vis.compute_coverage_gain(d[0], h, d[2]).grad_fn is None
    assert vis.compute_coverage_gain(*d).grad_fn is None


def _oracle_param_grads(fn, sd, loss_of):
    """Gradients of loss_of(fn(sd_with_grad)) w.r.t. every entry of the state dict, through the oracle's restatement of the
    reference arithmetic (torch CPU autograd)."""
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    loss = loss_of(fn(sd))
    loss.backward()
    return loss.item(), {k: v.grad for k, v in sd.items() if v.grad is not None}


def test_sconevis_training_step_matches_oracle_autograd(cuda_device):
    """One iteration of trainers/pretrain_scone_vis.py:162-225 in miniature: SconeVis.forward -> compute_coverage_gain ->
    MSE loss -> backward -> Adam step, on the mirrored classes; loss and every parameter gradient against autograd through
    the oracle (reference arithmetic, CPU)."""
    import contextlib
    import io
    from oracle import scone_nets as o_nets
    with contextlib.redirect_stdout(io.StringIO()):
        vis = SconeVis()
    sd = synth.seeded_state_dict(vis.state_dict(), 5)
    vis.load_state_dict(sd)
    vis = vis.to(cuda_device).train()
    pts, vh = synth.sconevis_inputs(1, 384, 41)
    cams = synth.fibonacci_cameras(24)[None].contiguous()
    target = torch.rand(1, 24, generator=torch.Generator().manual_seed(1))
    opt = torch.optim.Adam(vis.parameters(), lr=1e-4)
    harm = vis(pts.to(cuda_device), view_harmonics=vh.to(cuda_device))
    assert harm.grad_fn is not None
    cov = vis.compute_coverage_gain(pts.to(cuda_device), harm, cams.to(cuda_device))
    loss = torch.nn.functional.mse_loss(cov, target.to(cuda_device))
    opt.zero_grad()
    loss.backward()

    def fwd(sd_g):
        h = o_nets.scone_vis_forward(sd_g, pts, vh)
        return sh_cov.coverage_gain(pts, h, cams)
    want_loss, want = _oracle_param_grads(fwd, sd, lambda c: torch.nn.functional.mse_loss(c, target))
    assert abs(loss.item() - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    got = {k: p.grad.cpu() for k, p in vis.named_parameters()}
    assert set(got) == set(want)
    worst = 0.0
    overall = max(v.abs().max().item() for v in want.values())
    for k in want:
        scale = want[k].abs().max().item()
        err = (got[k] - want[k]).abs().max().item()
        worst = max(worst, err / max(scale, 1e-6 * overall))
        assert err <= 2e-3 * scale + 1e-6 * overall, (k, err, scale)
    print("SconeVis training step: loss %.6f (oracle %.6f); worst relative gradient error %.2e over %d tensors"
          % (loss.item(), want_loss, worst, len(want)))
    before = {k: p.detach().clone() for k, p in vis.named_parameters()}
    opt.step()
    assert any(not torch.equal(before[k], p.detach()) for k, p in vis.named_parameters())
    # the packed weights follow the optimizer step (version counters): the next forward uses the new parameters
    with torch.no_grad():
        new = vis(pts.to(cuda_device), view_harmonics=vh.to(cuda_device)).cpu()
        ref = o_nets.scone_vis_forward({k: v.detach().cpu() for k, v in vis.state_dict().items()}, pts, vh)
    assert (new - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert not torch.equal(new, harm.detach().cpu())


def test_sconeocc_gradients_match_oracle_autograd(cuda_device):
    """trainers/pretrain_scone_occ.py:158 in miniature: SconeOcc.forward -> MSE on the occupancy -> backward."""
    import contextlib
    import io
    from macarons_b200.networks.SconeOcc import SconeOcc
    from oracle import scone_nets as o_nets
    with contextlib.redirect_stdout(io.StringIO()):
        occ = SconeOcc()
    sd = synth.seeded_state_dict(occ.state_dict(), 5)
    occ.load_state_dict(sd)
    occ = occ.to(cuda_device).train()
    pc, x, vh = synth.sconeocc_inputs(1, 700, 160, 77)
    target = torch.rand(1, 160, 1, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(9)
    out = occ(pc.to(cuda_device), x.to(cuda_device), vh.to(cuda_device))
    assert out.shape == (1, 160, 1) and out.grad_fn is not None
    loss = torch.nn.functional.mse_loss(out, target.to(cuda_device))
    loss.backward()
    torch.manual_seed(9)
    want_loss, want = _oracle_param_grads(lambda sd_g: o_nets.scone_occ_forward(sd_g, pc, x, vh), sd,
                                          lambda o: torch.nn.functional.mse_loss(o, target))
    assert abs(loss.item() - want_loss) <= 1e-4 * max(1.0, abs(want_loss))
    got = {k: p.grad.cpu() for k, p in occ.named_parameters() if p.grad is not None}
    assert set(got) == set(want)
    worst = 0.0
    overall = max(v.abs().max().item() for v in want.values())
    for k in want:
        scale = want[k].abs().max().item()
        err = (got[k] - want[k]).abs().max().item()
        worst = max(worst, err / max(scale, 1e-6 * overall))
        # (key biases have a mathematically zero gradient -- softmax is shift invariant -- hence the absolute floor;
        # a neighbour swapped at a kNN rounding tie moves a few entries)
        assert err <= 2e-2 * scale + 1e-6 * overall, (k, err, scale)
    print("SconeOcc backward: loss %.6f (oracle %.6f); worst relative gradient error %.2e over %d tensors"
          % (loss.item(), want_loss, worst, len(want)))


def test_gradients_that_are_not_implemented_are_refused(cuda_device):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        vis = SconeVis().to(cuda_device).eval()
    pts, harm, cams = synth.covgain_inputs(1, 64, 4, seed=6)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    with pytest.raises(NotImplementedError):
        vis.compute_coverage_gain(d[0].clone().requires_grad_(True), d[1], d[2])
    with pytest.raises(NotImplementedError):
        vis.compute_coverage_gain(d[0], d[1], d[2].clone().requires_grad_(True))
    vh = torch.zeros(1, 64, 64, device=cuda_device)
    assert vis(d[0], view_harmonics=vh).grad_fn is not None   # parameters require grad: the training path
    with torch.no_grad():
        assert vis(d[0], view_harmonics=vh).shape == (1, 64, 64)
    vis.requires_grad_(False)
    assert vis(d[0], view_harmonics=vh).grad_fn is None    # frozen module: inference path
    with contextlib.redirect_stdout(io.StringIO()):
        drop = SconeVis(dropout=0.1).to(cuda_device).train()
    with pytest.raises(NotImplementedError):
        drop(d[0], view_harmonics=vh)                      # dropout is not implemented in the fused forward


def test_backward_edge_shapes(cuda_device):
    """One point, one camera, point counts that are not a multiple of the block, more cameras than one staging pass."""
    for B, P, C in ((1, 1, 1), (3, 129, 2), (1, 40, 1100)):
        pts, harm, cams = synth.covgain_inputs(B, P, C, seed=P + C)
        g = torch.randn(B, C, generator=torch.Generator().manual_seed(1))
        got = ops.coverage_gain_backward(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device), g.to(cuda_device)).cpu()
        truth = _f64_grad(pts, harm, cams, g, True, False)
        assert np.abs(got.numpy() - truth).max() <= GRAD_RTOL * np.abs(truth).max()
    with pytest.raises(ValueError):
        ops.coverage_gain_backward(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device), g.to(cuda_device)[:, :5])
