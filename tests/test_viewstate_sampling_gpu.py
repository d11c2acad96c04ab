"""GPU parity of rows a10-a13: view-state binning, histogram -> SH projection, bin permutation and proxy sampling,
CUDA path through the C ABI vs the oracle and the fixtures generated from the unmodified reference."""
import math

import numpy as np
import pytest
import torch

import synth
from conftest import load_golden
from macarons_b200 import ops
from macarons_b200.utility import scone_utils
from oracle import sampling as o_sampling
from oracle import view_state as o_vs

pytestmark = pytest.mark.gpu


def _boundary_margin(pts, X_view, n_elev, n_azim):
    """(B,P) smallest float64 distance (rad) of any ray's elevation / azimuth to a bin boundary (half-way between
    bin centres), i.e. how far the point is from a place where fp32 rounding may legitimately flip a bin."""
    d = X_view.double().view(1, 1, -1, 3) - pts[..., :3].double().unsqueeze(2)
    r = d.norm(dim=-1)
    elev = torch.asin((d[..., 1] / r).clamp(-1, 1))
    azim = torch.atan2(d[..., 0], d[..., 2])
    es, az = math.pi / (n_elev + 1), 2 * math.pi / n_azim
    me = ((elev / es) % 1.0 - 0.5).abs() * es
    ma = ((azim / az) % 1.0 - 0.5).abs() * az
    return torch.minimum(me, ma).min(dim=-1)[0]


@pytest.mark.parametrize("name", ["view_state_small", "view_state_10views"])
def test_view_state_and_harmonics(name, cuda_device):
    g = load_golden(name)
    B, P, V = int(g["B"]), int(g["P"]), int(g["V"])
    pts, _ = synth.view_state_inputs(B, P, V, int(g["seed"]))
    X_view = torch.from_numpy(g["X_view"])
    want = torch.from_numpy(np.unpackbits(g["state_bits"], axis=-1)[..., :98].astype(np.float32))
    n0 = ops.launch_count()
    got = scone_utils.compute_view_state(pts.to(cuda_device), X_view.to(cuda_device), 7, 14)
    assert ops.launch_count() == n0 + 1
    got_c = got.cpu()
    assert got_c.shape == (B, P, 98) and set(got_c.unique().tolist()) <= {0.0, 1.0}
    # identical bins wherever no ray sits within 1e-5 rad of a bin boundary (fp32 asin / acos differ by an ulp
    # between libm and CUDA); elsewhere at most a neighbouring bin
    safe = _boundary_margin(pts, X_view, 7, 14) > 1e-5
    assert safe.float().mean().item() > 0.99
    assert torch.equal(got_c[safe], want[safe])
    assert (got_c != want).any(-1).float().mean().item() <= 1e-3
    # a11: histogram -> SH coordinates
    base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, cuda_device)
    gb = load_golden("view_base_harmonics")
    assert np.abs(base.cpu().numpy() - gb["base"]).max() <= 2e-6
    vh = scone_utils.compute_view_harmonics(got, base, h_polar, h_azim, 7, 14).cpu()
    same = ~(got_c != want).any(-1)
    assert np.abs(vh.numpy() - g["view_harmonics"])[same.numpy()].max() <= 2e-6
    # general (non 0/1) states, e.g. the accumulated states of Scene.update_proxy_view_states
    acc = torch.rand(B, P, 98, generator=torch.Generator().manual_seed(3))
    ob, ohp, _ = o_vs.bin_centre_harmonics(8, 7, 14)
    want_acc = o_vs.view_harmonics(acc, ob, ohp, 7, 14)
    got_acc = scone_utils.compute_view_harmonics(acc.to(cuda_device), base, h_polar, h_azim, 7, 14).cpu()
    assert (got_acc - want_acc).abs().max().item() <= 1e-5


def test_pole_cameras_wrap_like_the_reference(cuda_device):
    """A camera straight above lands in bins 0-13, straight below in 84-97."""
    pts = torch.zeros(1, 4, 3)
    pts[0, :, 0] = torch.tensor([0.01, -0.02, 0.03, 0.0])
    pts[0, :, 2] = torch.tensor([0.02, 0.01, -0.03, 0.01])
    for y, lo, hi in ((1.5, 0, 14), (-1.5, 84, 98)):
        X_view = torch.tensor([[0.0, y, 0.0]])
        got = scone_utils.compute_view_state(pts.to(cuda_device), X_view.to(cuda_device), 7, 14).cpu()
        want = o_vs.view_state(pts, X_view, 7, 14)
        assert torch.equal(got, want)
        assert torch.all(got[..., lo:hi].sum(-1) == 1)


def test_move_view_state_to_view_space(cuda_device):
    class Cam:
        pass
    state = torch.rand(2, 300, 98, generator=torch.Generator().manual_seed(1))
    k = 3                                     # rotate about +y by k azimuth bins
    a = 2 * math.pi * k / 14
    cam = Cam()
    cam.R = torch.tensor([[[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]]])
    idx = scone_utils.view_space_bin_permutation(cam.R[0], 7, 14)
    i, j = torch.arange(98) // 14, torch.arange(98) % 14
    assert torch.equal(idx // 14, i)
    shift = (idx % 14 - j) % 14
    # all bins move by k azimuth steps, up to the reference's own fp32 rounding for directions exactly on a boundary
    assert (shift == shift.mode()[0]).float().mean() > 0.9 and int(shift.mode()[0]) in (k, 14 - k)
    got = scone_utils.move_view_state_to_view_space(state.to(cuda_device), cam, 7, 14).cpu()
    assert torch.equal(got, state[..., idx])
    cam.R = torch.eye(3).view(1, 3, 3)
    assert torch.equal(scone_utils.move_view_state_to_view_space(state.to(cuda_device), cam, 7, 14).cpu(), state)


def test_move_view_state_matches_reference_golden(cuda_device):
    """Row a12 against the output of the reference's own move_view_state_to_view_space (tests/golden/move_view_state.npz,
    24 camera poses on the stand-in pytorch3d camera) and against the oracle on a non-binary state."""
    from oracle import cameras as o_cams
    g = load_golden("move_view_state")
    aa, T = torch.from_numpy(g["axis_angle"]), torch.from_numpy(g["T"])
    state = torch.from_numpy(np.unpackbits(g["state_bits"], axis=-1)[..., :98].astype(np.float32))
    probe = torch.arange(98, dtype=torch.float32).view(1, 1, 98).to(cuda_device)
    dense = torch.rand(2, 77, 98, generator=torch.Generator().manual_seed(4))
    for i in range(aa.shape[0]):
        cam = o_cams.FoVPerspectiveCameras(R=o_cams.axis_angle_to_matrix(aa[i:i + 1]), T=T[i:i + 1], zfar=100.0)
        got = scone_utils.move_view_state_to_view_space(probe, cam, 7, 14).cpu().view(98)
        assert np.array_equal(got.numpy().astype(np.int64), g["indices"][i]), i
        if i % 6 == 0:
            want = o_vs.move_view_state_to_view_space(dense, cam, 7, 14)
            assert torch.equal(scone_utils.move_view_state_to_view_space(dense.to(cuda_device), cam, 7, 14).cpu(), want)
    k = int(g["moved_camera"])
    cam = o_cams.FoVPerspectiveCameras(R=o_cams.axis_angle_to_matrix(aa[k:k + 1]), T=T[k:k + 1], zfar=100.0)
    moved = scone_utils.move_view_state_to_view_space(state.to(cuda_device), cam, 7, 14).cpu()
    assert np.array_equal(moved.numpy(), np.unpackbits(g["moved_bits"], axis=-1)[..., :98].astype(np.float32))


def test_fused_view_state_harmonics_is_bitwise_the_two_kernel_path(cuda_device):
    """mac_viewstate_harm_f32 == compute_view_harmonics(compute_view_state(...)) bit for bit (ragged sizes, 0..40
    views, pts_dim 3 and 4), and within 2e-6 of the reference golden where the bins agree."""
    base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, cuda_device)
    for B, P, V, seed, D in ((1, 1, 1, 1, 4), (2, 333, 3, 2, 4), (1, 4099, 10, 3, 3), (3, 65, 40, 4, 4), (1, 50, 0, 5, 4)):
        pts, X_view = synth.view_state_inputs(B, P, max(V, 1), seed)
        pts, X_view = pts[..., :D].contiguous().to(cuda_device), X_view[:V].to(cuda_device)
        two = scone_utils.compute_view_harmonics(scone_utils.compute_view_state(pts, X_view, 7, 14), base, h_polar, h_azim, 7, 14)
        n0 = ops.launch_count()
        one = scone_utils.compute_view_state_harmonics(pts, X_view, base, h_polar, h_azim, 7, 14)
        assert ops.launch_count() == n0 + (1 if V > 0 else 0)
        assert torch.equal(one, two), (B, P, V)
    g = load_golden("view_state_10views")
    pts, _ = synth.view_state_inputs(int(g["B"]), int(g["P"]), int(g["V"]), int(g["seed"]))
    X_view = torch.from_numpy(g["X_view"])
    got = scone_utils.compute_view_state_harmonics(pts.to(cuda_device), X_view.to(cuda_device), base, h_polar, h_azim, 7, 14).cpu()
    safe = _boundary_margin(pts, X_view, 7, 14) > 1e-5
    assert np.abs(got.numpy() - g["view_harmonics"])[safe.numpy()].max() <= 2e-6


def test_view_state_bins_against_float64(cuda_device):
    """Index work vs the exact answer: bins of the CUDA kernel and of the fp32 reference arithmetic (oracle) against the
    mathematical rule evaluated in float64 (nearest bin centre in elevation / azimuth, clamps and wrap of
    scone_utils.py:838-849).  Two sources of disagreement with the exact rule exist IN THE REFERENCE and are reproduced
    by the kernel: (i) rays within fp32 rounding of a bin boundary, (ii) for negative azimuths the reference's float
    floor division (x - x % d) / d lands a hair above or below the integer and the later `.long()` truncates it
    (about 4 % of the rows move to the neighbouring azimuth bin).  The test prints all three rates, requires the
    kernel to equal the exact rule wherever the reference does and the ray is not on a boundary, and bounds
    kernel-vs-reference by the boundary cases."""
    B, P, V = 1, 50000, 10
    pts, X_view = synth.view_state_inputs(B, P, V, 77)
    got = scone_utils.compute_view_state(pts.to(cuda_device), X_view.to(cuda_device), 7, 14).cpu()
    ref32 = o_vs.view_state(pts, X_view, 7, 14)
    d = X_view.double().view(1, 1, V, 3) - pts[..., :3].double().unsqueeze(2)
    r = d.norm(dim=-1)
    elev = torch.asin((d[..., 1] / r).clamp(-1, 1))
    azim = torch.atan2(d[..., 0], d[..., 2])
    es, az = math.pi / 8, 2 * math.pi / 14
    ie = torch.floor(elev / es + 0.5).clamp(-4, 6) + 3
    ia = torch.floor(azim / az + 0.5)
    ia = torch.where(ia > 7, torch.full_like(ia, -7.0), ia)
    ia = torch.where(ia < 0, ia + 14, ia)
    idx = (ie.long() * 14 + ia.long()) % 98
    truth = torch.zeros(B, P, 98).scatter_(2, idx, 1.0)
    cuda_bad = (got != truth).any(-1)
    ref_bad = (ref32 != truth).any(-1)
    both = (got != ref32).any(-1)
    print("view-state rows differing from the exact float64 rule: CUDA %.3e, reference fp32 %.3e; CUDA vs reference fp32 %.3e"
          % (cuda_bad.float().mean().item(), ref_bad.float().mean().item(), both.float().mean().item()))
    assert both.float().mean().item() <= 1e-3
    assert abs(cuda_bad.float().mean().item() - ref_bad.float().mean().item()) <= 1e-3
    margin = _boundary_margin(pts, X_view, 7, 14)
    clear = (~ref_bad) & (margin > 1e-5)          # the reference is exact here and no ray is near a boundary
    assert clear.float().mean().item() > 0.9
    assert torch.equal(got[clear], truth[clear])


def test_sample_proxy_draws_against_float64(cuda_device):
    """Index work vs the exact answer: the 2048 draws of the CUDA sampler and of the reference's fp32 arithmetic
    (fp32 cumsum of fp32 quotients, scone_utils.py:1046-1056) against an inverse-CDF look-up in float64.  A draw can
    only differ where its uniform is within fp32 rounding of a CDF step; both sides' mismatch rates against float64 are
    printed and bounded."""
    gen = torch.Generator().manual_seed(909)
    N, n_sample = 100000, 2048
    X = torch.rand(N, 3, generator=gen) - 0.5
    preds = torch.rand(N, 1, generator=gen)
    vh = torch.randn(N, 64, generator=gen)
    bad_cuda = bad_ref = total = 0
    for rep in range(8):
        u = torch.rand(n_sample, generator=gen)
        res, _, inv = scone_utils.sample_proxy_points(X.to(cuda_device), preds.to(cuda_device), vh.to(cuda_device), n_sample,
                                                      0.1, return_index=True, samples=u.to(cuda_device))
        want_res, _, want_inv = o_sampling.sample_proxy_points(X, preds, vh, n_sample, 0.1, u=u)
        keep = preds[:, 0] > 0.1
        p64 = preds[keep, 0].double()
        cdf64 = torch.cumsum(p64 / p64.sum(), dim=0)
        pick = torch.searchsorted(cdf64, u.double()).clamp(max=cdf64.numel() - 1)
        truth = torch.cat((X[keep][pick], preds[keep][pick]), dim=-1)
        bad_cuda += (res.cpu()[inv.cpu()] != truth).any(-1).sum().item()
        bad_ref += (want_res[want_inv] != truth).any(-1).sum().item()
        total += n_sample
    print("proxy draws differing from the float64 inverse CDF: CUDA %.2e, reference fp32 %.2e (of %d draws)"
          % (bad_cuda / total, bad_ref / total, total))
    assert bad_cuda / total <= 5e-3 and bad_ref / total <= 5e-3
    assert bad_cuda <= 2 * bad_ref + 8       # the kernel is not further from the exact draw than the reference is


def test_sample_proxy_points_edge_cases(cuda_device):
    gen = torch.Generator().manual_seed(8)
    X = torch.rand(50, 3, generator=gen)
    vh = torch.randn(50, 64, generator=gen)
    # one point above the threshold: every draw picks it
    preds = torch.full((50, 1), 0.05)
    preds[17] = 0.9
    u = torch.rand(64, generator=gen)
    res, res_h, inv = scone_utils.sample_proxy_points(X.to(cuda_device), preds.to(cuda_device), vh.to(cuda_device), 64, 0.1,
                                                      return_index=True, samples=u.to(cuda_device))
    assert res.shape == (1, 4) and torch.equal(res.cpu()[0, :3], X[17]) and torch.all(inv == 0)
    # u beyond the last CDF value falls back to index 0 of the kept points (reference: all-negative row -> argmin 0)
    preds = torch.rand(50, 1, generator=gen) + 0.2
    u = torch.tensor([1.5, 0.0, 0.999999])
    res, _, inv = scone_utils.sample_proxy_points(X.to(cuda_device), preds.to(cuda_device), vh.to(cuda_device), 3, 0.1,
                                                  return_index=True, samples=u.to(cuda_device))
    want_res, _, want_inv = o_sampling.sample_proxy_points(X, preds, vh, 3, 0.1, u=u)
    assert torch.equal(res.cpu(), want_res) and torch.equal(inv.cpu(), want_inv)
    # nothing above the threshold: empty result
    res, res_h, inv = scone_utils.sample_proxy_points(X.to(cuda_device), torch.zeros(50, 1, device=cuda_device),
                                                      vh.to(cuda_device), 8, 0.1, return_index=True,
                                                      samples=torch.rand(8, device=cuda_device))
    assert res.shape == (0, 4) and res_h.shape == (0, 64)
