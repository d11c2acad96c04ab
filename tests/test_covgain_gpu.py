"""GPU parity tests of the coverage-gain kernel (rows a1-a5, a16): CUDA path through the C ABI vs the
oracle / golden vectors on identical seeded inputs, plus size-independent properties at full size."""
import numpy as np
import pytest
import torch

import synth
from conftest import load_golden
from macarons_b200 import _lib, ops, parallel
from macarons_b200.networks.Macarons import Macarons
from macarons_b200.networks.SconeVis import SconeVis
from oracle import sh_cov
from tolerances import COVERAGE_ATOL, assert_visibility_close

pytestmark = pytest.mark.gpu

CASES = ["covgain_ragged_sigmoid", "covgain_cfg2_sigmoid", "covgain_ragged_relu",
         "covgain_bigcoef_sigmoid", "covgain_single_cam"]


def _inputs(g):
    return synth.covgain_inputs(int(g["B"]), int(g["P"]), int(g["C"]), int(g["seed"]),
                                pts_dim=int(g["pts_dim"]), coef_scale=float(g["coef_scale"]))


@pytest.fixture(scope="module")
def vis(cuda_device):
    torch.manual_seed(5)
    return SconeVis().to(cuda_device)


@pytest.mark.parametrize("name", CASES)
def test_coverage_gain_matches_golden(name, vis, cuda_device):
    g = load_golden(name)
    pts, harm, cams = _inputs(g)
    vis.use_sigmoid = bool(g["use_sigmoid"])
    try:
        n0 = ops.launch_count()
        got = vis.compute_coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)).cpu().numpy()
        assert ops.launch_count() == n0 + 1, "the CUDA kernel did not run"
    finally:
        vis.use_sigmoid = True
    scale = 1.0 if g["use_sigmoid"] else max(1.0, float(np.abs(g["coverage"]).max()))
    assert got.shape == g["coverage"].shape and got.dtype == np.float32
    assert np.abs(got - g["coverage"]).max() <= COVERAGE_ATOL * scale
    assert np.array_equal(np.argmax(got, -1), g["argmax"])   # NBV index identical to the reference
    truth = sh_cov.coverage_gain_f64(pts.numpy(), harm.numpy(), cams.numpy(), use_sigmoid=bool(g["use_sigmoid"]))
    assert np.abs(got - truth).max() <= 2e-6 * scale


@pytest.mark.parametrize("name", CASES)
def test_visibility_gains_match_golden(name, vis, cuda_device):
    g = load_golden(name)
    pts, harm, cams = _inputs(g)
    sig = bool(g["use_sigmoid"])
    vis.use_sigmoid = sig
    try:
        got = vis.compute_visibilities(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)).cpu().numpy()
    finally:
        vis.use_sigmoid = True
    assert got.shape == g["visibility"].shape
    truth = sh_cov.visibility_gains_f64(pts.numpy(), harm.numpy(), cams.numpy(), use_sigmoid=sig)
    scale = float(g["coef_scale"]) * (1.0 if sig else 8.0)
    assert_visibility_close(got, g["visibility"], truth, pts.numpy(), cams.numpy(), coef_scale=scale)
    if sig:
        mac = Macarons(None, None, vis)
        again = mac.compute_visibility_gains(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device))
        assert np.array_equal(again.cpu().numpy(), got)   # same kernel behind the Macarons wrapper


def test_macarons_wrapper_refuses_relu(vis, cuda_device):
    mac = Macarons(None, None, vis)
    vis.use_sigmoid = False
    try:
        with pytest.raises(NameError):
            mac.compute_visibility_gains(torch.zeros(1, 4, 4, device=cuda_device),
                                         torch.zeros(1, 4, 64, device=cuda_device),
                                         torch.ones(1, 1, 3, device=cuda_device))
    finally:
        vis.use_sigmoid = True


def test_coverage_gain_multiple_matches_golden(vis, cuda_device):
    pts, harm, cams = synth.covgain_inputs(1, 128, 5, 106)
    for n_cam in (2, 3):
        g = load_golden("covgain_multiple_n%d" % n_cam)
        val, tuples = vis.compute_coverage_gain_multiple(pts.to(cuda_device), harm.to(cuda_device),
                                                         cams.to(cuda_device), n_cam)
        assert np.array_equal(tuples.numpy(), g["tuples"])
        assert np.abs(val.cpu().numpy() - g["coverage"]).max() <= COVERAGE_ATOL
    with pytest.raises(NameError):
        vis.compute_coverage_gain_multiple(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device), 4)


@pytest.mark.parametrize("B,P,C,D", [(1, 1, 1, 3), (1, 31, 3, 3), (2, 33, 129, 4), (1, 4097, 257, 5), (5, 64, 32, 4)])
def test_edge_shapes_against_oracle(B, P, C, D, cuda_device):
    """Ragged / minimal shapes: one point, P and C off the 32-lane and 128-camera tile sizes, pts_dim 3..5."""
    pts, harm, cams = synth.covgain_inputs(B, P, C, seed=1000 + P + C, pts_dim=D)
    got = ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)).cpu()
    want = sh_cov.coverage_gain(pts, harm, cams, cam_chunk=32)
    # the reference's own ill-conditioned rays (tolerances.py) are not averaged away when P is small
    assert (got - want).abs().max().item() <= COVERAGE_ATOL + 2e-3 / P
    truth = sh_cov.coverage_gain_f64(pts.numpy(), harm.numpy(), cams.numpy())
    assert np.abs(got.numpy() - truth).max() <= 2e-6
    assert np.array_equal(got.numpy().argmax(-1), truth.argmax(-1))
    per_point = ops.visibility_gains(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device))
    assert (per_point.mean(-1).cpu() - got).abs().max().item() <= 1e-6


def test_known_answers(cuda_device):
    """Zero coefficients -> sigmoid(0) = 0.5 exactly; only the l=0 coefficient -> a constant."""
    pts, harm, cams = synth.covgain_inputs(1, 1000, 17, seed=3)
    zero = torch.zeros_like(harm)
    got = ops.coverage_gain(pts.to(cuda_device), zero.to(cuda_device), cams.to(cuda_device)).cpu()
    assert torch.equal(got, torch.full_like(got, 0.5))
    l0 = zero.clone()
    l0[..., 0] = 2.0
    want = 1.0 / (1.0 + np.exp(-2.0 * 0.5 / np.sqrt(np.pi)))     # Y_00 = 1 / (2 sqrt(pi))
    got = ops.coverage_gain(pts.to(cuda_device), l0.to(cuda_device), cams.to(cuda_device)).cpu().numpy()
    assert np.abs(got - want).max() < 5e-7
    relu = ops.coverage_gain(pts.to(cuda_device), (-l0).to(cuda_device), cams.to(cuda_device), use_sigmoid=False)
    assert torch.equal(relu.cpu(), torch.zeros(1, 17))


def test_camera_partition_is_bitwise_neutral(cuda_device):
    """Scoring the camera axis in slices (the multi-GPU partition) gives bit-identical scores for every
    world size, and identical NBV index: per-camera sums are exact integer sums."""
    pts, harm, cams = synth.covgain_inputs(2, 5000, 100, seed=21)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    full = ops.coverage_gain(*d)
    for world in (2, 3, 4, 8):
        out = torch.zeros_like(full)
        for r in range(world):
            ops.coverage_gain(*d, cam_range=parallel.camera_partition(100, world, r), out=out)
        assert torch.equal(out, full), world
    again = ops.coverage_gain(*d)
    assert torch.equal(again, full)          # run-to-run deterministic despite atomics
    perm = torch.randperm(100)
    shuffled = ops.coverage_gain(d[0], d[1], d[2][:, perm.to(cuda_device)].contiguous())
    assert torch.equal(shuffled, full[:, perm.to(cuda_device)])   # camera order does not matter


def test_workspace_left_clean_and_stream_safe(cuda_device):
    pts, harm, cams = synth.covgain_inputs(1, 3000, 40, seed=22)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    ref = ops.coverage_gain(*d)
    s = torch.cuda.Stream(device=cuda_device)
    s.wait_stream(torch.cuda.current_stream(cuda_device))
    with torch.cuda.stream(s):
        outs = [ops.coverage_gain(*d) for _ in range(5)]
    s.synchronize()
    assert all(torch.equal(o, ref) for o in outs)


def test_argument_errors(cuda_device):
    pts, harm, cams = [t.to(cuda_device) for t in synth.covgain_inputs(1, 64, 4, seed=1)]
    with pytest.raises(ValueError):
        ops.coverage_gain(pts, harm[:, :, :32].contiguous(), cams)
    with pytest.raises(ValueError):
        ops.coverage_gain(pts, harm, cams, cam_range=(3, 9))
    with pytest.raises(TypeError):
        ops.coverage_gain(pts.double(), harm, cams)
    lib = _lib.load()
    rc = lib.mac_covgain_f32(pts.data_ptr(), 4, harm.data_ptr(), cams.data_ptr(), pts.data_ptr(), 1, 64, 4, 0, 4, 1,
                             None, 0, None)
    assert rc == -3 and b"workspace" in lib.mac_last_error()


def test_host_buffer_entry_point(cuda_device):
    pts, harm, cams = synth.covgain_inputs(2, 700, 12, seed=8)
    got = ops.coverage_gain_host(pts, harm, cams, device=0)
    dev = ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)).cpu().numpy()
    assert np.array_equal(got, dev)
    part = ops.coverage_gain_host(pts.numpy(), harm.numpy(), cams.numpy(), cam_range=(4, 9), device=0)
    assert np.array_equal(part[:, 4:9], dev[:, 4:9]) and not part[:, :4].any() and not part[:, 9:].any()


def test_host_buffer_entry_point_sliced_pipeline(cuda_device):
    """One large cloud goes through mac_covgain_host in 8 point slices (H2D copies overlapped with the kernel, partial
    sums kept in the fixed-point workspace): same scores as the single device call, pinned or pageable host memory."""
    pts, harm, cams = synth.covgain_inputs(1, 70001, 40, seed=12)
    dev = ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)).cpu().numpy()
    got = ops.coverage_gain_host(pts, harm, cams, device=0)
    assert np.abs(got - dev).max() <= 1e-7
    pinned = ops.coverage_gain_host(pts.pin_memory().numpy(), harm.pin_memory().numpy(), cams.numpy(), device=0)
    assert np.array_equal(pinned, got)
    part = ops.coverage_gain_host(pts, harm, cams, cam_range=(7, 33), device=0)
    assert np.abs(part[:, 7:33] - dev[:, 7:33]).max() <= 1e-7 and not part[:, :7].any() and not part[:, 33:].any()
    assert np.array_equal(ops.coverage_gain_host(pts, harm, cams, device=0), got)   # the workspace is left clean


def test_full_size_cfg5_properties(cuda_device):
    """BASELINE config 5 shape (200 704 points x 512 cameras): checked through properties that do not
    need the full oracle -- a random subset of cameras against the float64 closed form over ALL points,
    partition invariance, and mean(per-point) == coverage on a camera slice."""
    P, C = 200704, 512
    pts, harm, _ = synth.covgain_inputs(1, P, 1, seed=5000)
    cams = synth.fibonacci_cameras(C)[None].contiguous()
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    full = ops.coverage_gain(*d)
    assert full.shape == (1, C) and torch.isfinite(full).all()
    # ALL 512 cameras over all 200 704 points against the float64 closed form (C + OpenMP restatement of
    # oracle/sh_cov.py::coverage_gain_f64, a few seconds): every score and the NBV argmax
    truth = sh_cov.coverage_gain_f64_c(pts.numpy(), harm.numpy(), cams.numpy())
    got = full.cpu().numpy()
    order = np.argsort(-truth[0])
    print("cfg5: max |score - f64| over 512 cameras %.2e; f64 top-1 minus top-2 %.2e; argmax %d (f64 %d)"
          % (np.abs(got - truth).max(), truth[0, order[0]] - truth[0, order[1]], got.argmax(), truth.argmax()))
    assert np.abs(got - truth).max() <= 1e-6
    assert int(got.argmax()) == int(truth.argmax())
    assert np.array_equal(np.argsort(-got[0])[:8], order[:8])      # the 8 best candidates in the same order
    pick = [0, 17, 255]   # the numpy closed form on 3 cameras: the C restatement and the numpy original agree
    truth_np = sh_cov.coverage_gain_f64(pts.numpy(), harm.numpy(), cams.numpy()[:, pick])
    assert np.abs(truth[:, pick] - truth_np).max() <= 1e-12
    out = torch.zeros_like(full)
    for r in range(8):
        ops.coverage_gain(*d, cam_range=parallel.camera_partition(C, 8, r), out=out)
    assert torch.equal(out, full)
    sl = (96, 104)
    per_point = ops.visibility_gains(*d, cam_range=sl)
    assert (per_point[:, sl[0]:sl[1]].double().mean(-1) - full[:, sl[0]:sl[1]].double()).abs().max().item() <= 1e-6
    # fp32 oracle (reference arithmetic) on 2 cameras over all points
    want = sh_cov.coverage_gain(pts, harm, cams[:, :2])
    assert (full.cpu()[:, :2] - want).abs().max().item() <= COVERAGE_ATOL


def test_non_finite_inputs_give_nan_scores(cuda_device):
    """A NaN / inf coefficient row or point makes every score of its cloud NaN (the reference's mean over the points
    propagates it), a NaN camera only its own score; the other clouds and the next call are unaffected."""
    pts, harm, cams = synth.covgain_inputs(3, 500, 40, seed=77)
    clean = ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device))
    h2 = harm.clone()
    h2[1, 123, 7] = float("nan")
    p2 = pts.clone()
    p2[2, 499, 1] = float("inf")
    got = ops.coverage_gain(p2.to(cuda_device), h2.to(cuda_device), cams.to(cuda_device))
    assert torch.equal(got[0], clean[0]) and torch.isnan(got[1]).all() and torch.isnan(got[2]).all()
    c2 = cams.clone()
    c2[0, 5, 2] = float("nan")
    got = ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), c2.to(cuda_device))
    assert torch.isnan(got[0, 5]) and torch.equal(got[1:], clean[1:])
    keep = torch.ones(40, dtype=torch.bool)
    keep[5] = False
    assert torch.equal(got[0, keep], clean[0, keep])
    assert torch.equal(ops.coverage_gain(pts.to(cuda_device), harm.to(cuda_device), cams.to(cuda_device)), clean)


def test_full_size_cfg4_batched_clouds(cuda_device):
    """BASELINE config 4 shape (32 clouds x 2048 proxy points x 256 cameras, 32 cameras per GPU at 8 GPUs): three of the
    clouds against the fp32 oracle and the float64 closed form, the 8-way camera partition bitwise neutral, NBV per cloud."""
    B, P, C = 32, 2048, 256
    pts, harm, cams = synth.covgain_inputs(B, P, C, seed=4004)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    full = ops.coverage_gain(*d)
    assert full.shape == (B, C) and torch.isfinite(full).all()
    for b in (0, 13, 31):
        want = sh_cov.coverage_gain(pts[b:b + 1], harm[b:b + 1], cams[b:b + 1], cam_chunk=32)
        assert (full[b:b + 1].cpu() - want).abs().max().item() <= COVERAGE_ATOL
        truth = sh_cov.coverage_gain_f64(pts[b:b + 1].numpy(), harm[b:b + 1].numpy(), cams[b:b + 1, :16].numpy())
        assert np.abs(full[b:b + 1, :16].cpu().numpy() - truth).max() <= 2e-6
        assert int(full[b].argmax()) == int(want[0].argmax())
    out = torch.zeros_like(full)
    for r in range(8):
        ops.coverage_gain(*d, cam_range=parallel.camera_partition(C, 8, r), out=out)
    assert torch.equal(out, full)


def test_fused_push_and_argmax_single_rank(cuda_device):
    """World size 1 of the fused gather: the kernel pushes into a local board, the wait kernel takes the argmax."""
    pts, harm, cams = synth.covgain_inputs(3, 900, 70, seed=31)
    d = [t.to(cuda_device) for t in (pts, harm, cams)]
    board = parallel.PeerScoreBoard(3, 70, cuda_device)
    want = ops.coverage_gain(*d)
    for _ in range(3):   # alternates between the two boards
        scores, best = board.step(*d)
        board.check()
        assert torch.equal(scores, want)
        assert torch.equal(best, want.argmax(-1))
    # point-partitioned form at world size 1: all points are "local"
    for _ in range(2):
        scores, best = board.step_points(*d, 900)
        board.check()
        assert torch.equal(scores, want) and torch.equal(best, want.argmax(-1))
    # ... and one cloud uploaded from pinned host memory in slices (accumulate + fused finish)
    p1, h1, c1 = synth.covgain_inputs(1, 4100, 21, seed=32)
    want1 = ops.coverage_gain(p1.to(cuda_device), h1.to(cuda_device), c1.to(cuda_device))
    board1 = parallel.PeerScoreBoard(1, 21, cuda_device)
    for n_slices in (1, 2, 5):
        s1, b1 = board1.step_points_from_host(p1.pin_memory(), h1.pin_memory(), c1.to(cuda_device), slices=n_slices)
        board1.check()
        assert torch.equal(s1, want1) and torch.equal(b1, want1.argmax(-1))
    # ties and NaN follow torch.argmax: first maximum, NaN is maximal
    s = torch.tensor([[0.5, 0.7, 0.7, 0.1], [0.2, float("nan"), 0.9, float("nan")]], device=cuda_device)
    flags = torch.full((1,), 5, dtype=torch.int32, device=cuda_device)
    best = torch.zeros(2, dtype=torch.int64, device=cuda_device)
    status = torch.zeros(4, dtype=torch.int32, device=cuda_device)
    ops.gather_wait_argmax(s, flags, 1, 5, best, status)
    assert best.tolist() == [1, 1] and status[0].item() == 0 and status[3].item() == 1
    ops.gather_wait_argmax(s, flags, 1, 6, best, status)   # epoch never arrives -> bounded wait, status 1
    assert status[0].item() == 1 and status[1].item() > 1_000_000_000      # waited ~2 s
    ops.gather_wait_argmax(s, flags, 1, 5, best, status)   # the status is sticky: a later good step does not hide the timeout
    assert best.tolist() == [1, 1] and status[0].item() == 1 and status[3].item() == 3
