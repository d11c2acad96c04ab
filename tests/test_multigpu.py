"""Two-GPU test of the sharded scoring step (needs >= 2 CUDA devices; skipped otherwise): NCCL all-gather
path and fused peer-memory push path must both reproduce the single-GPU score matrix bit for bit."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from macarons_b200 import ops, parallel
        pts, harm, cams = synth.covgain_inputs(2, 3000, 101, seed=77)
        d = [t.to(dev) for t in (pts, harm, cams)]
        full = ops.coverage_gain(*d)

        def score(p, h, c, cam_range):
            return ops.coverage_gain(p, h, c, cam_range=cam_range)

        scores, best = parallel.sharded_coverage_gain(score, *d)
        ok_nccl = torch.equal(scores, full) and torch.equal(best, full.argmax(-1))
        board = parallel.PeerScoreBoard(2, 101, dev)
        ok_fused = True
        for _ in range(4):
            s, b = board.step(*d)
            board.check()
            ok_fused = ok_fused and torch.equal(s, full) and torch.equal(b, full.argmax(-1))
        # B = 32 clouds (BASELINE config 4 shape, reduced): sharded == single GPU bit for bit
        pts4, harm4, cams4 = synth.covgain_inputs(8, 512, 64, seed=78)
        d4 = [t.to(dev) for t in (pts4, harm4, cams4)]
        full4 = ops.coverage_gain(*d4)
        board4 = parallel.PeerScoreBoard(8, 64, dev)
        s4, b4 = board4.step(*d4)
        board4.check()
        ok_fused = ok_fused and torch.equal(s4, full4) and torch.equal(b4, full4.argmax(-1))
        # the step partitioned over the points (each rank holds only its rows): exact partial sums -> bitwise equal
        p0, p1 = parallel.point_partition(3000, world, rank)
        for _ in range(3):
            sp, bp = board.step_points(d[0][:, p0:p1].contiguous(), d[1][:, p0:p1].contiguous(), d[2], 3000)
            board.check()
            ok_fused = ok_fused and torch.equal(sp, full) and torch.equal(bp, full.argmax(-1))
        # one cloud from pinned host memory, uploaded in slices by each rank (its own rows only)
        pts1, harm1, cams1 = synth.covgain_inputs(1, 5000, 33, seed=80)
        full1 = ops.coverage_gain(pts1.to(dev), harm1.to(dev), cams1.to(dev))
        board1 = parallel.PeerScoreBoard(1, 33, dev)
        for n_slices in (1, 3, 4):
            sh, bh = board1.step_points_from_host(pts1.pin_memory(), harm1.pin_memory(), cams1.to(dev), slices=n_slices)
            board1.check()
            ok_fused = ok_fused and torch.equal(sh, full1) and torch.equal(bh, full1.argmax(-1))
        s_again, _ = board.step(*d)          # the two kinds of step can be mixed on one board
        ok_fused = ok_fused and torch.equal(s_again, full)
        # online loop: clouds of the SconeVis forward sharded over the ranks + camera-sharded scoring reproduces the
        # single-GPU loop (same chosen camera sequence, same final scores)
        import contextlib
        import io
        from macarons_b200 import nbv
        from macarons_b200.networks.SconeVis import SconeVis
        from macarons_b200.utility import scone_utils
        with contextlib.redirect_stdout(io.StringIO()):
            vis = SconeVis()
        vis.load_state_dict(synth.seeded_state_dict(vis.state_dict(), 5))
        vis = vis.to(dev).eval()
        ptsL = synth.covgain_inputs(1, 5 * 256, 1, seed=79)[0].to(dev)
        camsL = synth.fibonacci_cameras(37)[None].contiguous().to(dev)
        view0 = synth.sphere_cameras(1, 1.5, torch.Generator().manual_seed(3)).to(dev)
        base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
        boardL = parallel.PeerScoreBoard(1, 37, dev)
        solo, solo_scores = nbv.scone_online_loop(vis, ptsL, camsL, view0, base, h_polar, h_azim, 6, seq_len=256)
        shard, shard_scores = nbv.scone_online_loop(vis, ptsL, camsL, view0, base, h_polar, h_azim, 6, seq_len=256,
                                                    score_step=lambda p, h, c: boardL.step(p, h, c), shard_clouds=True)
        boardL.check()
        ok_loop = torch.equal(solo, shard) and torch.equal(solo_scores, shard_scores)
        out = [None] * world
        dist.all_gather_object(out, (ok_nccl, ok_fused and ok_loop))
        if rank == 0:
            ret.put(out)
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_scoring():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in out), out
