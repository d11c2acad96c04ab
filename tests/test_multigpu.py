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
        out = [None] * world
        dist.all_gather_object(out, (ok_nccl, ok_fused))
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
