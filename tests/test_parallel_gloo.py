"""Host-side logic of the N>1 path on CPU: camera partition + score all-gather + replicated argmax,
world_size 2 and 3 over gloo.  The local scorer is injected (the oracle stands in for the CUDA
kernel, which cannot run here); the product code under test is macarons_b200.parallel."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from macarons_b200 import parallel
from oracle import sh_cov


def test_camera_partition_covers_axis():
    for C in (1, 2, 7, 64, 512, 513):
        for W in (1, 2, 3, 4, 8):
            blocks = [parallel.camera_partition(C, W, r) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == C
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.camera_partition(4, 2, 2)


def _oracle_slice_scorer(pts, harm, cams, cam_range):
    c0, c1 = cam_range
    out = torch.zeros(pts.shape[0], cams.shape[1])
    if c1 > c0:
        out[:, c0:c1] = sh_cov.coverage_gain(pts, harm, cams[:, c0:c1])
    return out


def _worker(rank, world, port, C, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        pts, harm, cams = synth.covgain_inputs(2, 64, C, seed=5)
        scores, best = parallel.sharded_coverage_gain(_oracle_slice_scorer, pts, harm, cams)
        full = sh_cov.coverage_gain(pts, harm, cams)
        ok = torch.equal(scores, full) and torch.equal(best, full.argmax(-1))
        # sharded upload: every rank copies 1/W of the rows, one all-gather completes the tensor (rows % W != 0 too)
        for rows in (10, 7, 1):
            host = torch.arange(rows * 3, dtype=torch.float32).view(rows, 3)
            ok = ok and torch.equal(parallel.upload_rows_sharded(host, torch.device("cpu")), host)
        buf = torch.empty((-(-7 // world) * world, 3))
        ok = ok and torch.equal(parallel.upload_rows_sharded(host.expand(7, 3).contiguous(), torch.device("cpu"), buf=buf),
                                host.expand(7, 3))
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, best.tolist()))
        if rank == 0:
            ret.put(gathered)
    finally:
        dist.destroy_process_group()


def test_point_partition_is_tile_aligned_and_covers():
    for P in (1, 31, 32, 33, 2048, 200704, 200705):
        for W in (1, 2, 3, 8):
            blocks = [parallel.point_partition(P, W, r) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(b0 % 32 == 0 for b0, b1 in blocks if b1 > b0)
            sizes = [b1 - b0 for b0, b1 in blocks]
            assert max(sizes) - min(sizes) <= 32 + 31


@pytest.mark.parametrize("world,C", [(2, 9), (3, 7), (2, 1)])
def test_sharded_coverage_gain_gloo(world, C):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() + 17 * world + C) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, C, ret)) for r in range(world)]
    for p in procs:
        p.start()
    out = ret.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in out)
    assert all(b == out[0][1] for _, b in out)   # every rank agrees on the NBV index
