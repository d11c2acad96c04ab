"""Multi-GPU layout of the NBV scoring path: the candidate-camera axis is partitioned across ranks
(one process per GPU); every rank holds the full point set and scores its own slice of cameras, then
one all-gather of the per-candidate scores gives every rank the full (B, C) matrix and a replicated
argmax (ties -> lowest index), so no second collective is needed.

The reference never shards inference (its multi-GPU is DDP over scenes, train.py:29-33); this is the
partition SURVEY.md section 8e derives from `compute_coverage_gain` (cameras are independent,
networks/SconeVis.py:230-250).  Because the kernel's per-camera sums are exact integer sums, the
gathered matrix is bitwise identical for every world size.
"""
import torch
import torch.distributed as dist


def camera_partition(n_cameras, world_size, rank):
    """Contiguous balanced block [c0, c1) of rank `rank`: the first n % W ranks get one extra camera."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    base, rem = divmod(int(n_cameras), world_size)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_scores(local, n_cameras, group=None):
    """`local` is (B, C) with the columns of this rank's partition filled in -> (B, C) with every
    column filled, identical on all ranks.  One all_gather of B * ceil(C/W) floats per rank."""
    rank, world = _world(group)
    if world == 1:
        return local
    B = local.shape[0]
    width = -(-int(n_cameras) // world)  # widest block
    c0, c1 = camera_partition(n_cameras, world, rank)
    send = local.new_zeros((B, width))
    send[:, :c1 - c0] = local[:, c0:c1]
    recv = local.new_empty((world, B, width))
    dist.all_gather(list(recv.unbind(0)), send, group=group)  # one collective; works on nccl and gloo
    full = torch.empty_like(local)
    for r in range(world):
        r0, r1 = camera_partition(n_cameras, world, r)
        full[:, r0:r1] = recv[r, :, :r1 - r0]
    return full


def sharded_coverage_gain(score_fn, pts, harmonics, X_cam, group=None):
    """Score the local camera slice with `score_fn(pts, harmonics, X_cam, cam_range=(c0, c1))`
    (e.g. `SconeVis.compute_coverage_gain`), all-gather, return ((B, C) scores, (B,) argmax)."""
    rank, world = _world(group)
    C = X_cam.shape[1]
    c0, c1 = camera_partition(C, world, rank)
    local = score_fn(pts, harmonics, X_cam, cam_range=(c0, c1))
    scores = gather_scores(local, C, group=group)
    return scores, nbv_argmax(scores)


def nbv_argmax(scores):
    """First maximum along the camera axis (reference testers/shapenet.py:172, testers/scene.py:454)."""
    return torch.argmax(scores, dim=-1)
