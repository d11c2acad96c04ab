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


def point_partition(n_points, world_size, rank, tile=32):
    """Contiguous block [p0, p1) of rank `rank` when the POINTS are partitioned (PeerScoreBoard.step_points): whole
    32-point tiles (the unit whose float sum the scoring kernel converts to exact fixed point), the first n_tiles % W
    ranks get one tile more.  Tile-aligned boundaries make the partial sums add up to bitwise the single-GPU result."""
    n_tiles = -(-int(n_points) // tile)
    t0, t1 = camera_partition(n_tiles, world_size, rank)
    return min(t0 * tile, n_points), min(t1 * tile, n_points)


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


def upload_rows_sharded(host, device, buf=None, group=None):
    """Replicate a host tensor (rows, ...) on every rank's device with ONE pass of the data over the host links:
    rank r copies rows [r*n, (r+1)*n) (n = ceil(rows / W)) from (pinned) host memory, then one all-gather over
    NVLink (NCCL) completes the tensor on every GPU.  The naive `host.to(device)` on every rank pushes W copies of
    the same bytes through the PCIe root complexes at once (measured at W = 8: 2.5 ms for 54.6 MB vs 1.5 ms on one
    GPU alone).  `buf`: optional (W*n, ...) device buffer to reuse.  -> device tensor (rows, ...) (a view of buf)."""
    rank, world = _world(group)
    rows = host.shape[0]
    if world == 1:
        return host.to(device, non_blocking=True)
    n = -(-rows // world)
    if buf is None:
        buf = torch.empty((world * n,) + tuple(host.shape[1:]), dtype=host.dtype, device=device)
    lo, hi = min(rank * n, rows), min((rank + 1) * n, rows)
    mine = buf[rank * n:(rank + 1) * n]
    if hi > lo:
        mine[:hi - lo].copy_(host[lo:hi], non_blocking=True)
    dist.all_gather_into_tensor(buf, mine, group=group)     # in place: `mine` is this rank's slot of `buf`
    return buf[:rows]


def nbv_argmax(scores):
    """First maximum along the camera axis (reference testers/shapenet.py:172, testers/scene.py:454)."""
    return torch.argmax(scores, dim=-1)


class PeerScoreBoard:
    """Fused compute + exchange for the sharded scoring step: the coverage-gain kernel of every rank stores
    its score columns directly into the (B, C) score board of every peer through NVLink peer mappings
    (torch symmetric memory) and raises a flag; a one-CTA kernel then waits for all ranks and takes the
    argmax.  Two kernels per step, no NCCL call on the data path.

    Two boards are used alternately (even / odd epochs) so a fast rank can never overwrite scores that a
    slow peer is still reading (a rank's step k+2 push is stream-ordered after its own step k+1 wait, which
    needs every peer's step k+1 push, which is stream-ordered after that peer's step k argmax).

    `scores` returned by `step()` is a view of the live board: consume (or clone) it before calling `step()`
    two more times."""

    def __init__(self, n_clouds, n_cameras, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.B, self.C = int(n_clouds), int(n_cameras)
        self.device = torch.device(device)
        self.rank, self.world = _world(group)
        self.epoch = 0
        n_scores = self.B * self.C
        self._flag_off = 2 * n_scores            # in 4-byte words: [board 0 | board 1 | flags(world) | pad | partials 0 | 1]
        # partial-sum regions of the point-sharded step: per parity world * B * C int64 + world * B * C u32
        self._part_off = (self._flag_off + 64 + 3) // 4 * 4                  # 16-byte aligned regions (int64 sums inside)
        self._part_words = (3 * self.world * n_scores + 3) // 4 * 4
        words = self._part_off + 2 * self._part_words
        if self.world > 1:
            grp = group if group is not None else dist.group.WORLD
            self._buf = symm_mem.empty(words, dtype=torch.float32, device=self.device)
            self._buf.zero_()
            self._hdl = symm_mem.rendezvous(self._buf, grp)
            bases = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)            # every board is zeroed before anyone pushes
        else:
            self._buf = torch.zeros(words, dtype=torch.float32, device=self.device)
            bases = [self._buf.data_ptr()]
        self._bases = bases
        self._flags = self._buf[self._flag_off:self._flag_off + 64].view(torch.int32)
        self.best = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        self.status = torch.zeros(4, dtype=torch.int32, device=self.device)   # [timeout flag, last wait ns, sum ns, steps]
        # per-parity C structs and board views, built once: step() only updates the epoch
        from . import _lib
        self._lib = _lib.load()
        self._plans = {}
        self._c_boards, self._views = [], []
        for parity in (0, 1):
            cb = _lib.PeerBoard()
            cb.world, cb.rank, cb.epoch = self.world, self.rank, 0
            for r, b in enumerate(bases):
                cb.scores[r] = b + 4 * parity * n_scores
                cb.flags[r] = b + 4 * self._flag_off
                cb.partials[r] = b + 4 * (self._part_off + parity * self._part_words)
            self._c_boards.append(cb)
            self._views.append(self._board(parity))

    def _board(self, parity):
        n = self.B * self.C
        return self._buf[parity * n:(parity + 1) * n].view(self.B, self.C)

    def step(self, pts, harmonics, X_cam, use_sigmoid=True, events=None):
        """-> ((B, C) scores, (B,) argmax), identical on every rank.  `events` = optional pair of
        torch.cuda.Event recorded around the scoring kernel (bench.py's roofline timing).
        One host call per step (mac_covgain_push_argmax_f32); the argument checks run once per distinct input set."""
        from . import _lib, ops
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.step(pts, harmonics, X_cam, use_sigmoid=use_sigmoid, events=events)
        self.epoch += 1
        parity = self.epoch & 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        key = (pts.data_ptr(), harmonics.data_ptr(), X_cam.data_ptr(), tuple(pts.shape), tuple(X_cam.shape), stream)
        plan = self._plans.get(key)
        if plan is not None and not (pts.is_contiguous() and harmonics.is_contiguous() and X_cam.is_contiguous()
                                     and pts.dtype == harmonics.dtype == X_cam.dtype == torch.float32):
            plan = None   # same address and shape as a validated input set, but a different layout / dtype: validate again
        if plan is None:
            p_c, h_c, x_c, B, P, D, C = ops._prep(pts, harmonics, X_cam)
            if (B, C) != (self.B, self.C):
                raise ValueError("board is (%d, %d), inputs are (%d, %d)" % (self.B, self.C, B, C))
            if not (p_c.data_ptr() == pts.data_ptr() and h_c.data_ptr() == harmonics.data_ptr()
                    and x_c.data_ptr() == X_cam.data_ptr()):
                raise ValueError("PeerScoreBoard.step needs contiguous inputs")
            c0, c1 = camera_partition(self.C, self.world, self.rank)
            ws = ops._workspace(self.device, B, C)
            plan = (P, D, c0, c1, ws)     # validated metadata only: the caller's tensors are alive for the call
            if len(self._plans) >= 8:
                self._plans.clear()
            self._plans[key] = plan
        P, D, c0, c1, ws = plan
        board = self._c_boards[parity]
        board.epoch = self.epoch & 0xFFFFFFFF
        ev0 = ev1 = None
        if events is not None:
            if not events[0].cuda_event or not events[1].cuda_event:   # torch creates the CUDA event on first record()
                events[0].record()
                events[1].record()
            ev0, ev1 = events[0].cuda_event, events[1].cuda_event
        _lib.check(self._lib.mac_covgain_push_argmax_f32(
            key[0], D, key[1], key[2], self.B, P, self.C, c0, c1, ops.ACT_SIGMOID if use_sigmoid else ops.ACT_RELU,
            ws.data_ptr(), ws.numel(), board, self.best.data_ptr(), self.status.data_ptr(), ev0, ev1, stream))
        return self._views[parity], self.best

    def step_points(self, pts_local, harmonics_local, X_cam, n_points_total, use_sigmoid=True):
        """The step partitioned over the POINTS: this rank holds only its rows `point_partition(n_points_total, W, rank)`
        of the clouds (pts_local (B, P_local, D), harmonics_local (B, P_local, 64)), integrates ALL cameras over them and
        exchanges exact fixed-point partial sums (mac_covgain_push_partial_argmax_f32).  -> ((B, C) scores, (B,) argmax),
        bitwise those of `step` / of one GPU.  Meant for inputs that arrive from the host: every rank uploads and reads
        1 / W of the rows, and nothing has to be replicated over NVLink first."""
        from . import _lib, ops
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.step_points(pts_local, harmonics_local, X_cam, n_points_total, use_sigmoid=use_sigmoid)
        p_c, h_c, x_c, B, P, D, C = ops._prep(pts_local, harmonics_local, X_cam)
        if (B, C) != (self.B, self.C):
            raise ValueError("board is (%d, %d), inputs are (%d, %d)" % (self.B, self.C, B, C))
        p0, p1 = point_partition(n_points_total, self.world, self.rank)
        if P != p1 - p0 or P == 0:
            raise ValueError("rank %d of %d must hold rows [%d, %d) of the %d points (got %d rows)"
                             % (self.rank, self.world, p0, p1, n_points_total, P))
        self.epoch += 1
        parity = self.epoch & 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ws = ops._workspace(self.device, B, C)
        board = self._c_boards[parity]
        board.epoch = self.epoch & 0xFFFFFFFF
        _lib.check(self._lib.mac_covgain_push_partial_argmax_f32(
            p_c.data_ptr(), D, h_c.data_ptr(), x_c.data_ptr(), B, P, int(n_points_total), C,
            ops.ACT_SIGMOID if use_sigmoid else ops.ACT_RELU, ws.data_ptr(), ws.numel(), board, self.best.data_ptr(),
            self.status.data_ptr(), stream))
        return self._views[parity], self.best

    def step_points_from_host(self, pts_host, harmonics_host, X_cam, use_sigmoid=True, slices=4, bufs=None):
        """`step_points` for ONE cloud (B = 1) whose points live in (pinned) host memory: pts_host (1, P, D) and
        harmonics_host (1, P, 64) are the FULL host tensors; this rank copies only its rows, in `slices` pieces on a copy
        stream, and integrates piece k (mac_covgain_accumulate_f32) while piece k+1 crosses PCIe; the last piece goes
        through the fused exchange.  `bufs`: optional (pts, harmonics) device buffers of at least this rank's row count.
        -> ((1, C) scores, (1,) argmax), bitwise those of `step`."""
        from . import _lib, ops
        if self.B != 1 or pts_host.shape[0] != 1:
            raise ValueError("step_points_from_host handles one cloud (B = 1)")
        P, D = pts_host.shape[1], pts_host.shape[2]
        p0, p1 = point_partition(P, self.world, self.rank)
        n = p1 - p0
        if n == 0:
            raise ValueError("rank %d of %d has no points (P = %d)" % (self.rank, self.world, P))
        with torch.cuda.device(self.device):
            if bufs is None:
                bufs = (torch.empty((1, n, D), dtype=torch.float32, device=self.device),
                        torch.empty((1, n, 64), dtype=torch.float32, device=self.device))
            d_pts, d_harm = bufs[0][:, :n], bufs[1][:, :n]
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream(self.device)
            per = max(32, -(-(-(-n // max(1, int(slices)))) // 32) * 32)          # whole tiles per slice
            bounds = list(range(0, n, per)) + [n]
            events = []
            self._copy_stream.wait_stream(main)            # the buffers may still be read by the previous step
            with torch.cuda.stream(self._copy_stream):
                for a, b in zip(bounds, bounds[1:]):
                    d_pts[:, a:b].copy_(pts_host[:, p0 + a:p0 + b], non_blocking=True)
                    d_harm[:, a:b].copy_(harmonics_host[:, p0 + a:p0 + b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                    events.append(ev)
            x_c = X_cam.contiguous()
            ws = ops._workspace(self.device, 1, self.C)
            act = ops.ACT_SIGMOID if use_sigmoid else ops.ACT_RELU
            for k, (a, b) in enumerate(zip(bounds, bounds[1:])):
                main.wait_event(events[k])
                last = k == len(bounds) - 2
                if not last:
                    _lib.check(self._lib.mac_covgain_accumulate_f32(d_pts[:, a:b].data_ptr(), D, d_harm[:, a:b].data_ptr(),
                                                                    x_c.data_ptr(), b - a, self.C, act, ws.data_ptr(),
                                                                    ws.numel(), main.cuda_stream))
                else:
                    self.epoch += 1
                    parity = self.epoch & 1
                    board = self._c_boards[parity]
                    board.epoch = self.epoch & 0xFFFFFFFF
                    _lib.check(self._lib.mac_covgain_push_partial_argmax_f32(
                        d_pts[:, a:b].data_ptr(), D, d_harm[:, a:b].data_ptr(), x_c.data_ptr(), 1, b - a, int(P), self.C, act,
                        ws.data_ptr(), ws.numel(), board, self.best.data_ptr(), self.status.data_ptr(), main.cuda_stream))
            return self._views[parity], self.best

    def reset_wait_stats(self):
        self.status[1:].zero_()

    def wait_stats(self):
        """-> (mean nanoseconds per step this rank spent waiting for its slowest peer inside the fused step, steps) since
        reset_wait_stats().  Synchronises."""
        _, _, total, n = self.status.tolist()
        return (total / n if n else 0.0), n

    def check(self):
        """Synchronises; raises if a wait timed out (a peer never arrived) in any step since the last check: the
        device-side status is sticky and is cleared here."""
        if int(self.status[0].item()) != 0:
            self.status[0] = 0
            raise RuntimeError("PeerScoreBoard: timed out waiting for a peer's scores")
