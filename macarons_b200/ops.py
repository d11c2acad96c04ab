"""Tensor-level wrappers over the C ABI.  Inputs must be CUDA fp32 tensors; outputs and workspaces
are allocated here with torch (the C side never allocates caller-visible memory) and kernels are
enqueued on torch's current stream."""
import torch

from . import _lib

ACT_RELU, ACT_SIGMOID = 0, 1
N_HARMONICS = 64

_workspaces = {}


def _require_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.MacaronsB200Error(
            "%s is on %s: macarons_b200 runs on CUDA (sm_100a) only and has no CPU fallback" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))


def wants_grad(*tensors, module=None):
    """True if autograd is recording and any of the tensors (or parameters of `module`) requires a gradient."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        return True
    return module is not None and any(p.requires_grad for p in module.parameters())


def refuse_grad(what, *tensors, module=None):
    """The fused forwards are inference kernels: they return tensors without a grad_fn.  Rather than silently
    training nothing, refuse when a gradient would be expected (reference trainers call these under autograd:
    trainers/pretrain_scone_vis.py:162-225, pretrain_scone_occ.py:158, train_macarons.py:423-444)."""
    if wants_grad(*tensors, module=module):
        raise NotImplementedError(
            "%s is an inference kernel without a backward pass: call it under torch.no_grad() (or freeze the module with "
            "requires_grad_(False) and pass inputs that do not require grad).  Differentiable on this path: "
            "compute_coverage_gain / compute_visibilities / compute_visibility_gains w.r.t. the harmonics "
            "(mac_covgain_backward_f32)." % what)


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(device, B, C):
    stream = _stream_ptr(device)
    key = (device.index, B, C, stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = _lib.load().mac_covgain_workspace_bytes(B, C)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)  # zero-filled once; kernels keep it zeroed
        _workspaces[key] = ws
    return ws


def _prep(pts, harmonics, X_cam):
    _require_cuda_f32("pts", pts)
    _require_cuda_f32("harmonics", harmonics)
    _require_cuda_f32("X_cam", X_cam)
    if pts.dim() != 3 or pts.shape[-1] < 3:
        raise ValueError("pts must have shape (n_clouds, seq_len, >=3), got %s" % (tuple(pts.shape),))
    B, P, D = pts.shape
    if tuple(harmonics.shape) != (B, P, N_HARMONICS):
        raise ValueError("harmonics must have shape (%d, %d, %d), got %s" % (B, P, N_HARMONICS, tuple(harmonics.shape)))
    if X_cam.dim() != 3 or X_cam.shape[0] != B or X_cam.shape[2] != 3:
        raise ValueError("X_cam must have shape (%d, n_camera_candidates, 3), got %s" % (B, tuple(X_cam.shape)))
    if not (pts.device == harmonics.device == X_cam.device):
        raise ValueError("pts, harmonics and X_cam must be on the same device")
    return pts.contiguous(), harmonics.contiguous(), X_cam.contiguous(), B, P, D, X_cam.shape[1]


def coverage_gain(pts, harmonics, X_cam, use_sigmoid=True, cam_range=None, out=None):
    """(B,P,>=3), (B,P,64), (B,C,3) -> (B,C) mean activated SH projection per camera.
    `cam_range=(c0, c1)` scores only that slice of cameras (other columns of `out` untouched)."""
    pts, harmonics, X_cam, B, P, D, C = _prep(pts, harmonics, X_cam)
    c0, c1 = (0, C) if cam_range is None else (int(cam_range[0]), int(cam_range[1]))
    if out is None:
        out = (torch.empty if cam_range is None else torch.zeros)((B, C), dtype=torch.float32, device=pts.device)
    else:
        _require_cuda_f32("out", out)
        if tuple(out.shape) != (B, C) or not out.is_contiguous():
            raise ValueError("out must be a contiguous (%d, %d) tensor" % (B, C))
    if P == 0 or C == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(pts.device):
        ws = _workspace(pts.device, B, C)
        _lib.check(lib.mac_covgain_f32(pts.data_ptr(), D, harmonics.data_ptr(), X_cam.data_ptr(), out.data_ptr(),
                                       B, P, C, c0, c1, ACT_SIGMOID if use_sigmoid else ACT_RELU,
                                       ws.data_ptr(), ws.numel(), _stream_ptr(pts.device)))
    return out


def visibility_gains(pts, harmonics, X_cam, use_sigmoid=True, cam_range=None, out=None):
    """Same inputs -> (B,C,P) per-point activated SH projection (no mean)."""
    pts, harmonics, X_cam, B, P, D, C = _prep(pts, harmonics, X_cam)
    c0, c1 = (0, C) if cam_range is None else (int(cam_range[0]), int(cam_range[1]))
    if out is None:
        out = (torch.empty if cam_range is None else torch.zeros)((B, C, P), dtype=torch.float32, device=pts.device)
    else:
        _require_cuda_f32("out", out)
        if tuple(out.shape) != (B, C, P) or not out.is_contiguous():
            raise ValueError("out must be a contiguous (%d, %d, %d) tensor" % (B, C, P))
    if P == 0 or C == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(pts.device):
        _lib.check(lib.mac_visibility_f32(pts.data_ptr(), D, harmonics.data_ptr(), X_cam.data_ptr(), out.data_ptr(),
                                          B, P, C, c0, c1, ACT_SIGMOID if use_sigmoid else ACT_RELU,
                                          _stream_ptr(pts.device)))
    return out


def coverage_gain_backward(pts, harmonics, X_cam, grad_out, use_sigmoid=True, per_point=False):
    """d loss / d harmonics (B,P,64) of coverage_gain (grad_out (B,C)) or visibility_gains (grad_out (B,C,P))."""
    pts, harmonics, X_cam, B, P, D, C = _prep(pts, harmonics, X_cam)
    _require_cuda_f32("grad_out", grad_out)
    want = (B, C, P) if per_point else (B, C)
    if tuple(grad_out.shape) != want:
        raise ValueError("grad_out must have shape %s, got %s" % (want, tuple(grad_out.shape)))
    grad_out = grad_out.contiguous()
    grad_h = torch.zeros((B, P, N_HARMONICS), dtype=torch.float32, device=pts.device)
    if P == 0 or C == 0 or B == 0:
        return grad_h
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mac_covgain_backward_f32(pts.data_ptr(), D, harmonics.data_ptr(), X_cam.data_ptr(),
                                                        grad_out.data_ptr(), grad_h.data_ptr(), B, P, C,
                                                        ACT_SIGMOID if use_sigmoid else ACT_RELU, int(bool(per_point)),
                                                        _stream_ptr(pts.device)))
    return grad_h


class _SHIntegration(torch.autograd.Function):
    """Differentiable wrapper of the two integration kernels: gradient with respect to the harmonics only (the points
    and cameras are data in the reference's trainers)."""

    @staticmethod
    def forward(ctx, pts, harmonics, X_cam, use_sigmoid, per_point):
        ctx.save_for_backward(pts, harmonics, X_cam)
        ctx.use_sigmoid, ctx.per_point = use_sigmoid, per_point
        fn = visibility_gains if per_point else coverage_gain
        return fn(pts.detach(), harmonics.detach(), X_cam.detach(), use_sigmoid=use_sigmoid)

    @staticmethod
    def backward(ctx, grad_out):
        pts, harmonics, X_cam = ctx.saved_tensors
        grad_h = coverage_gain_backward(pts, harmonics, X_cam, grad_out.to(torch.float32), use_sigmoid=ctx.use_sigmoid,
                                        per_point=ctx.per_point)
        return None, grad_h, None, None, None


def sh_integration(pts, harmonics, X_cam, use_sigmoid=True, per_point=False, cam_range=None):
    """coverage_gain / visibility_gains with autograd: when a gradient with respect to `harmonics` is wanted the call
    goes through _SHIntegration (forward kernel + mac_covgain_backward_f32); gradients with respect to the points or
    the camera positions are not implemented and are refused rather than silently dropped."""
    if wants_grad(pts, X_cam):
        raise NotImplementedError("coverage / visibility gains are differentiable with respect to the harmonics only; "
                                  "pts and X_cam must not require grad (they are data in the reference's trainers)")
    if wants_grad(harmonics):
        if cam_range is not None:
            raise NotImplementedError("cam_range (the multi-GPU camera slice) is an inference option")
        return _SHIntegration.apply(pts, harmonics, X_cam, bool(use_sigmoid), bool(per_point))
    fn = visibility_gains if per_point else coverage_gain
    return fn(pts, harmonics, X_cam, use_sigmoid=use_sigmoid, cam_range=cam_range)


def coverage_gain_push(pts, harmonics, X_cam, use_sigmoid, cam_range, score_ptrs, flag_ptrs, rank, epoch):
    """Score cameras [c0, c1) and store them, from inside the kernel, into the (B, C) score board of every
    rank (`score_ptrs[r]`, `flag_ptrs[r]`: rank r's board as mapped into this process), then raise this rank's
    arrival flag (= epoch) on every board.  See include/macarons_b200.h, "Fused score all-gather"."""
    pts, harmonics, X_cam, B, P, D, C = _prep(pts, harmonics, X_cam)
    board = _lib.PeerBoard()
    board.world, board.rank, board.epoch = len(score_ptrs), int(rank), int(epoch) & 0xFFFFFFFF
    for r, (sp, fp) in enumerate(zip(score_ptrs, flag_ptrs)):
        board.scores[r] = sp
        board.flags[r] = fp
    lib = _lib.load()
    with torch.cuda.device(pts.device):
        ws = _workspace(pts.device, B, C)
        _lib.check(lib.mac_covgain_push_f32(pts.data_ptr(), D, harmonics.data_ptr(), X_cam.data_ptr(), B, P, C,
                                            int(cam_range[0]), int(cam_range[1]),
                                            ACT_SIGMOID if use_sigmoid else ACT_RELU, ws.data_ptr(), ws.numel(),
                                            board, _stream_ptr(pts.device)))


def gather_wait_argmax(scores, flags, world, epoch, best, status):
    """Enqueue the wait-for-all-ranks + argmax kernel on the local score board."""
    B, C = scores.shape
    lib = _lib.load()
    with torch.cuda.device(scores.device):
        _lib.check(lib.mac_gather_wait_argmax(scores.data_ptr(), flags.data_ptr(), int(world), int(epoch) & 0xFFFFFFFF,
                                              B, C, best.data_ptr(), status.data_ptr(), _stream_ptr(scores.device)))


def coverage_gain_host(pts, harmonics, X_cam, use_sigmoid=True, cam_range=None, device=0, out=None):
    """Host-buffer entry point (numpy float32 arrays or CPU tensors in, numpy out): what a caller
    without torch-on-GPU uses; copies are done inside the C call."""
    import numpy as np

    def as_np(x):
        if isinstance(x, torch.Tensor):
            x = x.numpy()
        return np.ascontiguousarray(x, dtype=np.float32)

    pts, harmonics, X_cam = as_np(pts), as_np(harmonics), as_np(X_cam)
    B, P, D = pts.shape
    C = X_cam.shape[1]
    if harmonics.shape != (B, P, N_HARMONICS) or X_cam.shape != (B, C, 3):
        raise ValueError("inconsistent shapes")
    c0, c1 = (0, C) if cam_range is None else (int(cam_range[0]), int(cam_range[1]))
    if out is None:
        out = np.zeros((B, C), dtype=np.float32)
    lib = _lib.load()
    _lib.check(lib.mac_covgain_host(pts.ctypes.data, D, harmonics.ctypes.data, X_cam.ctypes.data, out.ctypes.data,
                                    B, P, C, c0, c1, ACT_SIGMOID if use_sigmoid else ACT_RELU, int(device)))
    return out


def launch_count():
    return int(_lib.load().mac_launch_count())


# ---- tensor-core linear layer (csrc/linear.cu) ----------------------------------------------------
ACT_NONE, ACT_LIN_RELU, ACT_GELU = 0, 1, 2


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def linear(x, packed, act=ACT_NONE, residual=None, out=None, ln=None, pool=0, K=None):
    """out = residual + act(x @ W^T + b) through mac_linear_f32.

    x (M, >=K) fp32 CUDA with row stride % 4 == 0; `packed` a packing.PackedLinear; `ln=(gamma, beta, eps)`
    additionally returns LayerNorm(out); `pool=16` returns (M/16, 2N) = [max | mean] over 16-row groups;
    `out` may be a column slice of a wider buffer (last-dim stride 1)."""
    _require_cuda_f32("x", x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be 2-D with unit column stride")
    M = x.shape[0]
    K = packed.K if K is None else K
    N = packed.N
    if x.shape[1] < K:
        raise ValueError("x has %d columns, the layer needs %d" % (x.shape[1], K))
    dev = x.device
    Np = (N + 3) // 4 * 4   # rows padded to 16 bytes: the epilogue stores float4, TMA reads need it as well
    if out is None:
        out = torch.empty((M // 16, 2 * N), dtype=torch.float32, device=dev) if pool else \
            torch.empty((M, Np), dtype=torch.float32, device=dev)[:, :N]
    ln_out = None
    if ln is not None:
        ln_out = torch.empty((M, Np), dtype=torch.float32, device=dev)[:, :N]
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mac_linear_f32(
            x.data_ptr(), x.stride(0), packed.hi.data_ptr(), _ptr(packed.lo), packed.ldw, _ptr(packed.bias),
            out.data_ptr(), out.stride(0), M, N, K, int(act),
            _ptr(residual), 0 if residual is None else residual.stride(0),
            _ptr(ln_out), 0 if ln_out is None else ln_out.stride(0),
            0 if ln is None else ln[0].data_ptr(), 0 if ln is None else ln[1].data_ptr(),
            0.0 if ln is None else float(ln[2]), int(pool), _stream_ptr(dev)))
    return (out, ln_out) if ln is not None else out


def linear_lnio(x, packed, act=ACT_NONE, residual=None, ln_in=None, stats_out=False, eps=1e-5):
    """mac_linear_lnio_f32: `ln_in=(stats (M,2), gamma (K), beta (K))` normalises the rows of x on load;
    `stats_out=True` also returns the (mean, rstd) (M,2) of the output rows."""
    _require_cuda_f32("x", x)
    M, K, N = x.shape[0], packed.K, packed.N
    dev = x.device
    out = torch.empty((M, (N + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :N]
    stats = torch.empty((M, 2), dtype=torch.float32, device=dev) if stats_out else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mac_linear_lnio_f32(
            x.data_ptr(), x.stride(0), packed.hi.data_ptr(), _ptr(packed.lo), packed.ldw, _ptr(packed.bias), out.data_ptr(),
            out.stride(0), M, N, K, int(act), _ptr(residual), 0 if residual is None else residual.stride(0),
            0 if ln_in is None else ln_in[0].data_ptr(), 0 if ln_in is None else ln_in[1].data_ptr(),
            0 if ln_in is None else ln_in[2].data_ptr(), _ptr(stats), float(eps), _stream_ptr(dev)))
    return (out, stats) if stats_out else out


# ---- kNN and the fused network forwards (csrc/pointnet.cu, csrc/scone_nets.cu) --------------------
_net_workspaces = {}


def _net_workspace(device, nbytes):
    key = (device.index, _stream_ptr(device))
    ws = _net_workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _net_workspaces[key] = ws
    return ws


def knn16(x, pc, return_dists=True):
    """x (B,Q,3), pc (B,N,3) -> idx (B,Q,16) int32 nearest first [, dists (B,Q,16)]."""
    _require_cuda_f32("x", x)
    _require_cuda_f32("pc", pc)
    if x.dim() != 3 or pc.dim() != 3 or x.shape[-1] != 3 or pc.shape[-1] != 3 or x.shape[0] != pc.shape[0]:
        raise ValueError("x must be (B,Q,3) and pc (B,N,3)")
    x, pc = x.contiguous(), pc.contiguous()
    B, Q, _ = x.shape
    N = pc.shape[1]
    idx = torch.empty((B, Q, 16), dtype=torch.int32, device=x.device)
    dists = torch.empty((B, Q, 16), dtype=torch.float32, device=x.device) if return_dists else None
    if Q > 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().mac_knn16_f32(x.data_ptr(), pc.data_ptr(), idx.data_ptr(), _ptr(dists), B, Q, N,
                                                 _stream_ptr(x.device)))
    return (idx, dists) if return_dists else idx


def sconevis_forward(w, pts, view_harmonics, lens=None):
    """Packed weights (netpack.SconeVisW), pts (B,S,4), view_harmonics (B,S,64) -> (B,S,64).
    `lens` (B,) int32 CUDA tensor: ragged batch, cloud b holds lens[b] tokens followed by zero padding rows."""
    import ctypes
    _require_cuda_f32("pts", pts)
    _require_cuda_f32("view_harmonics", view_harmonics)
    B, S, D = pts.shape
    if D != 4 or tuple(view_harmonics.shape) != (B, S, N_HARMONICS):
        raise ValueError("pts must be (B,S,4) and view_harmonics (B,S,64); got %s, %s"
                         % (tuple(pts.shape), tuple(view_harmonics.shape)))
    pts, view_harmonics = pts.contiguous(), view_harmonics.contiguous()
    out = torch.empty((B, S, N_HARMONICS), dtype=torch.float32, device=pts.device)
    if B * S == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(pts.device):
        ws = _net_workspace(pts.device, lib.mac_sconevis_workspace_bytes(B, S))
        if lens is None:
            _lib.check(lib.mac_sconevis_forward_f32(ctypes.byref(w), pts.data_ptr(), view_harmonics.data_ptr(), out.data_ptr(),
                                                    B, S, ws.data_ptr(), ws.numel(), _stream_ptr(pts.device)))
        else:
            if lens.dtype != torch.int32 or lens.device != pts.device or lens.numel() != B:
                raise ValueError("lens must be an int32 tensor of %d entries on %s" % (B, pts.device))
            lens = lens.contiguous()
            _lib.check(lib.mac_sconevis_forward_ragged_f32(ctypes.byref(w), pts.data_ptr(), view_harmonics.data_ptr(),
                                                           out.data_ptr(), B, S, lens.data_ptr(), ws.data_ptr(), ws.numel(),
                                                           _stream_ptr(pts.device)))
    return out


def sconeocc_forward(w, pc_global, pc_scales, x, view_harmonics, chunk=65536):
    """Packed weights (netpack.SconeOccW); pc_global (B,Sg,3); pc_scales: list of 3 (B,N_s,3) clouds;
    x (B,Q,3); view_harmonics (B,Q,64) -> (B,Q,1)."""
    import ctypes
    for name, t in (("pc_global", pc_global), ("x", x), ("view_harmonics", view_harmonics)):
        _require_cuda_f32(name, t)
    B, Q, _ = x.shape
    if tuple(view_harmonics.shape) != (B, Q, N_HARMONICS) or pc_global.shape[0] != B or pc_global.shape[2] != 3:
        raise ValueError("inconsistent SconeOcc input shapes")
    pc_global, x, view_harmonics = pc_global.contiguous(), x.contiguous(), view_harmonics.contiguous()
    pc_scales = [p.contiguous() for p in pc_scales]
    for p in pc_scales:
        _require_cuda_f32("pc_scale", p)
    out = torch.empty((B, Q, 1), dtype=torch.float32, device=x.device)
    if Q == 0:
        return out
    chunk = int(min(chunk, Q))
    Sg = pc_global.shape[1]
    ptrs = (ctypes.c_void_p * len(pc_scales))(*[p.data_ptr() for p in pc_scales])
    counts = (ctypes.c_int * len(pc_scales))(*[p.shape[1] for p in pc_scales])
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws = _net_workspace(x.device, lib.mac_sconeocc_workspace_bytes(B, Sg, chunk, Q))
        _lib.check(lib.mac_sconeocc_forward_f32(ctypes.byref(w), pc_global.data_ptr(), Sg, ptrs, counts, x.data_ptr(),
                                                view_harmonics.data_ptr(), out.data_ptr(), B, Q, chunk, ws.data_ptr(),
                                                ws.numel(), _stream_ptr(x.device)))
    return out


def sconeocc_forward_cells(w, pc_global, lens_g, pc_scales, scale_offs, x, view_harmonics, q_off, cell_of_q, max_q,
                           chunk=65536):
    """Ragged batch of cells through mac_sconeocc_forward_cells_f32 (see include/macarons_b200.h).  pc_global
    (n_cells,Sg,3); lens_g (n_cells) int32; pc_scales: 3 concatenated clouds (total_s,3); scale_offs: 3 int32 offset
    vectors (n_cells+1); x (Qtot,3); view_harmonics (Qtot,64); q_off (n_cells+1) int32; cell_of_q (Qtot) int32
    -> (Qtot,) occupancy values."""
    import ctypes
    for name, t in (("pc_global", pc_global), ("x", x), ("view_harmonics", view_harmonics)):
        _require_cuda_f32(name, t)
    n_cells, Sg, _ = pc_global.shape
    Qtot = x.shape[0]
    dev = x.device
    if tuple(x.shape) != (Qtot, 3) or tuple(view_harmonics.shape) != (Qtot, N_HARMONICS) or len(pc_scales) != 3 or len(scale_offs) != 3:
        raise ValueError("inconsistent SconeOcc cell-batch shapes")
    for t, n in ((lens_g, n_cells), (q_off, n_cells + 1), (cell_of_q, Qtot)) + tuple((o, n_cells + 1) for o in scale_offs):
        if t.dtype != torch.int32 or t.device != dev or t.numel() != n or not t.is_contiguous():
            raise ValueError("offset / length vectors must be contiguous int32 CUDA tensors of the right size")
    pc_global, x, view_harmonics = pc_global.contiguous(), x.contiguous(), view_harmonics.contiguous()
    pc_scales = [p.contiguous() for p in pc_scales]
    for p in pc_scales:
        _require_cuda_f32("pc_scale", p)
    out = torch.empty((Qtot,), dtype=torch.float32, device=dev)
    if Qtot == 0:
        return out
    chunk = int(min(chunk, Qtot))
    ptrs = (ctypes.c_void_p * 3)(*[p.data_ptr() for p in pc_scales])
    offs = (ctypes.c_void_p * 3)(*[o.data_ptr() for o in scale_offs])
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = _net_workspace(dev, lib.mac_sconeocc_cells_workspace_bytes(n_cells, Sg, chunk, Qtot))
        _lib.check(lib.mac_sconeocc_forward_cells_f32(ctypes.byref(w), n_cells, pc_global.data_ptr(), Sg, lens_g.data_ptr(),
                                                      ptrs, offs, x.data_ptr(), view_harmonics.data_ptr(), q_off.data_ptr(),
                                                      cell_of_q.data_ptr(), int(max_q), out.data_ptr(), Qtot, chunk,
                                                      ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
    return out


# ---- view state and proxy sampling (csrc/viewstate.cu, csrc/sampling.cu) --------------------------
def view_state(pts, X_view, n_elev, n_azim):
    """pts (B,P,>=3), X_view (V,3) -> (B,P,n_elev*n_azim) fp32 {0,1}."""
    _require_cuda_f32("pts", pts)
    _require_cuda_f32("X_view", X_view)
    if pts.dim() != 3 or pts.shape[-1] < 3 or X_view.dim() != 2 or X_view.shape[-1] != 3:
        raise ValueError("pts must be (B,P,>=3) and X_view (V,3)")
    pts, X_view = pts.contiguous(), X_view.contiguous()
    B, P, D = pts.shape
    out = torch.empty((B, P, n_elev * n_azim), dtype=torch.float32, device=pts.device)
    if B * P == 0:
        return out
    if X_view.shape[0] == 0:      # no camera visited yet: empty histogram
        return out.zero_()
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mac_view_state_f32(pts.data_ptr(), D, X_view.data_ptr(), out.data_ptr(), B, P,
                                                  X_view.shape[0], int(n_elev), int(n_azim), _stream_ptr(pts.device)))
    return out


def view_harmonics(state, base, h_polar, n_elev, n_azim):
    """state (B,P,n_bins), base (64,n_bins), h_polar (n_bins) -> (B,P,64)."""
    for name, t in (("view_state", state), ("base_harmonics", base), ("h_polar", h_polar)):
        _require_cuda_f32(name, t)
    n_bins = n_elev * n_azim
    if state.dim() != 3 or state.shape[-1] != n_bins or tuple(base.shape) != (N_HARMONICS, n_bins) or h_polar.numel() != n_bins:
        raise ValueError("view_state must be (B,P,%d), base_harmonics (64,%d), h_polar (%d,)" % (n_bins, n_bins, n_bins))
    state, base, h_polar = state.contiguous(), base.contiguous(), h_polar.contiguous()
    B, P, _ = state.shape
    out = torch.empty((B, P, N_HARMONICS), dtype=torch.float32, device=state.device)
    if B * P == 0:
        return out
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().mac_view_harmonics_f32(state.data_ptr(), base.data_ptr(), h_polar.data_ptr(), out.data_ptr(),
                                                      B, P, int(n_elev), int(n_azim), _stream_ptr(state.device)))
    return out


def view_state_harmonics(pts, X_view, base, h_polar, n_elev, n_azim):
    """Fused compute_view_state + compute_view_harmonics: pts (B,P,>=3), X_view (V,3), base (64,n_bins),
    h_polar (n_bins) -> (B,P,64), bitwise equal to view_harmonics(view_state(...)) without the (B,P,n_bins) tensor."""
    for name, t in (("pts", pts), ("X_view", X_view), ("base_harmonics", base), ("h_polar", h_polar)):
        _require_cuda_f32(name, t)
    n_bins = n_elev * n_azim
    if pts.dim() != 3 or pts.shape[-1] < 3 or X_view.dim() != 2 or X_view.shape[-1] != 3:
        raise ValueError("pts must be (B,P,>=3) and X_view (V,3)")
    if tuple(base.shape) != (N_HARMONICS, n_bins) or h_polar.numel() != n_bins:
        raise ValueError("base_harmonics must be (64,%d) and h_polar (%d,)" % (n_bins, n_bins))
    pts, X_view, base, h_polar = pts.contiguous(), X_view.contiguous(), base.contiguous(), h_polar.contiguous()
    B, P, D = pts.shape
    out = torch.empty((B, P, N_HARMONICS), dtype=torch.float32, device=pts.device)
    if B * P == 0:
        return out
    if X_view.shape[0] == 0:
        return out.zero_()
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mac_viewstate_harm_f32(pts.data_ptr(), D, X_view.data_ptr(), base.data_ptr(), h_polar.data_ptr(),
                                                      out.data_ptr(), B, P, X_view.shape[0], int(n_elev), int(n_azim),
                                                      _stream_ptr(pts.device)))
    return out


def gather_bins(state, index):
    """state (B,P,n_bins), index (n_bins) int -> state[..., index]."""
    _require_cuda_f32("view_state", state)
    state = state.contiguous()
    index = index.to(device=state.device, dtype=torch.int32).contiguous()
    B, P, n_bins = state.shape
    out = torch.empty_like(state)
    if B * P == 0:
        return out
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().mac_gather_bins_f32(state.data_ptr(), index.data_ptr(), out.data_ptr(), B, P, n_bins,
                                                   _stream_ptr(state.device)))
    return out


def sample_proxy_points(X_world, preds, view_harmonics, u, min_occ):
    """(N,3), (N,1), (N,64), u (n_sample,) -> (res (U,4), res_harmonics (U,64), inverse (n_sample,) int64)."""
    for name, t in (("X_world", X_world), ("preds", preds), ("view_harmonics", view_harmonics), ("u", u)):
        _require_cuda_f32(name, t)
    N = X_world.shape[0]
    n_sample = u.numel()
    if tuple(X_world.shape) != (N, 3) or preds.numel() != N or tuple(view_harmonics.shape) != (N, N_HARMONICS):
        raise ValueError("expected X_world (N,3), preds (N,1), view_harmonics (N,64)")
    dev = X_world.device
    X_world, preds, view_harmonics, u = X_world.contiguous(), preds.contiguous(), view_harmonics.contiguous(), u.contiguous()
    res = torch.empty((n_sample, 4), dtype=torch.float32, device=dev)
    res_h = torch.empty((n_sample, N_HARMONICS), dtype=torch.float32, device=dev)
    inverse = torch.empty((n_sample,), dtype=torch.int64, device=dev)
    counts = torch.zeros((2,), dtype=torch.int32, device=dev)
    if N == 0:
        return res[:0], res_h[:0], inverse.zero_()
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = torch.empty(lib.mac_sample_proxy_workspace_bytes(N), dtype=torch.uint8, device=dev)
        _lib.check(lib.mac_sample_proxy_points_f32(X_world.data_ptr(), preds.data_ptr(), view_harmonics.data_ptr(),
                                                   u.data_ptr(), N, n_sample, float(min_occ), res.data_ptr(),
                                                   res_h.data_ptr(), inverse.data_ptr(), counts.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), _stream_ptr(dev)))
    n_unique = int(counts[1].item())   # the result has a data-dependent length (as torch.unique in the reference)
    return res[:n_unique], res_h[:n_unique], inverse


def fov_sample_proxy(X_world, preds, view_harmonics, cams, ndc_bounds, fov_range, min_occ, u):
    """Batched field-of-view / occupancy selection + proxy sampling for C candidate cameras (mac_fov_sample_proxy_f32).
    X_world (N,3), preds (N,1), view_harmonics (N,64), cams (C,36) [full projection | world-to-view | centre | pad],
    ndc_bounds 4 floats, u (C,n_sample) -> res (C,n_sample,4) and res_harmonics (C,n_sample,64) zero-padded after the
    unique picks, inverse (C,n_sample) int64, counts (C,2) int32 [kept, unique], volume (C,)."""
    import ctypes
    for name, t in (("X_world", X_world), ("preds", preds), ("view_harmonics", view_harmonics), ("cams", cams), ("u", u)):
        _require_cuda_f32(name, t)
    N = X_world.shape[0]
    C, n_sample = u.shape
    if tuple(X_world.shape) != (N, 3) or preds.numel() != N or tuple(view_harmonics.shape) != (N, N_HARMONICS) or \
            tuple(cams.shape) != (C, 36):
        raise ValueError("expected X_world (N,3), preds (N,1), view_harmonics (N,64), cams (C,36), u (C,n_sample)")
    dev = X_world.device
    X_world, preds, view_harmonics, cams, u = (t.contiguous() for t in (X_world, preds, view_harmonics, cams, u))
    res = torch.zeros((C, n_sample, 4), dtype=torch.float32, device=dev)
    res_h = torch.zeros((C, n_sample, N_HARMONICS), dtype=torch.float32, device=dev)
    inverse = torch.zeros((C, n_sample), dtype=torch.int64, device=dev)
    counts = torch.zeros((C, 2), dtype=torch.int32, device=dev)
    volume = torch.zeros((C,), dtype=torch.float32, device=dev)
    if N == 0 or C == 0:
        return res, res_h, inverse, counts, volume
    ndc = (ctypes.c_float * 4)(*[float(v) for v in ndc_bounds])
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = _net_workspace(dev, lib.mac_fov_sample_proxy_workspace_bytes(N, C))
        _lib.check(lib.mac_fov_sample_proxy_f32(X_world.data_ptr(), preds.data_ptr(), view_harmonics.data_ptr(), cams.data_ptr(),
                                                ctypes.cast(ndc, ctypes.c_void_p), -1.0 if fov_range is None else float(fov_range),
                                                float(min_occ), u.data_ptr(), N, C, n_sample, res.data_ptr(), res_h.data_ptr(),
                                                inverse.data_ptr(), counts.data_ptr(), volume.data_ptr(), ws.data_ptr(),
                                                ws.numel(), _stream_ptr(dev)))
    return res, res_h, inverse, counts, volume


def points_in_fov(X_world, cam_row, ndc_bounds, fov_range):
    """X_world (N,3), cam_row (36,) [full projection | world-to-view | centre | pad] -> bool mask (N,)."""
    import ctypes
    _require_cuda_f32("X_world", X_world)
    _require_cuda_f32("cam_row", cam_row)
    if X_world.dim() != 2 or X_world.shape[1] != 3 or cam_row.numel() != 36:
        raise ValueError("expected X_world (N,3) and a 36-float camera row")
    X_world, cam_row = X_world.contiguous(), cam_row.contiguous()
    N = X_world.shape[0]
    mask = torch.empty((N,), dtype=torch.uint8, device=X_world.device)
    ndc = (ctypes.c_float * 4)(*[float(v) for v in ndc_bounds])
    with torch.cuda.device(X_world.device):
        _lib.check(_lib.load().mac_points_in_fov_f32(X_world.data_ptr(), cam_row.data_ptr(), ctypes.cast(ndc, ctypes.c_void_p),
                                                     -1.0 if fov_range is None else float(fov_range), N, mask.data_ptr(),
                                                     _stream_ptr(X_world.device)))
    return mask.bool()


def cell_min_dist(pts, slot, stored, off):
    """pts (M,3) f32, slot (M) int32, stored (E,3) f32, off (n_slots+1) int32 -> (M,) float64 distance of every point to
    the nearest stored point of its own cell slot (+inf for empty slots)."""
    _require_cuda_f32("pts", pts)
    M = pts.shape[0]
    out = torch.empty((M,), dtype=torch.float64, device=pts.device)
    if M == 0:
        return out
    pts, slot, off = pts.contiguous(), slot.contiguous(), off.contiguous()
    stored = stored.contiguous() if stored.numel() else torch.zeros((1, 3), dtype=torch.float32, device=pts.device)
    if slot.dtype != torch.int32 or off.dtype != torch.int32 or stored.dtype != torch.float32:
        raise TypeError("slot / off must be int32 and stored float32")
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mac_cell_min_dist_f64(pts.data_ptr(), slot.data_ptr(), stored.data_ptr(), off.data_ptr(),
                                                     out.data_ptr(), M, _stream_ptr(pts.device)))
    return out


def unproject_depth(depth, cams, H, W):
    """depth (B, H*W) metric depth, cams (B, 18) [inverse full projection | f1 | f2] -> world points (B, H*W, 3)."""
    _require_cuda_f32("depth", depth)
    _require_cuda_f32("cams", cams)
    B = depth.shape[0]
    if depth.numel() != B * H * W or tuple(cams.shape) != (B, 18):
        raise ValueError("expected depth with B*H*W elements and cams (B, 18)")
    depth, cams = depth.contiguous(), cams.contiguous()
    out = torch.empty((B, H * W, 3), dtype=torch.float32, device=depth.device)
    with torch.cuda.device(depth.device):
        _lib.check(_lib.load().mac_unproject_depth_f32(depth.data_ptr(), cams.data_ptr(), out.data_ptr(), B, int(H), int(W),
                                                       _stream_ptr(depth.device)))
    return out


def signed_distance(pts, depth_maps, mask, cams, H, W, fill):
    """pts (P,3), depth_maps (n,H,W), mask (n,H,W) bool, cams (n,32) [full projection | world-to-view] -> (n, P)."""
    for name, t in (("pts", pts), ("depth_maps", depth_maps), ("cams", cams)):
        _require_cuda_f32(name, t)
    n, P = depth_maps.shape[0], pts.shape[0]
    if tuple(pts.shape) != (P, 3) or depth_maps.numel() != n * H * W or mask.numel() != n * H * W or tuple(cams.shape) != (n, 32):
        raise ValueError("expected pts (P,3), depth_maps / mask (n,H,W), cams (n,32)")
    pts, depth_maps, cams = pts.contiguous(), depth_maps.contiguous(), cams.contiguous()
    mask = mask.to(device=pts.device, dtype=torch.uint8).contiguous()
    out = torch.empty((n, P), dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        _lib.check(_lib.load().mac_signed_distance_f32(pts.data_ptr(), depth_maps.data_ptr(), mask.data_ptr(), cams.data_ptr(),
                                                       out.data_ptr(), n, P, int(H), int(W), float(fill), _stream_ptr(pts.device)))
    return out


def manydepth_forward(w, x, x_alpha, cam):
    """Packed weights (netpack.ManyDepthW); x (B,3,H,W), x_alpha (B,n_alpha,3,H,W), cam (B,1+n_alpha,13)
    -> (disp1 (B,1,H,W), disp2, disp3, disp4)."""
    import ctypes
    for name, t in (("x", x), ("x_alpha", x_alpha), ("cam", cam)):
        _require_cuda_f32(name, t)
    B, C, H, W = x.shape
    n_alpha = x_alpha.shape[1]
    if C != 3 or tuple(x_alpha.shape) != (B, n_alpha, 3, H, W) or tuple(cam.shape) != (B, 1 + n_alpha, 13):
        raise ValueError("expected x (B,3,H,W), x_alpha (B,n_alpha,3,H,W), cam (B,1+n_alpha,13)")
    x, x_alpha, cam = x.contiguous(), x_alpha.contiguous(), cam.contiguous()
    dev = x.device
    sizes = [(H, W)] + [(H // d, W // d + (1 if W % d else 0)) for d in (2, 4, 8)]
    disps = [torch.empty((B, 1, h, w_), dtype=torch.float32, device=dev) for h, w_ in sizes]
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = lib.mac_manydepth_workspace_bytes(ctypes.byref(w), B, n_alpha, H, W)
        if need == 0:
            raise ValueError("unsupported depth input shape %s" % (tuple(x.shape),))
        ws = _net_workspace(dev, need)
        _lib.check(lib.mac_manydepth_forward_f32(ctypes.byref(w), x.data_ptr(), x_alpha.data_ptr(), cam.data_ptr(),
                                                 disps[0].data_ptr(), disps[1].data_ptr(), disps[2].data_ptr(),
                                                 disps[3].data_ptr(), B, n_alpha, H, W, ws.data_ptr(), ws.numel(),
                                                 _stream_ptr(dev)))
    return tuple(disps)
