"""SconeVis: visibility-gain prediction (reference macarons/networks/SconeVis.py:6-303).

Same constructor, attribute tree / state_dict keys and method signatures as the reference class.
The three SH-integration methods run the hand-written sm_100a coverage-gain kernel
(csrc/covgain.cu through the C ABI); they need CUDA tensors and raise otherwise (no CPU path).
"""
import torch
from torch import nn

from .Attention import Embedding, Encoder
from .. import netpack, ops
from ..utility.spherical_harmonics import clear_spherical_harmonics_cache


class SconeVis(nn.Module):
    def __init__(self, pts_dim=4, seq_len=2048, pts_embedding_dim=256, n_heads=4, n_code=3, n_harmonics=64,
                 max_harmonic_rank=8, FF=True, gelu=True, dropout=None, use_view_state=True,
                 use_global_feature=True, view_state_mode="end", concatenate_input=True, k_for_knn=0,
                 alt=False, use_sigmoid=True):
        super().__init__()
        if n_harmonics != 64 or max_harmonic_rank != 8:
            # the reference hard-codes 64 in its projection (SconeVis.py:241); so does the kernel
            raise ValueError("SconeVis on sm_100a supports n_harmonics=64, max_harmonic_rank=8")
        self.pts_dim, self.seq_len, self.pts_embedding_dim = pts_dim, seq_len, pts_embedding_dim
        self.n_heads, self.n_code, self.dropout = n_heads, n_code, dropout
        self.n_harmonics, self.max_harmonic_rank = n_harmonics, max_harmonic_rank
        self.use_view_state, self.use_global_feature = use_view_state, use_global_feature
        self.view_state_mode, self.alt, self.use_sigmoid = view_state_mode, alt, use_sigmoid
        print("Use sigmoid in model." if use_sigmoid else "Use ReLU for output in model.")

        extra = n_harmonics if (use_view_state and view_state_mode == "start") else 0
        self.embedding = Embedding(pts_dim, pts_embedding_dim, gelu=gelu, global_feature=use_global_feature,
                                   additional_feature_dim=extra, concatenate_input=concatenate_input,
                                   k_for_knn=k_for_knn, dropout=None)
        self.encoders = nn.ModuleList(
            Encoder(seq_len=seq_len, embedding_dim=pts_embedding_dim, qk_dim=pts_embedding_dim // 4,
                    n_heads=n_heads, dropout=dropout, gelu=gelu, FF=FF) for _ in range(n_code))
        self.norm = nn.LayerNorm(pts_embedding_dim)

        if alt:
            fc1_in, factor = pts_embedding_dim + n_harmonics, 4
        else:
            fc1_in = pts_embedding_dim
            factor = 3 if (use_view_state and view_state_mode == "end") else 4
        self.fc1 = nn.Linear(fc1_in, factor * n_harmonics)
        self.nonlinear1 = nn.GELU()
        self.fc2 = nn.Linear(4 * n_harmonics, 2 * n_harmonics)
        self.nonlinear2 = nn.GELU()
        self.fc3 = nn.Linear(2 * n_harmonics, n_harmonics)

    # ---- a6: per-point SH coefficients of the visibility-gain function (SconeVis.py:121-162) ----
    def forward(self, pts, mask=None, view_harmonics=None):
        """pts (B,S,4) [xyz, occupancy], view_harmonics (B,S,64) -> (B,S,64): one fused CUDA forward
        (csrc/scone_nets.cu: tcgen05 linear layers + dense attention kernel)."""
        if mask is not None:
            raise NotImplementedError("attention masks are never used on the NBV path (SURVEY.md A.4)")
        if view_harmonics is None:
            raise NameError("view_harmonics is required (use_view_state=True, view_state_mode='end')")
        if self.training and self.dropout is not None:
            raise NotImplementedError("the fused SconeVis forward has no dropout: call .eval() (dropout=%r)" % self.dropout)
        if ops.wants_grad(pts, view_harmonics, module=self):
            # training: CUDA forward, backward by recomputation under torch.autograd (networks/_backward.py)
            from ._backward import SconeVisFunction
            return SconeVisFunction.apply(self, pts, view_harmonics, *self.parameters())
        return ops.sconevis_forward(netpack.pack_sconevis(self), pts, view_harmonics)

    # ---- a3/a4: SH integration over candidate cameras: CUDA kernel --------------------------------
    def compute_visibilities(self, pts, harmonics, X_cam):
        """(B,P,pts_dim), (B,P,64), (B,C,3) -> (B,C,P)   [reference SconeVis.py:164-208]"""
        clear_spherical_harmonics_cache()
        return ops.sh_integration(pts, harmonics, X_cam, use_sigmoid=self.use_sigmoid, per_point=True)

    def compute_coverage_gain(self, pts, harmonics, X_cam, cam_range=None):
        """(B,P,3|4), (B,P,64), (B,C,3) -> (B,C)   [reference SconeVis.py:210-252].
        `cam_range` (extension) scores a slice of the camera axis, see macarons_b200.parallel."""
        clear_spherical_harmonics_cache()
        return ops.sh_integration(pts, harmonics, X_cam, use_sigmoid=self.use_sigmoid, cam_range=cam_range)

    def compute_coverage_gain_multiple(self, pts, harmonics, X_cam, n_cam):
        """Coverage of every ordered n_cam-tuple of cameras: per-point max over the tuple, then mean
        -> ((B, C**n_cam), (C**n_cam, n_cam))   [reference SconeVis.py:254-303]."""
        if n_cam not in (2, 3):
            raise NameError("n_cam is too large.")
        z = self.compute_visibilities(pts, harmonics, X_cam)
        single = torch.arange(0, X_cam.shape[1])
        tuples = torch.cartesian_prod(*([single] * n_cam))
        best = z[:, tuples.to(z.device)].max(dim=-2)[0]
        return best.sum(dim=-1) / pts.shape[1], tuples
