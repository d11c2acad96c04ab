"""SconeOcc: occupancy-probability field (reference macarons/networks/SconeOcc.py:7-347).

Same constructors, attribute tree / state_dict keys and forward signature as the reference classes.  The
forward pass draws the reference's random sub-samples on the host (torch.randperm on the global CPU generator,
in the reference's order, SconeOcc.py:269 and :311) and then runs ONE fused CUDA forward through the C ABI
(csrc/scone_nets.cu): global transformer, 3 x (kNN + 16-token neighbourhood transformer), x embedding and the
occupancy head.  CUDA tensors only; there is no CPU path.
"""
import numpy as np
import torch
from torch import nn

from .Attention import Embedding, Encoder, _FusedOnly, _activation
from .. import netpack, ops


class XEmbedding(_FusedOnly):
    """reference SconeOcc.py:7-42: 3 -> d/4 -> d/2 -> d with a non-linearity after every layer."""

    def __init__(self, x_dim, x_embedding_dim, dropout=None, gelu=True):
        super().__init__()
        self.linear1 = nn.Linear(x_dim, x_embedding_dim // 4)
        self.linear2 = nn.Linear(x_embedding_dim // 4, x_embedding_dim // 2)
        self.linear3 = nn.Linear(x_embedding_dim // 2, x_embedding_dim)
        self.non_linear1, self.non_linear2, self.non_linear3 = _activation(gelu), _activation(gelu), _activation(gelu)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None


class PCTransformer(_FusedOnly):
    """reference SconeOcc.py:45-130: embedding -> n_code encoders -> LayerNorm -> linear0 -> [max | avg] pool."""

    def __init__(self, seq_len, pts_dim=3, pts_embedding_dim=256, feature_dim=512, concatenate_input=True,
                 n_code=2, n_heads=4, FF=True, gelu=True, dropout=None):
        super().__init__()
        self.seq_len, self.pts_dim, self.pts_embedding_dim = seq_len, pts_dim, pts_embedding_dim
        self.n_code, self.n_heads, self.FF, self.gelu = n_code, n_heads, FF, gelu
        self.feature_dim, self.dropout = feature_dim, dropout
        self.embedding = Embedding(input_dim=pts_dim, output_dim=pts_embedding_dim, dropout=None, gelu=gelu,
                                   global_feature=False, additional_feature_dim=0,
                                   concatenate_input=concatenate_input, k_for_knn=0)
        self.encoders = nn.ModuleList(
            Encoder(seq_len=seq_len, embedding_dim=pts_embedding_dim, qk_dim=pts_embedding_dim // 4, n_heads=n_heads,
                    dropout=dropout, gelu=gelu, FF=FF) for _ in range(n_code))
        self.norm = nn.LayerNorm(pts_embedding_dim)
        self.linear0 = nn.Linear(pts_embedding_dim, feature_dim // 2)


class SconeOcc(nn.Module):
    def __init__(self, seq_len=2048, pts_dim=3, pts_embedding_dim=128, concatenate_input=True, n_code=2, n_heads=4,
                 FF=True, gelu=True, global_feature_dim=512, n_scale=3, local_feature_dim=256, k_for_knn=16,
                 x_dim=3, x_embedding_dim=512, n_harmonics=64, output_dim=1, dropout=None, offset=True):
        super().__init__()
        self.seq_len, self.pts_dim, self.pts_embedding_dim = seq_len, pts_dim, pts_embedding_dim
        self.n_code, self.n_heads, self.FF, self.gelu = n_code, n_heads, FF, gelu
        self.n_scale, self.x_dim, self.x_embedding_dim = n_scale, x_dim, x_embedding_dim
        self.output_dim, self.dropout, self.encoding_dim = output_dim, dropout, pts_embedding_dim
        self.k_for_knn, self.offset = k_for_knn, offset
        if offset:
            print("Offset set to True.")
        self.global_feature_dim, self.local_feature_dim = global_feature_dim, local_feature_dim
        self.all_feature_size = x_embedding_dim + n_scale * local_feature_dim + global_feature_dim + n_harmonics

        self.global_transformer = PCTransformer(seq_len=seq_len, pts_dim=pts_dim, pts_embedding_dim=pts_embedding_dim,
                                                feature_dim=global_feature_dim, concatenate_input=concatenate_input,
                                                n_code=n_code, n_heads=n_heads, FF=FF, gelu=gelu, dropout=dropout)
        self.local_transformers = nn.ModuleList(
            PCTransformer(seq_len=k_for_knn, pts_dim=pts_dim, pts_embedding_dim=pts_embedding_dim,
                          feature_dim=local_feature_dim, concatenate_input=concatenate_input, n_code=n_code,
                          n_heads=n_heads, FF=FF, gelu=gelu, dropout=dropout) for _ in range(n_scale))
        self.x_embedding = XEmbedding(x_dim=x_dim, x_embedding_dim=x_embedding_dim, dropout=dropout, gelu=gelu)
        self.linear1 = nn.Linear(self.all_feature_size, 512)
        self.linear2 = nn.Linear(512, 256)
        self.linear3 = nn.Linear(256, output_dim)
        self.non_linear1, self.non_linear2, self.non_linear3 = _activation(gelu), _activation(gelu), _activation(gelu)
        # internal chunking of the fused forward (results do not depend on it): 65536 queries = 1 M neighbourhood tokens
        # per launch, ~4.6 GB of workspace; measured 77.1 ms per 64^3 grid vs 79.3 ms at 16384 (fewer launch ramps / tails)
        self.queries_per_pass = 65536

    def draw_subsamples(self, full_seq_len):
        """The random sub-samples of one forward call, drawn like the reference does (CPU generator, same
        sizes, same order): indices for the global transformer (SconeOcc.py:269) and for the successive
        down-samplings of the kNN cloud (:282-288, :311; ds_factor truncation with 0 -> 2)."""
        global_idx = torch.randperm(full_seq_len)[:self.seq_len]
        if self.n_scale > 1:
            ds_factor = int(np.power(full_seq_len / (self.k_for_knn * 8), 1. / (self.n_scale - 1)))
            if ds_factor == 0:
                ds_factor = 2
        else:
            ds_factor = 1
        scale_idx, n = [], full_seq_len
        for _ in range(self.n_scale - 1):
            scale_idx.append(torch.randperm(n)[:n // ds_factor])
            n = n // ds_factor
        return global_idx, scale_idx

    def forward(self, pc, x, view_harmonics, mask=None, verbose=False):
        """pc (B,N,3), x (B,Q,3), view_harmonics (B,Q,64) -> (B,Q,1)   [reference SconeOcc.py:250-347]"""
        if mask is not None:
            raise NotImplementedError("attention masks are never used on the NBV path (SURVEY.md A.4)")
        if self.training and self.dropout is not None:
            raise NotImplementedError("the fused SconeOcc forward has no dropout: call .eval() (dropout=%r)" % self.dropout)
        global_idx, scale_idx = self.draw_subsamples(pc.shape[1])
        clouds = [pc]
        for idx in scale_idx:
            clouds.append(clouds[-1][:, idx.to(pc.device)])
        pc_global = pc[:, global_idx.to(pc.device)]
        if ops.wants_grad(pc, x, view_harmonics, module=self):
            # training: CUDA forward, backward by recomputation under torch.autograd (networks/_backward.py)
            from ._backward import SconeOccFunction
            out = SconeOccFunction.apply(self, pc_global.contiguous(), *[c.contiguous() for c in clouds], x, view_harmonics,
                                         *self.parameters())
            return out.view(pc.shape[0], x.shape[1], self.output_dim)
        out = ops.sconeocc_forward(netpack.pack_sconeocc(self), pc_global, clouds, x, view_harmonics,
                                   chunk=self.queries_per_pass)
        return out.view(pc.shape[0], x.shape[1], self.output_dim)

    def forward_cells(self, clouds, queries, view_harmonics):
        """Batched form of a loop of forward calls over cells with their own clouds (extension; the reference calls the
        network once per occupied cell, utility/macarons_utils.py:1443-1518): `clouds[c]` (N_c, 3), `queries[c]`
        (Q_c, 3), `view_harmonics[c]` (Q_c, 64) -> list of (Q_c, 1) outputs, those of `self(clouds[c][None],
        queries[c][None], view_harmonics[c][None])[0]` called in the same order: the random sub-samples of every cell
        are drawn here on the host in exactly that order (`draw_subsamples`), then ONE ragged CUDA forward runs."""
        n_cells = len(clouds)
        if n_cells == 0:
            return []
        if self.training and self.dropout is not None:
            raise NotImplementedError("the fused SconeOcc forward has no dropout: call .eval() (dropout=%r)" % self.dropout)
        ops.refuse_grad("SconeOcc.forward_cells", *clouds, *queries, *view_harmonics, module=self)
        dev = clouds[0].device
        sizes = [int(c.shape[0]) for c in clouds]
        n_q = [int(q.shape[0]) for q in queries]
        cloud_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        Sg = min(self.seq_len, max(sizes))
        # host side: index arithmetic only (which rows of the concatenated cloud every sub-sample consists of)
        g_rows = np.zeros((n_cells, Sg), dtype=np.int64)
        g_len = np.zeros(n_cells, dtype=np.int32)
        scale_rows = [[], []]
        scale_off = [np.zeros(n_cells + 1, dtype=np.int64) for _ in range(2)]
        for c, n in enumerate(sizes):
            global_idx, scale_idx = self.draw_subsamples(n)
            g_len[c] = global_idx.numel()
            g_rows[c, :g_len[c]] = cloud_off[c] + global_idx.numpy()
            rows = np.arange(n, dtype=np.int64)
            for s, idx in enumerate(scale_idx):
                rows = rows[idx.numpy()]
                scale_rows[s].append(cloud_off[c] + rows)
                scale_off[s][c + 1] = scale_off[s][c] + rows.shape[0]
        as_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
        all_pc = torch.cat([c.reshape(-1, 3) for c in clouds]).to(torch.float32)
        pc_global = all_pc.index_select(0, as_dev(g_rows.reshape(-1), torch.int64)).view(n_cells, Sg, 3)
        pad = np.arange(Sg)[None, :] >= g_len[:, None]
        if pad.any():
            pc_global = pc_global * as_dev(~pad, torch.float32).unsqueeze(-1)          # zero rows behind lens_g[c]
        pc_scales = [all_pc] + [all_pc.index_select(0, as_dev(np.concatenate(r), torch.int64)) for r in scale_rows]
        scale_offs = [as_dev(cloud_off, torch.int32)] + [as_dev(o, torch.int32) for o in scale_off]
        q_off = np.concatenate(([0], np.cumsum(n_q))).astype(np.int64)
        x = torch.cat([q.reshape(-1, 3) for q in queries]).to(torch.float32)
        vh = torch.cat([v.reshape(-1, view_harmonics[0].shape[-1]) for v in view_harmonics]).to(torch.float32)
        cell_of_q = as_dev(np.repeat(np.arange(n_cells, dtype=np.int32), n_q), torch.int32)
        out = ops.sconeocc_forward_cells(netpack.pack_sconeocc(self), pc_global, as_dev(g_len, torch.int32), pc_scales,
                                         scale_offs, x, vh, as_dev(q_off, torch.int32), cell_of_q, max(n_q),
                                         chunk=self.queries_per_pass)
        return [out[q_off[c]:q_off[c + 1]].view(-1, self.output_dim) for c in range(n_cells)]
