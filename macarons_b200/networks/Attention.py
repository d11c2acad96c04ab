"""Building blocks shared by SconeOcc and SconeVis: point embedding, multi-head self-attention,
feed-forward and the pre-LayerNorm encoder (reference macarons/networks/Attention.py:39-300).

These classes hold the PARAMETERS with the reference's attribute names and shapes, so reference
checkpoints load unchanged:

  Embedding:  linear1 (inner, in)  linear2 (feat, inner)
  Encoder:    norm1, mhsa.{w_q, w_k, w_v, out}, norm2, ff.{linear1, linear2}

The arithmetic does not live here: SconeOcc.forward / SconeVis.forward pack these parameters
(macarons_b200/netpack.py) and run the whole network through the C ABI (csrc/scone_nets.cu: tcgen05 linear
layers + attention kernels).  There is no torch / CPU implementation of the blocks in this package; calling a
block on its own raises.
"""
from torch import nn


class _FusedOnly(nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "%s is a parameter container: its arithmetic runs inside the fused CUDA forward of SconeOcc / SconeVis "
            "(macarons_b200 has no stand-alone torch implementation of the block)" % type(self).__name__)


def _activation(gelu):
    return nn.GELU() if gelu else nn.ReLU(inplace=False)


class Embedding(_FusedOnly):
    """Per-point MLP embedding, optionally concatenated with a max-pooled global feature, an
    additional per-point feature and the raw input (reference Attention.py:39-128)."""

    def __init__(self, input_dim, output_dim, dropout=None, gelu=True, global_feature=False,
                 additional_feature_dim=0, concatenate_input=True, k_for_knn=0):
        super().__init__()
        if k_for_knn > 0:
            # reference option (pytorch3d knn_points inside the embedding); never enabled on the NBV path
            raise NotImplementedError("Embedding(k_for_knn>0) is outside the NBV scoring path")
        self.use_knn = False
        self.k = 0
        self.input_dim = input_dim
        self.global_feature = global_feature
        self.additional_feature_dim = additional_feature_dim
        self.concatenate_input = concatenate_input
        self.gelu = gelu

        feat, inner = output_dim, output_dim // 2
        if additional_feature_dim > 0:
            feat -= additional_feature_dim
            inner = feat
        if concatenate_input:
            feat -= input_dim
            inner = feat
        if global_feature:
            feat //= 2
            inner = feat
        self.inner_dim, self.feature_dim = inner, feat

        self.linear1 = nn.Linear(input_dim, inner)
        self.linear2 = nn.Linear(inner, feat)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.nonlinear = _activation(gelu)


class MultiHeadSelfAttention(_FusedOnly):
    """reference Attention.py:131-204"""

    def __init__(self, n_heads, in_dim, qk_dim, dropout=None):
        super().__init__()
        self.n_heads, self.in_dim, self.qk_dim, self.v_dim = n_heads, in_dim, qk_dim, in_dim
        self.qk_dim_per_head = qk_dim // n_heads
        self.v_dim_per_head = in_dim // n_heads
        self.w_q = nn.Linear(in_dim, qk_dim)
        self.w_k = nn.Linear(in_dim, qk_dim)
        self.w_v = nn.Linear(in_dim, in_dim)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        if n_heads > 1:
            self.out = nn.Linear(in_dim, in_dim)


class FeedForward(_FusedOnly):
    """reference Attention.py:207-236"""

    def __init__(self, input_dim, inner_dim, gelu=True, dropout=None):
        super().__init__()
        self.gelu = gelu
        self.linear1 = nn.Linear(input_dim, inner_dim)
        self.linear2 = nn.Linear(inner_dim, input_dim)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.nonlinear = _activation(gelu)


class Encoder(_FusedOnly):
    """Pre-LN transformer encoder block (reference Attention.py:239-300)."""

    def __init__(self, seq_len, qk_dim, embedding_dim=128, n_heads=1, dropout=None, gelu=True, FF=True):
        super().__init__()
        self.seq_len, self.embedding_dim, self.n_heads, self.qk_dim = seq_len, embedding_dim, n_heads, qk_dim
        self.dropout, self.FF = dropout, FF
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.mhsa = MultiHeadSelfAttention(n_heads=n_heads, in_dim=embedding_dim, qk_dim=qk_dim, dropout=dropout)
        self.dropout1 = None if dropout is None else nn.Dropout(dropout)
        if FF:
            self.norm2 = nn.LayerNorm(embedding_dim)
            self.ff = FeedForward(embedding_dim, 2 * embedding_dim, gelu=gelu, dropout=dropout)
            self.dropout2 = None if dropout is None else nn.Dropout(dropout)
