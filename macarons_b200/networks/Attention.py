"""Building blocks shared by SconeOcc and SconeVis: point embedding, multi-head self-attention,
feed-forward and the pre-LayerNorm encoder.  Mirrors the module tree (attribute names, parameter
shapes) of reference macarons/networks/Attention.py:39-300 so reference checkpoints load unchanged:

  Embedding:  linear1 (inner, in)  linear2 (feat, inner)
  Encoder:    norm1, mhsa.{w_q, w_k, w_v, out}, norm2, ff.{linear1, linear2}

Numerical contract reproduced from the reference (Attention.py:8-36): scores = q k^T, optional
mask fill with -1e3 BEFORE the 1/sqrt(d) scale, softmax over keys, then @ v.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F


def attention(q, k, v, mask=None, dropout=None):
    """q,k (..., N, d), v (..., N, dv) -> (..., N, dv)   [reference Attention.py:8-36]"""
    logits = torch.matmul(q, k.transpose(-2, -1))
    if mask is not None:
        logits = logits.masked_fill(mask == 0, -1e3)
    weights = F.softmax(logits / math.sqrt(q.shape[-1]), dim=-1)
    if dropout is not None:
        weights = dropout(weights)
    return torch.matmul(weights, v)


def _activation(gelu):
    return nn.GELU() if gelu else nn.ReLU(inplace=False)


class Embedding(nn.Module):
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

    def forward(self, x, additional_feature=None):
        y = self.nonlinear(self.linear1(x))
        if self.dropout is not None:
            y = self.dropout(y)
        y = self.linear2(y)
        parts = [y]
        if self.global_feature:
            parts.append(y.max(dim=1, keepdim=True)[0].expand(-1, x.shape[1], -1))
        if self.additional_feature_dim > 0:
            parts.append(additional_feature)
        if self.concatenate_input:
            parts.append(x)
        return torch.cat(parts, dim=-1) if len(parts) > 1 else y


class MultiHeadSelfAttention(nn.Module):
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

    def forward(self, x, mask=None):
        n = x.shape[0]
        heads = lambda t, d: t.reshape(n, -1, self.n_heads, d).transpose(1, 2)
        q = heads(self.w_q(x), self.qk_dim_per_head)
        k = heads(self.w_k(x), self.qk_dim_per_head)
        v = heads(self.w_v(x), self.v_dim_per_head)
        y = attention(q, k, v, mask, self.dropout).transpose(1, 2).reshape(n, -1, self.v_dim)
        return self.out(y) if self.n_heads > 1 else y


class FeedForward(nn.Module):
    """reference Attention.py:207-236"""

    def __init__(self, input_dim, inner_dim, gelu=True, dropout=None):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, inner_dim)
        self.linear2 = nn.Linear(inner_dim, input_dim)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.nonlinear = _activation(gelu)

    def forward(self, x):
        y = self.nonlinear(self.linear1(x))
        if self.dropout is not None:
            y = self.dropout(y)
        return self.linear2(y)


class Encoder(nn.Module):
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

    def forward(self, x, mask=None):
        y = self.mhsa(self.norm1(x), mask=mask)
        if self.dropout1 is not None:
            y = self.dropout1(y)
        x = x + y
        if self.FF:
            y = self.ff(self.norm2(x))
            if self.dropout2 is not None:
                y = self.dropout2(y)
            x = x + y
        return x
