"""Oracle for rows a6-a9 of SURVEY.md section 8: SconeVis.forward, SconeOcc.forward, get_knn_points and the
chunked occupancy inference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional restatement over a plain `state_dict`, using the
same torch CPU kernels in the same operation order as the reference modules, so that with equal weights,
inputs and torch RNG state it agrees with the reference bit for bit on the same machine
(`tests/golden/make_golden.py` asserts that before writing the fixtures).

Reference (paths relative to /root/reference/macarons):
  networks/Attention.py:8-36      attention           networks/Attention.py:100-128   Embedding.forward
  networks/Attention.py:182-204   MultiHeadSelfAttention.forward   :231-236 FeedForward.forward
  networks/Attention.py:281-300   Encoder.forward
  networks/SconeOcc.py:36-42      XEmbedding.forward  networks/SconeOcc.py:105-130    PCTransformer.forward
  networks/SconeOcc.py:250-347    SconeOcc.forward    networks/SconeVis.py:121-162    SconeVis.forward
  utility/utils.py:1497-1509      get_knn_points      utility/scone_utils.py:965-998  compute_occupancy_probability
"""
import numpy as np
import torch
import torch.nn.functional as F

N_HEADS = 4


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def attention(q, k, v):
    """networks/Attention.py:8-36 with mask=None, dropout=None."""
    scores = q.matmul(k.transpose(-2, -1))
    scores = scores / np.sqrt(q.shape[-1])
    scores = F.softmax(scores, dim=-1)
    return scores.matmul(v)


def embedding(sd, prefix, x, global_feature):
    """networks/Attention.py:100-128 (gelu, concatenate_input=True, no kNN, no additional feature)."""
    n_clouds, seq_len, _ = x.shape
    res = F.gelu(_lin(sd, prefix + ".linear1", x))
    res = _lin(sd, prefix + ".linear2", res)
    if global_feature:
        feat = res.shape[-1]
        g = F.max_pool1d(input=res.transpose(-1, -2), kernel_size=seq_len).view(n_clouds, 1, feat)
        res = torch.cat((res, g.expand(-1, seq_len, -1)), dim=-1)
    return torch.cat((res, x), dim=-1)


def mhsa(sd, prefix, x):
    """networks/Attention.py:182-204."""
    q, k, v = _lin(sd, prefix + ".w_q", x), _lin(sd, prefix + ".w_k", x), _lin(sd, prefix + ".w_v", x)
    bs = x.shape[0]
    q = q.reshape(bs, -1, N_HEADS, q.shape[-1] // N_HEADS).transpose(1, 2)
    k = k.reshape(bs, -1, N_HEADS, k.shape[-1] // N_HEADS).transpose(1, 2)
    v = v.reshape(bs, -1, N_HEADS, v.shape[-1] // N_HEADS).transpose(1, 2)
    s = attention(q, k, v)
    s = s.transpose(1, 2).contiguous().view(bs, -1, x.shape[-1])
    return _lin(sd, prefix + ".out", s)


def encoder(sd, prefix, x):
    """networks/Attention.py:281-300 (pre-LayerNorm, FF=True, gelu)."""
    res = x + mhsa(sd, prefix + ".mhsa", _ln(sd, prefix + ".norm1", x))
    h = _ln(sd, prefix + ".norm2", res)
    h = _lin(sd, prefix + ".ff.linear2", F.gelu(_lin(sd, prefix + ".ff.linear1", h)))
    return res + h


def _n_encoders(sd, prefix):
    n = 0
    prefix = prefix + "." if prefix else ""
    while (prefix + "encoders.%d.norm1.weight" % n) in sd:
        n += 1
    return n


def pc_transformer(sd, prefix, pc):
    """networks/SconeOcc.py:105-130 -> (n_clouds, feature_dim) = [max | avg] pooled linear0 features."""
    n_clouds, seq_len = pc.shape[0], pc.shape[1]
    x = embedding(sd, prefix + ".embedding", pc, global_feature=False)
    for i in range(_n_encoders(sd, prefix)):
        x = encoder(sd, prefix + ".encoders.%d" % i, x)
    f = _lin(sd, prefix + ".linear0", _ln(sd, prefix + ".norm", x))
    f = f.transpose(dim0=-1, dim1=-2)
    f = torch.cat((F.max_pool1d(input=f, kernel_size=seq_len), F.avg_pool1d(input=f, kernel_size=seq_len)), dim=-2)
    return f.view(n_clouds, -1)


def knn_points(X, pc, k):
    """utility/utils.py:1497-1509: cdist + topk(smallest) + gather -> (points (B,Q,k,3), dists, idx)."""
    dists = torch.cdist(X, pc)
    min_dists, idx = torch.topk(dists, k=k, dim=-1, largest=False)
    n, m, u = pc.shape
    _, l, _ = idx.shape
    pts = pc[:, :, None].expand(n, m, k, u).gather(1, idx[..., None].expand(n, l, k, u))
    return pts, min_dists, idx


def scone_occ_subsamples(n_points, seq_len=2048, k=16, n_scale=3):
    """The random sub-samples SconeOcc.forward draws from the global CPU generator, in its order
    (networks/SconeOcc.py:269, 282-288, 311): -> (global_idx, [scale_1_idx, scale_2_idx])."""
    global_idx = torch.randperm(n_points)[:seq_len]
    ds = int(np.power(n_points / (k * 8), 1. / (n_scale - 1)))
    if ds == 0:
        ds = 2
    scale_idx, n = [], n_points
    for _ in range(n_scale - 1):
        scale_idx.append(torch.randperm(n)[:n // ds])
        n = n // ds
    return global_idx, scale_idx


def scone_occ_forward(sd, pc, x, view_harmonics, seq_len=2048, k=16, n_scale=3):
    """networks/SconeOcc.py:250-347.  Consumes torch.randperm exactly like the reference."""
    n_clouds, full = pc.shape[0], pc.shape[1]
    n_sample = x.shape[1]
    g_pc = pc[:, torch.randperm(full)[:seq_len]]
    g_feat = pc_transformer(sd, "global_transformer", g_pc)
    ds = int(np.power(full / (k * 8), 1. / (n_scale - 1)))
    if ds == 0:
        ds = 2
    cur, local = pc, []
    for s in range(n_scale):
        nb, _, _ = knn_points(x, cur, k)
        nb = nb - x.view(n_clouds, n_sample, 1, 3)
        local.append(pc_transformer(sd, "local_transformers.%d" % s, nb.view(-1, k, 3)))
        n_cur = cur.shape[1]
        if s < n_scale - 1:
            cur = cur[:, torch.randperm(n_cur)[:n_cur // ds]]
    local = torch.cat(local, dim=-1).view(n_clouds, n_sample, -1)
    xf = F.gelu(_lin(sd, "x_embedding.linear1", x))
    xf = F.gelu(_lin(sd, "x_embedding.linear2", xf))
    xf = F.gelu(_lin(sd, "x_embedding.linear3", xf))
    g = g_feat.view(n_clouds, 1, -1).expand(-1, n_sample, -1)
    res = torch.cat((g, local, xf.view(n_clouds, n_sample, -1), view_harmonics), dim=-1)
    res = F.gelu(_lin(sd, "linear1", res))
    res = F.gelu(_lin(sd, "linear2", res))
    res = F.gelu(_lin(sd, "linear3", res))          # GELU on the output as well (SconeOcc.py:342)
    return res.view(n_clouds, n_sample, -1)


def compute_occupancy_probability(sd, pc, X, view_harmonics, max_points_per_pass=20000, **kw):
    """utility/scone_utils.py:965-998: chunk the queries, one forward (with fresh sub-samples) per chunk."""
    n_clouds, n_sample = pc.shape[0], X.shape[1]
    p = max_points_per_pass // n_clouds
    preds = []
    for lo in range(0, n_sample, p):
        preds.append(scone_occ_forward(sd, pc, X[:, lo:lo + p], view_harmonics[:, lo:lo + p], **kw))
    return torch.cat(preds, dim=1) if preds else torch.zeros(n_clouds, 0, 1)


def scone_vis_forward(sd, pts, view_harmonics):
    """networks/SconeVis.py:121-162 (default: global feature, view harmonics concatenated at the end)."""
    n_clouds, seq_len = pts.shape[0], pts.shape[1]
    x = embedding(sd, "embedding", pts, global_feature=True)
    for i in range(_n_encoders(sd, "")):
        x = encoder(sd, "encoders.%d" % i, x)
    res = F.gelu(_lin(sd, "fc1", _ln(sd, "norm", x)))
    res = torch.cat((res, view_harmonics), dim=-1)
    res = F.gelu(_lin(sd, "fc2", res))
    res = _lin(sd, "fc3", res)
    return res.view(n_clouds, seq_len, -1)
