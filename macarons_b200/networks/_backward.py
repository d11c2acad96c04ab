"""Training support for SconeVis / SconeOcc (reference trainers/pretrain_scone_vis.py:162-225,
pretrain_scone_occ.py:158, train_macarons.py:423-444, 1159-1162).

The FORWARD value of a differentiable call is produced by the fused CUDA kernels, exactly as in inference.  The BACKWARD
pass recomputes the network from its saved inputs with differentiable torch operations ON THE GPU (torch.autograd through
cuBLAS / ATen kernels) and back-propagates through that recomputation: activations are never stored between forward and
backward ("recompute in backward"), and the gradients are those of the same function up to the fp32 rounding difference
between the two evaluations (~1e-5 relative).  The hand-written backward kernel of this package is
mac_covgain_backward_f32 (the SH integration); the transformer backward is NOT hand-written CUDA -- this module says so
explicitly rather than hiding it.  Dropout (train mode with dropout != None) is not supported and is refused by the callers.

The functions below walk the parameter containers of networks/Attention.py, SconeVis.py and SconeOcc.py
(reference networks/Attention.py:8-300, SconeVis.py:121-162, SconeOcc.py:36-42, 105-130, 250-347)."""
import torch
import torch.nn.functional as F

from .. import netpack, ops


def _linear(layer, x):
    return F.linear(x, layer.weight, layer.bias)


def _norm(layer, x):
    return F.layer_norm(x, (x.shape[-1],), layer.weight, layer.bias, layer.eps)


def _encoder(enc, x):
    """Pre-LayerNorm block: x + MHSA(norm1 x), then + FF(norm2 .) (Attention.py:281-300, 182-204, 231-236)."""
    m, heads = enc.mhsa, enc.mhsa.n_heads
    n = _norm(enc.norm1, x)
    B, S, _ = n.shape
    q = _linear(m.w_q, n).view(B, S, heads, -1).transpose(1, 2)
    k = _linear(m.w_k, n).view(B, S, heads, -1).transpose(1, 2)
    v = _linear(m.w_v, n).view(B, S, heads, -1).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-2, -1) / (q.shape[-1] ** 0.5), dim=-1) @ v
    x = x + _linear(m.out, a.transpose(1, 2).reshape(B, S, -1))
    return x + _linear(enc.ff.linear2, F.gelu(_linear(enc.ff.linear1, _norm(enc.norm2, x))))


def _embedding(emb, x):
    """Embedding.forward (Attention.py:100-128): two-layer MLP [| global max feature] | raw input."""
    e = _linear(emb.linear2, F.gelu(_linear(emb.linear1, x)))
    if emb.global_feature:
        e = torch.cat((e, e.max(dim=1, keepdim=True)[0].expand_as(e)), dim=-1)
    return torch.cat((e, x), dim=-1)


def sconevis_math(vis, pts, view_harmonics):
    x = _embedding(vis.embedding, pts)
    for enc in vis.encoders:
        x = _encoder(enc, x)
    h = F.gelu(_linear(vis.fc1, _norm(vis.norm, x)))
    h = F.gelu(_linear(vis.fc2, torch.cat((h, view_harmonics), dim=-1)))
    return _linear(vis.fc3, h)


def _pct(pct, tokens):
    """PCTransformer.forward (SconeOcc.py:105-130): (n, S, 3) -> (n, feature_dim) = [max | mean] over the tokens."""
    x = _embedding(pct.embedding, tokens)
    for enc in pct.encoders:
        x = _encoder(enc, x)
    f = _linear(pct.linear0, _norm(pct.norm, x))
    return torch.cat((f.max(dim=1)[0], f.mean(dim=1)), dim=-1)


def sconeocc_math(occ, pc_global, clouds, knn_idx, x, view_harmonics, chunk=8192):
    """SconeOcc.forward (SconeOcc.py:250-347) given the sub-sampled clouds and the (non-differentiable) neighbour indices."""
    B, Q, _ = x.shape
    g = _pct(occ.global_transformer, pc_global)                                   # (B, 512)
    outs = []
    for q0 in range(0, Q, chunk):                                                 # bounds the (B * q * 16, 128) activations
        xq = x[:, q0:q0 + chunk]
        nq = xq.shape[1]
        feats = [g.view(B, 1, -1).expand(-1, nq, -1)]
        for s, pct in enumerate(occ.local_transformers):
            idx = knn_idx[s][:, q0:q0 + chunk].long()                             # (B, nq, 16)
            nb = torch.gather(clouds[s][:, None].expand(-1, nq, -1, -1), 2, idx[..., None].expand(-1, -1, -1, 3))
            feats.append(_pct(pct, (nb - xq[:, :, None, :]).reshape(B * nq, -1, 3)).view(B, nq, -1))
        xe = occ.x_embedding
        e = F.gelu(_linear(xe.linear3, F.gelu(_linear(xe.linear2, F.gelu(_linear(xe.linear1, xq))))))
        h = torch.cat(feats + [e, view_harmonics[:, q0:q0 + chunk]], dim=-1)
        h = F.gelu(_linear(occ.linear1, h))
        h = F.gelu(_linear(occ.linear2, h))
        outs.append(F.gelu(_linear(occ.linear3, h)))
    return torch.cat(outs, dim=1)


def _backprop(y_fn, tensors, needs, params, grad_out):
    """Recompute y = y_fn(*inputs) under autograd and return (grads of the inputs that need one, grads of params)."""
    with torch.enable_grad():
        ins = [t.detach().requires_grad_(bool(n)) for t, n in zip(tensors, needs)]
        y = y_fn(*ins)
        wanted = [t for t, n in zip(ins, needs) if n] + [p for p in params if p.requires_grad]
        grads = list(torch.autograd.grad(y, wanted, grad_out, allow_unused=True)) if wanted else []
    in_grads = [grads.pop(0) if n else None for n in needs]
    p_grads = [grads.pop(0) if p.requires_grad else None for p in params]
    return in_grads, p_grads


class SconeVisFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pts, view_harmonics, *params):
        ctx.module = module
        ctx.save_for_backward(pts, view_harmonics)
        return ops.sconevis_forward(netpack.pack_sconevis(module), pts.detach(), view_harmonics.detach())

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        in_grads, p_grads = _backprop(lambda p, v: sconevis_math(module, p, v), ctx.saved_tensors, ctx.needs_input_grad[1:3],
                                      list(module.parameters()), grad_out)
        return (None, *in_grads, *p_grads)


class SconeOccFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pc_global, cloud0, cloud1, cloud2, x, view_harmonics, *params):
        ctx.module = module
        clouds = [cloud0, cloud1, cloud2]
        out = ops.sconeocc_forward(netpack.pack_sconeocc(module), pc_global.detach(), [c.detach() for c in clouds], x.detach(),
                                   view_harmonics.detach(), chunk=module.queries_per_pass)
        # the neighbour indices are data for the backward pass (the kNN kernel again: exactly the neighbours the forward used)
        idx = [ops.knn16(x.detach(), c.detach(), return_dists=False) for c in clouds]
        ctx.save_for_backward(pc_global, cloud0, cloud1, cloud2, x, view_harmonics, *idx)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        saved = ctx.saved_tensors
        tensors, idx = saved[:6], saved[6:]

        def y_fn(pc_global, c0, c1, c2, x, vh):
            return sconeocc_math(module, pc_global, [c0, c1, c2], idx, x, vh)
        in_grads, p_grads = _backprop(y_fn, tensors, ctx.needs_input_grad[1:7], list(module.parameters()), grad_out)
        return (None, *in_grads, *p_grads)
