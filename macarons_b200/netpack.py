"""Pack the parameters of SconeOcc / SconeVis modules into the C structs of include/macarons_b200.h
(mac_linear_w_t, mac_encoder_w_t, mac_pct_w_t, mac_sconeocc_w_t, mac_sconevis_w_t).

Packing = TF32 hi/lo split of every Linear weight (packing.split_tf32), K padded to a multiple of 4 floats,
w_q / w_k / w_v stacked into one (qk+qk+v, d) matrix, and Embedding.linear2 of the PCTransformers extended with
identity rows so that the GEMM also concatenates the raw input (Attention.py:125).  A pack is cached per
module (outside the module, see _PACKS) and rebuilt when any parameter's storage or version counter changes
(load_state_dict, .to(), optimizer steps); `invalidate_pack` covers writes through `.data`.
"""
import ctypes

import torch

from . import packing

MAX_ENCODERS, MAX_SCALES = 4, 4
_fp = ctypes.c_void_p


class LinearW(ctypes.Structure):
    _fields_ = [("hi", _fp), ("lo", _fp), ("bias", _fp), ("N", ctypes.c_int), ("K", ctypes.c_int), ("ldw", ctypes.c_int)]


class EncoderW(ctypes.Structure):
    _fields_ = [("ln1_g", _fp), ("ln1_b", _fp), ("ln2_g", _fp), ("ln2_b", _fp),
                ("qkv", LinearW), ("out", LinearW), ("ff1", LinearW), ("ff2", LinearW)]


class PctW(ctypes.Structure):
    _fields_ = [("emb1_w", _fp), ("emb1_b", _fp), ("in_dim", ctypes.c_int), ("inner", ctypes.c_int),
                ("emb2", LinearW), ("n_enc", ctypes.c_int), ("d_model", ctypes.c_int), ("dqk", ctypes.c_int),
                ("dv", ctypes.c_int), ("enc", EncoderW * MAX_ENCODERS), ("ln_g", _fp), ("ln_b", _fp),
                ("linear0", LinearW)]


class SconeOccW(ctypes.Structure):
    _fields_ = [("global_pct", PctW), ("n_scale", ctypes.c_int), ("local_pct", PctW * MAX_SCALES),
                ("xemb1_w", _fp), ("xemb1_b", _fp), ("xemb1_n", ctypes.c_int), ("xemb2", LinearW), ("xemb3", LinearW),
                ("lin1_wg", _fp), ("lin1_wg_ld", ctypes.c_int), ("global_dim", ctypes.c_int), ("lin1_b", _fp),
                ("lin1", LinearW), ("lin2", LinearW), ("lin3", LinearW)]


class SconeVisW(ctypes.Structure):
    _fields_ = [("emb1_w", _fp), ("emb1_b", _fp), ("in_dim", ctypes.c_int), ("inner", ctypes.c_int),
                ("emb2", LinearW), ("n_enc", ctypes.c_int), ("d_model", ctypes.c_int), ("dqk", ctypes.c_int),
                ("dv", ctypes.c_int), ("enc", EncoderW * MAX_ENCODERS), ("ln_g", _fp), ("ln_b", _fp),
                ("fc1", LinearW), ("fc2", LinearW), ("fc3", LinearW)]


class _Packer:
    def __init__(self):
        self.keep = []   # tensors the structs point into

    def plain(self, t):
        t = t.detach().to(torch.float32).contiguous()
        self.keep.append(t)
        return t.data_ptr()

    def linear(self, weight, bias, dst):
        pk = packing.PackedLinear(weight, bias)
        self.keep.append(pk)
        dst.hi, dst.lo = pk.hi.data_ptr(), pk.lo.data_ptr()
        dst.bias = 0 if pk.bias is None else pk.bias.data_ptr()
        dst.N, dst.K, dst.ldw = pk.N, pk.K, pk.ldw

    def encoder(self, enc, dst):
        if not enc.FF or enc.n_heads != 4:
            raise NotImplementedError("the fused encoder supports FF=True and 4 heads")
        dst.ln1_g, dst.ln1_b = self.plain(enc.norm1.weight), self.plain(enc.norm1.bias)
        dst.ln2_g, dst.ln2_b = self.plain(enc.norm2.weight), self.plain(enc.norm2.bias)
        m = enc.mhsa
        self.linear(torch.cat((m.w_q.weight, m.w_k.weight, m.w_v.weight), dim=0),
                    torch.cat((m.w_q.bias, m.w_k.bias, m.w_v.bias), dim=0), dst.qkv)
        self.linear(m.out.weight, m.out.bias, dst.out)
        self.linear(enc.ff.linear1.weight, enc.ff.linear1.bias, dst.ff1)
        self.linear(enc.ff.linear2.weight, enc.ff.linear2.bias, dst.ff2)

    def encoders(self, encoders, dst):
        if len(encoders) > MAX_ENCODERS:
            raise NotImplementedError("at most %d encoders" % MAX_ENCODERS)
        dst.n_enc = len(encoders)
        e0 = encoders[0]
        dst.d_model, dst.dqk, dst.dv = e0.embedding_dim, e0.mhsa.qk_dim_per_head, e0.mhsa.v_dim_per_head
        for i, enc in enumerate(encoders):
            self.encoder(enc, dst.enc[i])

    def pct(self, pct, dst):
        emb = pct.embedding
        if not (emb.concatenate_input and not emb.global_feature and emb.additional_feature_dim == 0 and emb.gelu):
            raise NotImplementedError("the fused PCTransformer supports the default embedding (gelu, concatenated input)")
        dst.emb1_w, dst.emb1_b = self.plain(emb.linear1.weight), self.plain(emb.linear1.bias)
        dst.in_dim, dst.inner = emb.input_dim, emb.inner_dim
        # [e | x] = [[W2, 0], [0, I]] [h | x] + [b2 | 0]
        feat, inner, d_in = emb.feature_dim, emb.inner_dim, emb.input_dim
        w2 = emb.linear2.weight.detach()
        aug = torch.zeros(feat + d_in, inner + d_in, dtype=torch.float32, device=w2.device)
        aug[:feat, :inner] = w2
        aug[feat:, inner:] = torch.eye(d_in, device=w2.device)
        b2 = torch.cat((emb.linear2.bias.detach(), torch.zeros(d_in, device=w2.device)))
        self.linear(aug, b2, dst.emb2)
        self.encoders(pct.encoders, dst)
        dst.ln_g, dst.ln_b = self.plain(pct.norm.weight), self.plain(pct.norm.bias)
        self.linear(pct.linear0.weight, pct.linear0.bias, dst.linear0)


# Packs live OUTSIDE the modules (a WeakKeyDictionary keyed by the module): they hold ctypes structures with raw
# pointers, which must never be seen by copy.deepcopy / torch.save(model) (the reference pickles whole models,
# ManyDepth.py:546, :690).
import os
import weakref

_PACKS = weakref.WeakKeyDictionary()
# The fingerprint (storage address, version counter, device of every parameter / buffer) catches load_state_dict,
# .to(), optimizer steps and every other in-place op on the parameters.  Writes through `.data` (p.data.copy_(),
# an EMA written as p.data.mul_()) keep both the address and the version counter: call `invalidate_pack(module)`
# after such a write, or set `verify_content = True` (env MAC_NETPACK_VERIFY=1) to add a float64 checksum of every
# tensor to the fingerprint (one extra reduction and a host synchronisation per forward).
verify_content = os.environ.get("MAC_NETPACK_VERIFY", "0") == "1"


def _tensors(module):
    return list(module.parameters()) + list(module.buffers())


def _fingerprint(module):
    ts = _tensors(module)
    fp = tuple((t.data_ptr(), t._version, t.device.index) for t in ts)
    if verify_content and ts:
        with torch.no_grad():
            sums = torch.stack([t.detach().double().sum() + 3.0 * t.detach().double().square().sum() for t in ts])
        fp += (tuple(sums.tolist()),)
    return fp


def invalidate_pack(module):
    """Drop the packed weights of `module` (and of its sub-modules): the next forward re-packs.  Needed only after
    writes that bypass autograd's version counter (`param.data.<op>_()`)."""
    for m in module.modules():
        _PACKS.pop(m, None)


def _cached(module, build):
    fp = _fingerprint(module)
    cache = _PACKS.get(module)
    if cache is None or cache[0] != fp:
        cache = (fp,) + build()
        _PACKS[module] = cache
    return cache[1]


def pack_sconevis(vis):
    """-> SconeVisW (kept alive, with its tensors, on the module)."""
    def build():
        emb = vis.embedding
        if not (vis.use_view_state and vis.view_state_mode == "end" and vis.use_global_feature and not vis.alt
                and emb.concatenate_input and emb.gelu and emb.additional_feature_dim == 0):
            raise NotImplementedError("the fused SconeVis forward supports the reference's default architecture "
                                      "(global feature, view harmonics concatenated at the end, alt=False)")
        pk, w = _Packer(), SconeVisW()
        w.emb1_w, w.emb1_b = pk.plain(emb.linear1.weight), pk.plain(emb.linear1.bias)
        w.in_dim, w.inner = emb.input_dim, emb.inner_dim
        pk.linear(emb.linear2.weight, emb.linear2.bias, w.emb2)
        pk.encoders(vis.encoders, w)
        w.ln_g, w.ln_b = pk.plain(vis.norm.weight), pk.plain(vis.norm.bias)
        pk.linear(vis.fc1.weight, vis.fc1.bias, w.fc1)
        pk.linear(vis.fc2.weight, vis.fc2.bias, w.fc2)
        pk.linear(vis.fc3.weight, vis.fc3.bias, w.fc3)
        return w, pk
    return _cached(vis, build)


def pack_sconeocc(occ):
    """-> SconeOccW (kept alive, with its tensors, on the module)."""
    def build():
        if occ.n_scale != 3 or not occ.offset or not occ.gelu or occ.k_for_knn != 16:
            raise NotImplementedError("the fused SconeOcc forward supports n_scale=3, k_for_knn=16, offset=True, gelu")
        pk, w = _Packer(), SconeOccW()
        pk.pct(occ.global_transformer, w.global_pct)
        w.n_scale = occ.n_scale
        for s, lt in enumerate(occ.local_transformers):
            pk.pct(lt, w.local_pct[s])
        xe = occ.x_embedding
        w.xemb1_w, w.xemb1_b, w.xemb1_n = pk.plain(xe.linear1.weight), pk.plain(xe.linear1.bias), xe.linear1.out_features
        pk.linear(xe.linear2.weight, xe.linear2.bias, w.xemb2)
        pk.linear(xe.linear3.weight, xe.linear3.bias, w.xemb3)
        g = occ.global_feature_dim
        w1 = occ.linear1.weight.detach()
        wg = w1[:, :g].to(torch.float32).contiguous()
        pk.keep.append(wg)
        w.lin1_wg, w.lin1_wg_ld, w.global_dim = wg.data_ptr(), wg.stride(0), g
        w.lin1_b = pk.plain(occ.linear1.bias)
        pk.linear(w1[:, g:], None, w.lin1)
        pk.linear(occ.linear2.weight, occ.linear2.bias, w.lin2)
        pk.linear(occ.linear3.weight, occ.linear3.bias, w.lin3)
        return w, pk
    return _cached(occ, build)


# ---- ManyDepth (include/macarons_b200.h: mac_conv_w_t ... mac_manydepth_w_t) ------------------------
ACT_NONE, ACT_RELU, ACT_GELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3, 4


class ConvW(ctypes.Structure):
    _fields_ = [("lin", LinearW), ("k", ctypes.c_int), ("stride", ctypes.c_int), ("pad", ctypes.c_int),
                ("reflect", ctypes.c_int), ("act", ctypes.c_int)]


class BlockW(ctypes.Structure):
    _fields_ = [("conv1", ConvW), ("conv2", ConvW), ("down", ConvW), ("has_down", ctypes.c_int)]


class ExpansionW(ctypes.Structure):
    _fields_ = [("upconv", ConvW), ("iconv", ConvW)]


class ManyDepthW(ctypes.Structure):
    _fields_ = [("conv1", ConvW), ("layer1", BlockW * 2), ("layer2", BlockW * 2), ("layer3", BlockW * 2),
                ("layer4", BlockW * 2), ("conv_reduce", ConvW), ("expansion", ExpansionW * 5), ("disp", ConvW * 4),
                ("n_depth", ctypes.c_int), ("d_min", ctypes.c_float), ("d_max", ctypes.c_float)]


def _fold_bn(weight, bias, bn):
    """Conv weight (Cout, Cin, kh, kw) followed by an eval-mode BatchNorm2d -> equivalent weight and bias (float64 math)."""
    w = weight.detach().double()
    b = torch.zeros(w.shape[0], dtype=torch.float64, device=w.device) if bias is None else bias.detach().double()
    if bn is not None:
        scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().double()) * scale + bn.bias.detach().double()
    return w, b


def _pack_conv(pk, dst, conv, bn, act, transposed=False):
    weight = conv.weight
    k = weight.shape[-1]
    if transposed:   # ConvTranspose2d(k, stride 1, padding p) == Conv2d with the flipped kernel, padding k - 1 - p
        if conv.stride != (1, 1) or conv.output_padding != (0, 0):
            raise NotImplementedError("only stride-1 transposed convolutions are on the depth path")
        weight = weight.detach().flip(-1, -2).transpose(0, 1)
        pad, stride, reflect = k - 1 - conv.padding[0], 1, 0
    else:
        pad, stride = conv.padding[0], conv.stride[0]
        reflect = 1 if conv.padding_mode == "reflect" else 0
        if conv.padding_mode not in ("zeros", "reflect"):
            raise NotImplementedError("padding mode %s" % conv.padding_mode)
    w, b = _fold_bn(weight, conv.bias, bn)
    mat = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.float32)   # (Cout, ky, kx, c)
    pk.linear(mat, b.to(torch.float32), dst.lin)
    dst.k, dst.stride, dst.pad, dst.reflect, dst.act = k, stride, pad, reflect, act


def _pack_block(pk, dst, block):
    _pack_conv(pk, dst.conv1, block.conv1, block.bn1, ACT_RELU)
    _pack_conv(pk, dst.conv2, block.conv2, block.bn2, ACT_RELU)     # ReLU after the residual add (res_first)
    dst.has_down = 0
    if getattr(block, "downsample", None) is not None:
        _pack_conv(pk, dst.down, block.downsample[0], block.downsample[1], ACT_NONE)
        dst.has_down = 1


def pack_manydepth(model):
    """-> ManyDepthW (kept alive on the module).  BatchNorm is folded with its running statistics (inference)."""
    def build():
        dd = model.depth_decoder
        pk, w = _Packer(), ManyDepthW()
        fe = dd.feature_extractor
        _pack_conv(pk, w.conv1, fe.conv1, fe.bn1, ACT_RELU)
        for dst, layer in ((w.layer1, fe.layer), (w.layer2, dd.resnet_layer_2), (w.layer3, dd.resnet_layer_3),
                           (w.layer4, dd.resnet_layer_4)):
            if len(layer) != 2:
                raise NotImplementedError("the depth path is built for ResNet-18 (2 blocks per layer)")
            for i in range(2):
                _pack_block(pk, dst[i], layer[i])
        cvb = dd.cost_volume_builder
        _pack_conv(pk, w.conv_reduce, cvb.conv_reduce, None, ACT_RELU)
        for i, exp in enumerate((dd.expansion5, dd.expansion4, dd.expansion3, dd.expansion2, dd.expansion1)):
            _pack_conv(pk, w.expansion[i].upconv, exp.upconv, None, ACT_ELU, transposed=True)
            _pack_conv(pk, w.expansion[i].iconv, exp.iconv, None, ACT_ELU)
        for i, disp in enumerate((dd.disp1, dd.disp2, dd.disp3, dd.disp4)):
            _pack_conv(pk, w.disp[i], disp.conv, None, ACT_SIGMOID)
        w.n_depth, w.d_min, w.d_max = int(cvb.n_depth), float(cvb.d_min), float(cvb.d_max)
        return w, pk

    return _cached(model, build)
