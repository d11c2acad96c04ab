"""GPU parity of the tcgen05 linear layer (csrc/linear.cu) against a float64 matmul of the same
operands: plain, bias + GELU + residual, fused LayerNorm output, 16-row max|mean pooling, ragged shapes."""
import pytest
import torch
import torch.nn.functional as F

from macarons_b200 import ops, packing

pytestmark = pytest.mark.gpu

# 3-term TF32 split: every product is exact to 2^-22, accumulation in fp32 (tolerance of an fp32 matmul)
RTOL_SPLIT = 2e-6
# single-pass TF32 keeps 10 mantissa bits
RTOL_TF32 = 2e-3


def _case(M, N, K, seed, ldx=None):
    g = torch.Generator().manual_seed(seed)
    ldx = ldx or (K + 3) // 4 * 4
    xbuf = torch.randn(M, ldx, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    return xbuf, w, b


def _scale(ref):
    return ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (2048, 256, 256), (300, 192, 128), (2048, 384, 256),
                                   (129, 125, 125), (64, 64, 32), (1000, 512, 256), (4096, 64, 1856), (17, 1, 256)])
def test_linear_matches_float64(M, N, K, cuda_device):
    xbuf, w, b = _case(M, N, K, seed=M + N + K)
    x = xbuf[:, :K]
    ref = x.double() @ w.double().t() + b.double()
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device))
    n0 = ops.launch_count()
    got = ops.linear(xbuf.to(cuda_device)[:, :K], pk).cpu()
    assert ops.launch_count() == n0 + 1
    assert got.shape == (M, N)
    assert (got.double() - ref).abs().max().item() <= RTOL_SPLIT * _scale(ref) * (K ** 0.5)


def test_single_pass_tf32_is_within_tf32_tolerance(cuda_device):
    xbuf, w, b = _case(512, 128, 256, seed=3)
    ref = xbuf.double() @ w.double().t() + b.double()
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device), split=False)
    got = ops.linear(xbuf.to(cuda_device), pk).cpu()
    err = (got.double() - ref).abs().max().item()
    assert err <= RTOL_TF32 * _scale(ref) * 4
    assert err > 1e-6 * _scale(ref)   # it really is the reduced-precision path


@pytest.mark.parametrize("act", [ops.ACT_LIN_RELU, ops.ACT_GELU])
def test_bias_activation_residual(act, cuda_device):
    M, N, K = 777, 256, 512
    xbuf, w, b = _case(M, N, K, seed=11)
    res = torch.randn(M, N, generator=torch.Generator().manual_seed(5))
    y = xbuf.double() @ w.double().t() + b.double()
    y = F.relu(y) if act == ops.ACT_LIN_RELU else F.gelu(y)
    ref = res.double() + y
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device))
    got = ops.linear(xbuf.to(cuda_device), pk, act=act, residual=res.to(cuda_device)).cpu()
    assert (got.double() - ref).abs().max().item() <= 4e-6 * _scale(ref) * (K ** 0.5)


@pytest.mark.parametrize("N", [128, 256, 125])
def test_fused_layernorm_output(N, cuda_device):
    M, K = 515, 128
    xbuf, w, b = _case(M, N, K, seed=N)
    g = torch.Generator().manual_seed(9)
    res = torch.randn(M, (N + 3) // 4 * 4, generator=g)
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    use_res = N % 4 == 0      # the residual path needs 16-byte aligned rows
    y = xbuf[:, :K].double() @ w.double().t() + b.double()
    if use_res:
        y = y + res[:, :N].double()
    ref_ln = F.layer_norm(y, (N,), gamma.double(), beta.double(), 1e-5)
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device))
    out = torch.empty(M, (N + 3) // 4 * 4, device=cuda_device)
    got, got_ln = ops.linear(xbuf.to(cuda_device)[:, :K], pk, residual=res.to(cuda_device)[:, :N] if use_res else None,
                             out=out[:, :N],
                             ln=(gamma.to(cuda_device), beta.to(cuda_device), 1e-5))
    assert (got.cpu().double() - y).abs().max().item() <= 4e-6 * _scale(y) * (K ** 0.5)
    assert (got_ln.cpu().double() - ref_ln).abs().max().item() <= 2e-5


def test_pool16_max_and_mean(cuda_device):
    M, N, K = 16 * 301, 128, 128
    xbuf, w, b = _case(M, N, K, seed=21)
    y = (xbuf.double() @ w.double().t() + b.double()).view(M // 16, 16, N)
    ref = torch.cat((y.max(dim=1)[0], y.mean(dim=1)), dim=-1)
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device))
    got = ops.linear(xbuf.to(cuda_device), pk, pool=16).cpu()
    assert got.shape == (M // 16, 2 * N)
    assert (got.double() - ref).abs().max().item() <= 4e-6 * _scale(ref) * (K ** 0.5)


def test_output_into_column_slice(cuda_device):
    M, N, K = 256, 128, 64
    xbuf, w, b = _case(M, N, K, seed=2)
    ref = xbuf.double() @ w.double().t() + b.double()
    pk = packing.PackedLinear(w.to(cuda_device), b.to(cuda_device))
    wide = torch.full((M, 512), -7.0, device=cuda_device)
    ops.linear(xbuf.to(cuda_device), pk, out=wide[:, 256:384])
    assert (wide[:, 256:384].cpu().double() - ref).abs().max().item() <= 4e-6 * _scale(ref) * 8
    assert torch.all(wide[:, :256] == -7.0) and torch.all(wide[:, 384:] == -7.0)


@pytest.mark.parametrize("M,N,K", [(515, 128, 128), (300, 256, 256), (2048, 192, 128), (130, 64, 512)])
def test_layernorm_on_load_and_row_statistics(M, N, K, cuda_device):
    """Two chained layers that pass (mean, rstd) instead of a normalised copy == LayerNorm + Linear in float64."""
    xbuf, w1, b1 = _case(M, K, K, seed=M + N)
    _, w2, b2 = _case(M, N, K, seed=M + N + 1)
    gen = torch.Generator().manual_seed(9)
    g, beta = 1 + 0.1 * torch.randn(K, generator=gen), 0.1 * torch.randn(K, generator=gen)
    res = torch.randn(M, K, generator=gen)
    y1 = res.double() + xbuf.double() @ w1.double().t() + b1.double()
    ln = F.layer_norm(y1, (K,), g.double(), beta.double(), 1e-5)
    ref = F.gelu(ln @ w2.double().t() + b2.double())
    pk1 = packing.PackedLinear(w1.to(cuda_device), b1.to(cuda_device))
    pk2 = packing.PackedLinear(w2.to(cuda_device), b2.to(cuda_device))
    if K <= 256:
        got1, stats = ops.linear_lnio(xbuf.to(cuda_device), pk1, residual=res.to(cuda_device), stats_out=True)
        assert (got1.cpu().double() - y1).abs().max().item() <= 4e-6 * _scale(y1) * (K ** 0.5)
        mean, rstd = y1.mean(-1), 1.0 / torch.sqrt(y1.var(-1, unbiased=False) + 1e-5)
        assert (stats[:, 0].cpu().double() - mean).abs().max().item() <= 1e-5 * _scale(y1)
        assert ((stats[:, 1].cpu().double() - rstd) / rstd).abs().max().item() <= 1e-5
    else:   # row statistics need N <= 256: take them from the float64 reference
        got1 = y1.float().to(cuda_device)
        stats = torch.stack((y1.mean(-1), 1.0 / torch.sqrt(y1.var(-1, unbiased=False) + 1e-5)), dim=-1).float().to(cuda_device)
    got = ops.linear_lnio(got1, pk2, act=ops.ACT_GELU, ln_in=(stats, g.to(cuda_device), beta.to(cuda_device)))
    assert (got.cpu().double() - ref).abs().max().item() <= 2e-5 * max(1.0, _scale(ref)) * (K ** 0.5) / 8
