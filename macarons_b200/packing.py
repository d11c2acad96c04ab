"""Weight packing for the tensor-core kernels.

`split_tf32(w)` returns the two TF32 halves the 3-term fp32-accurate product needs
(csrc/linear.cu): w_hi = tf32(w) with round-to-nearest (ties away, the same rounding as PTX
`cvt.rna.tf32.f32`), w_lo = tf32(w - w_hi).  `pad_cols` pads the K dimension to a multiple of 4 floats so
that rows are 16-byte aligned for TMA.
"""
import torch


def round_tf32(x):
    """fp32 -> nearest TF32 value (10 mantissa bits), ties away from zero, as an fp32 tensor."""
    bits = x.contiguous().view(torch.int32)
    mag = bits & 0x7FFFFFFF
    finite = mag < 0x7F800000
    rounded = ((mag + 0x1000) & ~0x1FFF) | (bits & -0x80000000)
    return torch.where(finite, rounded, bits).view(torch.float32)


def split_tf32(w):
    w = w.detach().to(torch.float32)
    hi = round_tf32(w)
    lo = round_tf32(w - hi)
    return hi, lo


def pad_cols(w, multiple=4):
    n, k = w.shape
    kp = (k + multiple - 1) // multiple * multiple
    if kp == k:
        return w.contiguous()
    out = torch.zeros(n, kp, dtype=w.dtype, device=w.device)
    out[:, :k] = w
    return out


class PackedLinear:
    """One nn.Linear packed for mac_linear_f32: (N, Kp) hi / lo halves and the bias, on the module's device."""

    def __init__(self, weight, bias=None, split=True):
        w = pad_cols(weight.detach().to(torch.float32))
        self.N, self.K = weight.shape
        self.ldw = w.shape[1]
        if split:
            self.hi, self.lo = split_tf32(w)
        else:
            self.hi, self.lo = w.contiguous(), None
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()
