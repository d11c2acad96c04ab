#!/usr/bin/env python3
"""Where the tensor pipe of the tcgen05 linear layer waits (debug build: MAC_EXTRA_NVCC_FLAGS=-DMAC_LINEAR_PROFILE python -m
macarons_b200.build --force).  Runs the four layer shapes of a SconeOcc neighbourhood transformer block on 1 M tokens and
prints, per layer, the time and the share of the UMMA thread's cycles spent waiting for (accumulator free | X ready | W ready)."""
import ctypes, os, sys, json
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import torch
from macarons_b200 import ops, _lib
from macarons_b200.packing import PackedLinear

dev = torch.device("cuda:0")
M = 1 << 20
lib = _lib.load()
prof = getattr(lib, "mac_linear_profile_read", None)
if prof is not None:
    prof.restype = ctypes.c_int
    prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
g = torch.Generator().manual_seed(1)


def layer(K, N):
    return PackedLinear((torch.randn(N, K, generator=g) / K ** 0.5).to(dev), (0.1 * torch.randn(N, generator=g)).to(dev))


def run(name, fn, nbytes):
    for _ in range(2):
        fn()
    buf = (ctypes.c_ulonglong * 12)()
    if prof is not None:
        prof(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    rec = {"layer": name, "us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}
    if prof is not None:
        prof(buf)
        tot = max(1, buf[3])
        rec.update({"wait_acc": round(buf[0] / tot, 3), "wait_x": round(buf[1] / tot, 3), "wait_w": round(buf[2] / tot, 3),
                    "cycles_per_tile": round(buf[3] / max(1, buf[4])),
                    "epi_cycles_per_tile": {k: round(buf[i] / max(1, buf[4])) for k, i in
                                            (("prologue", 5), ("wait_acc", 6), ("chunks", 7), ("stats", 8))}})
    print(json.dumps(rec))


x128 = torch.randn(M, 128, generator=g).to(dev)
x256 = torch.randn(M, 256, generator=g).to(dev)
res = torch.randn(M, 128, generator=g).to(dev)
stats = torch.stack((x128.mean(1), 1.0 / x128.std(1)), 1).contiguous()
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
l_qkv, l_out, l_ff1, l_ff2 = layer(128, 192), layer(128, 128), layer(128, 256), layer(256, 128)
run("qkv   K128 N192 ln-on-load", lambda: ops.linear_lnio(x128, l_qkv, ln_in=(stats, gam, bet)), M * (128 + 192) * 4)
run("out   K128 N128 +res +stats", lambda: ops.linear_lnio(x128, l_out, residual=res, stats_out=True), M * (128 * 3) * 4)
run("ff1   K128 N256 gelu ln-on-load", lambda: ops.linear_lnio(x128, l_ff1, act=ops.ACT_GELU, ln_in=(stats, gam, bet)), M * (128 + 256) * 4)
run("ff2   K256 N128 +res +stats", lambda: ops.linear_lnio(x256, l_ff2, residual=res, stats_out=True), M * (256 + 256) * 4)
l_d256 = layer(256, 256)
res256 = torch.randn(M // 4, 256, generator=g).to(dev)
run("d256  K256 N256 +res +stats (M/4)", lambda: ops.linear_lnio(x256[:M // 4], l_d256, residual=res256, stats_out=True), (M // 4) * (256 * 3) * 4)
run("plain K128 N128", lambda: ops.linear(x128, l_out), M * 256 * 4)
