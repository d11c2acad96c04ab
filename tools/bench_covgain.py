#!/usr/bin/env python3
"""Kernel-only timing of mac_covgain_f32 for a few shapes (CUDA events, inputs rotated to defeat L2)."""
import os, sys, json, zlib
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from macarons_b200 import ops

dev = torch.device("cuda:0")
shapes = [(1, 200704, 512, (0, 512)), (1, 200704, 512, (0, 64)), (1, 200704, 512, (0, 8)), (1, 2048, 64, (0, 64)),
          (32, 2048, 256, (0, 256)), (32, 2048, 256, (0, 32))]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for B, P, C, rng in shapes:
    nsets = 5 if P > 10000 else 40
    sets = []
    for i in range(nsets):
        pts, harm, _ = synth.covgain_inputs(B, P, 1, seed=i)
        sets.append((pts.to(dev), harm.to(dev)))
    cams = synth.fibonacci_cameras(C)[None].expand(B, -1, -1).contiguous().to(dev)
    out = torch.zeros(B, C, device=dev)
    for i in range(5):
        ops.coverage_gain(*sets[i % nsets], cams, cam_range=rng, out=out)
    torch.cuda.synchronize()
    n = 30
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        ev[i][0].record()
        ops.coverage_gain(*sets[i % nsets], cams, cam_range=rng, out=out)
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    med = ts[len(ts) // 2]
    pairs = B * P * (rng[1] - rng[0])
    byts = B * P * 272 + B * (rng[1] - rng[0]) * 16
    print(json.dumps({"B": B, "P": P, "C_local": rng[1] - rng[0], "median_us": round(med * 1e3, 2), "min_us": round(ts[0] * 1e3, 2),
                      "Gpairs_per_s": round(pairs / med / 1e6, 2), "GBps": round(byts / med / 1e6, 1),
                      "evals_per_s": round(B * (rng[1] - rng[0]) / med * 1e3),
                      "variant": os.environ.get("MAC_COVGAIN_VARIANT", "0"),
                      "out_crc": zlib.crc32(ops.coverage_gain(*sets[0], cams, cam_range=rng, out=out).cpu().numpy().tobytes())}))
