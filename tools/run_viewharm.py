#!/usr/bin/env python3
"""A few launches of the fused view-state/harmonics kernel at the cfg5 shape (for ncu)."""
import os, sys
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from macarons_b200.utility import scone_utils
dev = torch.device("cuda:0")
base, hp, ha = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
big = (torch.rand(1, 200704, 3) - 0.5).to(dev)
views = synth.sphere_cameras(10, 1.5, torch.Generator().manual_seed(1)).to(dev)
for _ in range(4):
    out = scone_utils.compute_view_state_harmonics(big, views, base, hp, ha, 7, 14)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    out = scone_utils.compute_view_state_harmonics(big, views, base, hp, ha, 7, 14)
b.record()
torch.cuda.synchronize()
print("fused view state + harmonics, 200704 points x 10 views: %.1f us per call (CUDA events, 50 calls)" % (a.elapsed_time(b) * 20))
print(out.abs().sum().item())
