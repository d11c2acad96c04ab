#!/usr/bin/env python3
"""One launch each of the kernels added in round 2 at representative shapes (for `ncu --set full -k regex:...`):
viewstate_harm (200 704 points x 10 views), covgain_bwd (32 x 2048 points x 256 cameras), the ragged SconeOcc forward
over 40 cells (knn16_cells, gather_rows and the linear / attention kernels it drives)."""
import contextlib, io, os, sys
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from macarons_b200 import ops
from macarons_b200.networks.SconeOcc import SconeOcc
from macarons_b200.utility import scone_utils

dev = torch.device("cuda:0")
base, hp, ha = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
big = (torch.rand(1, 200704, 3) - 0.5).to(dev)
views = synth.sphere_cameras(10, 1.5, torch.Generator().manual_seed(1)).to(dev)
for _ in range(3):
    scone_utils.compute_view_state_harmonics(big, views, base, hp, ha, 7, 14)

pts, harm, cams = synth.covgain_inputs(32, 2048, 256, seed=3)
g = torch.randn(32, 256)
for _ in range(2):
    ops.coverage_gain_backward(pts.to(dev), harm.to(dev), cams.to(dev), g.to(dev))

with contextlib.redirect_stdout(io.StringIO()):
    occ = SconeOcc()
occ.load_state_dict(synth.seeded_state_dict(occ.state_dict(), 5))
occ = occ.to(dev).eval()
gen = torch.Generator().manual_seed(5)
clouds = [(torch.rand(int(n), 3, generator=gen) - 0.5).to(dev) for n in torch.randint(300, 6000, (40,), generator=gen)]
queries = [(torch.rand(int(q), 3, generator=gen) - 0.5).to(dev) for q in torch.randint(200, 3000, (40,), generator=gen)]
vhs = [0.3 * torch.randn(q.shape[0], 64, device=dev) for q in queries]
with torch.no_grad():
    for _ in range(2):
        out = occ.forward_cells(clouds, queries, vhs)
torch.cuda.synchronize()
print(sum(o.abs().sum().item() for o in out))
