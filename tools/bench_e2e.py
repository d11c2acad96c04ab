#!/usr/bin/env python3
"""Where the end-to-end time of the host-buffer scoring call goes (cfg5): raw H2D copy, kernel, mac_covgain_host."""
import os, sys, time
import torch
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from macarons_b200 import ops

dev = torch.device("cuda:0")
pts, harm, _ = synth.covgain_inputs(1, 200704, 1, seed=1)
cams = synth.fibonacci_cameras(512)[None].contiguous()
pp, hp = pts.pin_memory(), harm.pin_memory()
dp, dh = torch.empty_like(pts, device=dev), torch.empty_like(harm, device=dev)

def wall(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n

print("H2D pinned 54.6 MB        %.3f ms" % wall(lambda: (dp.copy_(pp, non_blocking=True), dh.copy_(hp, non_blocking=True))))
print("H2D pageable              %.3f ms" % wall(lambda: (dp.copy_(pts), dh.copy_(harm)), 5))
dc = cams.to(dev)
print("kernel (device resident)  %.3f ms" % wall(lambda: ops.coverage_gain(dp, dh, dc)))
pn, hn, cn = pp.numpy(), hp.numpy(), cams.numpy()
print("mac_covgain_host pinned   %.3f ms" % wall(lambda: ops.coverage_gain_host(pn, hn, cn, device=0)))
print("mac_covgain_host pageable %.3f ms" % wall(lambda: ops.coverage_gain_host(pts.numpy(), harm.numpy(), cn, device=0), 5))
