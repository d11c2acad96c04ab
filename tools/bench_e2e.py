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

if len(sys.argv) > 1 and sys.argv[1] == "quick":
    pn, hn, cn = pp.numpy(), hp.numpy(), cams.numpy()
    ops.coverage_gain_host(pn, hn, cn, device=0)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        ops.coverage_gain_host(pn, hn, cn, device=0)
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    print("mac_covgain_host median %.3f ms min %.3f ms" % (ts[len(ts) // 2], ts[0]))
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "alt":
    # alternating two pinned input sets: the C path vs plain torch copies of the same bytes, interleaved
    pts2, harm2, _ = synth.covgain_inputs(1, 200704, 1, seed=2)
    pp2, hp2 = pts2.pin_memory(), harm2.pin_memory()
    sets_t = [(pp, hp), (pp2, hp2)]
    sets_n = [(pp.numpy(), hp.numpy()), (pp2.numpy(), hp2.numpy())]
    cn = cams.numpy()
    for name, n_sets in (("one set", 1), ("two sets alternating", 2)):
        for _ in range(3):
            ops.coverage_gain_host(*sets_n[0], cn, device=0)
        tc, tt = [], []
        for i in range(24):
            t0 = time.perf_counter()
            ops.coverage_gain_host(*sets_n[i % n_sets], cn, device=0)
            tc.append(1e3 * (time.perf_counter() - t0))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dp.copy_(sets_t[i % n_sets][0], non_blocking=True)
            dh.copy_(sets_t[i % n_sets][1], non_blocking=True)
            torch.cuda.synchronize()
            tt.append(1e3 * (time.perf_counter() - t0))
        tc.sort(); tt.sort()
        print("%-22s mac_covgain_host median %.3f ms (min %.3f) | torch pinned copy median %.3f ms (min %.3f)"
              % (name, tc[len(tc) // 2], tc[0], tt[len(tt) // 2], tt[0]))
    sys.exit(0)
print("H2D pinned 54.6 MB        %.3f ms" % wall(lambda: (dp.copy_(pp, non_blocking=True), dh.copy_(hp, non_blocking=True))))
print("H2D pageable              %.3f ms" % wall(lambda: (dp.copy_(pts), dh.copy_(harm)), 5))
dc = cams.to(dev)
print("kernel (device resident)  %.3f ms" % wall(lambda: ops.coverage_gain(dp, dh, dc)))
pn, hn, cn = pp.numpy(), hp.numpy(), cams.numpy()
print("mac_covgain_host pinned   %.3f ms" % wall(lambda: ops.coverage_gain_host(pn, hn, cn, device=0)))
print("mac_covgain_host pageable %.3f ms" % wall(lambda: ops.coverage_gain_host(pts.numpy(), harm.numpy(), cn, device=0), 5))

# two pinned input sets used alternately (what bench.py does)
pts2, harm2, _ = synth.covgain_inputs(1, 200704, 1, seed=2)
pp2, hp2 = pts2.pin_memory(), harm2.pin_memory()
sets = [(pn, hn), (pp2.numpy(), hp2.numpy())]
state = {"i": 0}
def alt():
    p, h = sets[state["i"] & 1]
    state["i"] += 1
    s = torch.from_numpy(ops.coverage_gain_host(p, h, cn, device=0))
    return s, torch.argmax(s, dim=-1)
print("alternating 2 pinned sets %.3f ms" % wall(alt))
big = [torch.empty(1, 200704, 64, device=dev) for _ in range(5)]   # bench.py keeps 5 resident input sets
print("with 5 resident sets      %.3f ms" % wall(alt))
