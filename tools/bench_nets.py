#!/usr/bin/env python3
"""Device timings of the fused SconeVis / SconeOcc forwards (CUDA events, after warm-up).
    python tools/bench_nets.py [--q 262144] [--n 4096] [--iters 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from macarons_b200 import ops  # noqa: E402
from macarons_b200.networks.SconeOcc import SconeOcc  # noqa: E402
from macarons_b200.networks.SconeVis import SconeVis  # noqa: E402


def timed(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters, (ops.launch_count() - n0) // iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=64 ** 3)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--s", type=int, default=2048)
    ap.add_argument("--clouds", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip-occ", action="store_true")
    ap.add_argument("--depth", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {}
    with torch.no_grad():
        vis = SconeVis()
        vis.load_state_dict(synth.seeded_state_dict(vis.state_dict(), 5))
        vis = vis.to(dev).eval()
        pts, vh = synth.sconevis_inputs(args.clouds, args.s, 1)
        pts, vh = pts.to(dev), vh.to(dev)
        ms, launches = timed(lambda: vis(pts, view_harmonics=vh), args.iters)
        res["sconevis_forward"] = {"B": args.clouds, "S": args.s, "ms": ms, "launches": launches,
                                   "gflop": 13.7 * args.clouds * args.s / 2048, "tflops": 13.7e-3 * args.clouds * args.s / 2048 / ms}
        if not args.skip_occ:
            occ = SconeOcc()
            occ.load_state_dict(synth.seeded_state_dict(occ.state_dict(), 5))
            occ = occ.to(dev).eval()
            occ.queries_per_pass = args.chunk
            pc, x, qvh = synth.sconeocc_inputs(1, args.n, 4096, 2)
            x = torch.rand(1, args.q, 3) - 0.5
            qvh = 0.3 * torch.randn(1, args.q, 64)
            pc, x, qvh = pc.to(dev), x.to(dev), qvh.to(dev)
            ms, launches = timed(lambda: occ(pc, x, qvh), max(1, args.iters // 2), warmup=1)
            flop = 26.54e6 * args.q + 3.76e9
            res["sconeocc_forward"] = {"N": args.n, "Q": args.q, "chunk": args.chunk, "ms": ms, "launches": launches,
                                       "tflop": flop / 1e12, "tflops": flop / 1e9 / ms,
                                       "queries_per_s": args.q / ms * 1e3}
        if args.depth:
            from macarons_b200.networks import ManyDepth as MD
            resnet = MD.ResNet18Trunk()
            depth = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet), None)
            depth.load_state_dict(synth.seeded_state_dict(depth.state_dict(), 5))
            depth = depth.to(dev).eval()
            t = [v.to(dev) for v in synth.depth_inputs(1, 256, 456, 7)]
            ms, launches = timed(lambda: depth(t[0], t[1], t[2], t[3], t[4], dev, gt_pose=t[5]), args.iters)
            res["manydepth_forward"] = {"H": 256, "W": 456, "n_alpha": 2, "ms": ms, "launches": launches,
                                        "gflop": 22.2, "tflops": 22.2e-3 / ms}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
