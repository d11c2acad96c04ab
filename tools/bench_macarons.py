#!/usr/bin/env python3
"""MACARONS candidate scoring (BASELINE.json configs[2]: 128 candidate poses over one scene's proxy points):
batched `predict_coverage_gains_for_cameras` vs the reference's call pattern (one
`predict_coverage_gain_for_single_camera` per candidate) vs the CPU oracle (= the reference's arithmetic).
    python tools/bench_macarons.py [--n 100000] [--c 128] [--iters 5] [--cpu-cands 2]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import macarons_case  # noqa: E402
import synth  # noqa: E402
from macarons_b200 import netpack, ops  # noqa: E402
from macarons_b200.networks.Macarons import Macarons  # noqa: E402
from macarons_b200.networks.SconeVis import SconeVis  # noqa: E402
from macarons_b200.utility import macarons_utils as mu  # noqa: E402


def timed(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--c", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-cands", type=int, default=2)
    ap.add_argument("--no-loop", action="store_true", help="skip the per-candidate loop (profiling runs)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    S = 2048
    s, params, cams, pred, camera, proxy_scene, surface_scene, nb, _ = macarons_case.build(args.n, args.c, 11, 17.0, S, device=dev)
    vis = SconeVis()
    sd = synth.seeded_state_dict(vis.state_dict(), 5)
    vis.load_state_dict(sd)
    macarons = Macarons(None, None, vis.to(dev).eval())
    X, vh, occ = s["X_world"].to(dev), s["vh"].to(dev), s["occ"].to(dev)
    from oracle import cameras as o_cams   # stands in for pytorch3d's FoVPerspectiveCameras on this box
    batch_cam = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=1000., device=dev)
    X_cams = batch_cam.get_camera_center()
    u = torch.rand(args.c, S, generator=torch.Generator().manual_seed(1)).to(dev)
    res = {"N": args.n, "C": args.c, "seq_len": S}
    with torch.no_grad():
        batched = lambda: mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, X, vh, occ, camera,
                                                                X_cams, batch_cam, prediction_camera=pred, samples=u)
        out = batched()
        res["n_unique_mean"] = out["n_unique"].float().mean().item()
        res["batched_ms"] = timed(batched, args.iters)
        res["batched_candidates_per_s"] = args.c / res["batched_ms"] * 1e3

        def loop():
            best, best_c = -1.0, 0
            for c in range(args.c):
                cov = mu.predict_coverage_gain_for_single_camera(params, macarons, proxy_scene, surface_scene, X, vh, occ, camera,
                                                                 X_cams[c:c + 1], cams[c], prediction_camera=pred)[3]
                if cov.shape[0] > 0 and cov > best:     # testers/scene.py:454
                    best, best_c = cov, c
            return best_c
        if not args.no_loop:
            res["per_candidate_loop_ms"] = timed(loop, max(1, args.iters // 2), warmup=1)

        # stages of the batched pass
        rows = mu._camera_rows(cams, dev)
        res["stage_fov_select_sample_ms"] = timed(lambda: ops.fov_sample_proxy(X, occ, vh, rows, nb, 70., 0.1, u), args.iters)
        w = netpack.pack_sconevis(macarons.visibility)
        res["stage_sconevis_ragged_ms"] = timed(lambda: ops.sconevis_forward(w, out["proxy_points"], out["view_harmonics"],
                                                                            lens=out["n_unique"]), args.iters)
        pts_s = torch.gather(out["proxy_points"], 1, out["sample_idx"][..., None].expand(-1, -1, 4))
        harm_s = torch.gather(out["harmonics"], 1, out["sample_idx"][..., None].expand(-1, -1, 64))
        res["stage_visibility_ms"] = timed(lambda: ops.visibility_gains(pts_s, harm_s, out["X_cam"].view(args.c, 1, 3)), args.iters)

    # CPU: the oracle (the reference's own torch arithmetic) on a few candidates
    if args.cpu_cands > 0:
        from oracle import macarons_cov as o_mcov
        s2, _, cams_cpu, pred_cpu, _, _, _, _, _ = macarons_case.build(args.n, args.c, 11, 17.0, S)
        t0 = time.perf_counter()
        with torch.no_grad():
            for c in range(args.cpu_cands):
                o_mcov.predict_coverage_gain_for_single_camera(sd, s2["X_world"], s2["vh"], s2["occ"], cams_cpu[c].get_camera_center(),
                                                               cams_cpu[c], pred_cpu, nb, s2["diag"], seq_len=S, u=u[c].cpu().view(-1, 1))
        res["cpu_oracle_ms_per_candidate"] = 1e3 * (time.perf_counter() - t0) / args.cpu_cands
        res["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
