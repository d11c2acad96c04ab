#!/usr/bin/env python3
"""Benchmark of the NBV scoring hot path (BASELINE.json metric: candidate-camera coverage-gain
evaluations per second; one evaluation = one (cloud, camera) score integrated over all P surface points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

A step = one pass of `SconeVis.compute_coverage_gain` over the workload's point set for this rank's slice of
the candidate cameras, the all-gather of the per-candidate scores and the replicated argmax (NBV index).
Prints ONE JSON line on rank 0 (contract in the task statement; keys documented in DESIGN.md section 6).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[4] / north_star target shape; also the N=1 workload (fits one GPU)
    "cfg5": dict(B=1, P=200704, C=512, desc="200704 surface points x 512 candidate cameras, 1 cloud"),
    "cfg2": dict(B=1, P=2048, C=64, desc="2048 proxy points x 64 candidate cameras, 1 cloud"),
    "cfg4": dict(B=32, P=2048, C=256, desc="32 clouds x 2048 proxy points x 256 candidate cameras"),
}
METRIC = "coverage_gain_evals_per_sec"
UNIT = "evals/s"
N_INPUT_SETS = 5          # distinct resident input sets rotated between steps (5 x 54.6 MB > 126 MB L2)
INSTR_PER_PAIR = 94.3     # SASS count of the covgain inner loop (profiles/r01_covgain_sass.md)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin <= t <= t_end + 0.1 and len(r) >= 7] or [r for _, r in self.rows[-3:]]
        if not rows:
            return None
        def num(x):
            try:
                return float(x)
            except ValueError:
                return float("nan")
        sm = [num(r[0]) for r in rows]
        reasons = [n for i, n in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"),
                                  (6, "sw_power_cap")) if any(r[i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": num(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(num(r[2]) for r in rows)}


def make_inputs(cfg, n_sets):
    import torch
    import synth
    sets = []
    for i in range(n_sets):
        pts, harm, _ = synth.covgain_inputs(cfg["B"], cfg["P"], 1, seed=5000 + i)
        sets.append((pts, harm))
    cams = synth.fibonacci_cameras(cfg["C"])[None].expand(cfg["B"], -1, -1).contiguous()
    return sets, cams


def run_reference(args, cfg):
    """CPU arm: the oracle port of the reference arithmetic (the reference is pure Python and is not on the
    GPU box), all host threads torch gives us, on a bounded camera sample of the same workload per step."""
    import torch
    from oracle import sh_cov
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    (pts, harm), = make_inputs(cfg, 1)[0][:1]
    cams = make_inputs(cfg, 0)[1]
    n_cam = max(1, min(cfg["C"], (2_000_000 // (cfg["B"] * cfg["P"])) or 1))   # ~2 M pairs per step
    chunk = max(1, 900_000 // (cfg["B"] * cfg["P"]))                            # <= ~230 MB basis temporary
    sample = cams[:, :n_cam].contiguous()
    for _ in range(max(1, args.warmup // 3)):
        sh_cov.coverage_gain(pts, harm, sample, cam_chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sh_cov.coverage_gain(pts, harm, sample, cam_chunk=chunk)
    dt = time.perf_counter() - t0
    value = cfg["B"] * n_cam * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + cfg["desc"], "B": cfg["B"], "P": cfg["P"], "C": cfg["C"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of %d cameras x all %d points per step (oracle/sh_cov.py, torch CPU fp32)"
                                       % (n_cam, cfg["C"], cfg["P"])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(cfg, budget_s=15.0):
    import torch
    from oracle import sh_cov
    (pts, harm), = make_inputs(cfg, 1)[0][:1]
    cams = make_inputs(cfg, 0)[1]
    chunk = max(1, 900_000 // (cfg["B"] * cfg["P"]))
    n_cam = max(1, min(cfg["C"], (1_000_000 // (cfg["B"] * cfg["P"])) or 1))
    t0 = time.perf_counter()
    sh_cov.coverage_gain(pts, harm, cams[:, :n_cam].contiguous(), cam_chunk=chunk)   # warm-up + calibration
    per_cam = (time.perf_counter() - t0) / n_cam
    n_cam = int(max(n_cam, min(cfg["C"], budget_s / max(per_cam, 1e-6) / 3)))
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        sh_cov.coverage_gain(pts, harm, cams[:, :n_cam].contiguous(), cam_chunk=chunk)
        times.append(time.perf_counter() - t0)
    return {"value": cfg["B"] * n_cam / statistics.median(times), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": "%d of %d cameras x all %d points, median of 3 (oracle/sh_cov.py, torch CPU fp32)"
                      % (n_cam, cfg["C"], cfg["P"])}


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from macarons_b200 import ops, parallel
    from macarons_b200.networks.SconeVis import SconeVis

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, P, C = cfg["B"], cfg["P"], cfg["C"]
    c0, c1 = parallel.camera_partition(C, world, rank)

    torch.manual_seed(5)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        vis = SconeVis().to(dev)

    host_sets, cams_h = make_inputs(cfg, N_INPUT_SETS)
    dev_sets = [(p.to(dev), h.to(dev)) for p, h in host_sets]
    cams = cams_h.to(dev)
    pinned = [(p.pin_memory(), h.pin_memory()) for p, h in host_sets[:2]]
    cams_pin = cams_h.pin_memory()
    local = torch.zeros(B, C, device=dev)

    def step(i, ev=None):
        pts, harm = dev_sets[i % N_INPUT_SETS]
        if ev is not None:
            ev[0].record()
        ops.coverage_gain(pts, harm, cams, use_sigmoid=vis.use_sigmoid, cam_range=(c0, c1), out=local)
        if ev is not None:
            ev[1].record()
        scores = parallel.gather_scores(local, C)
        return scores, parallel.nbv_argmax(scores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    barrier()
    wall0 = time.perf_counter()
    t_begin.record()
    for i in range(args.steps):
        scores, best = step(args.warmup + i, kev[i])
    t_end.record()
    barrier()
    wall1 = time.perf_counter()
    launches = ops.launch_count() - n0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    elapsed_ms = t_begin.elapsed_time(t_end)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    if world > 1:
        t = torch.tensor([elapsed_ms, kernel_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = t.tolist()

    # ---- end to end through the public API: pinned host inputs -> device, score, gather, argmax -> host ----
    e2e_steps = max(3, min(args.steps, 10))
    h2d = sum(t.numel() * 4 for t in pinned[0]) + cams_pin.numel() * 4
    d2h = B * C * 4 + B * 8

    def e2e_step(i):
        p_h, h_h = pinned[i % 2]
        pts = p_h.to(dev, non_blocking=True)
        harm = h_h.to(dev, non_blocking=True)
        cam_d = cams_pin.to(dev, non_blocking=True)
        s, b = parallel.sharded_coverage_gain(vis.compute_coverage_gain, pts, harm, cam_d)
        return s.cpu(), b.cpu()

    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        s_host, b_host = e2e_step(i)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()

    # ---- parity of what was timed: last step's argmax vs the float64 closed form on a camera subset ----
    check = None
    if rank == 0:
        import numpy as np
        from oracle import sh_cov
        pts_h, harm_h = host_sets[(args.warmup + args.steps - 1) % N_INPUT_SETS]
        got = scores.cpu().numpy()
        top = np.argsort(-got[0])[:4].tolist()
        truth = sh_cov.coverage_gain_f64(pts_h.numpy(), harm_h.numpy(), cams_h.numpy()[:, top])
        check = {"nbv_index": int(best[0]), "max_abs_err_vs_f64_top4": float(np.abs(got[:, top] - truth).max()),
                 "top1_is_argmax_of_truth": bool(np.argmax(truth[0]) == 0)}

    if rank == 0:
        peaks, peak_src = load_peaks()
        n_local = c1 - c0
        alg_bytes = B * P * (4 * vis.pts_dim + 256) + B * n_local * 12 + B * n_local * 4
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = 148 * 4 * sm_mhz * 1e6                    # warp instructions / s at the sampled clock
        issue_rate = B * P * n_local * INSTR_PER_PAIR / 32 / (kernel_ms * 1e-3)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "covgain_traffic.json")
        if os.path.exists(tpath) and world == 1 and args.workload == "cfg5":
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": B * C * args.steps / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + cfg["desc"], "B": B, "P": P, "C": C,
                       "parallelism": "camera axis sharded x%d (points replicated), 1 all-gather of scores" % world,
                       "cameras_per_gpu": n_local,
                       "l2": "inputs rotate over %d distinct resident sets (%.0f MB > 126 MB L2)"
                             % (N_INPUT_SETS, N_INPUT_SETS * B * P * 272 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": B * C * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "pinned host tensors -> SconeVis.compute_coverage_gain -> parallel.gather_scores -> argmax -> .cpu()"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                         "kernel": "covgain_kernel<sigmoid,reduce>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "kernel is fp32-issue bound at >= 8 cameras per point pass (DESIGN.md 4.3)",
                         "fp32_issue": {"instr_per_pair": INSTR_PER_PAIR, "warp_instr_per_s": issue_rate,
                                        "peak_warp_instr_per_s": issue_peak, "frac": issue_rate / issue_peak,
                                        "sm_mhz": sm_mhz}},
            "parity": check,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
