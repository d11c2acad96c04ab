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
FMA_CYCLES_PER_PAIR = 74.0  # FMA-pipe cycles per (point, camera) pair: 65 FFMA + 4 FADD + 4 FMUL + 1 FADD of the tile sum
                            # (SASS of the sweep loop); 1 per cycle per SM sub-partition


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU through NVML every ~2 ms while the timed region
    runs (nvidia-smi's own polling loop is too coarse for a region of a few tens of milliseconds)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self._nv = None

    def _loop(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h), nv.nvmlDeviceGetPowerUsage(h) / 1e3))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self, t_begin, t_end):
        if self._thread is None:
            return None
        self._stop.set()
        self._thread.join(timeout=1.0)
        nv = self._nv
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end] or self.rows[-3:]
        if not rows:
            return None
        bits = 0
        for r in rows:
            bits |= r[2]
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksEventReasonHwPowerBrakeSlowdown"))
        reasons = [n for n, attr in names if bits & getattr(nv, attr, 0)]
        try:
            sm_max = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception:
            sm_max = None
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_mhz_min": min(r[1] for r in rows), "sm_max_mhz": sm_max,
                "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[3] for r in rows), "source": "NVML, 2 ms period"}


def bind_to_gpu_numa_node(gpu_index):
    """Run this process on the CPU cores of the NUMA node the GPU hangs off, BEFORE any host buffer is allocated: pinned
    host memory is placed by first touch, and a buffer that lands on the other socket feeds the GPU's PCIe link through
    the inter-socket fabric (measured on the 2-socket GPU boxes: 1.7-2.0 ms instead of 1.15 ms for the 54.6 MB of one
    step).  This is the host-side placement any deployment of a host-buffer API does (numactl --cpunodebind --membind).
    -> a short description for the JSON line (None when the platform exposes no NUMA information)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(gpu_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return "GPU %s on NUMA node %d: process bound to its %d allowed cores" % (bdf, node, len(allowed))
    except Exception:
        return None


def make_inputs(cfg, n_sets):
    import torch
    import synth
    sets = []
    for i in range(n_sets):
        pts, harm, _ = synth.covgain_inputs(cfg["B"], cfg["P"], 1, seed=5000 + i)
        sets.append((pts, harm))
    cams = synth.fibonacci_cameras(cfg["C"])[None].expand(cfg["B"], -1, -1).contiguous()
    return sets, cams


def make_config(args, cfg):
    """The `config` object of the JSON line: identical, key for key, for the GPU arm and for `--impl reference` (the
    driver compares the two); what is specific to one arm lives outside it (`exchange`, `cpu_baseline.sample`)."""
    world = max(1, int(args.gpus))
    return {"workload": args.workload + ": " + cfg["desc"], "B": cfg["B"], "P": cfg["P"], "C": cfg["C"],
            "parallelism": "GPU arm: camera axis sharded x%d (points replicated), one exchange of the per-camera scores; "
                           "CPU arm: rank 0 only, all host threads" % world,
            "cameras_per_gpu": -(-cfg["C"] // world),
            "l2": "GPU arm: inputs rotate over %d distinct resident sets (%.0f MB > 126 MB L2); CPU arm: one input set"
                  % (N_INPUT_SETS, N_INPUT_SETS * cfg["B"] * cfg["P"] * 272 / 1e6)}


def host_threads():
    """All host cores this process may run on, regardless of OMP_NUM_THREADS (torchrun sets it
    to 1 in every rank)."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference(args, cfg):
    """CPU arm: the oracle port of the reference arithmetic (the reference is pure Python and is not on the
    GPU box), all host threads, on a bounded camera sample of the same workload per step: the sample is calibrated
    so that the whole `--steps K --warmup W` run takes about two minutes (all C cameras when that fits)."""
    import torch
    from oracle import sh_cov
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    (pts, harm), = make_inputs(cfg, 1)[0][:1]
    cams = make_inputs(cfg, 0)[1]
    chunk = max(1, 900_000 // (cfg["B"] * cfg["P"]))                            # <= ~230 MB basis temporary
    n_cal = max(1, min(cfg["C"], (1_000_000 // (cfg["B"] * cfg["P"])) or 1))
    sh_cov.coverage_gain(pts, harm, cams[:, :n_cal].contiguous(), cam_chunk=chunk)          # page in / thread pool
    t0 = time.perf_counter()
    sh_cov.coverage_gain(pts, harm, cams[:, :n_cal].contiguous(), cam_chunk=chunk)
    per_cam = (time.perf_counter() - t0) / n_cal
    n_warm = max(1, args.warmup // 3)
    budget_s = float(os.environ.get("MAC_BENCH_REFERENCE_BUDGET_S", "120"))
    n_cam = int(max(1, min(cfg["C"], budget_s / (args.steps + n_warm) / max(per_cam, 1e-9))))
    sample = cams[:, :n_cam].contiguous()
    for _ in range(n_warm):
        sh_cov.coverage_gain(pts, harm, sample, cam_chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sh_cov.coverage_gain(pts, harm, sample, cam_chunk=chunk)
    dt = time.perf_counter() - t0
    value = cfg["B"] * n_cam * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": make_config(args, cfg),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of %d cameras x all %d points per step (oracle/sh_cov.py = the reference's torch "
                                       "CPU fp32 arithmetic, bit-pinned to it; evals/s is per camera, so the sample size "
                                       "does not bias it)" % (n_cam, cfg["C"], cfg["P"])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(cfg, budget_s=15.0):
    import torch
    from oracle import sh_cov
    host_threads()
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


def path_stages(dev):
    """Device times (CUDA events, 1 warm-up + mean of 3) of the other stages of the NBV path at their BASELINE.json
    shapes, on rank 0 at N = 1 only; informational (the headline metric is the coverage-gain scoring stage)."""
    import contextlib
    import io
    import torch
    import synth
    from macarons_b200.networks import ManyDepth as MD
    from macarons_b200.networks.SconeOcc import SconeOcc
    from macarons_b200.networks.SconeVis import SconeVis
    from macarons_b200.utility import scone_utils

    def timed(fn, iters=3):
        fn()
        fn()   # two warm-up calls: the depth forward captures its CUDA graph on the second call with a given shape
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        vis, occ = SconeVis(), SconeOcc()
        vis.load_state_dict(synth.seeded_state_dict(vis.state_dict(), 5))
        occ.load_state_dict(synth.seeded_state_dict(occ.state_dict(), 5))
        vis, occ = vis.to(dev).eval(), occ.to(dev).eval()
        pts, vh = (t.to(dev) for t in synth.sconevis_inputs(1, 2048, 1))
        out["sconevis_forward_2048_tokens_ms"] = timed(lambda: vis(pts, view_harmonics=vh))
        pc = synth.sconeocc_inputs(1, 4096, 8, 2)[0].to(dev)
        x = (torch.rand(1, 64 ** 3, 3) - 0.5).to(dev)
        xvh = (0.3 * torch.randn(1, 64 ** 3, 64)).to(dev)
        out["sconeocc_forward_64cube_queries_ms"] = timed(lambda: occ(pc, x, xvh), iters=2)
        base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
        big = (torch.rand(1, 200704, 3) - 0.5).to(dev)
        views = synth.sphere_cameras(10, 1.5, torch.Generator().manual_seed(1)).to(dev)
        out["view_state_plus_harmonics_200704_pts_10_views_ms"] = timed(lambda: scone_utils.compute_view_harmonics(
            scone_utils.compute_view_state(big, views, 7, 14), base, h_polar, h_azim, 7, 14))
        out["view_state_harmonics_fused_200704_pts_10_views_ms"] = timed(lambda: scone_utils.compute_view_state_harmonics(
            big, views, base, h_polar, h_azim, 7, 14), iters=10)
        resnet = MD.ResNet18Trunk()
        depth = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet), None)
        depth.load_state_dict(synth.seeded_state_dict(depth.state_dict(), 5))
        depth = depth.to(dev).eval()
        d = [t.to(dev) for t in synth.depth_inputs(1, 256, 456, 7)]
        out["manydepth_forward_256x456_ms"] = timed(lambda: depth(d[0], d[1], d[2], d[3], d[4], dev, gt_pose=d[5]))
    try:
        out.update(chained_configs(dev, timed, vis, occ, depth_256=None))
    except Exception as exc:   # informational
        out["chained_configs_error"] = "%s: %s" % (type(exc).__name__, exc)
    out["note"] = "fp32-accurate (3xTF32) tcgen05 linear layers; reference on 8 CPU threads: SconeVis 140 ms, SconeOcc 64^3 ~77 s (SURVEY section 6)"
    return out


def online_loop(args, cfg, dev, vis, board, world, rank, n_steps=50, seq_len=2048):
    """50-step online NBV loop at the cfg5 shape (macarons_b200.nbv.scone_online_loop); device time by CUDA events,
    max over ranks; the chosen camera sequence must be identical on every rank and (rank 0) to a single-GPU rerun."""
    import torch
    import torch.distributed as dist
    import synth
    from macarons_b200 import nbv
    from macarons_b200.utility import scone_utils
    B, P, C = cfg["B"], cfg["P"], cfg["C"]
    pts_h, _, _ = synth.covgain_inputs(1, P, 1, seed=5100)         # xyz ~ U[-0.5, 0.5]^3, occupancy ~ U[0.1, 1]
    pts = pts_h.to(dev)
    cams = synth.fibonacci_cameras(C)[None].contiguous().to(dev)
    X_view = synth.sphere_cameras(1, 1.5, torch.Generator().manual_seed(51)).to(dev)
    vis.load_state_dict(synth.seeded_state_dict(vis.state_dict(), 5))
    base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
    shard = world > 1

    def score(p, h, c):
        if board is not None:
            return board.step(p, h, c, use_sigmoid=vis.use_sigmoid)
        s = vis.compute_coverage_gain(p, h, c)
        return s, s.argmax(-1)

    def run():
        return nbv.scone_online_loop(vis, pts, cams, X_view, base, h_polar, h_azim, n_steps, score_step=score,
                                     seq_len=seq_len, shard_clouds=shard)

    nbv.scone_online_loop(vis, pts, cams, X_view, base, h_polar, h_azim, 2, score_step=score, seq_len=seq_len,
                          shard_clouds=shard)                      # warm-up (workspaces, packs)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    chosen, _ = run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    chosen = chosen.clone()
    same = True
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        ref = chosen.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([int(torch.equal(ref, chosen))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    out = {"steps": n_steps, "ms_per_step": ms / n_steps, "evals_per_s": C * n_steps / (ms * 1e-3),
           "clouds": P // seq_len, "tokens_per_cloud": seq_len, "chosen_first8": chosen[:8].tolist(),
           "distinct_cameras_chosen": int(chosen.unique().numel()),
           "stages": "fused view state+harmonics (1 launch) -> SconeVis.forward on %d clouds%s -> scoring kernel with fused "
                     "exchange + argmax -> device-side append of the chosen camera; no host synchronisation inside the loop"
                     % (P // seq_len, " sharded over the ranks + NCCL all-gather of the harmonics" if shard else "")}
    if world > 1:
        out["same_sequence_on_all_ranks"] = same
        if rank == 0:   # single-GPU rerun of the same loop on rank 0: the sharded loop must reproduce it bit for bit
            solo, _ = nbv.scone_online_loop(vis, pts, cams, X_view, base, h_polar, h_azim, n_steps, seq_len=seq_len)
            out["same_sequence_as_single_gpu"] = bool(torch.equal(solo, chosen))
    return out


def chained_configs(dev, timed, vis, occ, depth_256=None):
    """BASELINE.json configs[1..3] as CHAINED steps through the package's public API (device time, CUDA events):
      cfg2  SCONE step: view harmonics -> SconeOcc on a 64^3 grid -> proxy sampling -> SconeVis -> 64 candidates -> argmax
      cfg3  full MACARONS NBV step: depth 256x256 -> partial cloud -> scene update -> occupancy field (one ragged SconeOcc
            forward over all occupied cells) -> 128 candidate poses -> argmax, on a scene populated by 3 earlier frames
      cfg4  32 clouds x (64^3 grid + 256 candidates): what every GPU runs when depth / occupancy are replicated and only
            the camera axis is sharded (north_star); at N GPUs the coverage part shrinks to 256 / N cameras per cloud."""
    import contextlib
    import io
    import types
    import torch
    import synth
    from macarons_b200 import nbv
    from macarons_b200.networks import ManyDepth as MD
    from macarons_b200.networks.Macarons import Macarons
    from macarons_b200.utility import cameras, scene, scone_utils
    out = {}
    gen = torch.Generator().manual_seed(2)
    base, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(8, 7, 14, dev)
    lin = (torch.arange(64, dtype=torch.float32) + 0.5) / 64 - 0.5
    X = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), dim=-1).view(1, -1, 3).to(dev)
    X_view = synth.sphere_cameras(2, 1.5, gen).to(dev)

    def scone_step(seed, n_cams):
        pc = synth.airplane_surface(4096, torch.Generator().manual_seed(seed))[None].to(dev)
        return nbv.scone_nbv_step(occ, vis, pc, X, X_view, synth.fibonacci_cameras(n_cams).to(dev), base, h_polar, h_azim,
                                  seq_len=2048, max_points_per_pass=300000)

    out["cfg2_scone_step_64cube_grid_64_candidates_ms"] = timed(lambda: scone_step(1, 64), iters=2)
    out["cfg4_32_clouds_64cube_grid_256_candidates_per_gpu_ms"] = timed(lambda: [scone_step(100 + b, 256) for b in range(32)],
                                                                        iters=1)

    # ---- cfg3 ----
    H = W = 256
    with contextlib.redirect_stdout(io.StringIO()):
        resnet = MD.ResNet18Trunk()
        depth = MD.ManyDepth(MD.DepthDecoder(MD.FeatureExtractor(resnet), resnet, input_height=H, input_width=W), None)
    depth.load_state_dict(synth.seeded_state_dict(depth.state_dict(), 5))
    macarons = Macarons(depth.to(dev).eval(), occ, vis)
    params = types.SimpleNamespace(harmonic_degree=8, view_state_n_elev=7, view_state_n_azim=14, n_view_state_cameras=98,
                                   n_harmonics=64, k_for_knn=16, prediction_neighborhood_size=3, jz=False, ddp=False,
                                   znear=0.5, zfar=750., gathering_factor=0.05, sensor_range=6.0, carving_tolerance=0.3,
                                   seq_len=2048, min_occ_for_proxy_points=0.1, use_occ_to_sample_proxy_points=True,
                                   distance_factor_th=17.0, image_height=H, image_width=W)
    x_min, x_max = torch.tensor([-3.5, -2.0, -4.0]), torch.tensor([3.5, 2.5, 3.0])
    common = dict(x_min=x_min.to(dev), x_max=x_max.to(dev), grid_l=5, grid_w=3, grid_h=5, n_proxy_points=100000, device=dev,
                  view_state_n_elev=7, view_state_n_azim=14)
    surface_scene = scene.Scene(cell_capacity=1000, cell_resolution=None, feature_dim=1, **common)
    proxy_scene = scene.Scene(cell_capacity=100000, cell_resolution=0.001, feature_dim=1, score_threshold=0.95, **common)
    proxy_scene.initialize_proxy_points()
    nb = synth.ndc_bounds(H, W)
    eyes = torch.tensor([[0.2, 0.4, -3.2], [2.0, 0.8, -2.0], [-2.2, 0.6, -1.5], [1.0, 1.0, 2.2], [-1.2, 0.5, 2.4]])
    ce = (torch.rand(128, 3, generator=gen) - 0.5) * torch.tensor([6.0, 3.5, 6.0]) + torch.tensor([0.0, 0.3, -0.5])
    cR, cT = cameras.look_at(ce, (torch.rand(128, 3, generator=gen) - 0.5) * 2.0)
    cand = cameras.FoVCamera(cR, cT, zfar=750., device=dev)
    timings, n_timed = {}, 0
    for i in range(eyes.shape[0]):
        R, T = cameras.look_at(eyes[i:i + 1], torch.zeros(1, 3))
        cam = cameras.FoVCamera(R, T, zfar=750., device=dev)
        camera = types.SimpleNamespace(image_height=H, image_width=W, zfar=750., gathering_factor=0.05, fov_camera=cam,
                                       fov_camera_0=cam if i == 0 else camera0, X_cam=eyes[i:i + 1].to(dev), min_ndc_x=nb[0],
                                       max_ndc_x=nb[1], min_ndc_y=nb[2], max_ndc_y=nb[3])
        if i == 0:
            camera0 = cam
        fr = synth.depth_inputs(1, H, W, 20 + i)
        frames = {k: v.to(dev) for k, v in zip(("x", "x_alpha", "R", "T", "zfar", "gt_pose"), fr)}
        t = timings if i >= 3 else None
        n_timed += i >= 3
        cov, best, st = nbv.macarons_nbv_step(params, macarons, camera, surface_scene, proxy_scene, frames, ce.to(dev), cand,
                                              timings=t)
    stages = {k: v / n_timed for k, v in timings.items()}
    out["cfg3_full_macarons_step_256x256_128_candidates_ms"] = sum(stages.values())
    out["cfg3_stages"] = dict(stages, proxy_points=100000, proxy_points_scored=int(st["X_world"].shape[0]),
                              scene_cells="5x3x5", frames_before_timing=3)
    return out


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
    numa = None if args.no_numa_bind else bind_to_gpu_numa_node(local_rank)
    # NCCL prints its version banner on the process' stdout: keep stdout clean for the ONE JSON line by sending
    # everything else to stderr until the line is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
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
    # Pinned staging buffers.  On the (virtualised, 2-socket) GPU hosts the guest sees ONE NUMA node, but a pinned buffer
    # may physically sit behind the other socket: the SAME 54.6 MB then take 1.5-1.9 ms instead of 1.0-1.15 ms to reach
    # the GPU, per buffer and reproducibly.  A caller of a host-buffer API picks its staging memory once, so: pin a few
    # candidates, time one upload of each (a bandwidth probe, outside every timed region) and keep the two best.
    n_cand = min(len(host_sets), 5) if (world == 1 and not args.no_staging_probe) else 2
    cand = [(p.pin_memory(), h.pin_memory()) for p, h in host_sets[:n_cand]]
    probe_ms = []
    for ci, (p_pin, h_pin) in enumerate(cand):
        best_ms = float("inf")
        for _ in range(3):
            torch.cuda.synchronize()
            t_p = time.perf_counter()
            dev_sets[ci][0].copy_(p_pin, non_blocking=True)
            dev_sets[ci][1].copy_(h_pin, non_blocking=True)
            torch.cuda.synchronize()
            best_ms = min(best_ms, 1e3 * (time.perf_counter() - t_p))
        probe_ms.append(best_ms)
    keep = sorted(sorted(range(n_cand), key=lambda c: probe_ms[c])[:2])
    pinned = [cand[c] for c in keep]
    pinned_src = keep          # pinned[j] holds host_sets[pinned_src[j]]
    del cand
    cams_pin = cams_h.pin_memory()
    local = torch.zeros(B, C, device=dev)
    board, exchange = None, "nccl all_gather + torch.argmax"
    if not args.nccl_gather:
        try:
            board = parallel.PeerScoreBoard(B, C, dev)
            exchange = ("fused: the finishing CTA of the scoring kernel pushes the scores to every peer's board (NVLink P2P), waits for "
                        "the peers' flags and takes the argmax: 1 launch per step")
        except Exception as exc:  # symmetric memory unavailable: fall back to the NCCL collective
            exchange += " (peer board unavailable: %s)" % type(exc).__name__

    def step(i, ev=None):
        pts, harm = dev_sets[i % N_INPUT_SETS]
        if board is not None:
            return board.step(pts, harm, cams, use_sigmoid=vis.use_sigmoid, events=ev)
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
        time.sleep(0.02)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for ev_a, ev_b in kev:      # torch creates the CUDA event at its first record(): do that outside the timed region
        ev_a.record()
        ev_b.record()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    if board is not None:
        board.reset_wait_stats()
    barrier()
    wall0 = time.perf_counter()
    t_begin.record()
    for i in range(args.steps):
        scores, best = step(args.warmup + i, kev[i])
    t_end.record()
    scores, best = scores.clone(), best.clone()   # the fused path returns views of a live score board
    barrier()
    wall1 = time.perf_counter()
    launches = ops.launch_count() - n0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    elapsed_ms = t_begin.elapsed_time(t_end)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    # time the finishing CTA of the fused step spent waiting for the slowest peer (measured inside the kernel with
    # %globaltimer): kernel_ms contains it, kernel_ms - peer_wait_ms is this rank's own scoring work
    peer_wait_ms = (board.wait_stats()[0] * 1e-6) if board is not None else 0.0
    if world > 1:
        t = torch.tensor([elapsed_ms, kernel_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = t.tolist()
        w = torch.tensor([peer_wait_ms], device=dev)
        w_min, w_max = w.clone(), w.clone()
        dist.all_reduce(w_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(w_max, op=dist.ReduceOp.MAX)
        peer_wait = {"min_over_ranks_ms": w_min.item(), "max_over_ranks_ms": w_max.item()}
    else:
        peer_wait = {"min_over_ranks_ms": peer_wait_ms, "max_over_ranks_ms": peer_wait_ms}

    # ---- end to end through the public API: pinned host inputs -> device, score, gather, argmax -> host ----
    e2e_steps = max(3, min(args.steps, 30))
    h2d = sum(t.numel() * 4 for t in pinned[0]) + cams_pin.numel() * 4
    d2h = B * C * 4 + B * 8

    pinned_np = [(p.numpy(), h.numpy()) for p, h in pinned]   # views of the pinned host buffers
    cams_np = cams_pin.numpy()

    def e2e_step(i):
        if world == 1:
            # the C-ABI call with HOST buffers (mac_covgain_host): H2D slices overlapped with the kernel, D2H of the scores
            p_h, h_h = pinned_np[i % 2]
            s = torch.from_numpy(ops.coverage_gain_host(p_h, h_h, cams_np, use_sigmoid=vis.use_sigmoid, device=local_rank))
            return s, parallel.nbv_argmax(s)
        # N > 1: the step partitioned over the POINTS: every rank uploads only its own rows (tile-aligned 1/N of every
        # cloud) from pinned host memory, integrates all C cameras over them and exchanges exact int64 partial sums
        # inside the scoring kernel (PeerScoreBoard.step_points): each input byte crosses one PCIe link once, nothing is
        # replicated over NVLink first, and the scores are bitwise those of the camera-sharded `value` step
        p_h, h_h = pinned[i % 2]
        cam_d = cams_pin.to(dev, non_blocking=True)
        if board is not None and B == 1:
            # one cloud: the rows arrive in slices on a copy stream, slice k is integrated while slice k+1 is in flight
            s, b = board.step_points_from_host(p_h, h_h, cam_d, use_sigmoid=vis.use_sigmoid, slices=e2e_slices, bufs=e2e_bufs)
        elif board is not None:
            e2e_bufs[0].copy_(p_h[:, p0:p1], non_blocking=True)
            e2e_bufs[1].copy_(h_h[:, p0:p1], non_blocking=True)
            s, b = board.step_points(e2e_bufs[0], e2e_bufs[1], cam_d, P, use_sigmoid=vis.use_sigmoid)
        else:   # no peer board: replicate the points (1/N upload + NCCL all-gather), camera-sharded step + NCCL gather
            pts = parallel.upload_rows_sharded(p_h.view(B * P, -1), dev, buf=e2e_bufs[0]).view(B, P, -1)
            harm = parallel.upload_rows_sharded(h_h.view(B * P, 64), dev, buf=e2e_bufs[1]).view(B, P, 64)
            s, b = parallel.sharded_coverage_gain(vis.compute_coverage_gain, pts, harm, cam_d)
        s_cpu = s.cpu()                      # one device -> host read; the argmax of the (replicated) scores on the host
        return s_cpu, parallel.nbv_argmax(s_cpu)

    e2e_bufs = [None, None]
    p0, p1 = parallel.point_partition(P, world, rank)
    # upload slices of >= ~8 MB: one at N = 8 (6.8 MB per rank), two at N = 4, three or four at N = 2
    e2e_slices = max(1, min(8, round((p1 - p0) * 272 / 8e6)))
    if world > 1 and board is not None:
        e2e_bufs = [torch.empty((B, p1 - p0, host_sets[0][0].shape[-1]), device=dev), torch.empty((B, p1 - p0, 64), device=dev)]
    elif world > 1:
        n_rows = -(-(B * P) // world) * world
        e2e_bufs = [torch.empty((n_rows, host_sets[0][0].shape[-1]), device=dev), torch.empty((n_rows, 64), device=dev)]

    for i in range(2):
        e2e_step(i)
    barrier()
    # the floor of this box's host link: the same bytes, pinned host -> device, plain async copies, nothing else
    floor_each = []
    if world == 1:
        dst = [torch.empty_like(t, device=dev) for t in pinned[0]]
        for i in range(7):
            torch.cuda.synchronize()
            t_i = time.perf_counter()
            for d_t, s_t in zip(dst, pinned[i % 2]):
                d_t.copy_(s_t, non_blocking=True)
            torch.cuda.synchronize()
            floor_each.append(1e3 * (time.perf_counter() - t_i))
        del dst
    e2e_sampler = ClockSampler(local_rank)
    if rank == 0:
        e2e_sampler.start()
    e2e_wall0 = time.perf_counter()
    t0 = time.perf_counter()
    e2e_each = []
    for i in range(e2e_steps):
        t_i = time.perf_counter()
        s_host, b_host = e2e_step(i)
        e2e_each.append(1e3 * (time.perf_counter() - t_i))
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    e2e_clocks = e2e_sampler.stop(e2e_wall0, time.perf_counter()) if rank == 0 else None
    if world == 1:   # the link floor again, after the timed loop (host memory / link state drifts between runs)
        dst = [torch.empty_like(t, device=dev) for t in pinned[0]]
        for i in range(6):
            torch.cuda.synchronize()
            t_i = time.perf_counter()
            for d_t, s_t in zip(dst, pinned[i % 2]):
                d_t.copy_(s_t, non_blocking=True)
            torch.cuda.synchronize()
            floor_each.append(1e3 * (time.perf_counter() - t_i))
        del dst
    e2e_median_ms = statistics.median(e2e_each)   # reported next to the mean: single steps occasionally take +0.4 ms (host jitter)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()

    # ---- parity of what was timed: the LAST timed step's scores for all C cameras against (i) the float64 closed
    # form of the reference integrand (C + OpenMP restatement, oracle/c/sh_cov_f64.c: seconds for 102.8 M pairs),
    # (ii) the reference's fp32 torch arithmetic (oracle/sh_cov.py) on a camera sample around the maximum, and
    # (iii) at N > 1 the same kernel run by rank 0 alone over all C cameras (bitwise) ----
    check = None
    single = None
    if world > 1:
        pts_d, harm_d = dev_sets[(args.warmup + args.steps - 1) % N_INPUT_SETS]
        single = ops.coverage_gain(pts_d, harm_d, cams, use_sigmoid=vis.use_sigmoid) if rank == 0 else None
        torch.cuda.synchronize()
    if rank == 0:
        import numpy as np
        from oracle import sh_cov
        host_threads()
        pts_h, harm_h = host_sets[(args.warmup + args.steps - 1) % N_INPUT_SETS]
        got = scores.cpu().numpy()
        if board is not None:
            board.check()
        t0 = time.perf_counter()
        truth = sh_cov.coverage_gain_f64_c(pts_h.numpy(), harm_h.numpy(), cams_h.numpy(), use_sigmoid=vis.use_sigmoid)
        t_truth = time.perf_counter() - t0
        order = np.argsort(-truth, axis=-1)
        check = {"nbv_index": [int(v) for v in best.reshape(-1).tolist()],
                 "cameras_checked_vs_f64": int(B * C),
                 "max_abs_err_vs_f64": float(np.abs(got - truth).max()),
                 "argmax_equals_f64_argmax": bool((got.argmax(-1) == truth.argmax(-1)).all()),
                 "f64_top1_minus_top2": float(np.min(np.take_along_axis(truth, order[:, :1], -1)
                                                     - np.take_along_axis(truth, order[:, 1:2], -1))),
                 "f64_check_seconds": t_truth}
        if B * P <= 250_000:   # fp32 reference arithmetic on the 4 best cameras of cloud 0 (a few seconds on the host)
            top = order[0, :4].tolist()
            ref32 = sh_cov.coverage_gain(pts_h[:1], harm_h[:1], cams_h[:1, top].contiguous(), use_sigmoid=vis.use_sigmoid,
                                         cam_chunk=max(1, 900_000 // P)).numpy()
            check["max_abs_err_vs_reference_fp32_top4"] = float(np.abs(got[:1, top] - ref32).max())
            check["reference_fp32_err_vs_f64_top4"] = float(np.abs(ref32 - truth[:1, top]).max())
        if single is not None:
            check["bitwise_equal_to_single_gpu"] = bool(torch.equal(single.cpu(), scores.cpu()))
            e2e_set = dev_sets[pinned_src[(e2e_steps - 1) % 2]]
            check["e2e_point_sharded_bitwise_equal_to_single_gpu"] = bool(torch.equal(
                ops.coverage_gain(e2e_set[0], e2e_set[1], cams, use_sigmoid=vis.use_sigmoid).cpu(), s_host))
            check["argmax_equal_to_single_gpu"] = bool(torch.equal(single.argmax(-1).cpu(), best.reshape(-1).cpu()))

    # ---- BASELINE.json configs[4] as the online loop SURVEY.md section 8d defines (cfg5 only): 50 NBV steps, each = view
    # harmonics of all points for the cameras visited so far -> SconeVis.forward on 98 clouds of 2048 tokens -> score all
    # C cameras (sharded) -> argmax -> the chosen camera joins the visited set.  Reported next to `value`.
    loop = None
    if args.workload == "cfg5" and not args.no_loop:
        try:
            loop = online_loop(args, cfg, dev, vis, board, world, rank)
        except Exception as exc:  # informational: never lose the headline line
            loop = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        peaks, peak_src = load_peaks()
        n_local = c1 - c0
        alg_bytes = B * P * (4 * vis.pts_dim + 256) + B * n_local * 12 + B * n_local * 4
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = 148 * 4 * sm_mhz * 1e6                    # FMA-pipe warp-cycles / s at the sampled clock
        issue_rate = B * P * n_local * FMA_CYCLES_PER_PAIR / 32 / (kernel_ms * 1e-3)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "covgain_traffic.json")
        if os.path.exists(tpath) and world == 1 and args.workload == "cfg5":
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": B * C * args.steps / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": make_config(args, cfg),
            "exchange": exchange,
            "clocks": clocks,
            "e2e": {"value": B * C * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "ms_per_step_median": e2e_median_ms,
                    "ms_each": [round(v, 3) for v in e2e_each],
                    "host_numa": numa,
                    "staging_probe_ms": [round(v, 3) for v in probe_ms], "staging_kept": keep,
                    "clocks": e2e_clocks,
                    "pinned_h2d_copy_alone_ms": statistics.median(floor_each[2:]) if floor_each else None,
                    "api": ("mac_covgain_host (C ABI, pinned HOST buffers in, host scores out; cameras + 12 point slices on a copy stream, "
                            "each slice integrated while the next one crosses PCIe) + argmax on the host") if world == 1 else
                           "pinned host tensors -> each rank uploads its own 1/N of the point rows -> point-partitioned "
                           "scoring step (PeerScoreBoard.step_points: all cameras over the local points, exact int64 partial "
                           "sums exchanged over NVLink inside the kernel, bitwise the scores of `value`) -> scores, argmax "
                           ".cpu(); h2d bytes are the total over all ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                         "kernel": "covgain_kernel<sigmoid,reduce>", "kernel_ms": kernel_ms,
                         "peer_wait_inside_kernel": dict(peer_wait, note="mean time per step the finishing CTA waited for the "
                                                         "slowest rank's scores (device %globaltimer); the rank that finishes "
                                                         "last waits least: kernel_ms - min = the slowest rank's own work"),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "kernel is bound by the fp32 FMA pipe, not HBM, at >= 8 cameras per point pass "
                                 "(DESIGN.md section 4); fp32_pipe is the binding roofline",
                         "fp32_pipe": {"fma_cycles_per_pair": FMA_CYCLES_PER_PAIR, "warp_cycles_per_s": issue_rate,
                                       "peak_warp_cycles_per_s": issue_peak, "frac": issue_rate / issue_peak,
                                       "sm_mhz": sm_mhz}},
            "parity": check,
        }
        if loop is not None:
            line["loop"] = loop
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        if world == 1 and not args.no_stages:
            try:
                line["path_stages"] = path_stages(dev)
            except Exception as exc:  # informational only: never lose the headline line
                line["path_stages"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
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
    ap.add_argument("--no-stages", action="store_true", help="skip the informational timings of the other path stages")
    ap.add_argument("--no-staging-probe", action="store_true", help="use the first two pinned buffers without a bandwidth probe")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-loop", action="store_true", help="skip the 50-step online NBV loop figure (cfg5)")
    ap.add_argument("--nccl-gather", action="store_true", help="use the NCCL all_gather instead of the fused peer push")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
