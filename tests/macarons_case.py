"""Shared set-up of the MACARONS candidate-scoring tests: the synthetic scene of a golden case, its cameras and the
`params` / `camera` / scene stand-ins that `predict_coverage_gain_for_single_camera` reads."""
import types

import torch

import synth
from oracle import cameras as o_cams

H, W = 256, 456


def uniforms(c, seq_len):
    """The uniforms the golden run fed to the reference's torch.rand (tests/golden/make_golden.py: seed 9000 + c)."""
    state = torch.get_rng_state()
    torch.manual_seed(9000 + c)
    u = torch.rand(seq_len, 1)
    torch.set_rng_state(state)
    return u


def build(N, C, seed, th, seq_len, device="cpu"):
    s = synth.macarons_scene(N, C, seed)
    s["T"][-1] = s["T"][-1] + 500.0     # the last candidate sees nothing (empty field of view branch)
    nb = synth.ndc_bounds(H, W)
    params = types.SimpleNamespace(sensor_range=70., min_occ_for_proxy_points=0.1, seq_len=seq_len,
                                   use_occ_to_sample_proxy_points=True, jz=False, ddp=False, distance_factor_th=th,
                                   image_height=H, image_width=W, k_for_knn=16, n_harmonics=64)
    cams = [o_cams.FoVPerspectiveCameras(R=s["R"][c:c + 1], T=s["T"][c:c + 1], zfar=1000., device=device) for c in range(C)]
    pred = o_cams.FoVPerspectiveCameras(R=s["pred_R"], T=s["pred_T"], zfar=1000., device=device)
    camera = types.SimpleNamespace(min_ndc_x=nb[0], max_ndc_x=nb[1], min_ndc_y=nb[2], max_ndc_y=nb[3], device=device,
                                   fov_camera_0=pred)
    proxy_scene = types.SimpleNamespace(x_min=s["x_min"].to(device), x_max=s["x_max"].to(device))
    surface_scene = types.SimpleNamespace(cell_resolution=0.5)
    u = torch.stack([uniforms(c, seq_len).view(-1) for c in range(C)])
    return s, params, cams, pred, camera, proxy_scene, surface_scene, nb, u


def threshold_from_golden(g):
    th = float(g["distance_factor_th"])
    return None if th < 0 else ("smooth" if th == 0 else th)
