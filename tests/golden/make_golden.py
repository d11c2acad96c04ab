#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (Anttwo/MACARONS at /root/reference).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
For every case the reference's own function is executed on seeded inputs (tests/synth.py), the
result is stored next to the inputs' seed + a checksum of the inputs, and the oracle (oracle/*.py)
is asserted to reproduce the reference BIT FOR BIT on this machine.  The reference has no tests or
golden vectors of its own, so these files are the pin.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402

ref_shim.install()

import synth  # noqa: E402
from macarons.networks.SconeVis import SconeVis  # noqa: E402  (the reference's)
from macarons.networks.Macarons import Macarons  # noqa: E402
from macarons.utility import scone_utils as ref_su  # noqa: E402
from macarons.networks.SconeOcc import SconeOcc  # noqa: E402
from macarons.utility import utils as ref_utils  # noqa: E402
from macarons.networks import ManyDepth as ref_md  # noqa: E402
from oracle import cameras as o_cams  # noqa: E402
from oracle import depth as o_depth  # noqa: E402
from oracle import depth_io as o_dio  # noqa: E402
from oracle import macarons_cov as o_mcov  # noqa: E402
from oracle import sampling as o_sampling  # noqa: E402
from oracle import scone_nets as o_nets  # noqa: E402
from oracle import sh_cov as o_cov  # noqa: E402
from oracle import view_state as o_vs  # noqa: E402


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()[:16]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def must_equal(a, b, what):
    if not torch.equal(a, b):
        raise SystemExit("oracle != reference for %s (max diff %g)" % (what, (a - b).abs().max().item()))


# (name, B, P, C, pts_dim, seed, coef_scale, use_sigmoid)
COVGAIN_CASES = [
    ("covgain_ragged_sigmoid", 2, 300, 7, 4, 101, 0.5, True),
    ("covgain_cfg2_sigmoid", 1, 2048, 64, 4, 102, 0.5, True),
    ("covgain_ragged_relu", 1, 257, 33, 3, 103, 0.5, False),
    ("covgain_bigcoef_sigmoid", 3, 96, 130, 4, 104, 3.0, True),
    ("covgain_single_cam", 1, 2048, 1, 4, 105, 1.0, True),
]


def covgain_goldens():
    torch.manual_seed(5)
    vis = SconeVis()
    for name, B, P, C, D, seed, scale, sig in COVGAIN_CASES:
        pts, harm, cams = synth.covgain_inputs(B, P, C, seed, pts_dim=D, coef_scale=scale)
        vis.use_sigmoid = sig
        cov = vis.compute_coverage_gain(pts, harm, cams)
        per_point = vis.compute_visibilities(pts, harm, cams)
        must_equal(cov, o_cov.coverage_gain(pts, harm, cams, use_sigmoid=sig), name + " coverage")
        must_equal(per_point, o_cov.visibility_gains(pts, harm, cams, use_sigmoid=sig, cam_chunk=16), name + " vis")
        if sig:  # Macarons.compute_visibility_gains is the same arithmetic behind another class
            mac = Macarons(None, None, vis)
            must_equal(mac.compute_visibility_gains(pts, harm, cams), per_point, name + " macarons")
        save(name, B=B, P=P, C=C, pts_dim=D, seed=seed, coef_scale=scale, use_sigmoid=int(sig),
             input_digest=digest(pts, harm, cams), coverage=cov, visibility=per_point.to(torch.float32),
             argmax=np.argmax(cov.numpy(), axis=-1))
    vis.use_sigmoid = True
    # n-tuple coverage (tiny: C^n tuples)
    pts, harm, cams = synth.covgain_inputs(1, 128, 5, 106)
    for n_cam in (2, 3):
        val, tuples = vis.compute_coverage_gain_multiple(pts, harm, cams, n_cam)
        v2, t2 = o_cov.coverage_gain_multiple(pts, harm, cams, n_cam)
        must_equal(val, v2, "coverage_gain_multiple")
        must_equal(tuples, t2, "coverage_gain_multiple idx")
        save("covgain_multiple_n%d" % n_cam, seed=106, B=1, P=128, C=5, n_cam=n_cam,
             input_digest=digest(pts, harm, cams), coverage=val, tuples=tuples)


def view_state_goldens():
    base, h_polar, h_azim = ref_su.get_all_harmonics_under_degree(8, 7, 14, "cpu")
    b2, hp2, ha2 = o_vs.bin_centre_harmonics(8, 7, 14)
    must_equal(base, b2, "base harmonics")
    must_equal(h_polar, hp2, "h_polar")
    must_equal(h_azim, ha2, "h_azim")
    save("view_base_harmonics", base=base, h_polar=h_polar, h_azim=h_azim)
    for name, B, P, V, seed in (("view_state_small", 2, 500, 3, 201), ("view_state_10views", 1, 4096, 10, 202)):
        pts, X_view = synth.view_state_inputs(B, P, V, seed)
        if name == "view_state_small":
            X_view[0] = torch.tensor([0.0, 1.5, 0.0])   # straight above: exercises the bin wrap-around
            X_view[1] = torch.tensor([0.0, -1.5, 0.0])  # straight below
        state = ref_su.compute_view_state(pts, X_view, 7, 14)
        must_equal(state, o_vs.view_state(pts, X_view, 7, 14), name)
        vh = ref_su.compute_view_harmonics(state, base, h_polar, h_azim, 7, 14)
        must_equal(vh, o_vs.view_harmonics(state, base, h_polar, 7, 14, point_chunk=777), name + " harmonics")
        save(name, B=B, P=P, V=V, seed=seed, X_view=X_view, input_digest=digest(pts, X_view),
             state_bits=np.packbits(state.numpy().astype(np.uint8), axis=-1), view_harmonics=vh)


def move_view_state_goldens():
    """Row a12: the reference's move_view_state_to_view_space (scone_utils.py:863-930) run on the stand-in camera
    (oracle/cameras.py; pytorch3d itself is not installed) for 24 seeded camera poses, incl. the identity and a pure
    y-rotation.  Stored: the per-camera gather indices (recovered by feeding an arange 'state') and one rotated state."""
    gen = torch.Generator().manual_seed(211)
    aa = 1.5 * torch.randn(24, 3, generator=gen)
    aa[0] = 0.0
    aa[1] = torch.tensor([0.0, 2 * np.pi * 3 / 14, 0.0])
    T = torch.randn(24, 3, generator=gen)
    probe = torch.arange(98, dtype=torch.float32).view(1, 1, 98)
    indices = []
    for i in range(24):
        cam = o_cams.FoVPerspectiveCameras(R=o_cams.axis_angle_to_matrix(aa[i:i + 1]), T=T[i:i + 1], zfar=100.0)
        got = ref_su.move_view_state_to_view_space(probe, cam, 7, 14)
        must_equal(got, o_vs.move_view_state_to_view_space(probe, cam, 7, 14), "move_view_state indices %d" % i)
        indices.append(got.view(98).long())
    pts, X_view = synth.view_state_inputs(2, 400, 6, 212)
    state = ref_su.compute_view_state(pts, X_view, 7, 14)
    cam = o_cams.FoVPerspectiveCameras(R=o_cams.axis_angle_to_matrix(aa[5:6]), T=T[5:6], zfar=100.0)
    moved = ref_su.move_view_state_to_view_space(state, cam, 7, 14)
    must_equal(moved, o_vs.move_view_state_to_view_space(state, cam, 7, 14), "move_view_state state")
    save("move_view_state", seed=211, state_seed=212, axis_angle=aa, T=T, indices=torch.stack(indices),
         state_bits=np.packbits(state.numpy().astype(np.uint8), axis=-1),
         moved_bits=np.packbits(moved.numpy().astype(np.uint8), axis=-1), moved_camera=5)


def sampling_goldens():
    gen = torch.Generator().manual_seed(301)
    N = 20000
    X = torch.rand(N, 3, generator=gen) - 0.5
    preds = torch.rand(N, 1, generator=gen)
    vh = torch.randn(N, 64, generator=gen)
    u = torch.rand(2048, 1, generator=gen)
    # drive the reference with the same uniforms through the global generator it reads
    state = torch.get_rng_state()
    torch.manual_seed(4242)
    u_global = torch.rand(2048, 1)
    torch.manual_seed(4242)
    res, res_h, inv = ref_su.sample_proxy_points(X, preds, vh, 2048, 0.1, return_index=True)
    torch.set_rng_state(state)
    r2, h2, i2 = o_sampling.sample_proxy_points(X, preds, vh, 2048, 0.1, u=u_global)
    must_equal(res, r2, "sampling points")
    must_equal(res_h, h2, "sampling harmonics")
    must_equal(inv, i2, "sampling inverse")
    del u
    save("sampling_20k", seed=301, N=N, n_sample=2048, min_occ=0.1, u=u_global, input_digest=digest(X, preds, vh),
         res=res, inverse=inv, res_h_checksum=res_h.double().sum(dim=0))


# (name, B, S, seed, row stride of the stored output)
SCONEVIS_CASES = [("sconevis_small", 2, 300, 401, 1), ("sconevis_2048", 1, 2048, 402, 8)]
# (name, B, N, Q, seed, grid)
SCONEOCC_CASES = [("sconeocc_small", 2, 700, 150, 501, False), ("sconeocc_cfg1", 1, 2048, 4096, 502, True)]
NET_WEIGHT_SEED = 5


def nets_goldens():
    """SconeVis.forward, SconeOcc.forward, get_knn_points and compute_occupancy_probability of the reference,
    with weights from synth.seeded_state_dict loaded through load_state_dict (state_dict compatibility)."""
    vis = SconeVis()
    vis_sd = synth.seeded_state_dict(vis.state_dict(), NET_WEIGHT_SEED)
    vis.load_state_dict(vis_sd)
    vis.eval()
    with torch.no_grad():
        for name, B, S, seed, stride in SCONEVIS_CASES:
            pts, vh = synth.sconevis_inputs(B, S, seed)
            out = vis(pts, view_harmonics=vh)
            must_equal(out, o_nets.scone_vis_forward(vis_sd, pts, vh), name)
            save(name, B=B, S=S, seed=seed, weight_seed=NET_WEIGHT_SEED, weights_digest=synth.state_dict_digest(vis_sd),
                 input_digest=digest(pts, vh), row_stride=stride, harmonics=out[:, ::stride])
    occ = SconeOcc()
    occ_sd = synth.seeded_state_dict(occ.state_dict(), NET_WEIGHT_SEED)
    occ.load_state_dict(occ_sd)
    occ.eval()
    with torch.no_grad():
        for name, B, N, Q, seed, grid in SCONEOCC_CASES:
            pc, x, vh = synth.sconeocc_inputs(B, N, Q, seed, grid=grid)
            torch.manual_seed(seed)
            out = occ(pc, x, vh)
            torch.manual_seed(seed)
            must_equal(out, o_nets.scone_occ_forward(occ_sd, pc, x, vh), name)
            pts_k, d_k, i_k = ref_utils.get_knn_points(x, pc, 16)
            p2, d2, i2 = o_nets.knn_points(x, pc, 16)
            must_equal(pts_k, p2, name + " knn points")
            must_equal(i_k, i2, name + " knn idx")
            save(name, B=B, N=N, Q=Q, seed=seed, grid=int(grid), weight_seed=NET_WEIGHT_SEED,
                 weights_digest=synth.state_dict_digest(occ_sd), input_digest=digest(pc, x, vh), occupancy=out,
                 knn_idx=i_k.to(torch.int32) if Q <= 512 else i_k[:, ::16].to(torch.int32), knn_dist=d_k if Q <= 512 else d_k[:, ::16])
        # chunked inference (utility/scone_utils.py:965-998): 3 forward calls with fresh sub-samples each
        pc, x, vh = synth.sconeocc_inputs(1, 600, 250, 503)
        torch.manual_seed(503)
        out = ref_su.compute_occupancy_probability(occ, pc, x, vh, max_points_per_pass=100)
        torch.manual_seed(503)
        must_equal(out, o_nets.compute_occupancy_probability(occ_sd, pc, x, vh, max_points_per_pass=100), "chunked occ")
        save("sconeocc_chunked", B=1, N=600, Q=250, seed=503, max_points_per_pass=100, weight_seed=NET_WEIGHT_SEED,
             weights_digest=synth.state_dict_digest(occ_sd), input_digest=digest(pc, x, vh), occupancy=out)



# (name, N proxy points, C candidate cameras, seed, distance_factor_th, seq_len)
MACARONS_COV_CASES = [("macarons_cov_th17", 20000, 6, 701, 17.0, 2048), ("macarons_cov_smooth", 6000, 4, 702, "smooth", 512),
                      ("macarons_cov_pixel", 6000, 3, 703, None, 512)]


def macarons_cov_goldens():
    """`predict_coverage_gain_for_single_camera` of the reference (utility/macarons_utils.py:1580-1738) with its own
    `Camera.get_points_in_fov`, `Macarons` wrapper and `SconeVis`, per candidate camera of a synthetic scene; the last
    candidate of every case is moved far outside the scene so that the empty-field-of-view branch runs too."""
    import types
    from macarons.utility import macarons_utils as ref_mu
    vis = SconeVis()
    vis_sd = synth.seeded_state_dict(vis.state_dict(), NET_WEIGHT_SEED)
    vis.load_state_dict(vis_sd)
    vis.eval()
    macarons = Macarons(None, None, vis)
    H, W = 256, 456
    nb = synth.ndc_bounds(H, W)
    for name, N, C, seed, th, seq_len in MACARONS_COV_CASES:
        s = synth.macarons_scene(N, C, seed)
        s["T"][-1] = s["T"][-1] + 500.0     # candidate C-1 sees nothing
        params = types.SimpleNamespace(sensor_range=70., min_occ_for_proxy_points=0.1, seq_len=seq_len,
                                       use_occ_to_sample_proxy_points=True, jz=False, ddp=False, distance_factor_th=th,
                                       image_height=H, image_width=W, k_for_knn=16, n_harmonics=64)
        fake_camera = types.SimpleNamespace(min_ndc_x=nb[0], max_ndc_x=nb[1], min_ndc_y=nb[2], max_ndc_y=nb[3], device="cpu")
        fake_camera.get_points_in_fov = types.MethodType(ref_mu.Camera.get_points_in_fov, fake_camera)
        proxy_scene = types.SimpleNamespace(x_min=s["x_min"], x_max=s["x_max"])
        surface_scene = types.SimpleNamespace(cell_resolution=0.5)
        pred_cam = o_cams.FoVPerspectiveCameras(R=s["pred_R"], T=s["pred_T"], zfar=1000.)
        diag = torch.linalg.norm(s["x_max"] - s["x_min"]).item()
        cov_all, vis_sum, nuniq = [], [], []
        with torch.no_grad():
            for c in range(C):
                cam = o_cams.FoVPerspectiveCameras(R=s["R"][c:c + 1], T=s["T"][c:c + 1], zfar=1000.)
                X_cam = cam.get_camera_center()
                state = torch.get_rng_state()
                torch.manual_seed(9000 + c)
                u = torch.rand(seq_len, 1)
                torch.manual_seed(9000 + c)
                ref = ref_mu.predict_coverage_gain_for_single_camera(
                    params, macarons, proxy_scene, surface_scene, s["X_world"].clone(), s["vh"].clone(), s["occ"].clone(),
                    fake_camera, X_cam, cam, prediction_camera=pred_cam)
                torch.set_rng_state(state)
                got = o_mcov.predict_coverage_gain_for_single_camera(
                    vis_sd, s["X_world"], s["vh"], s["occ"], X_cam, cam, pred_cam, nb, diag, sensor_range=70., min_occ=0.1,
                    seq_len=seq_len, distance_factor_th=th, image_height=H, image_width=W, cell_resolution=0.5, u=u)
                for a, b, what in zip(ref, got, ("proxy points", "view harmonics", "visibility gains", "coverage gain")):
                    must_equal(a, b, "%s camera %d %s" % (name, c, what))
                cov_all.append(ref[3].view(-1))
                vis_sum.append(ref[2].double().sum().view(1))
                nuniq.append(ref[0].shape[1])
        save(name, N=N, C=C, seed=seed, seq_len=seq_len, distance_factor_th=-1.0 if th is None else (0.0 if th == "smooth" else th),
             weight_seed=NET_WEIGHT_SEED, weights_digest=synth.state_dict_digest(vis_sd),
             input_digest=digest(s["X_world"], s["vh"], s["occ"], s["R"], s["T"]), coverage=torch.cat(cov_all),
             visibility_sum=torch.cat(vis_sum), n_returned=np.asarray(nuniq))



SCENE_FIELD_CASES = [("scene_field_s31", 31, 6000, 9000), ("scene_field_s32", 32, 3000, 5000)]


def scene_field_goldens():
    """the reference's `Scene` / `Cell` bookkeeping (utility/macarons_utils.py:2503-2932) and
    its `compute_scene_occupancy_probability_field` (:1395-1540) -- one SconeOcc call per occupied cell -- on the synthetic
    three-frame scene of tests/scene_case.py; the oracle (oracle/scene.py) must reproduce it bit for bit."""
    import copy
    from macarons.utility import macarons_utils as ref_mu
    import scene_case
    from oracle import scene as o_scene
    occ = SconeOcc()
    occ_sd = synth.seeded_state_dict(occ.state_dict(), NET_WEIGHT_SEED)
    occ.load_state_dict(occ_sd)
    occ.eval()
    macarons = Macarons(None, occ, None)
    params = scene_case.params()
    pred = scene_case.prediction_camera()
    for name, seed, n_proxy, n_surface in SCENE_FIELD_CASES:
        surface_scene, proxy_scene = scene_case.build(ref_mu.Scene, "cpu", seed, n_proxy=n_proxy, n_surface=n_surface)
        state_digest, counts = scene_case.scene_digest(surface_scene, proxy_scene)
        o_surface, o_proxy = copy.deepcopy(surface_scene), copy.deepcopy(proxy_scene)
        with torch.no_grad():
            torch.manual_seed(seed + 1000)
            X_world, vh, probs = ref_mu.compute_scene_occupancy_probability_field(params, macarons, None, surface_scene,
                                                                                  proxy_scene, "cpu", prediction_camera=pred)
            torch.manual_seed(seed + 1000)
            oX, ovh, oprobs = o_scene.scene_occupancy_field(params, occ_sd, o_surface, o_proxy, pred)
        must_equal(X_world, oX, name + " X_world")
        must_equal(vh, ovh, name + " view harmonics")
        must_equal(probs, oprobs, name + " occupancy")
        must_equal(proxy_scene.proxy_proba, o_proxy.proxy_proba, name + " proxy_proba")
        n_oof = int((proxy_scene.out_of_field > 0.).sum())
        save(name, seed=seed, n_proxy=n_proxy, n_surface=n_surface, weight_seed=NET_WEIGHT_SEED,
             weights_digest=synth.state_dict_digest(occ_sd), scene_digest=state_digest, cell_counts=np.asarray(counts),
             n_points=X_world.shape[0], n_out_of_field=n_oof, X_world_digest=digest(X_world),
             X_world_head=X_world[:64], view_harmonics_stride8=vh[:X_world.shape[0] - n_oof:8],
             occupancy=probs[:X_world.shape[0] - n_oof, 0], proxy_proba=proxy_scene.proxy_proba[:, 0])


def depth_io_goldens():
    """`Camera.project_depth_in_3D`, `compute_partial_point_cloud` and `get_signed_distance_to_depth_maps` of the reference
    (utility/macarons_utils.py:2339-2500), called unbound on a stand-in for `self` (the Camera constructor needs a renderer)."""
    import types
    from macarons.utility import macarons_utils as ref_mu
    H, W, seed = 48, 80, 801
    s = synth.depth_io_inputs(H, W, seed)
    cam = o_cams.FoVPerspectiveCameras(R=s["R"], T=s["T"], zfar=100.)
    cam1 = o_cams.FoVPerspectiveCameras(R=s["R"][:1], T=s["T"][:1], zfar=100.)
    nx, ny = o_dio.ndc_tables(H, W)
    fake = types.SimpleNamespace(image_height=H, image_width=W, ndc_x_tab=nx, ndc_y_tab=ny, fov_camera=cam1, zfar=100.,
                                 gathering_factor=0.05)
    fake.project_depth_in_3D = types.MethodType(ref_mu.Camera.project_depth_in_3D, fake)
    fake.get_points_zbuf = types.MethodType(ref_mu.Camera.get_points_zbuf, fake)
    world = ref_mu.Camera.project_depth_in_3D(fake, s["depth"], fov_cameras=cam)
    must_equal(world, o_dio.project_depth_in_3D(s["depth"], cam, H, W), "depth unprojection")
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    pc, col = ref_mu.Camera.compute_partial_point_cloud(fake, s["depth"][:1], s["mask"][:1], images=s["images"][:1],
                                                        fov_cameras=cam1, fov_range=9.0)
    torch.manual_seed(seed)
    n_sel = int((s["mask"][:1].view(1, -1) * (s["depth"][:1] < 9.0).view(1, -1)).sum())
    perm = torch.randperm(n_sel)
    torch.set_rng_state(state)
    pc2, col2 = o_dio.compute_partial_point_cloud(s["depth"][:1], s["mask"][:1], cam1, H, W, 0.05, images=s["images"][:1],
                                                  fov_range=9.0, perm=perm)
    must_equal(pc, pc2, "partial point cloud")
    must_equal(col, col2, "partial point cloud colours")
    sd = ref_mu.Camera.get_signed_distance_to_depth_maps(fake, s["pts"], s["depth"], s["mask"], fov_camera=cam)
    must_equal(sd, o_dio.signed_distance_to_depth_maps(s["pts"], s["depth"], s["mask"], cam, H, W, 100.), "signed distances")
    save("depth_io_48x80", H=H, W=W, seed=seed, input_digest=digest(s["depth"], s["pts"], s["R"], s["T"]),
         world_points=world[:, ::7], partial_pc=pc, perm=perm, signed_distance=sd)


# (name, B, H, W, seed, stored row stride)
DEPTH_CASES = [("depth_64x96", 1, 64, 96, 601, 1), ("depth_96x160_b2", 2, 96, 160, 602, 2)]


def reference_many_depth(H, W):
    """The reference's ManyDepth on a locally built resnet18 (torch.hub is unreachable; torchvision's resnet18 has the
    topology of pytorch/vision:v0.8.2) -- ManyDepth.py:764-790 without the download."""
    import torchvision
    resnet = torchvision.models.resnet18(weights=None)
    fe = ref_md.FeatureExtractor(resnet)
    dd = ref_md.DepthDecoder(fe, resnet, input_height=H, input_width=W, input_channels=3)
    return ref_md.ManyDepth(depth_decoder=dd, pose_decoder=None).eval()


def depth_goldens():
    import warnings
    warnings.filterwarnings("ignore", message="Default grid_sample")
    for name, B, H, W, seed, stride in DEPTH_CASES:
        model = reference_many_depth(H, W)
        sd = synth.seeded_state_dict(model.state_dict(), NET_WEIGHT_SEED)
        model.load_state_dict(sd)
        x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(B, H, W, seed)
        with torch.no_grad():
            pose, d1, d2, d3, d4 = model(x, x_alpha, R, T, zfar, "cpu", gt_pose=gt_pose)
            o = o_depth.many_depth_forward(sd, x, x_alpha, R, T, zfar, gt_pose)
        for a, b, what in zip((pose, d1, d2, d3, d4), o, ("pose", "disp1", "disp2", "disp3", "disp4")):
            must_equal(a, b, name + " " + what)
        save(name, B=B, H=H, W=W, seed=seed, weight_seed=NET_WEIGHT_SEED, weights_digest=synth.state_dict_digest(sd),
             input_digest=digest(x, x_alpha, gt_pose), row_stride=stride, disp1=d1[..., ::stride, ::stride], disp2=d2,
             disp3=d3, disp4=d4)


# full-resolution cases (BASELINE.json configs[2]: 256x256; reference default 256x456): the four disparity maps are
# stored sub-sampled (row / column strides 8, 4, 2, 1 -> 32 x W/8 values each) to keep the fixtures small
DEPTH_FULL_CASES = [("depth_256x456", 1, 256, 456, 603), ("depth_256x256", 1, 256, 256, 604)]
DEPTH_FULL_STRIDES = (8, 4, 2, 1)


def depth_full_goldens():
    import warnings
    warnings.filterwarnings("ignore", message="Default grid_sample")
    for name, B, H, W, seed in DEPTH_FULL_CASES:
        model = reference_many_depth(H, W)
        sd = synth.seeded_state_dict(model.state_dict(), NET_WEIGHT_SEED)
        model.load_state_dict(sd)
        x, x_alpha, R, T, zfar, gt_pose = synth.depth_inputs(B, H, W, seed)
        with torch.no_grad():
            pose, d1, d2, d3, d4 = model(x, x_alpha, R, T, zfar, "cpu", gt_pose=gt_pose)
            o = o_depth.many_depth_forward(sd, x, x_alpha, R, T, zfar, gt_pose)
        for a, b, what in zip((pose, d1, d2, d3, d4), o, ("pose", "disp1", "disp2", "disp3", "disp4")):
            must_equal(a, b, name + " " + what)
        st = DEPTH_FULL_STRIDES
        save(name, B=B, H=H, W=W, seed=seed, weight_seed=NET_WEIGHT_SEED, weights_digest=synth.state_dict_digest(sd),
             input_digest=digest(x, x_alpha, gt_pose), strides=np.asarray(st),
             disp1=d1[..., ::st[0], ::st[0]], disp2=d2[..., ::st[1], ::st[1]], disp3=d3[..., ::st[2], ::st[2]],
             disp4=d4[..., ::st[3], ::st[3]])


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--depth-only" in sys.argv:
        depth_goldens()
        raise SystemExit(0)
    if "--scene-only" in sys.argv:
        scene_field_goldens()
        raise SystemExit(0)
    if "--depth-full-only" in sys.argv:
        depth_full_goldens()
        raise SystemExit(0)
    if "--depth-io-only" in sys.argv:
        depth_io_goldens()
        raise SystemExit(0)
    if "--macarons-only" in sys.argv:
        macarons_cov_goldens()
        raise SystemExit(0)
    if "--move-view-state-only" in sys.argv:
        move_view_state_goldens()
        raise SystemExit(0)
    if "--nets-only" in sys.argv:
        nets_goldens()
        raise SystemExit(0)
    covgain_goldens()
    view_state_goldens()
    move_view_state_goldens()
    sampling_goldens()
    nets_goldens()
    macarons_cov_goldens()
    scene_field_goldens()
    depth_io_goldens()
    depth_goldens()
    depth_full_goldens()
    print("all oracle == reference checks passed (bitwise)")
