"""The functions of reference macarons/utility/macarons_utils.py that sit on the NBV scoring path, over the CUDA
kernels of this package (same names, argument meaning and return values):

    compute_occupancy_probability               reference :1194-1230
    compute_scene_occupancy_probability_field   reference :1395-1540 (all occupied cells in ONE ragged SconeOcc forward)
    predict_coverage_gain_for_single_camera     reference :1580-1738
    get_distance_factor / _threshold / _smooth  reference :1741-1788
    load_images_for_depth_model / save_frame    reference :763-803, :2317-2335 (frames stay resident on the device)

plus `predict_coverage_gains_for_cameras`, the batched form the reference does not have: the testers / trainers call
`predict_coverage_gain_for_single_camera` once per neighbouring pose (testers/scene.py:434-456,
trainers/train_macarons.py:291-315); here all C candidates go through ONE field-of-view / occupancy selection +
sampling launch (csrc/sampling.cu), ONE ragged SconeVis forward (C clouds of <= seq_len unique points each) and ONE
per-point visibility launch.  The single-camera function is the C = 1 case of it.

Everything outside those functions (Camera / Scene / Memory classes, renderers, data loading, optimisers) is the
reference's control plane and is not mirrored; the functions below only need duck-typed
arguments: `camera` with `min_ndc_x/max_ndc_x/min_ndc_y/max_ndc_y` and `fov_camera_0`, cameras with pytorch3d's
`get_full_projection_transform() / get_world_to_view_transform() / get_camera_center()` and `.fov`.
"""
import numpy as np
import torch

from .. import netpack, ops
from .scone_utils import normalize_points_in_prediction_box


def compute_occupancy_probability(macarons, pc, X, view_harmonics, mask=None, max_points_per_pass=20000):
    """pc (n_clouds, seq_len, 3), X (n_clouds, n_sample, 3), view_harmonics (n_clouds, n_sample, 64)
    -> (n_clouds, n_sample, 1) through `macarons(mode='occupancy', ...)`, `max_points_per_pass // n_clouds` queries per
    forward as in the reference (:1211-1228; the O(passes^2) torch.cat growth is replaced by one output buffer)."""
    n_clouds, n_sample = pc.shape[0], X.shape[1]
    p = max_points_per_pass // n_clouds
    preds = torch.empty(n_clouds, n_sample, 1, dtype=torch.float32, device=X.device)
    for lo in range(0, n_sample, p):
        up = min(lo + p, n_sample)
        preds[:, lo:up] = macarons(mode='occupancy', partial_point_cloud=pc, proxy_points=X[:, lo:up],
                                   view_harmonics=view_harmonics[:, lo:up]).view(n_clouds, up - lo, -1)
    return preds


# ---- distance factors (reference :1741-1788); pts (..., n_points, 3), X_cam broadcastable (..., 1, 3) ----
def _pixel_distance_threshold(params, fov_camera, cell_resolution):
    fov = fov_camera.fov if isinstance(fov_camera.fov, torch.Tensor) else torch.tensor(float(fov_camera.fov))
    focal_length = 1. / torch.tan(np.pi / 180. * fov / 2.)
    pixel_size = 2. / min(params.image_height, params.image_width)
    epsilon = np.sqrt(np.pi) / 2. * cell_resolution
    return focal_length, pixel_size, epsilon, focal_length * epsilon / pixel_size


def get_distance_factor(params, pts, X_cam, fov_camera, cell_resolution):
    """1 up to the distance at which one surface cell covers one pixel, (threshold / d)^2 beyond.  pts (n_points, 3),
    X_cam (1, 3) -> (n_points, 1)."""
    focal_length, pixel_size, epsilon, distance_th = _pixel_distance_threshold(params, fov_camera, cell_resolution)
    focal_length, distance_th = focal_length.to(pts.device), distance_th.to(pts.device)
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    far = epsilon ** 2 * (focal_length / pixel_size / dists) ** 2
    return torch.where(dists > distance_th, far, torch.ones_like(dists))


def get_distance_factor_threshold(pts, X_cam, distance_th=17.):
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    return torch.where(dists > distance_th, distance_th ** 2 / dists ** 2, torch.ones_like(dists))


def get_distance_factor_smooth(params, pts, X_cam, fov_camera, cell_resolution):
    _, _, _, distance_th = _pixel_distance_threshold(params, fov_camera, cell_resolution)
    dists = torch.linalg.norm(pts - X_cam.view(1, 3), dim=-1, keepdim=True)
    return 1. / (1. + (dists / distance_th.to(pts.device)) ** 2)


def _as_list(fov_cameras):
    return list(fov_cameras) if isinstance(fov_cameras, (list, tuple)) else None


def _camera_fovs(fov_cameras, C, device):
    """(C,) field-of-view angles in degrees of a batched camera or a list of single cameras."""
    cams = _as_list(fov_cameras)
    fov = fov_cameras.fov if cams is None else torch.cat([torch.as_tensor(c.fov, dtype=torch.float32).reshape(-1)[:1] for c in cams])
    return torch.as_tensor(fov, dtype=torch.float32, device=device).reshape(-1).expand(C)


def _distance_factors(params, world_xyz, X_cam_world, fov_cameras, cell_resolution):
    """Batched: world_xyz (C, S, 3), X_cam_world (C, 3) -> (C, S), selecting the rule like the reference (:1685-1701)."""
    C = world_xyz.shape[0]
    dists = torch.linalg.norm(world_xyz - X_cam_world.view(C, 1, 3), dim=-1)
    if params.distance_factor_th is not None and params.distance_factor_th != 'smooth':
        th = params.distance_factor_th
        return torch.where(dists > th, th ** 2 / dists ** 2, torch.ones_like(dists))          # get_distance_factor_threshold
    focal_length = (1. / torch.tan(np.pi / 180. * _camera_fovs(fov_cameras, C, world_xyz.device) / 2.)).view(C, 1)
    pixel_size = 2. / min(params.image_height, params.image_width)
    epsilon = np.sqrt(np.pi) / 2. * cell_resolution
    distance_th = focal_length * epsilon / pixel_size
    if params.distance_factor_th == 'smooth':
        return 1. / (1. + (dists / distance_th) ** 2)                                          # get_distance_factor_smooth
    far = epsilon ** 2 * (focal_length / pixel_size / dists) ** 2                              # get_distance_factor
    return torch.where(dists > distance_th, far, torch.ones_like(dists))


def _camera_rows(fov_cameras, device):
    """(C, 36) rows [full projection 4x4 | world-to-view 4x4 | centre 3 | 0] of one batched pytorch3d-style camera
    (C entries: three batched calls) or of a list of C single cameras (the reference's call pattern, C x 3 calls)."""
    cams = _as_list(fov_cameras)
    if cams is None:
        proj = fov_cameras.get_full_projection_transform().get_matrix()
        view = fov_cameras.get_world_to_view_transform().get_matrix()
        centre = fov_cameras.get_camera_center()
    else:
        proj = torch.cat([c.get_full_projection_transform().get_matrix().reshape(1, 4, 4) for c in cams])
        view = torch.cat([c.get_world_to_view_transform().get_matrix().reshape(1, 4, 4) for c in cams])
        centre = torch.cat([c.get_camera_center().reshape(1, 3) for c in cams])
    C = proj.shape[0]
    rows = torch.cat((proj.reshape(C, 16), view.reshape(C, 16), centre.reshape(C, 3), centre.new_zeros(C, 1)), dim=1)
    return rows.to(device=device, dtype=torch.float32).contiguous()


def _scone_vis(macarons, params):
    model = macarons.module if (getattr(params, "jz", False) or getattr(params, "ddp", False)) else macarons
    return model.visibility


def predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, X_world, proxy_view_harmonics,
                                       occ_probs, camera, X_cams_world, fov_cameras, prediction_camera=None, samples=None):
    """Coverage gains of C candidate cameras in one pass (see the module docstring).

    X_world (N,3), proxy_view_harmonics (N,64), occ_probs (N,1): the scene's proxy points; X_cams_world (C,3) and
    `fov_cameras` (ONE batched camera of C entries -- preferred, 3 matrix calls in total -- or a sequence of C single
    cameras): the candidates; `samples` (C, seq_len) injects the uniforms of the proxy
    sampling (default: torch.rand on the device, like the reference per call).
    -> dict with coverage_gain (C,1), visibility_gains (C,1,seq_len), proxy_points_world (C,seq_len,4),
       view_harmonics (C,seq_len,64), n_points_in_fov (C,) int32, n_unique (C,) int32, fov_proxy_volume (C,)."""
    if not params.use_occ_to_sample_proxy_points:
        raise NotImplementedError("the batched path implements occupancy-weighted sampling (the reference default)")
    dev = X_world.device
    X_cams_world = X_cams_world.reshape(-1, 3).to(torch.float32)
    C = X_cams_world.shape[0]
    S = int(params.seq_len)
    vis_model = _scone_vis(macarons, params)
    if not vis_model.use_sigmoid:
        raise NameError("WARNING! ReLU has been used in visibility model.")     # Macarons.compute_visibility_gains :176
    if prediction_camera is None:
        if camera is None:
            raise NameError("Both camera and prediction_camera are equal to None.")
        prediction_camera = camera.fov_camera_0
    if samples is None:
        samples = torch.rand(C, S, device=dev)
    ndc = [float(v) for v in (camera.min_ndc_x, camera.max_ndc_x, camera.min_ndc_y, camera.max_ndc_y)]

    # field of view + occupancy threshold + inverse-CDF sampling + unique, all candidates at once  (:1603-1628)
    res, res_h, inverse, counts, volume = ops.fov_sample_proxy(
        X_world, occ_probs, proxy_view_harmonics, _camera_rows(fov_cameras, dev), ndc, params.sensor_range,
        params.min_occ_for_proxy_points, samples.reshape(C, S).to(torch.float32))
    n_unique = counts[:, 1].contiguous()
    valid = torch.arange(S, device=dev).view(1, S) < n_unique.view(C, 1)

    # prediction box: centre of the unique points, everything moved to the prediction camera's view space (:1632-1659)
    xyz = res[..., :3]
    big = torch.finfo(torch.float32).max
    hi = torch.where(valid[..., None], xyz, xyz.new_full((), -big)).amax(dim=1)
    lo = torch.where(valid[..., None], xyz, xyz.new_full((), big)).amin(dim=1)
    centre = torch.where((n_unique > 0).view(C, 1), (hi + lo) / 2., torch.zeros_like(hi))
    view_transform = prediction_camera.get_world_to_view_transform()
    box_centre = view_transform.transform_points(centre).view(C, 3)
    diag = torch.linalg.norm(proxy_scene.x_max - proxy_scene.x_min).item()
    pts = torch.cat((normalize_points_in_prediction_box(view_transform.transform_points(xyz.reshape(-1, 3)).view(C, S, 3),
                                                        box_centre.view(C, 1, 3), diag), res[..., 3:]), dim=-1)
    pts = torch.where(valid[..., None], pts, torch.zeros_like(pts))      # padding rows must be finite (zeros)
    X_cam = normalize_points_in_prediction_box(view_transform.transform_points(X_cams_world).view(C, 3), box_centre, diag)

    # visibility-gain harmonics of every candidate's point set: one ragged forward  (:1663)
    harm = ops.sconevis_forward(netpack.pack_sconevis(vis_model), pts, res_h, lens=n_unique)

    # Monte-Carlo set with duplicates (:1669-1672), per-point gains (:1676-1683), distance factor, integration (:1685-1703)
    gather = lambda t: torch.gather(t, 1, inverse[..., None].expand(-1, -1, t.shape[-1]))
    pts_s, harm_s, world_s, vh_s = gather(pts), gather(harm), gather(res), gather(res_h)
    gains = ops.visibility_gains(pts_s, harm_s, X_cam.view(C, 1, 3), use_sigmoid=True)
    gains = gains * _distance_factors(params, world_s[..., :3], X_cams_world, fov_cameras,
                                      getattr(surface_scene, "cell_resolution", None)).view(C, 1, S)
    empty = (counts[:, 0] == 0).view(C, 1)
    coverage = torch.where(empty, torch.zeros(C, 1, device=dev), torch.mean(gains, dim=-1) * volume.view(C, 1))
    return {"coverage_gain": coverage, "visibility_gains": gains, "proxy_points_world": world_s, "view_harmonics": vh_s,
            "n_points_in_fov": counts[:, 0], "n_unique": n_unique, "fov_proxy_volume": volume,
            "proxy_points": pts, "harmonics": harm, "sample_idx": inverse, "X_cam": X_cam}


def predict_coverage_gain_for_single_camera(params, macarons, proxy_scene, surface_scene, X_world, proxy_view_harmonics,
                                            occ_probs, camera, X_cam_world, fov_camera, prediction_camera=None):
    """Reference signature and return values (:1580-1738): (proxy_points_world (1,seq_len,4), view_harmonics
    (1,seq_len,64), visibility_gains (1,1,seq_len), coverage_gain (1,1)); when no proxy point with occupancy above
    the threshold is in the field of view, the reference's dummy 16-point pass with a zero gain."""
    out = predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, X_world, proxy_view_harmonics,
                                             occ_probs, camera, X_cam_world.view(1, 3), [fov_camera],
                                             prediction_camera=prediction_camera)
    if int(out["n_points_in_fov"][0].item()) > 0:
        return out["proxy_points_world"], out["view_harmonics"], out["visibility_gains"], out["coverage_gain"].view(-1, 1)
    dev = X_world.device
    model = macarons.module if (getattr(params, "jz", False) or getattr(params, "ddp", False)) else macarons
    dummy_pts = torch.zeros(1, params.k_for_knn, 4, device=dev)
    dummy_vh = torch.zeros(1, params.k_for_knn, params.n_harmonics, device=dev)
    dummy_harm = macarons(mode='visibility', proxy_points=dummy_pts, view_harmonics=dummy_vh)
    gains = model.compute_visibility_gains(pts=dummy_pts, harmonics=dummy_harm, X_cam=X_cam_world.view(1, -1, 3))
    return dummy_pts, dummy_vh, gains, (torch.mean(gains, dim=-1) * 0.).view(-1, 1)


# ---- occupancy probability field of the whole scene (reference :1395-1540) ----------------------------------------------
def _scone_occ(macarons, params):
    model = macarons.module if (getattr(params, "jz", False) or getattr(params, "ddp", False)) else macarons
    return model.occupancy


def compute_scene_occupancy_probability_field(params, macarons, camera, surface_scene, proxy_scene, device,
                                              use_supervision_occ_mask=True, prediction_camera=None,
                                              use_supervision_occ_instead_of_predicted=False):
    """Reference signature and return values (:1395-1540): (X_world (N,3), view_harmonics (N,64), occ_probs (N,1)) =
    every proxy point that has been in a field of view and is not carved away, cell by cell (cells in sorted index
    order, points in ascending proxy index inside a cell), followed by the out-of-field points with zero harmonics and
    their stored probability; `proxy_scene.proxy_proba` is updated like the reference does.

    The reference calls the occupancy network once per occupied cell (and per 20 000-query pass inside a cell), each
    call on the cell's own 27-cell neighbourhood cloud.  Here the host loop only collects index tensors; the bin
    permutation (`move_view_state_to_view_space`: one permutation for the whole scene), the view harmonics and the
    normalisation run once over all kept proxy points, and ALL cells go through ONE ragged SconeOcc forward
    (`SconeOcc.forward_cells`, csrc/scone_nets.cu) whose random sub-samples are drawn in the reference's call order."""
    from . import scone_utils
    occ_mask = (proxy_scene.proxy_supervision_occ > 0.)[..., 0]
    all_fov_mask = (proxy_scene.out_of_field < 1.)[..., 0]
    keep_mask = occ_mask * all_fov_mask if use_supervision_occ_mask else all_fov_mask
    fovs_proxy_points = proxy_scene.proxy_points[keep_mask]
    proxy_scene.proxy_proba[occ_mask * all_fov_mask] = 0.
    proxy_cells = proxy_scene.get_englobing_cells(fovs_proxy_points)
    base_harmonics, h_polar, h_azim = scone_utils.get_all_harmonics_under_degree(
        params.harmonic_degree, params.view_state_n_elev, params.view_state_n_azim, device)
    if prediction_camera is None:
        if camera is None:
            raise NameError("Both camera and prediction_camera are equal to None.")
        prediction_camera = camera.fov_camera_0
    view_transform = prediction_camera.get_world_to_view_transform()
    k_min = 2 * 2 * params.k_for_knn
    max_pass = 20000                       # max_points_per_pass of the reference's per-cell call (:1508)

    # ---- which proxy points and which surface points belong to every occupied cell: ONE host synchronisation ----
    # (the reference does this cell by cell with ~20 small tensor ops and three synchronisations each, :1443-1460)
    cell_list = proxy_cells.cpu().tolist()                       # sorted cell indices (torch.unique), the loop order
    grid = (surface_scene.grid_l, surface_scene.grid_w, surface_scene.grid_h)
    stored = [proxy_scene.cells[str(c)].cell_features.view(-1).long() for c in cell_list]   # proxy indices held by each cell
    n_stored = [int(t.numel()) for t in stored]
    cell_rows, cell_clouds, cell_centres, cell_diags = [], [], [], []
    if sum(n_stored) > 0:
        idx_all = torch.cat(stored)
        owner = torch.repeat_interleave(torch.arange(len(cell_list), device=idx_all.device),
                                        torch.tensor(n_stored, device=idx_all.device))
        if use_supervision_occ_mask:
            keep = occ_mask[idx_all]
            idx_all, owner = idx_all[keep], owner[keep]
        # ascending proxy index inside every cell (what indexing with a boolean mask gives); a cell holds an index once
        order = torch.argsort(owner * proxy_scene.n_proxy_points + idx_all)
        idx_all, owner = idx_all[order], owner[order]
        n_rows = torch.bincount(owner, minlength=len(cell_list)).tolist()                    # the synchronisation
        row_off = 0
        for c, n_r in zip(cell_list, n_rows):
            rows = idx_all[row_off:row_off + n_r]
            row_off += n_r
            # surface points of the (up to 27) neighbouring cells, clamped to the grid, in sorted unique order (:2706-2718)
            neigh = sorted({(min(max(c[0] + a, 0), grid[0] - 1), min(max(c[1] + b, 0), grid[1] - 1),
                             min(max(c[2] + d, 0), grid[2] - 1)) for a in (-1, 0, 1) for b in (-1, 0, 1) for d in (-1, 0, 1)})
            parts = [surface_scene.cells[str(list(k))].cell_pts for k in neigh]
            n_cloud = sum(int(t.shape[0]) for t in parts)
            if n_cloud > k_min and n_r > 0:
                cell = proxy_scene.cells[str(c)]
                cell_rows.append(rows)
                cell_clouds.append(torch.cat(parts))
                cell_centres.append(cell.center.view(1, 3))
                cell_diags.append(params.prediction_neighborhood_size * torch.linalg.norm(cell.x_max - cell.x_min))

    n_harm = params.n_harmonics
    if cell_rows:
        counts = [int(r.numel()) for r in cell_rows]
        rows_all = torch.cat(cell_rows)
        X_cells = proxy_scene.proxy_points[rows_all]
        # prediction view space, one centre / diagonal per cell (:1466-1484)
        centres = view_transform.transform_points(torch.cat(cell_centres)).view(-1, 3)
        diags = torch.stack([d.reshape(()) for d in cell_diags]).to(torch.float32)
        q_centre = torch.repeat_interleave(centres, torch.tensor(counts, device=centres.device), dim=0)
        q_diag = torch.repeat_interleave(diags, torch.tensor(counts, device=diags.device)).view(-1, 1)
        X_norm = normalize_points_in_prediction_box(view_transform.transform_points(X_cells).view(-1, 3), q_centre, q_diag)
        n_pc = [int(c.shape[0]) for c in cell_clouds]
        pc_all = view_transform.transform_points(torch.cat(cell_clouds)).view(-1, 3)
        pc_all = normalize_points_in_prediction_box(
            pc_all, torch.repeat_interleave(centres, torch.tensor(n_pc, device=centres.device), dim=0),
            torch.repeat_interleave(diags, torch.tensor(n_pc, device=diags.device)).view(-1, 1))
        # view states -> prediction view space -> harmonics, all cells at once (:1487-1498)
        states = scone_utils.move_view_state_to_view_space(
            proxy_scene.view_states[rows_all].view(1, -1, params.n_view_state_cameras), prediction_camera,
            n_elev=params.view_state_n_elev, n_azim=params.view_state_n_azim)
        harmonics = scone_utils.compute_view_harmonics(states, base_harmonics, h_polar, h_azim, params.view_state_n_elev,
                                                       params.view_state_n_azim).view(-1, n_harm)
        if use_supervision_occ_instead_of_predicted:
            probs = proxy_scene.proxy_supervision_occ[rows_all]
        else:
            # one entry per (cell, pass of <= 20 000 queries): every pass of the reference is a separate forward with its
            # own sub-samples of the cell's cloud
            clouds, queries, vhs = [], [], []
            q0 = p0 = 0
            for n_q, n_p in zip(counts, n_pc):
                for lo in range(0, n_q, max_pass):
                    up = min(lo + max_pass, n_q)
                    clouds.append(pc_all[p0:p0 + n_p])
                    queries.append(X_norm[q0 + lo:q0 + up])
                    vhs.append(harmonics[q0 + lo:q0 + up])
                q0, p0 = q0 + n_q, p0 + n_p
            # groups of cells: while the GPU runs the ragged forward of one group, the host draws the random sub-samples
            # (three torch.randperm of up to ~27 000 entries per cell, in the reference's order) of the next one
            occ_net, group, outs = _scone_occ(macarons, params), 12, []
            for g0 in range(0, len(clouds), group):
                outs.extend(occ_net.forward_cells(clouds[g0:g0 + group], queries[g0:g0 + group], vhs[g0:g0 + group]))
            probs = torch.cat(outs).view(-1, 1)
        proxy_scene.proxy_proba[rows_all] = probs
    else:
        X_cells = torch.zeros(0, 3, device=device)
        harmonics = torch.zeros(0, n_harm, device=device)
        probs = torch.zeros(0, 1, device=device)

    # ---- out-of-field points: zero harmonics, stored (default) probability (:1522-1538) ----
    oof_mask = (proxy_scene.out_of_field > 0.)[..., 0]
    oof_X = proxy_scene.proxy_points[oof_mask]
    X_world = torch.vstack((X_cells.view(-1, 3), oof_X))
    view_harmonics = torch.vstack((harmonics, torch.zeros(len(oof_X), n_harm, device=device)))
    occ_probs = torch.vstack((probs, proxy_scene.proxy_proba[oof_mask]))
    return X_world, view_harmonics, occ_probs


def compute_depth_from_disparity(params, disp):
    """depth = 1 / (a disp + b), a = 1/znear - 1/zfar, b = 1/zfar   [reference utility/depth_model_utils.py:844-848]"""
    a = 1. / params.znear - 1. / params.zfar
    b = 1. / params.zfar
    return 1. / (a * disp + b)


# ---- depth-side helpers (reference: methods of `Camera`, macarons_utils.py:2339-2500) -----------------------------------
# They take the reference's Camera object (or anything with image_height, image_width, zfar, gathering_factor, fov_camera)
# as first argument, so a maintainer binds them back as methods: `Camera.project_depth_in_3D = project_depth_in_3D`, ...
def _unproject_rows(fov_cameras, device):
    """(B, 18) rows [inverse full projection 4x4 | f1 | f2] for mac_unproject_depth_f32 (pytorch3d unproject_points)."""
    inv = fov_cameras.get_full_projection_transform().inverse().get_matrix()
    K = fov_cameras.get_projection_transform().get_matrix()
    B = inv.shape[0]
    rows = torch.cat((inv.reshape(B, 16), K[:, 2, 2].reshape(B, 1), K[:, 3, 2].reshape(B, 1)), dim=1)
    return rows.to(device=device, dtype=torch.float32).contiguous()


def project_depth_in_3D(camera, depth, fov_cameras=None):
    """depth (batch_size, height, width, 1) -> world points (batch_size, height*width, 3)   [reference :2339-2360]"""
    if fov_cameras is None:
        fov_cameras = camera.fov_camera
    B = depth.shape[0]
    return ops.unproject_depth(depth.reshape(B, -1), _unproject_rows(fov_cameras, depth.device), camera.image_height,
                               camera.image_width)


def compute_partial_point_cloud(camera, depth, mask, images=None, fov_cameras=None, gathering_factor=None, fov_range=None):
    """Partial point cloud seen by one camera: masked pixels un-projected, a random fraction kept  [reference :2362-2398].
    depth, mask (1, height, width, 1); the permutation is drawn like the reference (torch.randperm, CPU generator)."""
    points_mask = mask.view(1, -1) if fov_range is None else mask.view(1, -1) * (depth < fov_range).view(1, -1)
    world_points = project_depth_in_3D(camera, depth, fov_cameras=fov_cameras)[points_mask]
    if gathering_factor is None:
        gathering_factor = camera.gathering_factor
    n_points = int(len(world_points) * gathering_factor)
    points_indices = torch.randperm(len(world_points))[:n_points].to(world_points.device)
    world_points = world_points[points_indices]
    if images is None:
        return world_points
    return world_points, (0. + images.view(1, -1, 3))[points_mask][points_indices]


def get_points_in_fov(camera, pts, return_mask=False, fov_camera=None, fov_range=None):
    """Points of pts (n_point, 3) inside the field of view of `fov_camera` (default: the camera's current one) and closer
    than `fov_range` to its centre  [reference Camera.get_points_in_fov :2400-2435]; one kernel (csrc/sampling.cu)."""
    own = fov_camera is None
    row = _camera_rows(camera.fov_camera if own else fov_camera, pts.device)[0]
    if own:    # the reference measures the range from self.X_cam for the camera's own field of view (:2412-2414)
        row = torch.cat((row[:32], camera.X_cam.reshape(-1)[:3].to(row), row[35:]))
    fov_mask = ops.points_in_fov(pts.to(torch.float32), row, (camera.min_ndc_x, camera.max_ndc_x, camera.min_ndc_y,
                                                               camera.max_ndc_y), fov_range)
    return (pts[fov_mask], fov_mask) if return_mask else pts[fov_mask]


def get_signed_distance_to_depth_maps(camera, pts, depth_maps, mask, fov_camera=None):
    """pts (n_points, 3), depth_maps / mask (n_depth, height, width, 1) -> (n_depth, n_points, 1): positive behind the
    surface seen in the depth map, negative in front of it  [reference :2451-2500]."""
    n_depths = depth_maps.shape[0]
    if fov_camera is None:
        fov_camera = camera.fov_camera
        if n_depths > 1:
            raise NameError("Too many depth maps provided for current camera; depth_maps should have shape (1, ...)"
                            "If you want to simultaneously process depth maps for multiple cameras, "
                            "please provide the corresponding fov_camera argument.")
    elif n_depths != fov_camera.R.shape[0]:
        raise NameError("Number of cameras should be the same as number of depths.")
    H, W = camera.image_height, camera.image_width
    full = fov_camera.get_full_projection_transform().get_matrix()
    view = fov_camera.get_world_to_view_transform().get_matrix()
    rows = torch.cat((full.reshape(n_depths, 16), view.reshape(n_depths, 16)), dim=1).to(device=pts.device, dtype=torch.float32)
    out = ops.signed_distance(pts, depth_maps.reshape(n_depths, H, W), mask.reshape(n_depths, H, W), rows, H, W,
                              1.1 * float(camera.zfar))
    return out.view(n_depths, -1, 1)


# ---- frame files (reference :763-803 and the save half of Camera.capture_image :2317-2335) ---------------------------------
# The reference writes every captured frame to `<dir>/<n>.pt` (torch.save of a dict rgb / zbuf / mask / R / T / zfar) and
# re-reads the last n_frames + n_alpha of them from disk for EVERY depth prediction.  Here a frame that was saved (or
# loaded once) stays resident on its device in a small per-directory cache, the files keep the reference's format (either
# side can read the other's), and the loader fills preallocated batches instead of growing them with torch.cat.
import os as _os

_FRAME_CACHE = {}            # absolute frame path -> dict of tensors on the camera's device
_FRAME_CACHE_LIMIT = 64      # frames kept resident (least recently used are dropped; the files remain)
_FRAME_KEYS = ("rgb", "zbuf", "mask", "R", "T", "zfar")


def _cache_frame(path, frame):
    _FRAME_CACHE.pop(path, None)
    _FRAME_CACHE[path] = frame
    while len(_FRAME_CACHE) > _FRAME_CACHE_LIMIT:
        _FRAME_CACHE.pop(next(iter(_FRAME_CACHE)))


def clear_frame_cache():
    _FRAME_CACHE.clear()


def save_frame(camera, images, depth, fov_camera=None, dir_path=None):
    """Store one captured frame as `<dir>/<camera.n_frames_captured>.pt` in the reference's format and advance the
    counter (reference Camera.capture_image :2317-2335).  images (1,H,W,3), depth (1,H,W,1) z-buffer (-1 = background).
    -> the frame's path (None when no directory is configured, like the reference which then saves nothing)."""
    if fov_camera is None:
        fov_camera = camera.fov_camera
    if dir_path is None:
        dir_path = getattr(camera, "save_dir_path", None)
    if dir_path is None:
        return None
    frame = {"rgb": images, "zbuf": depth, "mask": depth > -1, "R": fov_camera.R, "T": fov_camera.T, "zfar": camera.zfar}
    path = _os.path.abspath(_os.path.join(dir_path, str(camera.n_frames_captured) + ".pt"))
    torch.save(frame, path)
    _cache_frame(path, frame)
    camera.n_frames_captured += 1
    return path


def load_images_for_depth_model(camera, n_frames, n_alpha, frame_nb=None, frames_dir_path=None, return_gt_zbuf=False):
    """The last n_frames + n_alpha frames up to `frame_nb` (default: the latest captured), oldest first
    -> (images (n,H,W,3), [zbuf (n,H,W,1),] mask (n,H,W,1) bool, R (n,3,3), T (n,3), zfar (n,))   [reference :763-803]."""
    current = camera.n_frames_captured - 1 if frame_nb is None else frame_nb
    if frames_dir_path is None:
        frames_dir_path = camera.save_dir_path
    n = n_frames + n_alpha
    dev, H, W = camera.device, camera.image_height, camera.image_width
    images = torch.empty(n, H, W, 3, device=dev)
    mask = torch.empty(n, H, W, 1, device=dev)
    R, T = torch.empty(n, 3, 3, device=dev), torch.empty(n, 3, device=dev)
    zbuf = torch.empty(n, H, W, 1, device=dev) if return_gt_zbuf else None
    for i in range(n):
        path = _os.path.abspath(_os.path.join(frames_dir_path, str(current - (n - 1) + i) + ".pt"))
        frame = _FRAME_CACHE.get(path)
        if frame is None or frame["rgb"].device != torch.device(dev):
            frame = torch.load(path, map_location=dev)
        _cache_frame(path, frame)
        images[i:i + 1], mask[i:i + 1], R[i:i + 1], T[i:i + 1] = frame["rgb"], frame["mask"], frame["R"], frame["T"]
        if return_gt_zbuf:
            zbuf[i:i + 1] = frame["zbuf"]
    zfar = torch.Tensor([camera.zfar]).to(dev).expand(n)
    if return_gt_zbuf:
        return images, zbuf, mask.bool(), R, T, zfar
    return images, mask.bool(), R, T, zfar
