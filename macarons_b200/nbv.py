"""One SCONE next-best-view scoring step, the loop body of reference macarons/testers/shapenet.py:126-172 written
over the mirrored API (every call below is the reference call of the same name; all arithmetic runs in the CUDA
kernels behind them):

    view state -> view harmonics -> occupancy field -> proxy sampling -> visibility harmonics -> coverage gain -> argmax
"""
import torch

from .utility import scone_utils


def scone_nbv_step(scone_occ, scone_vis, pc, X, X_view, X_cam, base_harmonics, h_polar, h_azim, n_elev=7, n_azim=14,
                   seq_len=2048, min_occ=0.1, max_points_per_pass=300000, true_monte_carlo_sampling=True, samples=None,
                   cam_range=None, return_stages=False):
    """pc (1,N,3) partial cloud, X (1,Q,3) proxy points, X_view (V,3) visited cameras, X_cam (C,3) candidates, all in the
    normalised prediction space -> (coverage (C,1), nbv index).  `samples` injects the uniforms of the proxy sampling
    (tests); `cam_range` scores a slice of the candidates (multi-GPU partition, see macarons_b200.parallel)."""
    with torch.no_grad():
        view_state = scone_utils.compute_view_state(X, X_view, n_elev, n_azim)                      # shapenet.py:126
        view_harmonics = scone_utils.compute_view_harmonics(view_state, base_harmonics, h_polar, h_azim, n_elev, n_azim)
        occ = scone_utils.compute_occupancy_probability(scone_occ=scone_occ, pc=pc, X=X, view_harmonics=view_harmonics,
                                                        max_points_per_pass=max_points_per_pass).view(-1, 1)   # :139
        proxy, proxy_vh, sample_idx = scone_utils.sample_proxy_points(X[0], occ, view_harmonics.squeeze(dim=0),
                                                                      n_sample=seq_len, min_occ=min_occ,
                                                                      return_index=True, samples=samples)       # :146
        proxy, proxy_vh = proxy.unsqueeze(0), proxy_vh.unsqueeze(0)
        harmonics = scone_vis(proxy, view_harmonics=proxy_vh)                                        # :157
        if true_monte_carlo_sampling:                                                                # :158-160
            proxy = proxy[0][sample_idx].unsqueeze(0)
            harmonics = harmonics[0][sample_idx].unsqueeze(0)
        if cam_range is None:
            cov = scone_vis.compute_coverage_gain(proxy, harmonics, X_cam.view(1, -1, 3)).view(-1, 1)  # :167
        else:
            cov = scone_vis.compute_coverage_gain(proxy, harmonics, X_cam.view(1, -1, 3), cam_range=cam_range).view(-1, 1)
        max_gain, max_idx = torch.max(cov, dim=0)                                                     # :172
    if return_stages:
        return cov, max_idx, {"view_state": view_state, "view_harmonics": view_harmonics, "occ": occ, "proxy": proxy,
                              "sample_idx": sample_idx, "harmonics": harmonics}
    return cov, max_idx


def sconevis_forward_clouds(scone_vis, pts, view_harmonics, seq_len, group=None):
    """SconeVis.forward over a large point set cut into clouds of `seq_len` tokens (the reference's network input
    size, networks/SconeVis.py:7): pts (1, P, 4) and view_harmonics (1, P, 64) with P % seq_len == 0 -> (1, P, 64).
    Every cloud is an independent forward, so with `group` (torch.distributed, one process per GPU) the clouds are
    partitioned across the ranks and one all-gather over NVLink (NCCL) completes the replicated result; the output is
    bitwise the same as the single-GPU batch (rows never mix across clouds)."""
    import torch.distributed as dist
    P = pts.shape[1]
    if pts.shape[0] != 1 or P % seq_len != 0:
        raise ValueError("pts must be (1, P, 4) with P a multiple of seq_len=%d (got %s)" % (seq_len, tuple(pts.shape)))
    n_clouds = P // seq_len
    clouds = pts.reshape(n_clouds, seq_len, pts.shape[-1])
    vh = view_harmonics.reshape(n_clouds, seq_len, view_harmonics.shape[-1])
    world = dist.get_world_size(group) if (group is not None or (dist.is_available() and dist.is_initialized())) else 1
    if world == 1:
        return scone_vis(clouds, view_harmonics=vh).reshape(1, P, -1)
    rank = dist.get_rank(group)
    per = -(-n_clouds // world)
    out = torch.empty((world * per, seq_len, 64), dtype=torch.float32, device=pts.device)
    lo, hi = min(rank * per, n_clouds), min((rank + 1) * per, n_clouds)
    mine = out[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[:hi - lo] = scone_vis(clouds[lo:hi], view_harmonics=vh[lo:hi])
    dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=group)
    return out[:n_clouds].reshape(1, P, 64)


def scone_online_loop(scone_vis, pts, X_cam, X_view, base_harmonics, h_polar, h_azim, n_steps, score_step=None,
                      n_elev=7, n_azim=14, seq_len=2048, shard_clouds=False):
    """The large-scene online NBV loop of BASELINE.json configs[4], the loop body of reference
    testers/shapenet.py:89-172 from the view state onwards on a fixed proxy-point set: for `n_steps` steps

        view state + view harmonics of all points for the cameras visited so far   (scone_utils.py:799-860, :934-960)
        -> SconeVis.forward on P / seq_len clouds                                     (networks/SconeVis.py:121-162)
        -> coverage gain of every candidate camera over all P points, argmax        (SconeVis.py:210-252, shapenet.py:172)
        -> the chosen camera joins the visited set                                  (shapenet.py:176-180)

    pts (1, P, 4) [normalised xyz, occupancy], X_cam (1, C, 3), X_view (V0, 3).  `score_step(pts, harmonics, X_cam) ->
    (scores (1, C), best (1,))` defaults to the single-GPU kernel; pass `PeerScoreBoard.step` for the camera-sharded
    form.  Nothing synchronises with the host inside the loop (the chosen camera is appended on the device).
    -> (chosen (n_steps,) int64, last scores (1, C))."""
    from . import ops
    if score_step is None:
        def score_step(p, h, c):
            s = ops.coverage_gain(p, h, c, use_sigmoid=scone_vis.use_sigmoid)
            return s, torch.argmax(s, dim=-1)
    with torch.no_grad():
        V0 = X_view.shape[0]
        views = torch.empty((V0 + n_steps, 3), dtype=torch.float32, device=pts.device)
        views[:V0] = X_view
        chosen = torch.empty((n_steps,), dtype=torch.int64, device=pts.device)
        scores = None
        for s in range(n_steps):
            vh = scone_utils.compute_view_state_harmonics(pts, views[:V0 + s], base_harmonics, h_polar, h_azim, n_elev, n_azim)
            if shard_clouds:
                harmonics = sconevis_forward_clouds(scone_vis, pts, vh, seq_len)
            else:
                n_clouds = pts.shape[1] // seq_len
                harmonics = scone_vis(pts.reshape(n_clouds, seq_len, -1),
                                      view_harmonics=vh.reshape(n_clouds, seq_len, 64)).reshape(1, -1, 64)
            scores, best = score_step(pts, harmonics, X_cam)
            views[V0 + s] = X_cam[0].index_select(0, best.reshape(-1)[:1])[0]
            chosen[s] = best.reshape(-1)[0]
    return chosen, scores


class _StageClock:
    """CUDA events around the stages of a step (optional instrumentation: `timings` dict name -> milliseconds)."""

    def __init__(self, timings, device):
        self.t, self.dev, self.marks = timings, device, []

    def mark(self, name):
        if self.t is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.dev))
        self.marks.append((name, ev))

    def finish(self):
        if self.t is None:
            return
        torch.cuda.synchronize(self.dev)
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            self.t[name] = self.t.get(name, 0.0) + a.elapsed_time(b)


def macarons_nbv_step(params, macarons, camera, surface_scene, proxy_scene, frames, X_cams_world, fov_cameras, samples=None,
                      depth_mask=None, timings=None):
    """One full MACARONS next-best-view step (BASELINE.json configs[2]; the loop body of reference testers/scene.py:352-456
    and trainers/train_macarons.py:228-315, without the simulator / renderer around it):

        depth forward (ManyDepth)                     -> depth map of the current frame         (macarons_utils.py:888-960)
        un-projection + 5 % sub-sample                -> partial surface cloud -> surface_scene (:2362-2398, testers/scene.py:373-384)
        field of view + signed distances of the proxy points -> view states, pseudo-GT occupancy, out-of-field flags (:390-414)
        occupancy probability field of the scene      -> X_world, view_harmonics, occ_probs     (:1395-1540, ONE ragged forward)
        coverage gain of every candidate camera pose  -> argmax                                 (:1580-1738, testers/scene.py:434-456)

    `frames` = dict(x (1,3,H,W), x_alpha (1,n_alpha,3,H,W), R, T, zfar, gt_pose) as fed to `macarons(mode='depth', ...)`;
    `camera`: the reference's Camera or any object with image_height / image_width / zfar / gathering_factor / fov_camera /
    fov_camera_0 / X_cam / min_ndc_* / max_ndc_*; `fov_cameras`: ONE batched camera holding the C candidate poses (or a list
    of C cameras), `X_cams_world` (C,3) their centres; `timings` (optional dict) receives the device time of every stage.
    -> (coverage_gain (C,1), nbv index, dict of the stage outputs)."""
    from .utility import macarons_utils as mu
    dev = proxy_scene.proxy_points.device
    H, W = camera.image_height, camera.image_width
    clock = _StageClock(timings, dev)
    with torch.no_grad():
        clock.mark("start")
        pose, disp1 = macarons(mode='depth', x=frames["x"], x_alpha=frames["x_alpha"], R=frames["R"], T=frames["T"],
                               zfar=frames["zfar"], device=dev, gt_pose=frames["gt_pose"])[:2]
        depth = mu.compute_depth_from_disparity(params, disp1).permute(0, 2, 3, 1).contiguous()        # (1,H,W,1)
        mask = torch.ones(1, H, W, 1, dtype=torch.bool, device=dev) if depth_mask is None else depth_mask
        part_pc = mu.compute_partial_point_cloud(camera, depth=depth, mask=mask, fov_cameras=camera.fov_camera,
                                                 gathering_factor=params.gathering_factor, fov_range=params.sensor_range)
        clock.mark("depth_forward_and_partial_cloud_ms")
        surface_scene.fill_cells(part_pc, features=torch.zeros(len(part_pc), 1, device=dev))
        fov_proxy_points, fov_proxy_mask = mu.get_points_in_fov(camera, proxy_scene.proxy_points, return_mask=True,
                                                                fov_camera=None, fov_range=params.sensor_range)
        proxy_scene.fill_cells(fov_proxy_points, features=proxy_scene.get_proxy_indices_from_mask(fov_proxy_mask).view(-1, 1))
        sgn_dists = mu.get_signed_distance_to_depth_maps(camera, pts=fov_proxy_points, depth_maps=depth, mask=mask,
                                                         fov_camera=None)
        proxy_scene.update_proxy_view_states(camera, fov_proxy_mask, signed_distances=sgn_dists, distance_to_surface=None,
                                             X_cam=None)
        proxy_scene.update_proxy_supervision_occ(fov_proxy_mask, sgn_dists, tol=params.carving_tolerance)
        proxy_scene.update_proxy_out_of_field(fov_proxy_mask)
        surface_scene.set_all_features_to_value(value=1.)
        clock.mark("scene_update_ms")
        X_world, view_harmonics, occ_probs = mu.compute_scene_occupancy_probability_field(params, macarons, camera,
                                                                                          surface_scene, proxy_scene, dev)
        clock.mark("occupancy_field_ms")
        out = mu.predict_coverage_gains_for_cameras(params, macarons, proxy_scene, surface_scene, X_world, view_harmonics,
                                                    occ_probs, camera, X_cams_world, fov_cameras, samples=samples)
        coverage = out["coverage_gain"]
        best = torch.argmax(coverage.view(-1))       # strict '>' running maximum of testers/scene.py:454 = first maximum
        clock.mark("candidate_scoring_ms")
    clock.finish()
    stages = {"disp1": disp1, "depth": depth, "part_pc": part_pc, "fov_proxy_mask": fov_proxy_mask, "sgn_dists": sgn_dists,
              "X_world": X_world, "view_harmonics": view_harmonics, "occ_probs": occ_probs, "candidates": out}
    return coverage, best, stages
