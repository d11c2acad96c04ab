"""Hot-path subset of reference macarons/utility/scone_utils.py: same function names, argument meaning and return
values; the per-point arithmetic runs in CUDA kernels through the C ABI (CUDA tensors only, no CPU path).

  get_all_harmonics_under_degree  (reference :714-738)   host-side set-up, 98 bin centres once per run
  get_cameras_on_sphere           (reference :741-785)   host-side set-up
  normalize_points_in_prediction_box (:788-796)
  compute_view_state              (reference :799-860)   -> csrc/viewstate.cu
  move_view_state_to_view_space   (reference :863-930)   98-entry permutation on the host + gather kernel
  compute_view_harmonics          (reference :934-960)   -> csrc/viewstate.cu
  compute_occupancy_probability   (reference :965-998)   chunked SconeOcc inference
  sample_proxy_points             (reference :1030-1076) -> csrc/sampling.cu
"""
import numpy as np
import torch

from .. import ops
from .CustomGeometry import get_cartesian_coords, get_spherical_coords
from .spherical_harmonics import clear_spherical_harmonics_cache, get_spherical_harmonics


def get_all_harmonics_under_degree(degree, n_elev, n_azim, device):
    """-> (base (degree^2, n_elev*n_azim), h_polar, h_azim): the SH basis at the bin centres, bins elevation-major,
    elev_i = -pi/2 + (i+1) pi / (n_elev+1), azim_j = 2 pi j / n_azim."""
    # 98 bin centres x 64 basis functions, once per run: evaluated on the host (the reference's torch CPU arithmetic,
    # ~1000 tiny element-wise ops) and copied to `device`, instead of ~1000 device launches of a few threads each
    h_elev = torch.Tensor([-np.pi / 2 + (i + 1) / (n_elev + 1) * np.pi for i in range(n_elev) for _ in range(n_azim)])
    h_polar = -h_elev + np.pi / 2
    h_azim = torch.Tensor([2 * np.pi * j / n_azim for _ in range(n_elev) for j in range(n_azim)])
    clear_spherical_harmonics_cache()
    z = torch.cat([get_spherical_harmonics(l, h_polar, h_azim) for l in range(degree)], dim=-1)
    clear_spherical_harmonics_cache()
    return z.transpose(dim0=0, dim1=1).contiguous().to(device), h_polar.to(device), h_azim.to(device)


def get_cameras_on_sphere(params, device, pole_cameras=False, n_elev=None, n_azim=None, camera_dist=None):
    """Candidate camera positions on a sphere -> (X_cam (n,3), dist (n,), elev (n,), azim (n,)) in degrees."""
    if n_elev is None or n_azim is None:
        n_elev, n_azim, n_camera = params.n_camera_elev, params.n_camera_azim, params.n_camera
    else:
        n_camera = n_elev * n_azim + (2 if pole_cameras else 0)
    if camera_dist is None:
        camera_dist = params.camera_dist
    dist = torch.Tensor([camera_dist for _ in range(n_camera)]).to(device)
    elev = [-90. + (i + 1) / (n_elev + 1) * 180. for i in range(n_elev) for _ in range(n_azim)]
    azim = [360. * j / n_azim for _ in range(n_elev) for j in range(n_azim)]
    if pole_cameras:
        elev, azim = [-89.9] + elev + [89.9], [0.] + azim + [0.]
    elev, azim = torch.Tensor(elev).to(device), torch.Tensor(azim).to(device)
    X_cam = get_cartesian_coords(r=dist.view(-1, 1), elev=elev.view(-1, 1), azim=azim.view(-1, 1), in_degrees=True)
    return X_cam, dist, elev, azim


def normalize_points_in_prediction_box(points, prediction_box_center, prediction_box_diag):
    return (points - prediction_box_center) / prediction_box_diag


def compute_view_state(pts, X_view, n_elev, n_azim):
    """pts (n_cloud, seq_len, >=3), X_view (n_view, 3) -> (n_cloud, seq_len, n_elev*n_azim): 1.0 in the bin of every
    visited camera as seen from every point."""
    return ops.view_state(pts, X_view, n_elev, n_azim)


def view_space_bin_permutation(R, n_elev, n_azim, T=None):
    """The n_elev*n_azim gather indices of move_view_state_to_view_space for a camera with rotation R (3,3) and
    translation T (3,) (pytorch3d row-vector convention X_view = X_world R + T).  Host-side set-up on 98 points that
    follows the reference statement by statement (:876-926) so that the indices are the reference's, bin for bin, also
    for directions that land on a bin boundary: unit bin directions -> inverse of the 4x4 world-to-view matrix
    (`torch.inverse`, fp32, as pytorch3d's Transform3d.inverse) -> minus the camera centre -> spherical coordinates ->
    nearest bin (NB the clamps here are +-(n_elev // 2), unlike compute_view_state)."""
    R = R.detach().to(device="cpu", dtype=torch.float32).view(3, 3)
    T = torch.zeros(3) if T is None else T.detach().to(device="cpu", dtype=torch.float32).view(3)
    n_view = n_elev * n_azim
    elev = torch.Tensor([-90. + (i + 1) / (n_elev + 1) * 180. for i in range(n_elev) for _ in range(n_azim)])
    azim = torch.Tensor([360. * j / n_azim for _ in range(n_elev) for j in range(n_azim)])
    X_ref = get_cartesian_coords(r=torch.ones(n_view, 1), elev=elev.view(-1, 1), azim=azim.view(-1, 1), in_degrees=True)
    M = torch.zeros(4, 4)
    M[:3, :3], M[3, :3], M[3, 3] = R, T, 1.0
    M_inv = torch.inverse(M.view(1, 4, 4))[0]

    def transform(points):   # Transform3d.transform_points: homogeneous row vectors, divide by w
        out = torch.cat((points, torch.ones(points.shape[0], 1)), dim=-1) @ M_inv
        return out[:, :3] / out[:, 3:]

    X_inv = transform(X_ref) - transform(torch.zeros(1, 3))
    elev_step, azim_step = np.pi / (n_elev + 1), 2 * np.pi / n_azim
    _, ray_elev, ray_azim = get_spherical_coords(X_inv.view(-1, 3))
    idx_elev = (ray_elev - ray_elev % elev_step) / elev_step
    idx_azim = (ray_azim - ray_azim % azim_step) / azim_step
    idx_elev[ray_elev % elev_step > elev_step / 2.] += 1
    idx_azim[ray_azim % azim_step > azim_step / 2.] += 1
    idx_elev[idx_elev > n_elev // 2] = n_elev // 2
    idx_elev[idx_elev < -(n_elev // 2)] = -(n_elev // 2)
    idx_azim[idx_azim > n_azim // 2] = -(n_azim // 2)
    idx_elev += n_elev // 2
    idx_azim[idx_azim < 0] += n_azim
    return idx_elev.long() * n_azim + idx_azim.long()


def move_view_state_to_view_space(view_state, fov_camera, n_elev, n_azim):
    """'Rotate' view states (n_cloud, seq_len, n_elev*n_azim) into the view space of `fov_camera` (any object with
    pytorch3d-style `.R` (1,3,3) and `.T` (1,3)): 98 gather indices on the host, one gather kernel."""
    T = getattr(fov_camera, "T", None)
    indices = view_space_bin_permutation(fov_camera.R[0], n_elev, n_azim, T=None if T is None else T[0])
    return ops.gather_bins(view_state, indices)


def compute_view_harmonics(view_state, base_harmonics, h_polar, h_azim, n_elev, n_azim):
    """(n_cloud, seq_len, n_elev*n_azim) -> (n_cloud, seq_len, n_harmonics): spherical L2 product of the histogram
    with every basis function (sum over bins of state * base * sin(polar) * polar_step * azim_step)."""
    return ops.view_harmonics(view_state, base_harmonics, h_polar, n_elev, n_azim)


def compute_view_state_harmonics(pts, X_view, base_harmonics, h_polar, h_azim, n_elev, n_azim):
    """compute_view_harmonics(compute_view_state(pts, X_view, ...), ...) in one kernel (extension; the pair of reference
    calls at testers/shapenet.py:126-131): same result bit for bit, the (n_cloud, seq_len, n_bins) histogram is never
    written to memory."""
    return ops.view_state_harmonics(pts, X_view, base_harmonics, h_polar, n_elev, n_azim)


def compute_occupancy_probability(scone_occ, pc, X, view_harmonics, mask=None, max_points_per_pass=20000):
    """pc (n_clouds, seq_len, 3), X (n_clouds, n_sample, 3), view_harmonics (n_clouds, n_sample, 64)
    -> (n_clouds, n_sample, 1).  Like the reference, the queries are cut into passes of
    `max_points_per_pass // n_clouds` and every pass is a separate `scone_occ` forward, i.e. it re-draws the random
    sub-samples and re-runs the global transformer (reference :982-996); only the O(passes^2) `torch.cat`
    growth is replaced by one preallocated output."""
    n_clouds, n_sample = pc.shape[0], X.shape[1]
    p = max_points_per_pass // n_clouds
    preds = torch.empty(n_clouds, n_sample, 1, dtype=torch.float32, device=X.device)
    for lo in range(0, n_sample, p):
        up = min(lo + p, n_sample)
        preds[:, lo:up] = scone_occ(pc, X[:, lo:up], view_harmonics[:, lo:up], verbose=False).view(n_clouds, up - lo, -1)
    return preds


def sample_proxy_points(X_world, preds, view_harmonics, n_sample, min_occ, use_occ_to_sample=True, return_index=False,
                        samples=None):
    """X_world (n_points, 3), preds (n_points, 1), view_harmonics (n_points, 64) -> (res (U,4), res_harmonics (U,64)
    [, inverse_idx (n_sample,)]): occupancy-weighted inverse-CDF draw of n_sample points among those with occupancy
    > min_occ, duplicates merged (sorted unique indices).  The uniforms are `torch.rand(n_sample, 1, device=...)`
    exactly as in the reference (:1052) unless `samples` injects them (tests)."""
    if not use_occ_to_sample:
        mask = preds[..., 0] > min_occ                    # reference :1063-1070, plain truncation, no arithmetic
        res = torch.cat((X_world[mask][:n_sample], preds[mask][:n_sample]), dim=-1)
        res_harmonics, inverse_idx = view_harmonics[mask][:n_sample], None
    else:
        if samples is None:
            samples = torch.rand(n_sample, 1, device=X_world.device)
        res, res_harmonics, inverse_idx = ops.sample_proxy_points(X_world, preds, view_harmonics,
                                                                  samples.reshape(-1).to(torch.float32), min_occ)
    if return_index:
        return res, res_harmonics, inverse_idx
    return res, res_harmonics
