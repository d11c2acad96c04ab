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
