"""Oracle for one SCONE NBV scoring step (reference macarons/testers/shapenet.py:126-172) chained from the stage
oracles.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import torch

from . import sampling, scone_nets, sh_cov, view_state


def scone_nbv_step(occ_sd, vis_sd, pc, X, X_view, X_cam, n_elev=7, n_azim=14, seq_len=2048, min_occ=0.1,
                   max_points_per_pass=300000, samples=None, occ_override=None):
    """-> (coverage (C,1), argmax, stages).  `occ_override` injects an occupancy field at the stage boundary."""
    base, h_polar, _ = view_state.bin_centre_harmonics(8, n_elev, n_azim)
    state = view_state.view_state(X, X_view, n_elev, n_azim)
    vh = view_state.view_harmonics(state, base, h_polar, n_elev, n_azim)
    if occ_override is None:
        occ = scone_nets.compute_occupancy_probability(occ_sd, pc, X, vh, max_points_per_pass=max_points_per_pass).view(-1, 1)
    else:
        occ = occ_override.view(-1, 1)
    proxy, proxy_vh, sample_idx = sampling.sample_proxy_points(X[0], occ, vh.squeeze(0), seq_len, min_occ, u=samples)
    harmonics = scone_nets.scone_vis_forward(vis_sd, proxy.unsqueeze(0), proxy_vh.unsqueeze(0))
    proxy_mc = proxy[sample_idx].unsqueeze(0)
    harmonics_mc = harmonics[0][sample_idx].unsqueeze(0)
    cov = sh_cov.coverage_gain(proxy_mc, harmonics_mc, X_cam.view(1, -1, 3)).view(-1, 1)
    return cov, torch.max(cov, dim=0)[1], {"view_harmonics": vh, "occ": occ, "proxy": proxy_mc, "harmonics": harmonics_mc,
                                           "sample_idx": sample_idx}
