"""Oracle for row a13 of SURVEY.md section 8: occupancy-weighted inverse-CDF sampling of proxy
points.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: /root/reference/macarons/utility/scone_utils.py:1030-1076 (sample_proxy_points).
The reference draws `torch.rand(n_sample, 1)` from the global generator of the tensors' device;
here the uniforms can be injected (`u`) so that CPU oracle and GPU path see the same draws.
"""
import torch


def sample_proxy_points(X_world, preds, view_harmonics, n_sample, min_occ, u=None, row_chunk=256):
    """(N,3), (N,1), (N,64) -> (res (U,4) [xyz, occ], res_harmonics (U,64), inverse_idx (n_sample,)).
    Keeps points with occupancy > min_occ, builds the CDF of occupancy / sum, and for each uniform
    takes the first index whose CDF value is >= u: the reference forms cdf - u, overwrites negative
    entries with 2 and takes argmin (scone_utils.py:1054-1056); if every entry is negative the
    argmin of the all-2 row is index 0.  Sorted unique indices + inverse map follow."""
    mask = preds[..., 0] > min_occ
    res_X, res_preds, res_h = X_world[mask], preds[mask], view_harmonics[mask]
    n_points = res_X.shape[0]
    probs = res_preds[..., 0] / torch.sum(res_preds)
    cdf = torch.cumsum(probs, dim=-1)
    if u is None:
        u = torch.rand(n_sample, 1)
    u = u.view(n_sample, 1)
    picks = []
    for r0 in range(0, n_sample, row_chunk):  # the reference materialises (n_sample, N) at once
        diff = cdf.view(1, n_points).expand(min(row_chunk, n_sample - r0), -1) - u[r0:r0 + row_chunk].expand(-1, n_points)
        diff = torch.where(diff < 0, torch.full_like(diff, 2.0), diff)
        picks.append(torch.argmin(diff, dim=-1))
    idx = torch.cat(picks)
    uniq, inverse = torch.unique(idx, dim=0, return_inverse=True)
    res = torch.cat((res_X[uniq], res_preds[uniq]), dim=-1)
    return res, res_h[uniq], inverse
