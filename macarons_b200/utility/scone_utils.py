"""Hot-path subset of reference macarons/utility/scone_utils.py, same function names and argument meaning.

  compute_occupancy_probability   (reference :965-998)   chunked SconeOcc inference
"""
import torch


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
