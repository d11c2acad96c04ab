"""Real spherical harmonics helper API of reference macarons/utility/spherical_harmonics.py.

On the NBV path the harmonics of camera rays are evaluated inside the CUDA coverage-gain kernel, so
the reference's global (l, m) memo table no longer exists; `clear_spherical_harmonics_cache` is kept
as a callable no-op because reference callers invoke it (networks/SconeVis.py:218,
utility/scone_utils.py:731).  `get_spherical_harmonics` is still needed by host-side set-up code
(`get_all_harmonics_under_degree`: 98 bin centres, once per run) and is provided as a small torch
function following spherical_harmonics.py:67-156 (Condon-Shortley phase, k = l*l+l+m ordering).
"""
import math

import torch


def clear_spherical_harmonics_cache():
    """No cache to clear (kept for API compatibility)."""
    return None


def _odd_double_factorial(n):
    out = 1.0
    for v in range(n, 1, -2):
        out *= v
    return out


def get_spherical_harmonics(l, theta, phi):
    """Tesseral harmonics of degree l at polar angle theta and azimuth phi -> (*theta.shape, 2l+1),
    last axis m = -l..l (sin for m<0, cos for m>0)."""
    x = torch.cos(theta)
    sin_t = torch.sqrt(torch.clamp(1 - x * x, min=0))
    cols = [None] * (2 * l + 1)
    for m in range(l + 1):
        # P_m^m, then climb to P_l^m
        p_prev = torch.zeros_like(x)
        p = ((-1) ** m * _odd_double_factorial(2 * m - 1)) * sin_t ** m if m else torch.ones_like(x)
        for ll in range(m + 1, l + 1):
            p, p_prev = ((2 * ll - 1) * x * p - (ll + m - 1) * p_prev) / (ll - m), p
        n = math.sqrt((2 * l + 1) / (4 * math.pi))
        if m == 0:
            cols[l] = n * p
        else:
            n *= math.sqrt(2.0 * math.factorial(l - m) / math.factorial(l + m))
            cols[l + m] = n * p * torch.cos(m * phi)
            cols[l - m] = n * p * torch.sin(m * phi)
    return torch.stack(cols, dim=-1)
