"""Coordinate helpers of reference macarons/utility/CustomGeometry.py (host-side set-up code: a few hundred
camera positions per run; the per-ray conversions of the hot path live inside the CUDA kernels)."""
import numpy as np
import torch


def get_cartesian_coords(r, elev, azim, in_degrees=False):
    """r, elev, azim of shape (N, 1) -> (N, 3) points, y up, azimuth measured from +z towards +x  [reference :5-24]"""
    factor = np.pi / 180. if in_degrees else 1
    e, a = factor * elev, factor * azim
    unit = torch.stack((torch.cos(e) * torch.sin(a), torch.sin(e), torch.cos(e) * torch.cos(a)), dim=2)
    return r * unit.view(-1, 3)


def get_spherical_coords(X):
    """(M,3) -> r, elev, azim (each (M,)); same clamps as the reference [:27-45]"""
    r = torch.linalg.norm(X, dim=1)
    sin_elev = X[:, 1] / r
    elev = torch.asin(sin_elev)
    elev = torch.where(sin_elev <= -1, torch.full_like(elev, -np.pi / 2), elev)
    elev = torch.where(sin_elev >= 1, torch.full_like(elev, np.pi / 2), elev)
    cos_azim = X[:, 2] / (r * torch.cos(elev))
    azim = torch.acos(cos_azim)
    azim = torch.where(cos_azim <= -1, torch.full_like(azim, np.pi), azim)
    azim = torch.where(cos_azim >= 1, torch.zeros_like(azim), azim)
    azim = torch.where(X[:, 0] < 0, -azim, azim)
    return r, elev, azim
