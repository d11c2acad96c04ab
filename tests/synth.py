"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.
Everything is drawn on the CPU from an explicit torch.Generator so that the oracle and the GPU path
see identical bits."""
import math

import torch


def sphere_cameras(n, radius, gen):
    """n camera centres uniform on the sphere of given radius, kept away from exact poles."""
    v = torch.randn(n, 3, generator=gen)
    v = v / v.norm(dim=-1, keepdim=True)
    return radius * v


def covgain_inputs(B, P, C, seed, pts_dim=4, coef_scale=0.5, radius=1.5):
    """pts (B,P,pts_dim) with xyz ~ U[-0.5,0.5]^3 and occupancy ~ U[0.1,1]; harmonics ~ N(0, coef_scale^2);
    cameras on the r=1.5 sphere (one set per cloud)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    pts = torch.rand(B, P, pts_dim, generator=gen) - 0.5
    if pts_dim > 3:
        pts[..., 3] = 0.1 + 0.9 * torch.rand(B, P, generator=gen)
    harm = coef_scale * torch.randn(B, P, 64, generator=gen)
    cams = torch.stack([sphere_cameras(C, radius, gen) for _ in range(B)])
    return pts.contiguous(), harm.contiguous(), cams.contiguous()


def fibonacci_cameras(n, radius=1.5):
    """Deterministic quasi-uniform camera layout (bench workload)."""
    i = torch.arange(n, dtype=torch.float64) + 0.5
    y = 1 - 2 * i / n
    r = torch.sqrt(torch.clamp(1 - y * y, min=0))
    phi = i * math.pi * (3 - math.sqrt(5))
    return (radius * torch.stack((r * torch.sin(phi), y, r * torch.cos(phi)), dim=-1)).float()


def view_state_inputs(B, P, V, seed, radius=1.5):
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    pts = torch.rand(B, P, 4, generator=gen) - 0.5
    X_view = sphere_cameras(V, radius, gen)
    return pts.contiguous(), X_view.contiguous()
