"""Oracle for rows a1-a5 and a16 of SURVEY.md section 8: ray -> spherical coordinates, the real
spherical-harmonic basis, per-point visibility gains, per-camera coverage gain and the NBV argmax.

TEST INFRASTRUCTURE (see oracle/__init__.py).  fp32 functions use torch CPU kernels in exactly
the operation order of the reference so that, on the same machine, they agree with it bit for
bit; `*_f64` functions are an independent trig-free closed form in numpy float64.

Reference (paths relative to /root/reference/macarons):
  utility/CustomGeometry.py:27-45       get_spherical_coords
  utility/spherical_harmonics.py:49-156 semifactorial, pochhammer, lpmv, get_spherical_harmonics(_element)
  networks/SconeVis.py:164-208          SconeVis.compute_visibilities
  networks/SconeVis.py:210-252          SconeVis.compute_coverage_gain
  networks/SconeVis.py:254-303          SconeVis.compute_coverage_gain_multiple
  networks/Macarons.py:138-178          Macarons.compute_visibility_gains
  testers/shapenet.py:172, testers/scene.py:454   argmax rule (first maximum wins)
"""
import math

import numpy as np
import torch

N_DEGREE = 8          # l = 0..7  (SconeVis.max_harmonic_rank, networks/SconeVis.py:11)
N_HARMONICS = 64      # sum_{l<8} (2l+1); hard-coded `64` at networks/SconeVis.py:241


# ----------------------------------------------------------------------------------------------
# a1  utility/CustomGeometry.py:27-45
# ----------------------------------------------------------------------------------------------
def spherical_coords(X):
    """(M,3) -> r, elev, azim, each (M,).  elev in [-pi/2, pi/2] measured from the xz-plane
    towards +y; azim = signed angle from +z towards +x.  Clamps exactly as the reference does
    (CustomGeometry.py:36-43); exact poles stay undefined (NaN / garbage), as upstream."""
    r = torch.linalg.norm(X, dim=1)
    sin_elev = X[:, 1] / r
    elev = torch.asin(sin_elev)
    elev = torch.where(sin_elev <= -1, torch.full_like(elev, -np.pi / 2), elev)
    elev = torch.where(sin_elev >= 1, torch.full_like(elev, np.pi / 2), elev)
    cos_azim = X[:, 2] / (r * torch.cos(elev))
    azim = torch.acos(cos_azim)
    azim = torch.where(cos_azim <= -1, torch.full_like(azim, np.pi), azim)
    azim = torch.where(cos_azim >= 1, torch.zeros_like(azim), azim)
    azim = torch.where(X[:, 0] < 0, azim * -1, azim)
    return r, elev, azim


# ----------------------------------------------------------------------------------------------
# a2  utility/spherical_harmonics.py
# ----------------------------------------------------------------------------------------------
def _double_factorial(n):
    """n!! for odd n >= -1 (spherical_harmonics.py:49-50)."""
    out = 1.0
    while n > 1:
        out *= n
        n -= 2
    return out


def _rising(x, k):
    """x (x+1) ... (x+k-1) as a float (spherical_harmonics.py:53-54)."""
    out = float(x)
    for v in range(x + 1, x + k):
        out *= v
    return out


def legendre_table(x, n_degree=N_DEGREE):
    """Associated Legendre functions with Condon-Shortley phase, P[l][m] for 0<=m<=l<n_degree,
    built with the reference's recurrences and operation order (spherical_harmonics.py:67-108):
      P_0^0 = 1;  P_m^m = (-1)^m (2m-1)!! * pow(1 - x*x, m/2)
      P_l^m = ((2l-1)/(l-m)) * x * P_{l-1}^m  [ - ((l+m-1)/(l-m)) * P_{l-2}^m  if l-m > 1 ]
    The reference memoises on (l, m) only and is cleared by its callers whenever x changes;
    building the whole table per call is the same thing without the global."""
    P = [[None] * (l + 1) for l in range(n_degree)]
    for m in range(n_degree):
        for l in range(m, n_degree):
            if l == 0:
                P[l][m] = torch.ones_like(x)
            elif l == m:
                P[l][m] = ((-1) ** m * _double_factorial(2 * m - 1)) * torch.pow(1 - x * x, m / 2)
            else:
                y = ((2 * l - 1) / (l - m)) * x * P[l - 1][m]
                if l - m > 1:
                    y -= ((l + m - 1) / (l - m)) * P[l - 2][m]
                P[l][m] = y
    return P


def sh_normalisation(l, m_abs):
    """N_lm of the real (tesseral) harmonics (spherical_harmonics.py:126,138)."""
    n = math.sqrt((2 * l + 1) / (4 * math.pi))
    if m_abs:
        n *= math.sqrt(2.0 / _rising(l - m_abs + 1, 2 * m_abs))
    return n


def real_sh_basis(theta, phi, n_degree=N_DEGREE):
    """theta = polar angle (colatitude), phi = azimuth, both (M,) -> (M, n_degree^2) with column
    k = l^2 + l + m, m = -l..l; sin(|m| phi) for m<0, cos(m phi) for m>0
    (spherical_harmonics.py:111-156; concatenation over l as in SconeVis.py:236-238)."""
    P = legendre_table(torch.cos(theta), n_degree)
    cols = []
    for l in range(n_degree):
        for m in range(-l, l + 1):
            ma = abs(m)
            n = math.sqrt((2 * l + 1) / (4 * math.pi))
            if m == 0:
                cols.append(n * P[l][0])
                continue
            y = torch.cos(m * phi) if m > 0 else torch.sin(ma * phi)
            y *= P[l][ma]
            n *= math.sqrt(2.0 / _rising(l - ma + 1, 2 * ma))
            y *= n
            cols.append(y)
    return torch.stack(cols, dim=-1)


# ----------------------------------------------------------------------------------------------
# a3-a5  networks/SconeVis.py:164-303, networks/Macarons.py:138-178
# ----------------------------------------------------------------------------------------------
def _activated_projection(pts, harmonics, X_cam, use_sigmoid):
    """(B,P,>=3), (B,P,64), (B,C,3) -> (B,C,P) activated SH projection of every camera ray."""
    B, P = pts.shape[0], pts.shape[1]
    C = X_cam.shape[1]
    X_pts = pts[..., :3]
    rays = (X_cam.view(B, C, 1, 3).expand(-1, -1, P, -1)
            - X_pts.view(B, 1, P, 3).expand(-1, C, -1, -1)).reshape(-1, 3)
    _, elev, azim = spherical_coords(rays)
    theta = -elev + np.pi / 2.
    basis = real_sh_basis(theta, azim).view(B, C, P, N_HARMONICS)
    z = torch.sum(basis * harmonics.view(B, 1, P, N_HARMONICS).expand(-1, C, -1, -1), dim=-1)
    return torch.sigmoid(z) if use_sigmoid else torch.relu(z)


def visibility_gains(pts, harmonics, X_cam, use_sigmoid=True, cam_chunk=None):
    """SconeVis.compute_visibilities / Macarons.compute_visibility_gains -> (B,C,P).
    `cam_chunk` evaluates cameras in slices; every reduction is per (camera, point), so slicing
    is bit-neutral (it is the one change BASELINE.md section 3 permits for the CPU baseline)."""
    C = X_cam.shape[1]
    step = C if not cam_chunk else int(cam_chunk)
    parts = [_activated_projection(pts, harmonics, X_cam[:, c0:c0 + step], use_sigmoid)
             for c0 in range(0, C, step)]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


def coverage_gain(pts, harmonics, X_cam, use_sigmoid=True, cam_chunk=None):
    """SconeVis.compute_coverage_gain -> (B,C): mean over the P points (SconeVis.py:250)."""
    C = X_cam.shape[1]
    P = pts.shape[1]
    step = C if not cam_chunk else int(cam_chunk)
    parts = []
    for c0 in range(0, C, step):
        z = _activated_projection(pts, harmonics, X_cam[:, c0:c0 + step], use_sigmoid)
        parts.append(torch.sum(z, dim=-1) / P)
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


def coverage_gain_multiple(pts, harmonics, X_cam, n_cam, use_sigmoid=True):
    """SconeVis.compute_coverage_gain_multiple: coverage of every ordered n_cam-tuple of cameras,
    per-point max over the tuple then mean -> ((B, C^n_cam), (C^n_cam, n_cam))."""
    if n_cam not in (2, 3):
        raise NameError("n_cam is too large.")
    z = _activated_projection(pts, harmonics, X_cam, use_sigmoid)
    single = torch.arange(0, X_cam.shape[1])
    tuples = torch.cartesian_prod(*([single] * n_cam))
    zz = z[:, tuples]                                  # (B, C^n, n, P)
    return torch.sum(torch.max(zz, dim=-2)[0], dim=-1) / pts.shape[1], tuples


def nbv_argmax(scores):
    """(C,) or (B,C) -> index of the first maximum (testers/shapenet.py:172 torch.max; the strict
    `>` running maximum of testers/scene.py:454 gives the same index)."""
    s = np.asarray(scores)
    return np.argmax(s, axis=-1)


# ----------------------------------------------------------------------------------------------
# independent float64 closed form (no asin / acos / cos(m*phi) / pow): SURVEY.md appendix A.1
# ----------------------------------------------------------------------------------------------
def sh_basis_closed_form_f64(d):
    """d (...,3) float64 ray vectors -> (...,64) real SH values in the reference's ordering.
    cos(theta) = y/r, sin(theta) = rho/r, cos(phi) = z/rho, sin(phi) = x/rho."""
    d = np.asarray(d, dtype=np.float64)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    rho2 = x * x + z * z
    r = np.sqrt(rho2 + y * y)
    rho = np.sqrt(rho2)
    ct, st = y / r, rho / r
    cp, sp = z / rho, x / rho
    cm = [np.ones_like(ct)]
    sm = [np.zeros_like(ct)]
    for m in range(1, N_DEGREE):
        cm.append(cm[-1] * cp - sm[-1] * sp)
        sm.append(sm[-1] * cp + cm[-2] * sp)
    out = np.empty(d.shape[:-1] + (N_HARMONICS,), dtype=np.float64)
    for m in range(N_DEGREE):
        pmm = (-1.0) ** m * _double_factorial(2 * m - 1) * st ** m
        prev2, prev1 = None, pmm
        for l in range(m, N_DEGREE):
            if l == m:
                p = pmm
            elif l == m + 1:
                p = (2 * m + 1) * ct * pmm
            else:
                p = ((2 * l - 1) * ct * prev1 - (l + m - 1) * prev2) / (l - m)
            if l > m:
                prev2, prev1 = prev1, p
            n = sh_normalisation(l, m)
            k = l * l + l
            if m == 0:
                out[..., k] = n * p
            else:
                out[..., k + m] = n * p * cm[m]
                out[..., k - m] = n * p * sm[m]
    return out


def visibility_gains_f64(pts, harmonics, X_cam, use_sigmoid=True):
    """float64 truth for a3-a5: (B,P,>=3), (B,P,64), (B,C,3) -> (B,C,P)."""
    pts = np.asarray(pts, dtype=np.float64)[..., :3]
    H = np.asarray(harmonics, dtype=np.float64)
    cam = np.asarray(X_cam, dtype=np.float64)
    out = np.empty((pts.shape[0], cam.shape[1], pts.shape[1]), dtype=np.float64)
    for c in range(cam.shape[1]):
        Y = sh_basis_closed_form_f64(cam[:, c, None, :] - pts)
        z = np.sum(Y * H, axis=-1)
        out[:, c] = 1.0 / (1.0 + np.exp(-z)) if use_sigmoid else np.maximum(z, 0.0)
    return out


def coverage_gain_f64(pts, harmonics, X_cam, use_sigmoid=True):
    return visibility_gains_f64(pts, harmonics, X_cam, use_sigmoid).mean(axis=-1)


# ----------------------------------------------------------------------------------------------
# the same float64 closed form in C + OpenMP (oracle/c/sh_cov_f64.c): full benchmark shapes in seconds
# ----------------------------------------------------------------------------------------------
def _c_call(name, pts, harmonics, X_cam, use_sigmoid, per_point):
    from . import cbuild
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float32))
    H = np.ascontiguousarray(np.asarray(harmonics, dtype=np.float32))
    cam = np.ascontiguousarray(np.asarray(X_cam, dtype=np.float32))
    B, P, D = pts.shape
    C = cam.shape[1]
    assert H.shape == (B, P, N_HARMONICS) and cam.shape == (B, C, 3)
    out = np.empty((B, C, P) if per_point else (B, C), dtype=np.float64)
    import os
    try:   # all cores this process may use, even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)
        cbuild.load().omp_set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    rc = getattr(cbuild.load(), name)(pts.ctypes.data, D, H.ctypes.data, cam.ctypes.data, B, P, C, int(bool(use_sigmoid)),
                                      out.ctypes.data)
    if rc != 0:
        raise ValueError("%s: bad arguments" % name)
    return out


def coverage_gain_f64_c(pts, harmonics, X_cam, use_sigmoid=True):
    """coverage_gain_f64 computed by the C restatement on all host threads (fp32 inputs, float64 arithmetic)."""
    return _c_call("mac_oracle_coverage_f64", pts, harmonics, X_cam, use_sigmoid, False)


def visibility_gains_f64_c(pts, harmonics, X_cam, use_sigmoid=True):
    return _c_call("mac_oracle_visibility_f64", pts, harmonics, X_cam, use_sigmoid, True)
