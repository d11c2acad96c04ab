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


# ---- networks (rows a6-a9) -----------------------------------------------------------------------
def seeded_state_dict(template, seed):
    """Deterministic weights for a SconeOcc / SconeVis `state_dict()` template (name -> tensor), independent of
    module construction order: every tensor is drawn from its own generator keyed by (seed, name).
    Scales follow the reference initialisers (utility/scone_utils.py:260-289, 399-428): Xavier-normal for
    w_q / w_k / w_v, Kaiming-normal (relu) for the other Linear weights, default-Linear-style uniform biases;
    LayerNorm weights / biases are perturbed away from (1, 0) so that they are actually exercised."""
    import zlib
    out = {}
    for name in sorted(template):
        shape = tuple(template[name].shape)
        gen = torch.Generator(device="cpu").manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        is_norm = ".norm" in name or name.startswith("norm")
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.long)
            continue
        if name.endswith("running_mean"):
            out[name] = (0.1 * torch.randn(shape, generator=gen)).to(torch.float32)
            continue
        if name.endswith("running_var"):
            out[name] = (0.5 + torch.rand(shape, generator=gen)).to(torch.float32)
            continue
        is_bn = ".bn" in name or "downsample.1" in name
        if len(shape) == 4:       # Conv2d (out, in, kh, kw) / ConvTranspose2d (in, out, kh, kw): Kaiming-normal fan-in
            fan_in = (shape[0] if ".upconv." in name else shape[1]) * shape[2] * shape[3]
            out[name] = (torch.randn(shape, generator=gen) * math.sqrt(2.0 / fan_in)).to(torch.float32)
            continue
        if is_bn:
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=gen)
            out[name] = t.to(torch.float32)
            continue
        if len(shape) == 2:
            fan_out, fan_in = shape
            if name.split(".")[-2] in ("w_q", "w_k", "w_v"):
                std = math.sqrt(2.0 / (fan_in + fan_out))
            else:
                std = math.sqrt(2.0 / fan_in)
            t = torch.randn(shape, generator=gen) * std
        elif is_norm and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif is_norm:
            t = 0.1 * torch.randn(shape, generator=gen)
        else:
            w = template[name[:-4] + "weight"]
            fan_in = w[0].numel() if w.dim() != 4 or ".upconv." not in name else w.shape[0] * w.shape[2] * w.shape[3]
            t = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(fan_in)
        out[name] = t.to(torch.float32)
    return out


def state_dict_digest(sd):
    import hashlib
    h = hashlib.sha256()
    for name in sorted(sd):
        h.update(name.encode())
        h.update(sd[name].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def airplane_surface(n, gen):
    """'Airplane-like' closed surface samples (ellipsoid fuselage + two flat wing boxes + tail fin), centred and
    scaled so that the bounding-box diagonal is 1 (like adjust_mesh_diagonally, utility/utils.py:633-648)."""
    n_f, n_w, n_t = n // 2, n // 3, n - n // 2 - n // 3
    v = torch.randn(n_f, 3, generator=gen)
    v = v / v.norm(dim=-1, keepdim=True)
    fus = v * torch.tensor([0.12, 0.12, 1.0])
    wing = (torch.rand(n_w, 3, generator=gen) - 0.5) * torch.tensor([1.6, 0.03, 0.35])
    face = torch.randint(0, 2, (n_w,), generator=gen).float() * 2 - 1
    wing[:, 1] = face * 0.015
    tail = (torch.rand(n_t, 3, generator=gen) - 0.5) * torch.tensor([0.03, 0.4, 0.25])
    tail[:, 0] = (torch.randint(0, 2, (n_t,), generator=gen).float() * 2 - 1) * 0.015
    tail = tail + torch.tensor([0.0, 0.25, -0.85])
    pts = torch.cat((fus, wing, tail))
    lo, hi = pts.min(dim=0)[0], pts.max(dim=0)[0]
    return ((pts - (lo + hi) / 2) / (hi - lo).norm()).contiguous()


def sconeocc_inputs(B, N, Q, seed, grid=False):
    """pc (B,N,3): the part of an airplane-like surface facing a sphere camera; x (B,Q,3) queries (a regular grid in
    [-0.5,0.5]^3 when `grid`, Q must be a cube); view_harmonics (B,Q,64) ~ N(0, 0.3^2)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    pcs = []
    for _ in range(B):
        surf = airplane_surface(8 * N, gen)
        cam = sphere_cameras(1, 1.5, gen)[0]
        order = torch.argsort((surf * cam).sum(-1), descending=True)   # the half facing the camera
        keep = order[:4 * N][torch.randperm(4 * N, generator=gen)[:N]]
        pcs.append(surf[keep])
    pc = torch.stack(pcs)
    if grid:
        n = round(Q ** (1 / 3))
        assert n ** 3 == Q
        ax = (torch.arange(n, dtype=torch.float32) + 0.5) / n - 0.5
        g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).view(1, Q, 3)
        x = g.expand(B, Q, 3).contiguous()
    else:
        x = torch.rand(B, Q, 3, generator=gen) - 0.5
    vh = 0.3 * torch.randn(B, Q, 64, generator=gen)
    return pc.contiguous(), x, vh.contiguous()


def sconevis_inputs(B, S, seed):
    """pts (B,S,4) = xyz ~ U[-0.5,0.5]^3 + occupancy ~ U[0.1,1]; view_harmonics (B,S,64) ~ N(0, 0.3^2)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    pts = torch.rand(B, S, 4, generator=gen) - 0.5
    pts[..., 3] = 0.1 + 0.9 * torch.rand(B, S, generator=gen)
    vh = 0.3 * torch.randn(B, S, 64, generator=gen)
    return pts.contiguous(), vh.contiguous()


def depth_inputs(B, H, W, seed, n_alpha=2):
    """Smooth synthetic frames (low-frequency colour fields, so that the plane sweep is not sampling white noise),
    identity target camera as in apply_depth_model (utility/macarons_utils.py:917-919) and small relative poses:
    -> x (B,3,H,W), x_alpha (B,n_alpha,3,H,W), R (B,3,3), T (B,3), zfar (B,), gt_pose (B,n_alpha,6)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))

    def frames(n):
        low = torch.rand(n, 3, H // 16 + 2, W // 16 + 2, generator=gen)
        img = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
        return (0.8 * img + 0.2 * torch.rand(n, 3, H, W, generator=gen)).clamp(0, 1)

    x = frames(B)
    x_alpha = frames(B * n_alpha).view(B, n_alpha, 3, H, W)
    R = torch.eye(3).view(1, 3, 3).expand(B, -1, -1).contiguous()
    T = torch.zeros(B, 3)
    zfar = torch.full((B,), 750.0)
    gt_pose = torch.cat((0.02 * (torch.rand(B, n_alpha, 3, generator=gen) - 0.5),       # translation / pose_factor
                         0.002 * (torch.rand(B, n_alpha, 3, generator=gen) - 0.5)), dim=-1)
    return x.contiguous(), x_alpha.contiguous(), R, T, zfar, gt_pose.contiguous()


# ---- MACARONS candidate scoring ---------------------------------------
def look_at_RT(eye, at, up=(0.0, 1.0, 0.0)):
    """pytorch3d `look_at_view_transform(eye=, at=)` convention restated: rows of X_view = X_world @ R + T,
    +Z towards `at`, +X left, +Y up.  eye, at (n,3) -> R (n,3,3), T (n,3)."""
    up = torch.tensor(up, dtype=eye.dtype).view(1, 3).expand_as(eye)
    z = torch.nn.functional.normalize(at - eye, eps=1e-5)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=1), eps=1e-5)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    R = torch.cat((x[:, None, :], y[:, None, :], z[:, None, :]), dim=1).transpose(1, 2)
    T = -torch.bmm(R.transpose(1, 2), eye[:, :, None])[:, :, 0]
    return R.contiguous(), T.contiguous()


def ndc_bounds(image_height=256, image_width=456):
    """(min_ndc_x, max_ndc_x, min_ndc_y, max_ndc_y) as `Camera.__init__` derives them (macarons_utils.py:1929-1938)."""
    m = min(image_width, image_height)
    max_x = image_width / m
    min_x = image_width / m - ((image_width - 1) / (m - 1)) * 2
    max_y = image_height / m
    min_y = image_height / m - ((image_height - 1) / (m - 1)) * 2
    f = lambda v: float(torch.tensor(v, dtype=torch.float32))
    return f(min_x), f(max_x), f(min_y), f(max_y)


def macarons_scene(N, C, seed, half_extent=(40.0, 15.0, 40.0)):
    """A synthetic MACARONS scoring state: N proxy points uniform in the scene box with a blobby occupancy field
    (a third of them below the 0.1 threshold), their view harmonics, C candidate cameras inside the box looking at
    random targets, and the prediction camera.  -> dict of CPU tensors."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    ext = torch.tensor(half_extent)
    X = (torch.rand(N, 3, generator=gen) * 2 - 1) * ext
    centres = (torch.rand(6, 3, generator=gen) * 2 - 1) * ext * 0.7
    d2 = ((X[:, None, :] - centres[None]) / (0.45 * ext)).pow(2).sum(-1)
    occ = (torch.exp(-d2).sum(-1) * (0.6 + 0.4 * torch.rand(N, generator=gen))).clamp(0, 1).view(N, 1)
    vh = 0.3 * torch.randn(N, 64, generator=gen)
    eye = (torch.rand(C, 3, generator=gen) * 2 - 1) * ext * 0.9
    at = (torch.rand(C, 3, generator=gen) * 2 - 1) * ext * 0.5
    R, T = look_at_RT(eye, at)
    pe = (torch.rand(1, 3, generator=gen) * 2 - 1) * ext * 0.5
    pR, pT = look_at_RT(pe, torch.zeros(1, 3))
    return {"X_world": X.contiguous(), "occ": occ.contiguous(), "vh": vh.contiguous(), "X_cam": eye.contiguous(),
            "R": R, "T": T, "pred_R": pR, "pred_T": pT, "x_min": -ext, "x_max": ext,
            "diag": torch.linalg.norm(2 * ext).item(), "u": torch.rand(C, 2048, 1, generator=gen)}


def depth_io_inputs(H, W, seed, n=2):
    """n depth maps with masks and colours, cameras looking at the origin, and 3-D probe points around it."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    eye = torch.tensor([[3., 2., -8.], [1., -1., -6.], [-4., 1., 5.]])[:n]
    R, T = look_at_RT(eye, torch.zeros(n, 3))
    depth = 2 + 10 * torch.rand(n, H, W, 1, generator=gen)
    mask = torch.rand(n, H, W, 1, generator=gen) > 0.3
    images = torch.rand(n, H, W, 3, generator=gen)
    pts = (torch.rand(700, 3, generator=gen) - 0.5) * 8
    return {"R": R, "T": T, "depth": depth, "mask": mask, "images": images, "pts": pts}
