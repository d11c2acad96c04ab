"""A minimal field-of-view perspective camera with the interface the NBV path reads from pytorch3d's
`FoVPerspectiveCameras` (reference environment.yml:142, pytorch3d 0.6.2; used at utility/macarons_utils.py:1603-1660,
:2339-2435, utility/scone_utils.py:896-897): row-vector convention `X_view = X_world @ R + T`, +X left, +Y up, +Z into the
scene, NDC in [-1, 1] for a square image, depth mapped by (zfar (z - znear)) / (z (zfar - znear)).

It exists so that the package can be driven (benchmarks, stand-alone use) on a machine without pytorch3d; callers that do
have pytorch3d pass its camera objects instead -- every function of this package only uses the methods below.  The
matrices are built once per call on the cameras' device in fp32; inverses of the rigid world-to-view transform are formed
in closed form (R^T, -T R^T) instead of a numerical 4x4 inverse."""
import math

import torch


class Transform:
    """Batch of 4x4 matrices acting on row vectors: p' = [p, 1] @ M, followed by the perspective divide."""

    def __init__(self, matrix, inverse_matrix=None):
        self._m = matrix
        self._inv = inverse_matrix

    def get_matrix(self):
        return self._m

    def compose(self, other):
        inv = None if (self._inv is None or other._inv is None) else other._inv @ self._inv
        return Transform(self._m @ other._m, inv)

    def inverse(self):
        return Transform(torch.inverse(self._m) if self._inv is None else self._inv, self._m)

    def transform_points(self, points, eps=None):
        flat = points.dim() == 2
        pts = points[None] if flat else points
        m = self._m[:, None]                                    # (N, 1, 4, 4); broadcast multiply-adds, no GEMM library call
        hom = pts[..., 0:1] * m[..., 0, :] + pts[..., 1:2] * m[..., 1, :] + pts[..., 2:3] * m[..., 2, :] + m[..., 3, :]
        w = hom[..., 3:]
        if eps is not None:
            w = torch.where(w.abs() < eps, torch.where(w < 0, -torch.ones_like(w), torch.ones_like(w)) * eps, w)
        out = hom[..., :3] / w
        return out[0] if (flat and out.shape[0] == 1) else out


class FoVCamera:
    """N cameras.  R (N,3,3), T (N,3); znear, zfar, fov (degrees), aspect_ratio: floats or (N,) tensors."""

    def __init__(self, R, T, znear=1.0, zfar=100.0, fov=60.0, aspect_ratio=1.0, device=None):
        device = R.device if device is None else torch.device(device)
        self.R, self.T = R.to(device=device, dtype=torch.float32), T.to(device=device, dtype=torch.float32)
        self.device = device
        n = self.R.shape[0]

        def vec(v):
            t = torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1)
            return t.expand(n) if t.numel() == 1 else t
        self.znear, self.zfar, self.fov, self.aspect_ratio = vec(znear), vec(zfar), vec(fov), vec(aspect_ratio)

    def __len__(self):
        return self.R.shape[0]

    def __getitem__(self, i):
        sl = slice(i, i + 1) if isinstance(i, int) else i
        return FoVCamera(self.R[sl], self.T[sl], self.znear[sl], self.zfar[sl], self.fov[sl], self.aspect_ratio[sl], self.device)

    def get_world_to_view_transform(self):
        n = len(self)
        m = torch.zeros(n, 4, 4, device=self.device)
        m[:, :3, :3], m[:, 3, :3], m[:, 3, 3] = self.R, self.T, 1.0
        inv = torch.zeros(n, 4, 4, device=self.device)
        Rt = self.R.transpose(1, 2)
        inv[:, :3, :3], inv[:, 3, :3], inv[:, 3, 3] = Rt, -(self.T[:, None, :] @ Rt)[:, 0], 1.0
        return Transform(m, inv)

    def get_camera_center(self):
        return -(self.T[:, None, :] @ self.R.transpose(1, 2))[:, 0]

    def get_projection_transform(self):
        n = len(self)
        t = torch.tan(self.fov * (math.pi / 360.0))          # tan(fov / 2)
        k = torch.zeros(n, 4, 4, device=self.device)
        k[:, 0, 0] = 1.0 / (t * self.aspect_ratio)
        k[:, 1, 1] = 1.0 / t
        k[:, 2, 2] = self.zfar / (self.zfar - self.znear)
        k[:, 3, 2] = -(self.zfar * self.znear) / (self.zfar - self.znear)
        k[:, 2, 3] = 1.0
        return Transform(k)

    def get_full_projection_transform(self):
        return self.get_world_to_view_transform().compose(self.get_projection_transform())


def look_at(eye, at, up=(0.0, 1.0, 0.0)):
    """Camera rotations / translations looking from `eye` (N,3) at `at` (N,3) -> R (N,3,3), T (N,3) in the convention above
    (what pytorch3d's look_at_view_transform returns for eye= / at=)."""
    up = torch.as_tensor(up, dtype=eye.dtype, device=eye.device).expand_as(eye)
    z = torch.nn.functional.normalize(at - eye, dim=-1, eps=1e-5)
    x = torch.nn.functional.normalize(torch.linalg.cross(up, z), dim=-1, eps=1e-5)
    y = torch.nn.functional.normalize(torch.linalg.cross(z, x), dim=-1, eps=1e-5)
    R = torch.stack((x, y, z), dim=-1)                       # columns = camera axes in world coordinates
    T = -(eye[:, None, :] @ R)[:, 0]
    return R.contiguous(), T.contiguous()
