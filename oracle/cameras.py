"""Camera / rotation conventions of pytorch3d 0.6.2 that the depth path relies on (reference environment.yml:142).

TEST INFRASTRUCTURE (see oracle/__init__.py).  pytorch3d is a third-party dependency that is NOT installed in the
build container and not vendored by the reference, so these semantics are RESTATED from its published source
(pytorch3d/renderer/cameras.py `FoVPerspectiveCameras`, pytorch3d/transforms/transform3d.py `Transform3d`,
pytorch3d/transforms/rotation_conversions.py) and could not be executed against the original here:
**parity unpinned for this file** (SURVEY.md section 8c / appendix B).  It is also what stands in for pytorch3d
when `tests/golden/make_golden.py` runs the reference's own ManyDepth code.

Conventions: row vectors, X_view = X_world @ R + T; +X left, +Y up, +Z into the scene in view and NDC space;
defaults znear = 1, zfar = 100, fov = 60 degrees, aspect ratio 1 (the reference only overrides R, T, zfar).
"""
import math

import torch


class Transform3d:
    def __init__(self, matrix):
        self._matrix = matrix                      # (N,4,4), row-vector convention: p' = [p, 1] @ M

    def get_matrix(self):
        return self._matrix

    def compose(self, other):
        return Transform3d(self._matrix @ other._matrix)

    def inverse(self):
        return Transform3d(torch.inverse(self._matrix))

    def transform_points(self, points, eps=None):
        ones = torch.ones(points.shape[:-1] + (1,), dtype=points.dtype, device=points.device)
        out = torch.cat([points, ones], dim=-1) @ self._matrix
        denom = out[..., 3:]
        if eps is not None:
            denom_sign = denom.sign() + (denom == 0.0).type_as(denom)
            denom = denom_sign * torch.clamp(denom.abs(), eps)
        res = out[..., :3] / denom
        if res.shape[0] == 1 and points.dim() == 2:   # pytorch3d: a (P,3) input with one transform comes back as (P,3)
            res = res.reshape(points.shape)
        return res


class FoVPerspectiveCameras:
    def __init__(self, znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees=True, R=None, T=None, device="cpu"):
        self.R, self.T = R.to(device), T.to(device)
        n = self.R.shape[0]
        as_t = lambda v: (v.to(device).to(torch.float32).view(-1) if isinstance(v, torch.Tensor)
                          else torch.full((n,), float(v), device=device))
        self.znear, self.zfar, self.aspect_ratio, self.fov = as_t(znear), as_t(zfar), as_t(aspect_ratio), as_t(fov)
        if self.zfar.numel() == 1 and n > 1:
            self.zfar = self.zfar.expand(n)
        self.degrees, self.device = degrees, device

    def get_world_to_view_transform(self):
        n = self.R.shape[0]
        M = torch.zeros(n, 4, 4, dtype=torch.float32, device=self.R.device)
        M[:, :3, :3] = self.R
        M[:, 3, :3] = self.T
        M[:, 3, 3] = 1.0
        return Transform3d(M)

    def get_camera_center(self):
        return self.get_world_to_view_transform().inverse().transform_points(torch.zeros(self.R.shape[0], 1, 3,
                                                                                          device=self.R.device))[:, 0]

    def get_projection_transform(self):
        n = self.R.shape[0]
        K = torch.zeros(n, 4, 4, dtype=torch.float32, device=self.R.device)
        fov = self.fov * math.pi / 180 if self.degrees else self.fov
        tan_half = torch.tan(fov / 2)
        max_y = tan_half * self.znear
        min_y = -max_y
        max_x = max_y * self.aspect_ratio
        min_x = -max_x
        K[:, 0, 0] = 2.0 * self.znear / (max_x - min_x)
        K[:, 1, 1] = 2.0 * self.znear / (max_y - min_y)
        K[:, 0, 2] = (max_x + min_x) / (max_x - min_x)
        K[:, 1, 2] = (max_y + min_y) / (max_y - min_y)
        K[:, 3, 2] = 1.0
        K[:, 2, 2] = self.zfar / (self.zfar - self.znear)
        K[:, 2, 3] = -(self.zfar * self.znear) / (self.zfar - self.znear)
        return Transform3d(K.transpose(1, 2).contiguous())

    def get_full_projection_transform(self):
        return self.get_world_to_view_transform().compose(self.get_projection_transform())

    def unproject_points(self, xy_depth, world_coordinates=True, scaled_depth_input=False):
        to_ndc = self.get_full_projection_transform() if world_coordinates else self.get_projection_transform()
        if scaled_depth_input:
            xy_sdepth = xy_depth
        else:
            K = self.get_projection_transform().get_matrix()
            unsq = (slice(None),) + (None,) * (xy_depth.dim() - 1)
            f1, f2 = K[:, 2, 2][unsq], K[:, 3, 2][unsq]
            sdepth = (f1 * xy_depth[..., 2:3] + f2) / xy_depth[..., 2:3]
            xy_sdepth = torch.cat((xy_depth[..., 0:2], sdepth), dim=-1)
        return to_ndc.inverse().transform_points(xy_sdepth)


# ---- rotation conversions (pytorch3d/transforms/rotation_conversions.py), real-first quaternions ----
def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_quaternion(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = angles * 0.5
    small = angles.abs() < 1e-6
    sin_half_over_angle = torch.empty_like(angles)
    sin_half_over_angle[~small] = torch.sin(half[~small]) / angles[~small]
    sin_half_over_angle[small] = 0.5 - (angles[small] * angles[small]) / 48
    return torch.cat([torch.cos(half), axis_angle * sin_half_over_angle], dim=-1)


def axis_angle_to_matrix(axis_angle):
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(matrix):
    batch = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                             1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    return quat_candidates[torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch + (4,))


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def quaternion_apply(quaternion, point):
    real = point.new_zeros(point.shape[:-1] + (1,))
    point_q = torch.cat((real, point), -1)
    inv = quaternion * quaternion.new_tensor([1, -1, -1, -1])
    out = quaternion_raw_multiply(quaternion_raw_multiply(quaternion, point_q), inv)
    return out[..., 1:]
