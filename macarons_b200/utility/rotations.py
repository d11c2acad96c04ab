"""Pose plumbing of the depth path (a handful of 3x3 matrices per step, host-side set-up math): what the reference
takes from pytorch3d.transforms in ManyDepth.forward (networks/ManyDepth.py:740-750)."""
import torch


def _mm3(A, B):
    """(..., n, 3) x (..., 3, 3) products written as a broadcast multiply + sum: a handful of 3x3 matrices per step do not
    warrant a GEMM library call."""
    return (A[..., :, :, None] * B[..., None, :, :]).sum(dim=-2)


def axis_angle_to_matrix(axis_angle):
    """(..., 3) rotation vectors -> (..., 3, 3) rotation matrices (Rodrigues; column-vector convention as pytorch3d)."""
    angle = torch.norm(axis_angle, dim=-1, keepdim=True)
    small = angle < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    a = torch.where(small, 1.0 - angle * angle / 6.0, torch.sin(safe) / safe)                     # sin(t)/t
    b = torch.where(small, 0.5 - angle * angle / 24.0, (1.0 - torch.cos(safe)) / (safe * safe))   # (1-cos t)/t^2
    x, y, z = axis_angle.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack((zero, -z, y, z, zero, -x, -y, x, zero), dim=-1).reshape(axis_angle.shape[:-1] + (3, 3))
    eye = torch.eye(3, dtype=axis_angle.dtype, device=axis_angle.device)
    return eye + a[..., None] * K + b[..., None] * _mm3(K, K)


def relative_cameras(R, T, pose, pose_factor):
    """Target camera (R (B,3,3), T (B,3)) and 6-vector relative poses (B, n_alpha, 6) = [translation, axis-angle] /
    pose_factor -> source cameras: R_alpha = R @ relative_R, T_alpha = relative_T + relative_R^T T (column form), the
    'correct formula' of ManyDepth.py:736-750."""
    B, n_alpha = pose.shape[0], pose.shape[1]
    rel_R = axis_angle_to_matrix(pose_factor * pose[..., 3:])
    rel_T = pose_factor * pose[..., :3]
    R_alpha = _mm3(R.view(B, 1, 3, 3), rel_R)
    T_alpha = rel_T + _mm3(T.view(B, 1, 1, 3), rel_R).squeeze(-2)
    return R_alpha, T_alpha
