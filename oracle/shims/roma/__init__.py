"""ORACLE SHIM (test infrastructure): CPU restatement of the seven functions of
roma==1.2.7 (environment.yml:28) that the reference's hot path calls
(robust_e_nerf/models/trajectories.py:87-89, robust_e_nerf/utils/tensor_ops.py:94,
111,165,172,175).  roma is NOT vendored under /root/reference and is not
installable here, so this follows the published roma 1.2.7 algorithm (XYZW
quaternions, Hamilton product, SciPy-style small-angle Taylor branches) as the
reference's own comments document them (tensor_ops.py:83-180).  PARITY UNPINNED:
there is no roma golden vector in the reference.
"""
import torch

from . import internal  # noqa: F401


def quat_conjugation(quat):
    """roma.quat_conjugation: negate the vector part (XYZW)."""
    inv = quat.clone()
    inv[..., :3] *= -1
    return inv


def quat_product(p, q):
    """roma.quat_product: Hamilton product p (x) q, XYZW convention."""
    vector = (
        p[..., None, 3] * q[..., :3]
        + q[..., None, 3] * p[..., :3]
        + torch.cross(p[..., :3], q[..., :3], dim=-1)
    )
    last = p[..., 3] * q[..., 3] - torch.sum(p[..., :3] * q[..., :3], dim=-1)
    return torch.cat((vector, last[..., None]), dim=-1)


def rotvec_to_unitquat(rotvec):
    """roma.rotvec_to_unitquat (adapted from SciPy Rotation.from_rotvec)."""
    rotvec, batch_shape = internal.flatten_batch_dims(rotvec, end_dim=-2)
    num_rotations, D = rotvec.shape
    assert D == 3
    norms = torch.norm(rotvec, dim=-1)
    small_angle = norms <= 1e-3
    large_angle = ~small_angle
    scale = torch.empty((num_rotations,), device=rotvec.device, dtype=rotvec.dtype)
    scale[small_angle] = (
        0.5 - norms[small_angle] ** 2 / 48 + norms[small_angle] ** 4 / 3840
    )
    scale[large_angle] = torch.sin(norms[large_angle] / 2) / norms[large_angle]
    quat = torch.empty((num_rotations, 4), device=rotvec.device, dtype=rotvec.dtype)
    quat[:, :3] = scale[:, None] * rotvec
    quat[:, 3] = torch.cos(norms / 2)
    return internal.unflatten_batch_dims(quat, batch_shape)


def unitquat_to_rotmat(quat):
    """roma.unitquat_to_rotmat (no re-normalisation; XYZW)."""
    x = quat[..., 0]
    y = quat[..., 1]
    z = quat[..., 2]
    w = quat[..., 3]
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    matrix = torch.empty(quat.shape[:-1] + (3, 3), dtype=quat.dtype, device=quat.device)
    matrix[..., 0, 0] = x2 - y2 - z2 + w2
    matrix[..., 1, 0] = 2 * (xy + zw)
    matrix[..., 2, 0] = 2 * (xz - yw)
    matrix[..., 0, 1] = 2 * (xy - zw)
    matrix[..., 1, 1] = -x2 + y2 - z2 + w2
    matrix[..., 2, 1] = 2 * (yz + xw)
    matrix[..., 0, 2] = 2 * (xz + yw)
    matrix[..., 1, 2] = 2 * (yz - xw)
    matrix[..., 2, 2] = -x2 - y2 + z2 + w2
    return matrix


def unitquat_slerp(q0, q1, steps, shortest_path=False):
    """roma.unitquat_slerp -- only used by dead code in the reference
    (tensor_ops.py:73, `slerp_uniform`); not restated."""
    raise NotImplementedError
