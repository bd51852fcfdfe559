"""Drop-in mirror of the reference's ``lib/se3.py`` surface used on the forward path.

``compute_rigid_transformation`` runs the batched Procrustes kernel (no host SVD, no sync); the
3x4 / 4x4 helpers are trivial tensor algebra and stay PyTorch with the reference's signatures.
"""
from __future__ import annotations

import torch

from . import autograd, ops

__all__ = ["decompose_trans", "integrate_trans", "torch_identity", "torch_inverse", "torch_concatenate",
           "torch_transform", "compute_rigid_transformation"]


def decompose_trans(trans):
    """lib/se3.py:14-26."""
    if len(trans.shape) == 3:
        return trans[:, :3, :3], trans[:, :3, 3:4]
    return trans[:3, :3], trans[:3, 3:4]


def integrate_trans(R, t):
    """lib/se3.py:29-52 (torch branch)."""
    if len(R.shape) == 3:
        trans = torch.eye(4, device=R.device)[None].repeat(R.shape[0], 1, 1)
        trans[:, :3, :3] = R
        trans[:, :3, 3:4] = t.view([-1, 3, 1])
    else:
        trans = torch.eye(4, device=R.device)
        trans[:3, :3] = R
        trans[:3, 3:4] = t
    return trans


def torch_identity(batch_size):
    """lib/se3.py:55-56."""
    return torch.eye(3, 4)[None, ...].repeat(batch_size, 1, 1)


def torch_inverse(g):
    """lib/se3.py:59-71."""
    rot = g[..., 0:3, 0:3]
    trans = g[..., 0:3, 3]
    return torch.cat([rot.transpose(-1, -2), rot.transpose(-1, -2) @ -trans[..., None]], dim=-1)


def torch_concatenate(a, b):
    """lib/se3.py:74-93."""
    rot1, trans1 = a[..., :3, :3], a[..., :3, 3]
    rot2, trans2 = b[..., :3, :3], b[..., :3, 3]
    return torch.cat([rot1 @ rot2, rot1 @ trans2[..., None] + trans1[..., None]], dim=-1)


def torch_transform(g, a, normals=None):
    """lib/se3.py:96-117."""
    R = g[..., :3, :3]
    p = g[..., :3, 3]
    if len(g.size()) != len(a.size()):
        raise NotImplementedError
    b = torch.matmul(a, R.transpose(-1, -2)) + p[..., None, :]
    if normals is not None:
        return b, normals @ R.transpose(-1, -2)
    return b


def compute_rigid_transformation(src, src_corr, weight):
    """lib/se3.py:256-289.  src, src_corr (B,3,n), weight (B,1,n) -> R (B,3,3), t (B,3,1).  Differentiable: when autograd
    records the call the backward kernel (``ogmm_rigid_transform_backward``) supplies the gradients."""
    if autograd.records(src, src_corr, weight):
        rot, t = autograd.RigidTransform.apply(src, src_corr, weight)
    else:
        with torch.no_grad():
            rot, t = ops.rigid_transform(src, src_corr, weight)
    return rot, t.unsqueeze(-1)


compute_rigid_transformation.ogmm_autograd_safe = True
