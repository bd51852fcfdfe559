"""Backward support for the differentiable pieces of the path the reference trains through (SURVEY.md section 8(f)
N4): the feature M-step ``node_feats = gmm_params(gamma, feats)[1]`` (lib/utils.py:289), the soft-correspondence head
``GMMSVD(is_sk=False)`` (models/dgcnn.py:96-115) and ``compute_rigid_transformation`` (lib/se3.py:256-289); and the
DeepGMR path: the xyz moments with sigma as a function of gamma and ``gmm_register`` (baseline/deepgmr.py:17-38, :71-75).

In the reference's training step (train.py:57-75) the Sinkhorn loop runs under ``no_grad`` and ``gamma`` is detached
(lib/utils.py:275-286), so autograd enters the clustering only through ``feats``:

    mu[b,j,:]  = sum_n gamma[b,n,j] feats[b,n,:] / npi[b,j],   npi = pi N + 1e-5
    dL/dfeats[b,n,:] = sum_j gamma[b,n,j] dL/dmu[b,j,:] / npi[b,j]

Forward = the streaming TMA kernel (``ogmm_gmm_moments_feat``), backward = ``ogmm_gmm_moments_feat_backward``; pi does
not depend on feats.  A gamma that itself requires grad is outside this function (``can_differentiate`` says so and
``install()`` then leaves that call to the reference's own function).
"""
from __future__ import annotations

import torch

from . import ops


class FeatureMoments(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gamma, feats):
        pi, mu = ops.gmm_moments(gamma, feats)
        ctx.save_for_backward(gamma, pi)
        ctx.like = feats                                   # shape / strides only are used; feats itself is not needed
        ctx.mark_non_differentiable(pi)
        return pi, mu

    @staticmethod
    def backward(ctx, grad_pi, grad_mu):
        gamma, pi = ctx.saved_tensors
        if grad_mu is None:
            return None, None
        return None, ops.gmm_moments_feat_backward(gamma, grad_mu.contiguous().float(), pi, ctx.like)


def can_differentiate(gamma, pts, return_sigma=False):
    """True when the call is the one this module differentiates: wide features, no sigma, gamma without grad."""
    return (not return_sigma) and pts.dim() == 3 and pts.shape[-1] > 4 and not gamma.requires_grad


def feature_moments(gamma, feats):
    """(pi, mu) with autograd history into ``feats``."""
    return FeatureMoments.apply(gamma.detach(), feats)


class SoftProcrustes(torch.autograd.Function):
    """GMMSVD(is_sk=False) -> (R, t, corr).  Backward = ``ogmm_soft_procrustes_backward`` (closed-form chain through the
    softmax correspondences, the weighted covariance and the 3x3 SVD); inputs are saved, nothing else."""

    @staticmethod
    def forward(ctx, src_mu, tgt_mu, src_desc, tgt_desc, temperature):
        rot, t, corr, _ = ops.soft_procrustes(src_mu, tgt_mu, src_desc, tgt_desc, temperature=temperature)
        ctx.save_for_backward(src_mu, tgt_mu, src_desc, tgt_desc)
        ctx.temperature = temperature
        return rot, t, corr

    @staticmethod
    def backward(ctx, grad_rot, grad_t, grad_corr):
        src_mu, tgt_mu, src_desc, tgt_desc = ctx.saved_tensors
        g = ops.soft_procrustes_backward(src_mu, tgt_mu, src_desc, tgt_desc, _f32(grad_rot), _f32(grad_t), _f32(grad_corr),
                                         temperature=ctx.temperature)
        return tuple(gi if need else None for gi, need in zip(g, ctx.needs_input_grad[:4])) + (None,)


class RigidTransform(torch.autograd.Function):
    """compute_rigid_transformation -> (R (B,3,3), t (B,3)).  Backward = ``ogmm_rigid_transform_backward``."""

    @staticmethod
    def forward(ctx, src, corr, weight):
        rot, t = ops.rigid_transform(src, corr, weight)
        ctx.save_for_backward(src, corr, weight)
        return rot, t

    @staticmethod
    def backward(ctx, grad_rot, grad_t):
        src, corr, weight = ctx.saved_tensors
        g = ops.rigid_transform_backward(src, corr, weight, _f32(grad_rot), _f32(grad_t))
        return tuple(gi if need else None for gi, need in zip(g, ctx.needs_input_grad))


class NarrowMoments(torch.autograd.Function):
    """gmm_params(gamma, pts[, return_sigma]) on 3-D points, differentiable with respect to gamma (how DeepGMR trains:
    gamma = softmax(logits), baseline/deepgmr.py:71-74).  Backward = ``ogmm_gmm_moments_backward``."""

    @staticmethod
    def forward(ctx, gamma, pts, return_sigma):
        out = ops.gmm_moments(gamma, pts, return_sigma)
        ctx.save_for_backward(pts, *out)
        ctx.like = gamma
        return out

    @staticmethod
    def backward(ctx, *grads):
        pts, pi, mu, *rest = ctx.saved_tensors
        sigma = rest[0] if rest else None
        g_pi, g_mu = grads[0], grads[1]
        g_sigma = grads[2] if len(grads) > 2 else None
        return ops.gmm_moments_backward(pts, pi, mu, sigma, _f32(g_pi), _f32(g_mu), _f32(g_sigma), ctx.like), None, None


class GmmRegister(torch.autograd.Function):
    """baseline/deepgmr.py:17-38 -> T (B,4,4).  Backward = ``ogmm_gmm_register_backward``."""

    @staticmethod
    def forward(ctx, pi_s, mu_s, mu_t, sigma_t):
        ctx.save_for_backward(pi_s, mu_s, mu_t, sigma_t)
        return ops.gmm_register(pi_s, mu_s, mu_t, sigma_t)

    @staticmethod
    def backward(ctx, grad_tf):
        g = ops.gmm_register_backward(*ctx.saved_tensors, grad_tf.float())
        return tuple(gi if need else None for gi, need in zip(g, ctx.needs_input_grad))


def can_differentiate_narrow(gamma, pts, return_sigma=False):
    """True for the xyz-moments call DeepGMR differentiates: 3-D points without grad, gamma with grad."""
    return pts.dim() == 3 and pts.shape[-1] == 3 and not pts.requires_grad


def _f32(g):
    return None if g is None else g.float()


def records(*tensors):
    """True when autograd would have to record a call on these tensors."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
