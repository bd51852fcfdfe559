"""Backward support for the differentiable piece of the path that the reference trains through first (SURVEY.md
section 8(f) N4): the feature M-step ``node_feats = gmm_params(gamma, feats)[1]`` (lib/utils.py:289).

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
