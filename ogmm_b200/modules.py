"""nn.Module mirrors of the reference's hot-path modules and DeepGMR's registration function.

  Clustering    models/gmmreg.py:19-29
  GMMSVD        models/dgcnn.py:90-115
  graph_features  models/dgcnn.py:135-137 (the kNN + edge-gather opening of DGCNN.forward)
  gmm_register  baseline/deepgmr.py:17-38
  deepgmr_em    baseline/deepgmr.py:71-74
"""
from __future__ import annotations

import torch
from torch import nn

from . import autograd, ops
from ._guard import forward_only
from .se3 import compute_rigid_transformation
from .utils import wkeans_plus

__all__ = ["Clustering", "GMMSVD", "graph_features", "gmm_register", "deepgmr_em", "edge_conv1", "dgcnn_forward",
           "angle_conv1", "position_encoding_forward"]


class Clustering(nn.Module):
    """models/gmmreg.py:19-29: transposed views into wkeans_plus(iters=10, tau=1.0)."""

    def __init__(self, n_clusters):
        super().__init__()
        self.n_clusters = n_clusters

    def forward(self, xyz, feats, o_scores):
        return wkeans_plus(xyz.transpose(-1, -2), feats.transpose(-1, -2), o_scores, self.n_clusters,
                           iters=10, tau=1.0)

    def ogmm_can_differentiate(self, *args, **kwargs):
        return True                      # wkeans_plus: clustering under no_grad, differentiable feature M-step


class GMMSVD(nn.Module):
    """models/dgcnn.py:90-115.  ``is_sk=False`` (how models/gmmreg.py:41 builds it) is one fused kernel."""

    def __init__(self, is_sk=True, epsilon=1e-3):
        super().__init__()
        self.is_sk = is_sk
        self.epsilon = epsilon

    def forward(self, src, tgt, src_desc, tgt_desc, src_pi, tgt_pi):
        batch_size = src.size(0)
        if not self.is_sk:
            if autograd.records(src, tgt, src_desc, tgt_desc):            # training: the head brings its own backward
                R, t, src_corr = autograd.SoftProcrustes.apply(src, tgt, src_desc, tgt_desc, 0.05)
            else:
                with torch.no_grad():
                    R, t, src_corr, _ = ops.soft_procrustes(src, tgt, src_desc, tgt_desc, temperature=0.05)
            return R, t.view(batch_size, 3), src_corr, tgt.transpose(-1, -2)
        return self._forward_sk(src, tgt, src_desc, tgt_desc, src_pi, tgt_pi)

    def ogmm_can_differentiate(self, *args, **kwargs):
        return not self.is_sk            # the softmax head has a backward kernel; the Sinkhorn variant does not

    @forward_only
    def _forward_sk(self, src, tgt, src_desc, tgt_desc, src_pi, tgt_pi):
        batch_size = src.size(0)
        similarity = ops.cos_similarity(src_desc, tgt_desc)
        scores = ops.sinkhorn(2.0 * (1.0 - similarity), src_pi, tgt_pi, 1e-2, 1e-2, 30)[0]
        scores = torch.nan_to_num(scores, 1e-4)
        scores = scores / torch.sum(scores, dim=-1, keepdim=True).clip(min=1e-4)
        src_corr = torch.einsum('bmd,bnm->bdn', tgt, scores)
        weight = scores.sum(dim=-1).unsqueeze(1)
        R, t = compute_rigid_transformation(src.transpose(-1, -2), src_corr, weight)
        return R, t.view(batch_size, 3), src_corr, tgt.transpose(-1, -2)


@forward_only
def graph_features(x, k=20):
    """models/dgcnn.py:135-137 in one launch: x (B,3,N) -> (B,6,N,k) edge tensor fed to conv1."""
    pts = x.transpose(-1, -2)
    return ops.knn_graph(pts, pts, k, want_edge=True)[2].permute(0, 3, 1, 2)


def gmm_register(pi_s, mu_s, mu_t, sigma_t):
    """baseline/deepgmr.py:17-38 -> (B,4,4); no host SVD and no hard-coded ``.cuda()``.  Differentiable (backward kernel
    ``ogmm_gmm_register_backward``) when autograd records the call."""
    if autograd.records(pi_s, mu_s, mu_t, sigma_t):
        return autograd.GmmRegister.apply(pi_s, mu_s, mu_t, sigma_t)
    with torch.no_grad():
        return ops.gmm_register(pi_s, mu_s, mu_t, sigma_t)


gmm_register.ogmm_autograd_safe = True


@forward_only
def deepgmr_em(logits, pts):
    """baseline/deepgmr.py:71-74 fused: logits (B,J,N), pts (B,3,N) -> gamma (B,J,N), pi, mu, sigma."""
    return ops.softmax_moments(logits, pts)


def _fold_bn(bn):
    """BatchNorm in eval mode as y * scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps) if bn.affine else 1.0 / torch.sqrt(bn.running_var + bn.eps)
    shift = (bn.bias if bn.affine else 0.0) - bn.running_mean * scale
    return scale.float(), shift.float()


@forward_only
def edge_conv1(x, idx, conv, bn, want_act=True):
    """models/dgcnn.py:137-141 in one launch, for inference: x (B,3,N), idx (B,N,k), conv = Conv2d(6, C, 1, bias=False),
    bn = BatchNorm2d(C) in eval mode -> (relu(bn(conv(edge))) (B,C,N,k) | None, its max over k (B,C,N,1))."""
    if bn.training:
        raise RuntimeError("edge_conv1 folds BatchNorm's running statistics: the module must be in eval() mode")
    scale, shift = _fold_bn(bn)
    act, pooled = ops.edge_conv_max(x, idx, conv.weight.reshape(conv.out_channels, -1), scale, shift, want_act)
    return act, pooled.unsqueeze(-1)


def dgcnn_forward(self, x):
    """Drop-in body for ``DGCNN.forward`` (models/dgcnn.py:133-154) in eval / no_grad mode: kNN graph (K1), then the edge
    gather fused with conv1 + bn1 + ReLU + max (N3); conv2..conv5 are the module's own PyTorch layers, unchanged."""
    import torch.nn.functional as F
    batch_size, _, num_points = x.size()
    pts = x.transpose(-1, -2)
    idx = ops.knn_graph(pts, pts, self.k)[0]
    x, x1 = edge_conv1(x, idx, self.conv1, self.bn1)
    x = F.relu(self.bn2(self.conv2(x)))
    x2 = x.max(dim=-1, keepdim=True)[0]
    x = F.relu(self.bn3(self.conv3(x)))
    x3 = x.max(dim=-1, keepdim=True)[0]
    x = F.relu(self.bn4(self.conv4(x)))
    x4 = x.max(dim=-1, keepdim=True)[0]
    x = torch.cat((x1, x2, x3, x4), dim=1)
    return F.relu(self.bn5(self.conv5(x))).view(batch_size, -1, num_points)


@forward_only
def angle_conv1(points, idx, conv, bn, slope=0.2, want_alpha=False):
    """models/attn.py:65-73 in one launch, for inference: points (B,3,N), idx (B,N,k) (the k-NN graph of the points),
    conv = Conv2d(1, C, 1, bias=False), bn = BatchNorm2d(C) in eval mode, LeakyReLU(slope)
    -> (alpha (B,1,N,k) | None, max over k of leaky(bn(conv(alpha))) (B,C,N))."""
    if bn.training:
        raise RuntimeError("angle_conv1 folds BatchNorm's running statistics: the module must be in eval() mode")
    scale, shift = _fold_bn(bn)
    centroid = torch.mean(points, dim=-1)                                   # models/attn.py:65
    alpha, pooled = ops.edge_angle_max(points, centroid, idx, conv.weight.reshape(-1), scale, shift, slope, want_alpha)
    return (alpha.unsqueeze(1) if alpha is not None else None), pooled


def position_encoding_forward(self, points, k=5):
    """Drop-in body for ``PositionEncoding.forward`` (models/attn.py:58-77) in eval / no_grad mode: the distance branch is
    the module's own PyTorch layers; the angle branch is the kNN graph (K1) and one fused kernel up to the max over k."""
    centroid = torch.mean(points, dim=-1, keepdim=True)
    g_dis = torch.square(points - centroid).sum(dim=1, keepdim=True)
    dis_feature = self.conv_dis(g_dis)
    pts = points.transpose(-1, -2)
    idx = ops.knn_graph(pts, pts, k)[0]
    act = self.conv_ang1[2]
    _, pooled = angle_conv1(points, idx, self.conv_ang1[0], self.conv_ang1[1], slope=getattr(act, "negative_slope", 0.2))
    ang_feature = self.conv_ang2(pooled)
    return torch.cat([dis_feature, ang_feature], dim=1)
