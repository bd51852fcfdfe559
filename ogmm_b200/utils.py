"""Drop-in mirror of the reference's ``lib/utils.py`` hot-path functions (same names, signatures,
argument meaning, return shapes and dtypes), computed by the sm_100a kernels.

Each function cites the reference lines it replaces.  Inputs must be CUDA float32 tensors (views
are fine); results come back on the same device and stream.  Forward / inference only: outputs do
not carry autograd history ("Autograd"), and a call that autograd would have
to record (grad mode on, an input requiring grad) raises instead of returning gradient-less tensors
(``_guard.forward_only``; ``knn`` and ``farthest_point_sample`` return indices and are exempt).
"""
from __future__ import annotations

import threading
import weakref

import torch

from . import autograd, ops
from ._guard import forward_only

__all__ = ["square_distance", "knn", "get_graph_feature", "sinkhorn", "index_points", "gmm_params",
           "og_params", "farthest_point_sample", "cos_similarity", "get_local_corrs", "get_anchor_corrs",
           "wkeans_plus"]


@forward_only
def square_distance(src, dst, normalize=False):
    """lib/utils.py:12-34.  The dense (B,N,M) matrix, for callers that want it (metrics, losses).

    The hot path never materialises this matrix -- ``knn`` fuses it with the selection -- but the stand-alone
    function runs the same arithmetic in its own kernel (``ogmm_square_distance``)."""
    return ops.square_distance(src, dst, normalize)


@torch.no_grad()
def knn(src, tgt, k, normalize=False):
    """lib/utils.py:37-44.  src (B,N,C), tgt (B,M,C) -> int64 (B,N,k), ascending distance.

    Ties resolve to the lowest index (torch.topk's order is unspecified)."""
    return ops.knn_graph(src, tgt, k, normalize)[0]


knn.ogmm_autograd_safe = True            # int64 indices: nothing to differentiate


@forward_only
def get_graph_feature(x, k=20, idx=None, extra_dim=False):
    """lib/utils.py:47-66.  x (B,C,N) -> (B,2C,N,k) view over (B,N,k,2C) memory, [x_j - x_i ; x_i].

    ``idx=None`` runs the fused distance + top-k + gather kernel.  Unlike the reference (:57) a
    caller-supplied ``idx`` is NOT modified (no reference caller reads it afterwards)."""
    nb, c, n = x.shape
    if idx is None:
        if extra_dim:
            pts = x[:, 6:].transpose(-1, -2)
            edge = ops.edge_gather(x, ops.knn_graph(pts, pts, k)[0])
        else:
            pts = x.transpose(-1, -2)
            edge = ops.knn_graph(pts, pts, k, want_edge=True)[2]
    else:
        edge = ops.edge_gather(x, idx)
    return edge.permute(0, 3, 1, 2)


@forward_only
def sinkhorn(cost, p=None, q=None, epsilon=1e-2, thresh=1e-2, max_iter=100):
    """lib/utils.py:74-108.  Log-domain Sinkhorn -> (gamma (B,N,M), mean_b sum gamma*cost)."""
    gamma, loss, _ = ops.sinkhorn(cost, p, q, epsilon, thresh, max_iter)
    return gamma, loss.mean()


def index_points(points, idx):
    """lib/utils.py:111-127.  points (B,N,C), idx (B,S) -> (B,S,C)."""
    nb = points.shape[0]
    view = [nb] + [1] * (idx.dim() - 1)
    rep = [1] + list(idx.shape[1:])
    bidx = torch.arange(nb, dtype=torch.long, device=points.device).view(view).repeat(rep)
    return points[bidx, idx, :]


class SharedMoments:
    """One feature M-step per (gamma, feats), shared by its two callers.

    ``wkeans_plus`` ends with ``gmm_params(gamma, feats)`` (lib/utils.py:289) and ``CluLoss.forward`` immediately
    recomputes the same product from the same two tensors (lib/loss.py:114-115): 2 x 537 MB
    of feature reads per forward at B = 256.  The last wide (D > 4) result of each host thread is remembered and handed
    out again when the SAME data comes back: the same ``gamma`` tensor object, the same feature storage, offset, shape
    and strides (a new ``transpose`` view of the same tensor qualifies), unchanged autograd version counters (any
    in-place write bumps them), the same device and stream.  Both operands are held by weak reference only, so the
    cache never extends a tensor's life, and a freed-and-reallocated buffer cannot alias a live entry.
    """

    def __init__(self):
        self._tls = threading.local()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _signature(gamma, pts):
        return (gamma._version, tuple(gamma.shape), tuple(gamma.stride()), gamma.data_ptr(),
                pts._version, pts.storage_offset(), tuple(pts.shape), tuple(pts.stride()), str(pts.device),
                torch.cuda.current_stream(pts.device).cuda_stream if pts.is_cuda else 0)

    def get(self, gamma, pts, compute):
        entry = getattr(self._tls, "entry", None)
        sig = self._signature(gamma, pts)
        if entry is not None:
            g_ref, s_ref, old_sig, result = entry
            if g_ref() is gamma and s_ref() is pts.untyped_storage() and old_sig == sig:
                self.hits += 1
                return result
        result = compute(gamma, pts)
        self.misses += 1
        self._tls.entry = (weakref.ref(gamma), weakref.ref(pts.untyped_storage()), sig, result)
        return result

    def clear(self):
        self._tls.entry = None


shared_moments = SharedMoments()


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def gmm_params(gamma, pts, return_sigma=False):
    """lib/utils.py:130-149.  gamma (B,N,J), pts (B,N,D) -> pi (B,J), mu (B,J,D) [, sigma (B,J,D,D)].

    Wide features (D > 4, no sigma) go through ``shared_moments``: the call ``CluLoss`` makes right after
    ``wkeans_plus`` on the same tensors returns the result already computed (the returned tensors are shared: treat
    them as read-only, as both reference callers do).  The same wide call is differentiable with respect to ``pts``, and
    the xyz call (3-D points, with or without sigma) with respect to ``gamma`` (``ogmm_b200/autograd.py``); any other call
    that autograd would have to record is refused."""
    if _needs_grad(gamma, pts):
        if autograd.can_differentiate(gamma, pts, return_sigma):
            return autograd.feature_moments(gamma, pts)
        if autograd.can_differentiate_narrow(gamma, pts, return_sigma):          # DeepGMR: moments of xyz as a function of gamma
            return autograd.NarrowMoments.apply(gamma, pts, bool(return_sigma))
        raise RuntimeError("ogmm_b200.gmm_params is forward-only for this call (sigma, narrow points or a gamma that requires "
                           "grad): its kernels have no backward.  Call it under torch.no_grad() or keep the reference function.")
    with torch.no_grad():
        if not return_sigma and pts.dim() == 3 and pts.shape[-1] > 4:
            return shared_moments.get(gamma, pts, ops.gmm_moments)
        return ops.gmm_moments(gamma, pts, return_sigma)


gmm_params.ogmm_can_differentiate = lambda gamma, pts, return_sigma=False: (
    autograd.can_differentiate(gamma, pts, return_sigma) or autograd.can_differentiate_narrow(gamma, pts, return_sigma))


@forward_only
def og_params(pts, gamma, o_score=None, feature=None):
    """lib/utils.py:152-167.  Overlap-guided moments with the extra (J+1)-th non-overlap component."""
    if o_score is not None:
        score = torch.cat([gamma * o_score.unsqueeze(-1), (1.0 - o_score).unsqueeze(-1)], dim=-1)
    else:
        score = gamma
    pi, mu = ops.gmm_moments(score, pts)
    if feature is not None:
        return pi, mu, ops.gmm_moments(score, feature)[1]
    return pi, mu


@torch.no_grad()
def farthest_point_sample(xyz, npoint, is_center=False):
    """lib/utils.py:170-198.  xyz (B,N,3) -> int64 (B,npoint).

    ``is_center=False`` draws the start index with the reference's own call
    (``torch.randint(0, N, (B,), dtype=torch.long)`` on the host, :190) so a seeded run matches."""
    start = None if is_center else torch.randint(0, xyz.shape[1], (xyz.shape[0],), dtype=torch.long)
    return ops.fps(xyz, npoint, start)[0]


farthest_point_sample.ogmm_autograd_safe = True


@forward_only
def cos_similarity(x, y):
    """lib/utils.py:222-226.  (B,N,D), (B,M,D) -> (B,N,M)."""
    return ops.cos_similarity(x, y)


def get_local_corrs(xyz, xyz_mu, feats):
    """lib/utils.py:244-254.  Feature of the point nearest to each anchor: 1-NN through the kNN kernel.

    Autograd: the reference's graph -- the index comes from ``topk`` (no gradient), the features from ``torch.gather``,
    so only ``feats`` receives a gradient (a scatter of the anchor gradients); the gather stays a torch op here too."""
    with torch.no_grad():
        idx = ops.knn_graph(xyz_mu.detach(), xyz.detach(), 1)[0]                       # (B,S,1)
    return torch.gather(feats, dim=1, index=idx.repeat(1, 1, feats.size(-1)))


get_local_corrs.ogmm_autograd_safe = True


def get_anchor_corrs(xyz, feats, num_clusters, dst='eu', iters=10, is_fast=True):
    """lib/utils.py:257-266, ``is_fast=True`` branch (the only one the model takes).  xyz (B,3,N), feats (B,D,N).

    Autograd as in the reference: FPS indices carry no gradient; the gathered features (and anchors, should ``xyz``
    require grad) are index selections that torch differentiates."""
    if not is_fast:
        raise NotImplementedError("get_anchor_corrs(is_fast=False) has no live caller in the reference")
    xt, ft = xyz.transpose(-1, -2), feats.transpose(-1, -2)
    start = torch.randint(0, xt.shape[1], (xt.shape[0],), dtype=torch.long)
    with torch.no_grad():
        ids, xyz_mu = ops.fps(xt.detach(), num_clusters, start, want_points=True)
    if _needs_grad(xt):
        xyz_mu = index_points(xt, ids)
    feats_pos = index_points(ft, ids).transpose(-1, -2)
    feats_anchor = get_local_corrs(xt, xyz_mu, ft).transpose(-1, -2)
    return feats_anchor, feats_pos, xyz_mu.transpose(-1, -2)


get_anchor_corrs.ogmm_autograd_safe = True


def wkeans_plus(xyz, feats, o_scores, n_clusters, iters=10, tau=1.0):
    """lib/utils.py:269-291.  xyz (B,N,3), feats (B,N,D) (a view of (B,D,N) is read in place), o (B,N)
    -> gamma (B,N,J), pi (B,J), node_xyz (B,J,3), node_feats (B,J,D).

    Autograd: exactly the reference's graph.  There the whole loop runs under ``no_grad`` on a detached ``o_scores``
    and ``gamma`` is detached (:275-286), so gamma, pi and node_xyz carry no history and ``node_feats`` is
    differentiable with respect to ``feats`` only (:289) -- which the feature M-step's own backward provides."""
    with torch.no_grad():
        gamma, pi, node_xyz, _ = ops.sinkhorn_cluster(xyz.detach(), o_scores.detach(), n_clusters, iters=iters, tau=tau)
    if _needs_grad(feats):
        if feats.shape[-1] <= 4:
            raise RuntimeError("ogmm_b200.wkeans_plus: differentiable features need D > 4")
        return gamma, pi, node_xyz, autograd.feature_moments(gamma, feats)[1]
    with torch.no_grad():
        if feats.shape[-1] > 4:
            node_feats = shared_moments.get(gamma, feats, ops.gmm_moments)[1]      # CluLoss asks for it again (lib/loss.py:115)
        else:
            node_feats = ops.gmm_moments(gamma, feats)[1]
    return gamma, pi, node_xyz, node_feats


wkeans_plus.ogmm_autograd_safe = True
