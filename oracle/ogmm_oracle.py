"""CPU oracle for the OGMM registration hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of every reference function that
SURVEY.md section 8(a) puts on the hot path.  It exists to CHECK the CUDA
product in ``ogmm_b200/`` -- it is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  Nothing under ``ogmm_b200/`` imports it, and the product
raises if its CUDA library is missing rather than falling back to this code.

Parity status: PINNED.  The reference ships no tests and no golden vectors
(SURVEY.md section 4), so the pin is the reference itself:
``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (with empty stub modules for the four absent third-party
imports), runs each hot-path function on seeded inputs, and stores
input/output pairs under ``tests/golden/``.  ``tests/test_oracle_golden.py``
requires every function below to reproduce those outputs bit-for-bit in fp32.

The reference is PyTorch, so the restatement is PyTorch as well: the same ATen
operators in the same order, which is what makes bit-equality on the CPU
possible (``torch.topk`` tie order, ``cdist``'s matmul form, ``logsumexp``'s
max-shift are all ATen behaviour, not ours to re-invent).  All functions are
dtype-generic: feeding float64 tensors gives the fp64 arbiter that SURVEY.md
section 8(c) asks for (the reference hard-codes float32 for Sinkhorn's default
marginals, ``lib/utils.py:79-83``; here they follow ``cost.dtype``, which is
identical in fp32).

Every function cites the reference lines it follows (paths are relative to the
reference checkout).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = [
    "pairwise_sqdist", "knn_indices", "edge_features", "sinkhorn_log",
    "gather_points", "gmm_moments", "overlap_moments", "fps_indices",
    "cosine_similarity", "nearest_anchor_feats", "anchor_corrs",
    "sinkhorn_kmeans", "rigid_from_corr", "soft_svd_head", "deepgmr_em",
    "deepgmr_register", "se3_decompose", "se3_integrate", "se3_inverse",
    "se3_concatenate", "se3_transform", "rotation_error_deg",
    "translation_error",
]


# --------------------------------------------------------------------------
# kNN graph + edge features
# --------------------------------------------------------------------------
def pairwise_sqdist(src, dst, normalize=False):
    """lib/utils.py:12-34 (square_distance).

    -2 * src.dst^T, then + |src|^2, then + |dst|^2, then clamp(min=1e-12) --
    the expanded form, in that order.  ``normalize`` returns 2 + (-2 s.d^T).
    """
    nb, n, _ = src.shape
    m = dst.shape[1]
    d = torch.matmul(src, dst.transpose(1, 2)) * -2
    if normalize:
        return d + 2.0
    d = d + (src * src).sum(-1).reshape(nb, n, 1)
    d = d + (dst * dst).sum(-1).reshape(nb, 1, m)
    return d.clamp(min=1e-12)


def knn_indices(src, tgt, k, normalize=False):
    """lib/utils.py:37-44 (knn): k smallest per row, ascending, int64."""
    d = pairwise_sqdist(src, tgt, normalize)
    return torch.topk(d, k, dim=-1, largest=False, sorted=True)[1]


def edge_features(x, k=20, idx=None, extra_dim=False):
    """lib/utils.py:47-66 (get_graph_feature).

    x (B,C,N) -> (B,2C,N,k) = [x_j - x_i ; x_i], returned as the permuted view
    of (B,N,k,2C) memory exactly like the reference.  The reference adds the
    batch offset into the caller's ``idx`` in place (:57); the oracle leaves
    the caller's tensor alone (no reference caller reads it afterwards).
    """
    nb, c, n = x.shape
    if idx is None:
        pts = x[:, 6:] if extra_dim else x
        idx = knn_indices(pts.transpose(1, 2), pts.transpose(1, 2), k)
    flat = (idx + torch.arange(nb, device=x.device).view(-1, 1, 1) * n).reshape(-1)
    xt = x.transpose(2, 1).contiguous()
    nbr = xt.reshape(nb * n, c)[flat].reshape(nb, n, k, c)
    ctr = xt.reshape(nb, n, 1, c).expand(nb, n, k, c)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------
# Sinkhorn E-step
# --------------------------------------------------------------------------
def sinkhorn_log(cost, p=None, q=None, epsilon=1e-2, thresh=1e-2, max_iter=100,
                 return_iters=False):
    """lib/utils.py:69-108 (log_boltzmann_kernel + sinkhorn).

    Log-domain Sinkhorn.  Row update over the last axis, column update over the
    middle axis, early exit when the BATCH MEAN of sum|du|+sum|dv| < thresh.
    Returns (exp(K), mean_b sum gamma*cost) [+ number of iterations run].
    """
    nb, nx, ny = cost.shape
    if p is None:
        p = torch.full((nb, nx), 1.0 / nx, dtype=cost.dtype, device=cost.device).squeeze()
    if q is None:
        q = torch.full((nb, ny), 1.0 / ny, dtype=cost.dtype, device=cost.device).squeeze()

    def kern(u, v):
        return (-cost + u.unsqueeze(-1) + v.unsqueeze(-2)) / epsilon

    u = torch.zeros_like(p)
    v = torch.zeros_like(q)
    iters = 0
    for _ in range(max_iter):
        u_prev, v_prev = u, v
        u = epsilon * (torch.log(p + 1e-8) - torch.logsumexp(kern(u, v), dim=-1)) + u
        v = epsilon * (torch.log(q + 1e-8)
                       - torch.logsumexp(kern(u, v).transpose(-2, -1), dim=-1)) + v
        iters += 1
        change = (u - u_prev).abs().sum(-1) + (v - v_prev).abs().sum(-1)
        if change.mean().item() < thresh:
            break
    gamma = torch.exp(kern(u, v))
    loss = (gamma * cost).sum(dim=(-2, -1)).mean()
    if return_iters:
        return gamma, loss, iters
    return gamma, loss


# --------------------------------------------------------------------------
# M-step
# --------------------------------------------------------------------------
def gather_points(points, idx):
    """lib/utils.py:111-127 (index_points): points (B,N,C), idx (B,S) -> (B,S,C)."""
    nb = points.shape[0]
    shape = [nb] + [1] * (idx.dim() - 1)
    rep = [1] + list(idx.shape[1:])
    bidx = torch.arange(nb, dtype=torch.long, device=points.device).view(shape).repeat(rep)
    return points[bidx, idx, :]


def gmm_moments(gamma, pts, return_sigma=False):
    """lib/utils.py:130-149 (gmm_params).

    pi = mean_n gamma; npi = pi*N + 1e-5; mu = gamma^T pts / npi; optional
    isotropic sigma = (sum_n gamma |x-mu|^2 / npi) * I  (NOT divided by D).
    """
    dim = pts.size(-1)
    pi = gamma.mean(dim=1)
    npi = pi * gamma.shape[1] + 1e-5
    mu = gamma.transpose(1, 2) @ pts / npi.unsqueeze(2)
    if not return_sigma:
        return pi, mu
    diff = pts.unsqueeze(2) - mu.unsqueeze(1)                       # (B,N,J,D)
    sq = (diff.unsqueeze(3) @ diff.unsqueeze(4)).squeeze()          # (B,N,J)
    eye = torch.eye(dim).unsqueeze(0).unsqueeze(1).to(gamma.device)
    sigma = ((sq * gamma).sum(dim=1) / npi).unsqueeze(2).unsqueeze(3) * eye
    return pi, mu, sigma


def overlap_moments(pts, gamma, o_score=None, feature=None):
    """lib/utils.py:152-167 (og_params): extra (J+1)-th non-overlap component."""
    if o_score is not None:
        score = torch.cat([torch.einsum('bnk,bn->bnk', gamma, o_score),
                           (1.0 - o_score).unsqueeze(-1)], dim=-1)
    else:
        score = gamma
    pi, mu = gmm_moments(score, pts)
    if feature is not None:
        return pi, mu, gmm_moments(score, feature)[1]
    return pi, mu


# --------------------------------------------------------------------------
# Farthest point sampling / anchors
# --------------------------------------------------------------------------
def fps_indices(xyz, npoint, is_center=False, start=None):
    """lib/utils.py:170-198 (farthest_point_sample).

    ``is_center=True`` starts from the point farthest from the centroid
    (deterministic).  Otherwise the reference draws ``torch.randint`` (:190);
    pass ``start`` (B,) to make that draw explicit, or leave None to draw it
    here with the same call.
    """
    nb, n, c = xyz.shape
    dev = xyz.device
    out = torch.zeros(nb, npoint, dtype=torch.long, device=dev)
    best = torch.ones(nb, n).to(xyz) * 1e10
    rows = torch.arange(nb, dtype=torch.long, device=dev)

    def relax(centre):
        d = ((xyz - centre) ** 2).sum(-1)
        closer = d < best
        best[closer] = d[closer]
        return best.max(-1)[1]

    if is_center:
        far = relax(xyz.mean(1).view(nb, 1, c))
    elif start is not None:
        far = start.to(dev).long()
    else:
        far = torch.randint(0, n, (nb,), dtype=torch.long).to(dev)
    for i in range(npoint):
        out[:, i] = far
        far = relax(xyz[rows, far, :].view(nb, 1, c))
    return out


def cosine_similarity(x, y):
    """lib/utils.py:222-226 (cos_similarity)."""
    return torch.einsum('bnd,bmd->bnm', F.normalize(x, dim=-1, p=2), F.normalize(y, dim=-1, p=2))


def nearest_anchor_feats(xyz, xyz_mu, feats):
    """lib/utils.py:244-254 (get_local_corrs): feature of the point nearest to each anchor."""
    idx = torch.topk(torch.cdist(xyz_mu, xyz), k=1, dim=2, largest=False)[1]
    idx = torch.nan_to_num(idx, nan=0)
    return torch.gather(feats, dim=1, index=idx.repeat(1, 1, feats.size(-1)))


def anchor_corrs(xyz, feats, num_clusters, start=None):
    """lib/utils.py:257-266 (get_anchor_corrs, is_fast=True branch): xyz (B,3,N), feats (B,D,N)."""
    xt, ft = xyz.transpose(-1, -2), feats.transpose(-1, -2)
    ids = fps_indices(xt, num_clusters, start=start)
    mu = gather_points(xt, ids)
    pos = gather_points(ft, ids).transpose(-1, -2)
    anc = nearest_anchor_feats(xt, mu, ft).transpose(-1, -2)
    return anc, pos, mu.transpose(-1, -2)


# --------------------------------------------------------------------------
# Overlap-guided Sinkhorn k-means (the E/M loop)
# --------------------------------------------------------------------------
def sinkhorn_kmeans(xyz, feats, o_scores, n_clusters, iters=10, tau=1.0, trace=None):
    """lib/utils.py:269-291 (wkeans_plus).

    FPS(is_center) init; ``iters`` x { cdist -> Sinkhorn(p=o/sum o, q=1/J,
    max_iter=10) -> nan_to_num -> row-normalise with clip 1e-3 -> M-step on
    xyz }; final M-step on feats.  ``trace`` (a list) receives the number of
    inner Sinkhorn iterations of each outer iteration (oracle-only extra).
    """
    nb, n, _ = xyz.shape
    node = gather_points(xyz, fps_indices(xyz, n_clusters, True))
    o = o_scores.detach()
    o = o / o.sum(dim=-1, keepdim=True).clip(min=1e-4)
    gamma, pi = torch.ones((nb, n, n_clusters), device=xyz.device, dtype=xyz.dtype), None
    with torch.no_grad():
        for _ in range(iters):
            cost = torch.cdist(xyz, node).clip(min=0.0) / tau
            g, _, it = sinkhorn_log(cost, p=o, q=None, max_iter=10, return_iters=True)
            if trace is not None:
                trace.append(it)
            g = torch.nan_to_num(g, nan=0.0).detach()
            gamma = g / g.sum(dim=-1, keepdim=True).clip(min=1e-3)
            pi, node = gmm_moments(gamma, xyz)
    node_feats = gmm_moments(gamma, feats)[1]
    return gamma, pi, node, node_feats


# --------------------------------------------------------------------------
# Weighted Procrustes and the two registration heads
# --------------------------------------------------------------------------
def rigid_from_corr(src, src_corr, weight):
    """lib/se3.py:256-289 (compute_rigid_transformation).

    src, src_corr (B,3,n), weight (B,1,n) -> R (B,3,3), t (B,3,1).
    cov = (src_c * w) corr_c^T; nan_to_num + 1e-5 I; SVD; R = V U^T with the
    third column of V flipped when det <= 0; t = -R c_s + c_t.
    """
    wsum = weight.sum(dim=2, keepdim=True)
    c_s = (src * weight).sum(dim=2, keepdim=True) / wsum
    c_t = (src_corr * weight).sum(dim=2, keepdim=True) / wsum
    nb, dim, _ = src.shape
    cov = torch.matmul((src - c_s) * weight, (src_corr - c_t).transpose(2, 1).contiguous())
    eye = torch.eye(dim).to(src).repeat(nb, 1, 1)
    try:
        cov = torch.nan_to_num(cov, nan=0.0) + 1e-5 * eye
        u, _, v = torch.svd(cov.cpu(), some=True, compute_uv=True)
    except Exception:  # the reference falls back to an identity covariance (:277-279)
        u, _, v = torch.svd(eye.cpu(), some=True, compute_uv=True)
    u, v = u.to(src), v.to(src_corr)
    r_pos = v @ u.transpose(-1, -2)
    v_flip = v.clone()
    v_flip[:, :, 2] *= -1
    r_neg = v_flip @ u.transpose(-1, -2)
    rot = torch.where(torch.det(r_pos)[:, None, None] > 0, r_pos, r_neg)
    t = torch.matmul(-rot, c_s.mean(dim=2, keepdim=True)) + c_t.mean(dim=2, keepdim=True)
    return rot, t


def soft_svd_head(src, tgt, src_desc, tgt_desc, src_pi=None, tgt_pi=None, is_sk=False):
    """models/dgcnn.py:96-115 (GMMSVD.forward).

    src/tgt (B,J,3) component means, *_desc (B,J,D).  is_sk=False (as built at
    models/gmmreg.py:41): scores = softmax(cos_sim / 0.05, dim=2).  is_sk=True:
    Sinkhorn(2(1-sim), p=src_pi, q=tgt_pi, max_iter=30), nan_to_num(nan=1e-4),
    row-normalise with clip 1e-4.
    Returns R (B,3,3), t (B,3), src_corr (B,3,J), tgt^T (B,3,J).
    """
    nb = src.size(0)
    sim = cosine_similarity(src_desc, tgt_desc)
    if is_sk:
        s = sinkhorn_log(2.0 * (1.0 - sim), p=src_pi, q=tgt_pi,
                         epsilon=1e-2, thresh=1e-2, max_iter=30)[0]
        s = torch.nan_to_num(s, 1e-4)
        scores = s / torch.sum(s, dim=-1, keepdim=True).clip(min=1e-4)
    else:
        scores = torch.softmax(sim / 0.05, dim=2)
    corr = torch.einsum('bmd,bnm->bdn', tgt, scores)
    w = scores.sum(dim=-1).unsqueeze(1)
    rot, t = rigid_from_corr(src.transpose(-1, -2), corr, w)
    return rot, t.view(nb, 3), corr, tgt.transpose(-1, -2)


def deepgmr_em(logits, pts):
    """baseline/deepgmr.py:71-74: gamma = softmax over J of logits (B,J,N); M-step with sigma.

    pts (B,3,N).  Returns gamma (B,J,N), pi (B,J), mu (B,J,3), sigma (B,J,3,3).
    """
    gamma = F.softmax(logits, dim=1)
    pi, mu, sigma = gmm_moments(gamma.transpose(-1, -2), pts.transpose(-1, -2), True)
    return gamma, pi, mu, sigma


def deepgmr_register(pi_s, mu_s, mu_t, sigma_t):
    """baseline/deepgmr.py:17-38 (gmm_register), without its hard-coded ``.cuda()`` (:30-31).

    c_s = pi_s mu_s, c_t = pi_s mu_t (pi_s for BOTH); M = sum_j pi_j (mu_s-c_s)(mu_t-c_t)^T Sigma_t^-1;
    SVD(nan_to_num(M) + 1e-4 on all nine entries); R = V diag(1,1,det(V U^T)) U^T;
    t = c_t - R c_s; returns the 4x4 transform.
    """
    c_s = pi_s.unsqueeze(1) @ mu_s
    c_t = pi_s.unsqueeze(1) @ mu_t
    m = torch.sum((pi_s.unsqueeze(2) * (mu_s - c_s)).unsqueeze(3)
                  @ (mu_t - c_t).unsqueeze(2) @ sigma_t.inverse(), dim=1)
    u, _, v = torch.svd(torch.nan_to_num(m, nan=0).cpu() + 1e-4)
    u, v = u.to(m.device), v.to(m.device)
    s = torch.eye(3, dtype=m.dtype).unsqueeze(0).repeat(u.shape[0], 1, 1).to(u.device)
    s[:, 2, 2] = torch.det(v @ u.transpose(1, 2))
    rot = v @ s @ u.transpose(1, 2)
    t = c_t.transpose(1, 2) - rot @ c_s.transpose(1, 2)
    bottom = torch.tensor([[[0, 0, 0, 1]]], dtype=m.dtype).repeat(rot.shape[0], 1, 1).to(rot.device)
    return torch.cat([torch.cat([rot, t], dim=2), bottom], dim=1)


# --------------------------------------------------------------------------
# SE(3) helpers and the evaluation metrics that the multi-GPU runner reduces
# --------------------------------------------------------------------------
def se3_decompose(trans):
    """lib/se3.py:14-26 (decompose_trans)."""
    if trans.dim() == 3:
        return trans[:, :3, :3], trans[:, :3, 3:4]
    return trans[:3, :3], trans[:3, 3:4]


def se3_integrate(rot, t):
    """lib/se3.py:29-52 (integrate_trans), torch branch."""
    if rot.dim() == 3:
        out = torch.eye(4)[None].repeat(rot.shape[0], 1, 1).to(rot.device)
        out[:, :3, :3] = rot
        out[:, :3, 3:4] = t.view([-1, 3, 1])
    else:
        out = torch.eye(4).to(rot.device)
        out[:3, :3] = rot
        out[:3, 3:4] = t
    return out


def se3_inverse(g):
    """lib/se3.py:59-71 (torch_inverse)."""
    rot = g[..., 0:3, 0:3]
    t = g[..., 0:3, 3]
    rt = rot.transpose(-1, -2)
    return torch.cat([rt, rt @ -t[..., None]], dim=-1)


def se3_concatenate(a, b):
    """lib/se3.py:74-93 (torch_concatenate): a @ b on 3x4 transforms."""
    r1, t1 = a[..., :3, :3], a[..., :3, 3]
    r2, t2 = b[..., :3, :3], b[..., :3, 3]
    return torch.cat([r1 @ r2, r1 @ t2[..., None] + t1[..., None]], dim=-1)


def se3_transform(g, a, normals=None):
    """lib/se3.py:96-117 (torch_transform)."""
    rot = g[..., :3, :3]
    p = g[..., :3, 3]
    if g.dim() != a.dim():
        raise NotImplementedError
    b = torch.matmul(a, rot.transpose(-1, -2)) + p[..., None, :]
    if normals is not None:
        return b, normals @ rot.transpose(-1, -2)
    return b


def rotation_error_deg(r1, r2):
    """lib/metric.py:85-88 (rotation_error)."""
    c = torch.einsum('bij,bij->b', r1, r2)
    return torch.arccos(torch.clamp((c - 1) / 2, -1.0, 1.0)) * 180 / torch.pi


def translation_error(t1, t2):
    """lib/metric.py:91-93 (translation_error)."""
    return torch.norm(t1 - t2, dim=1)
