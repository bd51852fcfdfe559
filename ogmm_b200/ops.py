"""Tensor-level wrappers over the C ABI: validate, allocate outputs, pass the current stream.

Every function takes CUDA float32 tensors (views welcome where the ABI takes strides) and returns
freshly allocated tensors on the same device, computed on the caller's current stream without
synchronising.  Anything else raises -- there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib


def _need_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise TypeError(f"{name}: ogmm_b200 kernels run on CUDA tensors only (got device {t.device}); "
                        "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32 like the reference, got {t.dtype}")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def knn_graph(src, dst, k, normalize=False, want_edge=False, want_dist=False):
    """src (B,N,C), dst (B,M,C) views -> idx (B,N,k) int64 [, dist (B,N,k)] [, edge (B,N,k,2C)]."""
    _need_cuda_f32("src", src); _need_cuda_f32("dst", dst)
    if src.dim() != 3 or dst.dim() != 3 or src.shape[0] != dst.shape[0] or src.shape[2] != dst.shape[2]:
        raise ValueError(f"knn_graph: incompatible shapes {tuple(src.shape)} and {tuple(dst.shape)}")
    B, N, C = src.shape
    M = dst.shape[1]
    k = int(k)
    idx = torch.empty((B, N, k), dtype=torch.int64, device=src.device)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=src.device) if want_dist else None
    edge = torch.empty((B, N, k, 2 * C), dtype=torch.float32, device=src.device) if want_edge else None
    with torch.cuda.device(src.device):
        st = _lib.load().ogmm_knn_graph(src.data_ptr(), *src.stride(), dst.data_ptr(), *dst.stride(),
                                        B, N, M, C, k, int(bool(normalize)), idx.data_ptr(), _ptr(dist), _ptr(edge),
                                        _stream(src))
    _lib.check(st, "ogmm_knn_graph")
    return idx, dist, edge


def square_distance(src, dst, normalize=False):
    """src (B,N,C), dst (B,M,C) views -> dist (B,N,M), the reference's expanded form with clamp 1e-12."""
    _need_cuda_f32("src", src); _need_cuda_f32("dst", dst)
    if src.dim() != 3 or dst.dim() != 3 or src.shape[0] != dst.shape[0] or src.shape[2] != dst.shape[2]:
        raise ValueError(f"square_distance: incompatible shapes {tuple(src.shape)} and {tuple(dst.shape)}")
    B, N, C = src.shape
    M = dst.shape[1]
    dist = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        st = _lib.load().ogmm_square_distance(src.data_ptr(), *src.stride(), dst.data_ptr(), *dst.stride(), B, N, M, C,
                                              int(bool(normalize)), dist.data_ptr(), _stream(src))
    _lib.check(st, "ogmm_square_distance")
    return dist


def knn_wide(src, dst, k, normalize=False, want_dist=False):
    """Tensor-core feature-space kNN (32 <= C <= 256): idx (B,N,k) int64 [, dist], fallback count (1,) int32."""
    _need_cuda_f32("src", src); _need_cuda_f32("dst", dst)
    B, N, C = src.shape
    M = dst.shape[1]
    k = int(k)
    idx = torch.empty((B, N, k), dtype=torch.int64, device=src.device)
    dist = torch.empty((B, N, k), dtype=torch.float32, device=src.device) if want_dist else None
    fallback = torch.zeros((1,), dtype=torch.int32, device=src.device)
    with torch.cuda.device(src.device):
        st = _lib.load().ogmm_knn_wide(src.data_ptr(), *src.stride(), dst.data_ptr(), *dst.stride(), B, N, M, C, k,
                                       int(bool(normalize)), idx.data_ptr(), _ptr(dist), fallback.data_ptr(), _stream(src))
    _lib.check(st, "ogmm_knn_wide")
    return idx, dist, fallback


def edge_gather(x, idx):
    """x (B,C,N) view, idx (B,N,k) int64 -> edge (B,N,k,2C) memory."""
    _need_cuda_f32("x", x)
    if idx.dtype != torch.int64 or not idx.is_cuda:
        raise TypeError("edge_gather: idx must be a CUDA int64 tensor")
    B, C, N = x.shape
    k = idx.shape[-1]
    idx = idx.reshape(B, N, k).contiguous()
    edge = torch.empty((B, N, k, 2 * C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.load().ogmm_edge_gather(x.data_ptr(), *x.stride(), idx.data_ptr(), B, C, N, k, edge.data_ptr(), _stream(x))
    _lib.check(st, "ogmm_edge_gather")
    return edge


def edge_conv_max(x, idx, weight, scale, shift, want_act=True):
    """x (B,3,N) view, idx (B,N,k) int64, weight (C,6), scale / shift (C) -> act (B,C,N,k) | None, pooled (B,C,N)."""
    _need_cuda_f32("x", x); _need_cuda_f32("weight", weight); _need_cuda_f32("scale", scale); _need_cuda_f32("shift", shift)
    if idx.dtype != torch.int64 or not idx.is_cuda:
        raise TypeError("edge_conv_max: idx must be a CUDA int64 tensor")
    B, three, N = x.shape
    k = idx.shape[-1]
    C = weight.shape[0]
    if three != 3 or tuple(idx.shape) != (B, N, k) or weight.numel() != C * 6 or scale.numel() != C or shift.numel() != C:
        raise ValueError("edge_conv_max: expected x (B,3,N), idx (B,N,k), weight (C,6), scale (C), shift (C)")
    idx, weight, scale, shift = idx.contiguous(), weight.reshape(C, 6).contiguous(), scale.contiguous(), shift.contiguous()
    act = torch.empty((B, C, N, k), dtype=torch.float32, device=x.device) if want_act else None
    pooled = torch.empty((B, C, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.load().ogmm_edge_conv_max(x.data_ptr(), *x.stride(), idx.data_ptr(), weight.data_ptr(), scale.data_ptr(),
                                            shift.data_ptr(), B, N, k, C, _ptr(act), pooled.data_ptr(), _stream(x))
    _lib.check(st, "ogmm_edge_conv_max")
    return act, pooled


def edge_angle_max(x, centroid, idx, weight, scale, shift, slope=0.2, want_alpha=False):
    """x (B,3,N) view, centroid (B,3), idx (B,N,k) int64, weight / scale / shift (C) -> alpha (B,N,k) | None, pooled (B,C,N):
    the angle feature of PositionEncoding (models/attn.py:65-73) through conv_ang1's 1x1 conv + BN + LeakyReLU + max."""
    _need_cuda_f32("x", x); _need_cuda_f32("centroid", centroid)
    _need_cuda_f32("weight", weight); _need_cuda_f32("scale", scale); _need_cuda_f32("shift", shift)
    if idx.dtype != torch.int64 or not idx.is_cuda:
        raise TypeError("edge_angle_max: idx must be a CUDA int64 tensor")
    B, three, N = x.shape
    k = idx.shape[-1]
    C = weight.numel()
    if three != 3 or tuple(idx.shape) != (B, N, k) or centroid.numel() != 3 * B or scale.numel() != C or shift.numel() != C:
        raise ValueError("edge_angle_max: expected x (B,3,N), centroid (B,3), idx (B,N,k), weight / scale / shift (C)")
    idx, centroid = idx.contiguous(), centroid.reshape(B, 3).contiguous()
    weight, scale, shift = weight.reshape(C).contiguous(), scale.contiguous(), shift.contiguous()
    alpha = torch.empty((B, N, k), dtype=torch.float32, device=x.device) if want_alpha else None
    pooled = torch.empty((B, C, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.load().ogmm_edge_angle_max(x.data_ptr(), *x.stride(), centroid.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                             scale.data_ptr(), shift.data_ptr(), float(slope), B, N, k, C, _ptr(alpha),
                                             pooled.data_ptr(), _stream(x))
    _lib.check(st, "ogmm_edge_angle_max")
    return alpha, pooled


def fps(xyz, npoint, start=None, want_points=False):
    """xyz (B,N,3) view -> ids (B,npoint) int64 [, points (B,npoint,3)].  start=None: is_center."""
    _need_cuda_f32("xyz", xyz)
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("fps: points must be 3-D")
    ids = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    pts = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if want_points else None
    if start is not None:
        start = start.to(device=xyz.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(xyz.device):
        st = _lib.load().ogmm_fps(xyz.data_ptr(), *xyz.stride(), B, N, int(npoint), _ptr(start), ids.data_ptr(),
                                  _ptr(pts), _stream(xyz))
    _lib.check(st, "ogmm_fps")
    return ids, pts


def sinkhorn_cluster(xyz, o_scores, n_clusters, iters=10, tau=1.0, epsilon=1e-2, thresh=1e-2, max_iter=10,
                     want_iters=False):
    """xyz (B,N,3) view, o_scores (B,N) -> gamma (B,N,J), pi (B,J), mu (B,J,3) [, inner iterations (iters) int32]."""
    _need_cuda_f32("xyz", xyz); _need_cuda_f32("o_scores", o_scores)
    B, N, C = xyz.shape
    if C != 3 or tuple(o_scores.shape) != (B, N):
        raise ValueError(f"sinkhorn_cluster: bad shapes {tuple(xyz.shape)} / {tuple(o_scores.shape)}")
    J = int(n_clusters)
    o = o_scores.contiguous()
    dev = xyz.device
    gamma = torch.empty((B, N, J), dtype=torch.float32, device=dev)
    pi = torch.empty((B, J), dtype=torch.float32, device=dev)
    mu = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
    run = torch.empty((int(iters),), dtype=torch.int32, device=dev) if want_iters else None
    lib = _lib.load()
    nbytes = lib.ogmm_sinkhorn_cluster_workspace(B, N, J, int(iters), int(max_iter))
    ws = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = lib.ogmm_sinkhorn_cluster(xyz.data_ptr(), *xyz.stride(), o.data_ptr(), B, N, J, int(iters), float(tau),
                                       float(epsilon), float(thresh), int(max_iter), gamma.data_ptr(), pi.data_ptr(),
                                       mu.data_ptr(), _ptr(run), ws.data_ptr(), ws.numel(), _stream(xyz))
    _lib.check(st, "ogmm_sinkhorn_cluster")
    return gamma, pi, mu, run


def sinkhorn(cost, p=None, q=None, epsilon=1e-2, thresh=1e-2, max_iter=100, want_iters=False):
    """cost (B,N,M) -> gamma (B,N,M), per-batch loss (B) [, iterations (1) int32]."""
    _need_cuda_f32("cost", cost)
    B, N, M = cost.shape
    cost = cost.contiguous()
    dev = cost.device
    if p is not None:
        p = _need_cuda_f32("p", p).reshape(B, N).contiguous()
    if q is not None:
        q = _need_cuda_f32("q", q).reshape(B, M).contiguous()
    gamma = torch.empty_like(cost)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    run = torch.empty((1,), dtype=torch.int32, device=dev) if want_iters else None
    lib = _lib.load()
    nbytes = lib.ogmm_sinkhorn_workspace(B, N, M, int(max_iter))
    ws = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = lib.ogmm_sinkhorn(cost.data_ptr(), _ptr(p), _ptr(q), B, N, M, float(epsilon), float(thresh), int(max_iter),
                               gamma.data_ptr(), loss.data_ptr(), _ptr(run), ws.data_ptr(), ws.numel(), _stream(cost))
    _lib.check(st, "ogmm_sinkhorn")
    return gamma, loss, run


def gmm_moments(gamma, pts, return_sigma=False):
    """gamma (B,N,J) view, pts (B,N,D) view -> pi (B,J), mu (B,J,D) [, sigma (B,J,D,D)]."""
    _need_cuda_f32("gamma", gamma); _need_cuda_f32("pts", pts)
    B, N, J = gamma.shape
    if pts.dim() != 3 or pts.shape[0] != B or pts.shape[1] != N:
        raise ValueError(f"gmm_moments: incompatible shapes {tuple(gamma.shape)} and {tuple(pts.shape)}")
    D = pts.shape[2]
    dev = gamma.device
    pi = torch.empty((B, J), dtype=torch.float32, device=dev)
    mu = torch.empty((B, J, D), dtype=torch.float32, device=dev)
    sigma = torch.empty((B, J, D, D), dtype=torch.float32, device=dev) if return_sigma else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        if D <= 4:
            st = lib.ogmm_gmm_moments(gamma.data_ptr(), *gamma.stride(), pts.data_ptr(), *pts.stride(), B, N, J, D,
                                      pi.data_ptr(), mu.data_ptr(), _ptr(sigma), _stream(gamma))
            what = "ogmm_gmm_moments"
        else:
            if return_sigma:
                raise NotImplementedError("gmm_moments: sigma is built for D <= 4 (the reference only uses it on xyz)")
            nbytes = int(lib.ogmm_gmm_moments_feat_workspace(B, N, J, D))       # > 0: few clouds, very many points
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev) if nbytes > 0 else None
            st = lib.ogmm_gmm_moments_feat_ws(gamma.data_ptr(), *gamma.stride(), pts.data_ptr(), *pts.stride(), B, N, J, D,
                                              pi.data_ptr(), mu.data_ptr(), _ptr(ws), nbytes, _stream(gamma))
            what = "ogmm_gmm_moments_feat"
    _lib.check(st, what)
    return (pi, mu, sigma) if return_sigma else (pi, mu)


def gmm_moments_feat_backward(gamma, grad_mu, pi, like):
    """gamma (B,N,J) view, grad_mu (B,J,D), pi (B,J) -> grad_feats with the shape AND memory layout of ``like`` (B,N,D):
    for the transposed view of a (B,D,N) tensor the gradient comes back as the same kind of view."""
    _need_cuda_f32("gamma", gamma); _need_cuda_f32("grad_mu", grad_mu); _need_cuda_f32("pi", pi)
    B, N, J = gamma.shape
    D = grad_mu.shape[2]
    if tuple(grad_mu.shape) != (B, J, D) or tuple(pi.shape) != (B, J) or tuple(like.shape) != (B, N, D):
        raise ValueError("gmm_moments_feat_backward: inconsistent shapes")
    grad_mu, pi = grad_mu.contiguous(), pi.contiguous()
    native = like.stride(1) == 1 and like.stride(2) >= N                  # a transposed view of (B,D,N)
    out = torch.empty((B, D, N), dtype=torch.float32, device=gamma.device).transpose(1, 2) if native else \
        torch.empty((B, N, D), dtype=torch.float32, device=gamma.device)
    with torch.cuda.device(gamma.device):
        st = _lib.load().ogmm_gmm_moments_feat_backward(gamma.data_ptr(), *gamma.stride(), grad_mu.data_ptr(), pi.data_ptr(),
                                                        B, N, J, D, out.data_ptr(), *out.stride(), _stream(gamma))
    _lib.check(st, "ogmm_gmm_moments_feat_backward")
    return out


def gmm_moments_backward(pts, pi, mu, sigma, grad_pi, grad_mu, grad_sigma, like):
    """dL/dgamma of the xyz moments ``gmm_moments(gamma, pts, return_sigma)`` (D = 3): pts (B,N,3) view, pi (B,J), mu (B,J,3),
    sigma (B,J,3,3) | None, upstream gradients (each may be None) -> grad_gamma with the shape AND strides of ``like``
    (B,N,J) (DeepGMR holds gamma as the transposed view of a (B,J,N) tensor)."""
    _need_cuda_f32("pts", pts); _need_cuda_f32("pi", pi); _need_cuda_f32("mu", mu)
    B, N, three = pts.shape
    J = pi.shape[1]
    if three != 3 or tuple(mu.shape) != (B, J, 3) or tuple(like.shape) != (B, N, J):
        raise ValueError("gmm_moments_backward: expected pts (B,N,3), pi (B,J), mu (B,J,3), gamma-like (B,N,J)")
    pi, mu = pi.contiguous(), mu.contiguous()
    sigma = sigma.contiguous() if sigma is not None else None
    gs = []
    for n_, g, shape in (("grad_pi", grad_pi, (B, J)), ("grad_mu", grad_mu, (B, J, 3)), ("grad_sigma", grad_sigma, (B, J, 3, 3))):
        if g is not None:
            _need_cuda_f32(n_, g)
            g = g.reshape(shape).contiguous()
        gs.append(g)
    transposed = like.stride(1) == 1 and like.stride(2) >= N               # a transposed view of (B,J,N)
    out = torch.empty((B, J, N), dtype=torch.float32, device=pts.device).transpose(1, 2) if transposed else \
        torch.empty((B, N, J), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        st = _lib.load().ogmm_gmm_moments_backward(pts.data_ptr(), *pts.stride(), pi.data_ptr(), mu.data_ptr(), _ptr(sigma),
                                                   _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), B, N, J, out.data_ptr(),
                                                   *out.stride(), _stream(pts))
    _lib.check(st, "ogmm_gmm_moments_backward")
    return out


def gmm_register_backward(pi_s, mu_s, mu_t, sigma_t, grad_transform):
    """Gradients of ``gmm_register``: grad_transform (B,4,4) -> grad_pi_s (B,J), grad_mu_s, grad_mu_t (B,J,3),
    grad_sigma_t (B,J,3,3)."""
    for n_, t_ in (("pi_s", pi_s), ("mu_s", mu_s), ("mu_t", mu_t), ("sigma_t", sigma_t), ("grad_transform", grad_transform)):
        _need_cuda_f32(n_, t_)
    B, J = pi_s.shape
    pi_s, mu_s, mu_t, sigma_t = (t_.contiguous() for t_ in (pi_s, mu_s, mu_t, sigma_t))
    grad_transform = grad_transform.reshape(B, 4, 4).contiguous()
    g_pi, g_ms, g_mt, g_sg = (torch.empty_like(t_) for t_ in (pi_s, mu_s, mu_t, sigma_t))
    with torch.cuda.device(pi_s.device):
        st = _lib.load().ogmm_gmm_register_backward(pi_s.data_ptr(), mu_s.data_ptr(), mu_t.data_ptr(), sigma_t.data_ptr(), B, J,
                                                    grad_transform.data_ptr(), g_pi.data_ptr(), g_ms.data_ptr(),
                                                    g_mt.data_ptr(), g_sg.data_ptr(), _stream(pi_s))
    _lib.check(st, "ogmm_gmm_register_backward")
    return g_pi, g_ms, g_mt, g_sg


def softmax_moments(logits, pts, want_gamma=True):
    """logits (B,J,N), pts (B,3,N) view -> gamma (B,J,N) | None, pi (B,J), mu (B,J,3), sigma (B,J,3,3)."""
    _need_cuda_f32("logits", logits); _need_cuda_f32("pts", pts)
    B, J, N = logits.shape
    if tuple(pts.shape) != (B, 3, N):
        raise ValueError(f"softmax_moments: pts must be (B,3,N), got {tuple(pts.shape)}")
    logits = logits.contiguous()
    dev = logits.device
    gamma = torch.empty_like(logits) if want_gamma else None
    pi = torch.empty((B, J), dtype=torch.float32, device=dev)
    mu = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
    sigma = torch.empty((B, J, 3, 3), dtype=torch.float32, device=dev)
    sb, sd, sn = pts.stride()
    with torch.cuda.device(dev):
        st = _lib.load().ogmm_softmax_moments(logits.data_ptr(), pts.data_ptr(), sb, sn, sd, B, N, J, _ptr(gamma),
                                              pi.data_ptr(), mu.data_ptr(), sigma.data_ptr(), _stream(logits))
    _lib.check(st, "ogmm_softmax_moments")
    return gamma, pi, mu, sigma


def rigid_transform(src, corr, weight):
    """src, corr (B,3,n) views, weight (B,1,n) view -> R (B,3,3), t (B,3)."""
    _need_cuda_f32("src", src); _need_cuda_f32("corr", corr); _need_cuda_f32("weight", weight)
    B, three, n = src.shape
    if three != 3 or tuple(corr.shape) != (B, 3, n) or tuple(weight.shape) != (B, 1, n):
        raise ValueError("rigid_transform: expected src/corr (B,3,n) and weight (B,1,n)")
    dev = src.device
    rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().ogmm_rigid_transform(src.data_ptr(), *src.stride(), corr.data_ptr(), *corr.stride(),
                                              weight.data_ptr(), weight.stride(0), weight.stride(2), B, n,
                                              rot.data_ptr(), t.data_ptr(), _stream(src))
    _lib.check(st, "ogmm_rigid_transform")
    return rot, t


def soft_procrustes(src_mu, tgt_mu, src_desc, tgt_desc, temperature=0.05, want_sim=False):
    """(B,Js,3), (B,Jt,3), (B,Js,D), (B,Jt,D) -> R (B,3,3), t (B,3), corr (B,3,Js) [, sim (B,Js,Jt)]."""
    for n_, t_ in (("src_mu", src_mu), ("tgt_mu", tgt_mu), ("src_desc", src_desc), ("tgt_desc", tgt_desc)):
        _need_cuda_f32(n_, t_)
    B, Js, _ = src_mu.shape
    Jt = tgt_mu.shape[1]
    D = src_desc.shape[2]
    if tuple(src_desc.shape) != (B, Js, D) or tuple(tgt_desc.shape) != (B, Jt, D) or tgt_mu.shape[2] != 3 or src_mu.shape[2] != 3:
        raise ValueError("soft_procrustes: inconsistent shapes")
    src_mu, tgt_mu, src_desc, tgt_desc = (t_.contiguous() for t_ in (src_mu, tgt_mu, src_desc, tgt_desc))
    dev = src_mu.device
    rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    corr = torch.empty((B, 3, Js), dtype=torch.float32, device=dev)
    sim = torch.empty((B, Js, Jt), dtype=torch.float32, device=dev) if want_sim else None
    with torch.cuda.device(dev):
        st = _lib.load().ogmm_soft_procrustes(src_mu.data_ptr(), tgt_mu.data_ptr(), src_desc.data_ptr(),
                                              tgt_desc.data_ptr(), B, Js, Jt, D, float(temperature), rot.data_ptr(),
                                              t.data_ptr(), corr.data_ptr(), _ptr(sim), _stream(src_mu))
    _lib.check(st, "ogmm_soft_procrustes")
    return rot, t, corr, sim


def rigid_transform_backward(src, corr, weight, grad_rot, grad_trans):
    """Gradients of ``rigid_transform``: src, corr (B,3,n) views, weight (B,1,n) view, grad_rot (B,3,3) | None,
    grad_trans (B,3) | None -> grad_src (B,3,n), grad_corr (B,3,n), grad_weight (B,1,n)."""
    _need_cuda_f32("src", src); _need_cuda_f32("corr", corr); _need_cuda_f32("weight", weight)
    B, three, n = src.shape
    if three != 3 or tuple(corr.shape) != (B, 3, n) or tuple(weight.shape) != (B, 1, n):
        raise ValueError("rigid_transform_backward: expected src/corr (B,3,n) and weight (B,1,n)")
    dev = src.device
    if grad_rot is not None:
        _need_cuda_f32("grad_rot", grad_rot)
        grad_rot = grad_rot.reshape(B, 3, 3).contiguous()
    if grad_trans is not None:
        _need_cuda_f32("grad_trans", grad_trans)
        grad_trans = grad_trans.reshape(B, 3).contiguous()
    g_src = torch.empty((B, 3, n), dtype=torch.float32, device=dev)
    g_corr = torch.empty((B, 3, n), dtype=torch.float32, device=dev)
    g_w = torch.empty((B, 1, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().ogmm_rigid_transform_backward(src.data_ptr(), *src.stride(), corr.data_ptr(), *corr.stride(),
                                                       weight.data_ptr(), weight.stride(0), weight.stride(2), B, n,
                                                       _ptr(grad_rot), _ptr(grad_trans), g_src.data_ptr(),
                                                       g_corr.data_ptr(), g_w.data_ptr(), _stream(src))
    _lib.check(st, "ogmm_rigid_transform_backward")
    return g_src, g_corr, g_w


def soft_procrustes_backward(src_mu, tgt_mu, src_desc, tgt_desc, grad_rot, grad_trans, grad_corr, temperature=0.05):
    """Gradients of ``soft_procrustes`` given dL/dR (B,3,3) | None, dL/dt (B,3) | None, dL/dcorr (B,3,Js) | None ->
    grad_src_mu (B,Js,3), grad_tgt_mu (B,Jt,3), grad_src_desc (B,Js,D), grad_tgt_desc (B,Jt,D)."""
    for n_, t_ in (("src_mu", src_mu), ("tgt_mu", tgt_mu), ("src_desc", src_desc), ("tgt_desc", tgt_desc)):
        _need_cuda_f32(n_, t_)
    B, Js, _ = src_mu.shape
    Jt = tgt_mu.shape[1]
    D = src_desc.shape[2]
    if tuple(src_desc.shape) != (B, Js, D) or tuple(tgt_desc.shape) != (B, Jt, D) or tgt_mu.shape[2] != 3 or src_mu.shape[2] != 3:
        raise ValueError("soft_procrustes_backward: inconsistent shapes")
    src_mu, tgt_mu, src_desc, tgt_desc = (t_.contiguous() for t_ in (src_mu, tgt_mu, src_desc, tgt_desc))
    grads = []
    for n_, g, shape in (("grad_rot", grad_rot, (B, 3, 3)), ("grad_trans", grad_trans, (B, 3)), ("grad_corr", grad_corr, (B, 3, Js))):
        if g is not None:
            _need_cuda_f32(n_, g)
            g = g.reshape(shape).contiguous()
        grads.append(g)
    dev = src_mu.device
    g_smu, g_tmu = torch.empty_like(src_mu), torch.empty_like(tgt_mu)
    g_sd, g_td = torch.empty_like(src_desc), torch.empty_like(tgt_desc)
    with torch.cuda.device(dev):
        st = _lib.load().ogmm_soft_procrustes_backward(src_mu.data_ptr(), tgt_mu.data_ptr(), src_desc.data_ptr(),
                                                       tgt_desc.data_ptr(), B, Js, Jt, D, float(temperature),
                                                       _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]), g_smu.data_ptr(),
                                                       g_tmu.data_ptr(), g_sd.data_ptr(), g_td.data_ptr(), _stream(src_mu))
    _lib.check(st, "ogmm_soft_procrustes_backward")
    return g_smu, g_tmu, g_sd, g_td


def cos_similarity(x, y):
    """x (B,N,D), y (B,M,D) -> (B,N,M) cosine similarity."""
    _need_cuda_f32("x", x); _need_cuda_f32("y", y)
    B, N, D = x.shape
    M = y.shape[1]
    x, y = x.contiguous(), y.contiguous()
    sim = torch.empty((B, N, M), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.load().ogmm_cos_similarity(x.data_ptr(), y.data_ptr(), B, N, M, D, sim.data_ptr(), _stream(x))
    _lib.check(st, "ogmm_cos_similarity")
    return sim


def gmm_register(pi_s, mu_s, mu_t, sigma_t):
    """(B,J), (B,J,3), (B,J,3), (B,J,3,3) -> T (B,4,4)."""
    for n_, t_ in (("pi_s", pi_s), ("mu_s", mu_s), ("mu_t", mu_t), ("sigma_t", sigma_t)):
        _need_cuda_f32(n_, t_)
    B, J = pi_s.shape
    pi_s, mu_s, mu_t, sigma_t = (t_.contiguous() for t_ in (pi_s, mu_s, mu_t, sigma_t))
    tf = torch.empty((B, 4, 4), dtype=torch.float32, device=pi_s.device)
    with torch.cuda.device(pi_s.device):
        st = _lib.load().ogmm_gmm_register(pi_s.data_ptr(), mu_s.data_ptr(), mu_t.data_ptr(), sigma_t.data_ptr(), B, J,
                                           tf.data_ptr(), _stream(pi_s))
    _lib.check(st, "ogmm_gmm_register")
    return tf
