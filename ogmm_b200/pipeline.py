"""The registration hot path as one call, and the pair-sharded multi-GPU plumbing.

``register_hot_path`` chains the four kernel stages exactly where ``GMMReg.forward`` has them
(models/gmmreg.py:52-53 kNN graph, :100-101 clustering, :102-103 soft SVD head).  The DGCNN
convolutions and the transformer overlap detector between them stay PyTorch in the reference and
are not part of this path: their outputs (point features, overlap scores) are inputs here.

Multi-GPU: pairs are independent, so ranks own contiguous blocks of pairs and never talk on the hot
path.  The only collective is one all-reduce of a small vector of metric
sums per evaluation (``reduce_metrics``).
"""
from __future__ import annotations

import torch

from . import ops

STAGES = ("knn_edge", "knn_wide", "cluster", "feat_moments", "procrustes")
DEEPGMR_STAGES = ("knn_edge", "softmax_em", "gmm_register")


def launches_per_step(iters=10):
    """Kernel launches of one ``register_hot_path`` call: 2 kNN, 2 clustering (redo rounds, when the early exit fires,
    are tail-launched from the device), 2 feature M-step, 1 head."""
    return 2 + 2 + 2 + 1


import os as _os

_side_streams = {}
_copy_streams = {}


def _copy_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=key)
    return _copy_streams[key]


_sm_counts = {}


def _split_schedule(src):
    """Four-stream schedule (graphs next to the clustering chains) or the two-stream chain per cloud.  OGMM_SCHEDULE =
    pair | split overrides for A/B runs."""
    mode = _os.environ.get("OGMM_SCHEDULE", "auto")
    if mode != "auto":
        return mode == "split"
    idx = src.device.index if src.device.index is not None else torch.cuda.current_device()
    if idx not in _sm_counts:
        _sm_counts[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return src.shape[-1] > 8192 or src.shape[0] < _sm_counts[idx]


def _side_stream(device, which=0):
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key[0])
    return _side_streams[key]


def _graph_chain(x, k, timers, wide=None):
    """kNN graph + edge features (and the feature-space graph of ``wide``) of one side on the current stream."""
    pts = x.transpose(-1, -2)
    with _Stage(timers, "knn_edge"):
        edge = ops.knn_graph(pts, pts, k, want_edge=True)[2].permute(0, 3, 1, 2)
    wide_idx = None
    if wide is not None:
        with _Stage(timers, "knn_wide"):
            wide_idx = ops.knn_graph(wide, wide, k)[0]
    return edge, wide_idx


def _cluster_chain(x, feats, o, n_clusters, iters, timers, feats_ready=None):
    """Clustering + feature M-step of one side on the current stream."""
    pts = x.transpose(-1, -2)
    with _Stage(timers, "cluster"):
        gam, pi, mu, _ = ops.sinkhorn_cluster(pts, o, n_clusters, iters=iters)
    if feats_ready is not None:
        torch.cuda.current_stream(x.device).wait_event(feats_ready)
    with _Stage(timers, "feat_moments"):
        nf = ops.gmm_moments(gam, feats.transpose(-1, -2))[1]
    return gam, pi, mu, nf


def _cloud_chain(x, feats, o, n_clusters, k, iters, timers, tag, feats_ready=None, wide=None):
    """kNN graph + edge features, clustering, feature M-step for one side (src or tgt) on the current stream.

    ``feats_ready``: optional CUDA event after which ``feats`` may be read (its host-to-device copy); only the
    feature M-step waits on it, the kNN graph and the clustering need xyz and the overlap scores alone.
    ``wide``: optional (B,N,C) wide point features whose feature-space kNN graph is built as well (the
    large-scale configuration, BASELINE.json configs[3]; tensor-core kernel for 32 <= C <= 256).
    """
    pts = x.transpose(-1, -2)
    with _Stage(timers, "knn_edge"):
        edge = ops.knn_graph(pts, pts, k, want_edge=True)[2].permute(0, 3, 1, 2)
    wide_idx = None
    if wide is not None:
        with _Stage(timers, "knn_wide"):
            wide_idx = ops.knn_graph(wide, wide, k)[0]
    with _Stage(timers, "cluster"):
        gam, pi, mu, _ = ops.sinkhorn_cluster(pts, o, n_clusters, iters=iters)
    if feats_ready is not None:
        torch.cuda.current_stream(x.device).wait_event(feats_ready)
    with _Stage(timers, "feat_moments"):
        nf = ops.gmm_moments(gam, feats.transpose(-1, -2))[1]
    return edge, gam, pi, mu, nf, wide_idx


@torch.no_grad()
def register_hot_path(src, tgt, src_feats, tgt_feats, src_o, tgt_o, n_clusters=16, k=20, iters=10, timers=None,
                      overlap=True, feats_ready=(None, None), wide=(None, None)):
    """src, tgt (B,3,N|M); *_feats (B,D,N|M); *_o (B,N|M)  ->  dict with rot (B,3,3), trans (B,3),
    edge_src/edge_tgt (B,6,N,k) views, and the GMM parameters of both clouds.

    The two clouds of a pair are independent until the registration head, so with ``overlap`` the target
    chain runs on a side stream next to the source chain (their kernels share the SMs; neither fills the
    GPU alone at moderate batch sizes).  ``timers``: optional dict stage -> list of (start, end) CUDA events
    recorded on the stream each stage runs on; pass ``overlap=False`` to time the kernels in isolation.
    ``feats_ready``: optional (src, tgt) CUDA events gating the first read of ``src_feats`` / ``tgt_feats``
    (``register_from_host`` copies them while the kNN graph and the clustering already run).
    """
    cur = torch.cuda.current_stream(src.device)
    split = overlap and _split_schedule(src)
    if split:
        # When a clustering launch cannot fill the GPU -- fewer clouds than SMs (one CTA per cloud), or large clouds (16
        # CTAs per cloud, most SMs idle while the kNN kernels would fill them) -- the kNN graphs, which need the points
        # only, run on streams of their own next to the two clustering chains (the critical path: clustering -> feature
        # M-step -> head), which are issued first.  Measured: 8.72 -> 7.03 ms per step at cfg 4, 0.247 -> 0.195 ms at
        # B = 1; no difference from B = 256 up, where the two-stream chain already keeps every SM at its two CTAs.
        s_t, g_s, g_t = (_side_stream(src.device, i) for i in range(3))
        for st in (s_t, g_s, g_t):
            st.wait_stream(cur)
        gam_s, pi_s, mu_s, nf_s = _cluster_chain(src, src_feats, src_o, n_clusters, iters, timers, feats_ready[0])
        with torch.cuda.stream(s_t):
            gam_t, pi_t, mu_t, nf_t = _cluster_chain(tgt, tgt_feats, tgt_o, n_clusters, iters, timers, feats_ready[1])
        with torch.cuda.stream(g_s):
            edge_s, wi_s = _graph_chain(src, k, timers, wide[0])
        with torch.cuda.stream(g_t):
            edge_t, wi_t = _graph_chain(tgt, k, timers, wide[1])
        for st in (s_t, g_s, g_t):
            cur.wait_stream(st)
        for t in (edge_s, wi_s, edge_t, wi_t, gam_t, pi_t, mu_t, nf_t):
            if t is not None:
                t.record_stream(cur)
    elif overlap:
        side = _side_stream(src.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            edge_t, gam_t, pi_t, mu_t, nf_t, wi_t = _cloud_chain(tgt, tgt_feats, tgt_o, n_clusters, k, iters, timers, "tgt",
                                                           feats_ready[1], wide[1])
        edge_s, gam_s, pi_s, mu_s, nf_s, wi_s = _cloud_chain(src, src_feats, src_o, n_clusters, k, iters, timers, "src",
                                                       feats_ready[0], wide[0])
        cur.wait_stream(side)
        for t in (edge_t, gam_t, pi_t, mu_t, nf_t, wi_t):
            if t is not None:
                t.record_stream(cur)
    else:
        edge_s, gam_s, pi_s, mu_s, nf_s, wi_s = _cloud_chain(src, src_feats, src_o, n_clusters, k, iters, timers, "src",
                                                       feats_ready[0], wide[0])
        edge_t, gam_t, pi_t, mu_t, nf_t, wi_t = _cloud_chain(tgt, tgt_feats, tgt_o, n_clusters, k, iters, timers, "tgt",
                                                       feats_ready[1], wide[1])
    with _Stage(timers, "procrustes"):
        rot, trans, corr, _ = ops.soft_procrustes(mu_s, mu_t, nf_s, nf_t, 0.05)
    extra = {} if wi_s is None else {"src_wide_idx": wi_s, "tgt_wide_idx": wi_t}
    return {**extra, "rot": rot, "trans": trans, "edge_src": edge_s, "edge_tgt": edge_t, "src_gamma": gam_s, "tgt_gamma": gam_t,
            "src_pi": pi_s, "tgt_pi": pi_t, "src_mu": mu_s, "tgt_mu": mu_t, "src_node_feats": nf_s,
            "tgt_node_feats": nf_t, "src_corr": corr}


class GraphedHotPath:
    """``register_hot_path`` over fixed device buffers, captured once into a CUDA graph and replayed.

    A step is nine kernel launches, two memsets and a fork / join over two streams -- about 0.45 ms of host work
    next to 1.4 ms of GPU work.  That margin is enough in a quiet process but not when the host hiccups (another
    rank's Python on the same socket, an ``nvidia-smi`` poll holding a driver lock): the GPU then waits for
    launches.  Replaying the captured step costs one launch and keeps the two-stream overlap exactly as captured.

    The inputs are STATIC: write new pairs into the tensors passed here (``copy_``) and call ``replay()``; the
    returned dict holds the same output tensors every time (clone what must outlive the next replay).
    """

    def __init__(self, src, tgt, src_feats, tgt_feats, src_o, tgt_o, n_clusters=16, k=20, iters=10, warmup=3, fn=None,
                 **kwargs):
        self.inputs = (src, tgt, src_feats, tgt_feats, src_o, tgt_o)
        self.args = (n_clusters, k, iters)
        fn = fn or (lambda: register_hot_path(*self.inputs, *self.args, **kwargs))
        dev = src.device
        cur = torch.cuda.current_stream(dev)
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):                 # lazy one-time setup (function attributes, driver entry points)
            for _ in range(max(int(warmup), 1)):
                fn()
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def replay(self):
        self.graph.replay()
        return self.out


@torch.no_grad()
def deepgmr_hot_path(src, tgt, src_logits, tgt_logits, k=20, timers=None, overlap=True):
    """DeepGMR's share of the path (baseline/deepgmr.py:64-79; BASELINE.json configs[2]): the DGCNN kNN graph + edge
    features of both clouds, the fused softmax E-step + M-step with sigma (:71-74), and gmm_register (:17-38).

    src, tgt (B,3,N|M); *_logits (B,J,N|M), the output of the PyTorch ``cluster`` CONV  ->  dict with transform
    (B,4,4), the edge tensors and the GMM parameters."""
    def side_chain(x, logits):
        pts = x.transpose(-1, -2)
        with _Stage(timers, "knn_edge"):
            edge = ops.knn_graph(pts, pts, k, want_edge=True)[2].permute(0, 3, 1, 2)
        with _Stage(timers, "softmax_em"):
            _, pi, mu, sigma = ops.softmax_moments(logits, x, want_gamma=False)
        return edge, pi, mu, sigma

    cur = torch.cuda.current_stream(src.device)
    if overlap and _split_schedule(src):
        # small batches / large clouds: the E+M kernels and the kNN graphs are independent, four streams (see register_hot_path)
        def em_chain(x, logits):
            with _Stage(timers, "softmax_em"):
                return ops.softmax_moments(logits, x, want_gamma=False)[1:]

        def graph_chain(x):
            pts = x.transpose(-1, -2)
            with _Stage(timers, "knn_edge"):
                return ops.knn_graph(pts, pts, k, want_edge=True)[2].permute(0, 3, 1, 2)

        s_t, g_s, g_t = (_side_stream(src.device, i) for i in range(3))
        for st in (s_t, g_s, g_t):
            st.wait_stream(cur)
        pi_s, mu_s, sg_s = em_chain(src, src_logits)
        with torch.cuda.stream(s_t):
            pi_t, mu_t, sg_t = em_chain(tgt, tgt_logits)
        with torch.cuda.stream(g_s):
            edge_s = graph_chain(src)
        with torch.cuda.stream(g_t):
            edge_t = graph_chain(tgt)
        for st in (s_t, g_s, g_t):
            cur.wait_stream(st)
        for t in (edge_s, edge_t, pi_t, mu_t, sg_t):
            t.record_stream(cur)
    elif overlap:
        side = _side_stream(src.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            edge_t, pi_t, mu_t, sg_t = side_chain(tgt, tgt_logits)
        edge_s, pi_s, mu_s, sg_s = side_chain(src, src_logits)
        cur.wait_stream(side)
        for t in (edge_t, pi_t, mu_t, sg_t):
            t.record_stream(cur)
    else:
        edge_s, pi_s, mu_s, sg_s = side_chain(src, src_logits)
        edge_t, pi_t, mu_t, sg_t = side_chain(tgt, tgt_logits)
    with _Stage(timers, "gmm_register"):
        tf = ops.gmm_register(pi_s, mu_s, mu_t, sg_t)
    return {"transform": tf, "rot": tf[:, :3, :3], "trans": tf[:, :3, 3], "edge_src": edge_s, "edge_tgt": edge_t,
            "src_pi": pi_s, "src_mu": mu_s, "src_sigma": sg_s, "tgt_pi": pi_t, "tgt_mu": mu_t, "tgt_sigma": sg_t}


class _Stage:
    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()          # on the current stream of the enclosing ``torch.cuda.stream`` context

    def __exit__(self, *exc):
        if self.timers is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self.timers.setdefault(self.name, []).append((self.start, end))
        return False


@torch.no_grad()
def register_from_host(host, device, n_clusters=16, k=20, iters=10, device_feats=None):
    """End-to-end call with HOST buffers: pinned tensors in, (rot, trans) back on the host.

    ``host`` maps src, tgt, src_feats, tgt_feats, src_o, tgt_o to pinned CPU tensors.  Returns
    (rot, trans) as CPU tensors plus the bytes moved in each direction.

    The small inputs (xyz, overlap scores: 8 MB at B=256) go first on the caller's stream; the two feature tensors
    (2 x 537 MB) follow on a copy stream, and only the feature M-step of each side waits for its tensor, so the kNN
    graph and the clustering of the WHOLE batch run under the copies.  The batch is not chunked: the Sinkhorn early
    exit is a mean over the batch (lib/utils.py:99-102) and chunking would change the iteration schedule.

    ``device_feats``: optional (src_feats, tgt_feats) already resident on ``device`` -- the model-boundary variant: in
    the real forward the 512-d point features are produced on the device by the PyTorch DGCNN + attention
    (models/gmmreg.py:52-97) and never exist on the host; only xyz and the overlap scores cross PCIe then.
    """
    names = ("src", "tgt", "src_o", "tgt_o") if device_feats is not None else ("src", "tgt", "src_feats", "tgt_feats", "src_o", "tgt_o")
    cur = torch.cuda.current_stream(device)
    dev = {n: host[n].to(device, non_blocking=True) for n in ("src", "tgt", "src_o", "tgt_o")}
    ready = (None, None)
    if device_feats is None:
        copy = _copy_stream(device)
        ready = []
        copy.wait_stream(cur)
        with torch.cuda.stream(copy):
            for n in ("src_feats", "tgt_feats"):
                dev[n] = host[n].to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                ready.append(ev)
        dev["src_feats"].record_stream(cur)
        dev["tgt_feats"].record_stream(_side_stream(device))
    else:
        dev["src_feats"], dev["tgt_feats"] = device_feats
    out = register_hot_path(dev["src"], dev["tgt"], dev["src_feats"], dev["tgt_feats"], dev["src_o"], dev["tgt_o"],
                            n_clusters, k, iters, feats_ready=tuple(ready))
    rot, trans = out["rot"].cpu(), out["trans"].cpu()
    h2d = sum(host[n].numel() * host[n].element_size() for n in names)
    d2h = rot.numel() * 4 + trans.numel() * 4
    return rot, trans, h2d, d2h


class HostBoundary:
    """The model-boundary call as a serving loop would run it: xyz and overlap scores arrive in pinned host buffers every
    step, the point features already live on the device (models/gmmreg.py:52-97 produces them there), (rot, trans) go back
    to the host.  The device side is one CUDA-graph launch over static buffers (``GraphedHotPath``); the host-to-device
    copies land in those buffers on the same stream, so a step costs four small copies, one graph launch and one
    read-back instead of seven eager launches plus allocator traffic.  Results are bit-identical to ``register_from_host``."""

    def __init__(self, host, device, src_feats, tgt_feats, n_clusters=16, k=20, iters=10):
        self.names = ("src", "tgt", "src_o", "tgt_o")
        self.static = {n: host[n].to(device) for n in self.names}
        self.graph = GraphedHotPath(self.static["src"], self.static["tgt"], src_feats, tgt_feats, self.static["src_o"],
                                    self.static["tgt_o"], n_clusters, k, iters)
        self.h2d = sum(host[n].numel() * host[n].element_size() for n in self.names)

    @torch.no_grad()
    def __call__(self, host):
        for n in self.names:
            self.static[n].copy_(host[n], non_blocking=True)
        out = self.graph.replay()
        rot, trans = out["rot"].cpu(), out["trans"].cpu()
        return rot, trans, self.h2d, rot.numel() * 4 + trans.numel() * 4


@torch.no_grad()
def deepgmr_from_host(host, device, k=20):
    """DeepGMR path end to end: pinned src, tgt (B,3,N) and src_logits, tgt_logits (B,J,N) in, the (B,4,4) transform
    back on the host, plus the bytes moved in each direction."""
    names = ("src", "tgt", "src_logits", "tgt_logits")
    dev = {n: host[n].to(device, non_blocking=True) for n in names}
    out = deepgmr_hot_path(dev["src"], dev["tgt"], dev["src_logits"], dev["tgt_logits"], k)
    tf = out["transform"].cpu()
    return tf, sum(host[n].numel() * host[n].element_size() for n in names), tf.numel() * 4


# ---- pair sharding -------------------------------------------------------------------------------------
def shard_range(total_pairs, rank, world_size):
    """Contiguous block of ceil(total/world) pairs per rank."""
    per = -(-total_pairs // world_size)
    lo = min(rank * per, total_pairs)
    return lo, min(lo + per, total_pairs)


METRIC_NAMES = ("err_r_deg_sum", "err_t_sum", "n_correct", "count")


def local_metrics(rot, trans, rot_gt, t_gt, r_th=1.0, t_th=0.1):
    """Per-shard metric sums (definitions: lib/metric.py:85-93 rotation_error / translation_error)."""
    c = torch.einsum('bij,bij->b', rot, rot_gt)
    err_r = torch.arccos(torch.clamp((c - 1) / 2, -1.0, 1.0)) * 180 / torch.pi
    err_t = torch.norm(trans - t_gt, dim=1)
    ok = ((err_r < r_th) & (err_t < t_th)).float()
    return torch.stack([err_r.sum(), err_t.sum(), ok.sum(), torch.tensor(float(rot.shape[0]), device=rot.device)])


def reduce_metrics(vec):
    """Sum the metric vector over ranks: the ONLY collective of the framework (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    out = dict(zip(METRIC_NAMES, vec.tolist()))
    n = max(out["count"], 1.0)
    return {"mean_err_r_deg": out["err_r_deg_sum"] / n, "mean_err_t": out["err_t_sum"] / n,
            "recall": out["n_correct"] / n, "count": out["count"]}
