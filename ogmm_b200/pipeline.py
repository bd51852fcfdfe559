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

STAGES = ("knn_edge", "cluster", "feat_moments", "procrustes")


def launches_per_step(iters=10):
    """Kernel launches of one ``register_hot_path`` call: 2 kNN, 2 x (1 + iters) clustering, 2 feature M-step, 1 head."""
    return 2 + 2 * (1 + iters) + 2 + 1


@torch.no_grad()
def register_hot_path(src, tgt, src_feats, tgt_feats, src_o, tgt_o, n_clusters=16, k=20, iters=10, timers=None):
    """src, tgt (B,3,N|M); *_feats (B,D,N|M); *_o (B,N|M)  ->  dict with rot (B,3,3), trans (B,3),
    edge_src/edge_tgt (B,6,N,k) views, and the GMM parameters of both clouds.

    ``timers``: optional dict stage -> list of (start_event, end_event) pairs, filled on the current stream.
    """
    def stage(name):
        return _Stage(timers, name)

    with stage("knn_edge"):
        ps, pt = src.transpose(-1, -2), tgt.transpose(-1, -2)
        edge_s = ops.knn_graph(ps, ps, k, want_edge=True)[2].permute(0, 3, 1, 2)
        edge_t = ops.knn_graph(pt, pt, k, want_edge=True)[2].permute(0, 3, 1, 2)
    with stage("cluster"):
        gam_s, pi_s, mu_s, _ = ops.sinkhorn_cluster(ps, src_o, n_clusters, iters=iters)
        gam_t, pi_t, mu_t, _ = ops.sinkhorn_cluster(pt, tgt_o, n_clusters, iters=iters)
    with stage("feat_moments"):
        nf_s = ops.gmm_moments(gam_s, src_feats.transpose(-1, -2))[1]
        nf_t = ops.gmm_moments(gam_t, tgt_feats.transpose(-1, -2))[1]
    with stage("procrustes"):
        rot, trans, corr, _ = ops.soft_procrustes(mu_s, mu_t, nf_s, nf_t, 0.05)
    return {"rot": rot, "trans": trans, "edge_src": edge_s, "edge_tgt": edge_t, "src_gamma": gam_s, "tgt_gamma": gam_t,
            "src_pi": pi_s, "tgt_pi": pi_t, "src_mu": mu_s, "tgt_mu": mu_t, "src_node_feats": nf_s,
            "tgt_node_feats": nf_t, "src_corr": corr}


class _Stage:
    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if self.timers is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self.timers.setdefault(self.name, []).append((self.start, end))
        return False


@torch.no_grad()
def register_from_host(host, device, n_clusters=16, k=20, iters=10):
    """End-to-end call with HOST buffers: pinned tensors in, (rot, trans) back on the host.

    ``host`` maps src, tgt, src_feats, tgt_feats, src_o, tgt_o to pinned CPU tensors.  Returns
    (rot, trans) as CPU tensors plus the bytes moved in each direction.
    """
    dev = {n: host[n].to(device, non_blocking=True) for n in ("src", "tgt", "src_feats", "tgt_feats", "src_o", "tgt_o")}
    out = register_hot_path(dev["src"], dev["tgt"], dev["src_feats"], dev["tgt_feats"], dev["src_o"], dev["tgt_o"],
                            n_clusters, k, iters)
    rot, trans = out["rot"].cpu(), out["trans"].cpu()
    h2d = sum(host[n].numel() * host[n].element_size() for n in dev)
    d2h = rot.numel() * 4 + trans.numel() * 4
    return rot, trans, h2d, d2h


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
