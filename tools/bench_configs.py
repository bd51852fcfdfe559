"""Throughput of the remaining BASELINE.json configurations on one GPU (run on the GPU box):

  cfg3  DeepGMR path: ICL-NUIM-shape pairs with density variation, 1024 pts, J=16 (kNN graph, softmax E-step +
        M-step with sigma, gmm_register)
  cfg4  large-scale pairs: 16384 pts, J=64, 64-d features kNN on the tensor cores
  cfg5  batch sweep of the flagship hot path, B = 1 .. 8192 pairs

    python tools/bench_configs.py [--cfg 3 4 5] > gpurun_out/configs.jsonl
"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ogmm_b200 as og
from ogmm_b200 import ops, pipeline, synth

dev = "cuda:0"


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def tile(t, count):
    reps = -(-count // t.shape[0])
    return t.repeat(reps, *([1] * (t.dim() - 1)))[:count].contiguous()


def cfg3(B=256, N=1024, J=16):
    src, tgt, _, _ = synth.icl_nuim_batch(0, 16, N)
    s, t = tile(torch.from_numpy(src).to(dev), B), tile(torch.from_numpy(tgt).to(dev), B)
    g = torch.Generator().manual_seed(3)
    ls = tile((torch.randn(16, J, N, generator=g) * 2).to(dev), B)
    lt = tile((torch.randn(16, J, N, generator=g) * 2).to(dev), B)

    def step():
        ps, pt = s.transpose(1, 2), t.transpose(1, 2)
        ops.knn_graph(ps, ps, 20, want_edge=True)
        ops.knn_graph(pt, pt, 20, want_edge=True)
        _, pi_s, mu_s, _ = ops.softmax_moments(ls, s)
        _, _, mu_t, sg_t = ops.softmax_moments(lt, t)
        return ops.gmm_register(pi_s, mu_s, mu_t, sg_t)
    ms = timed(step, 20)

    def em_only():
        ops.softmax_moments(ls, s)
        ops.softmax_moments(lt, t)
    em = timed(em_only, 50)
    nbytes = 4 * (J * N + 3 * N + 13 * J) * 2 * B
    return {"cfg": 3, "workload": "DeepGMR path, ICL-NUIM-shape pairs", "pairs": B, "n_points": N, "J": J, "ms_per_step": ms,
            "pairs_per_s": B / ms * 1e3, "softmax_em_ms": em, "softmax_em_gbs": nbytes / em / 1e6,
            "softmax_em_algorithmic_bytes": nbytes}


def cfg4(B=4, N=16384, J=64, C=64, D=512):
    g = torch.Generator().manual_seed(4)
    src, _, _, _ = synth.modelnet_batch(0, 2, N)
    xyz = tile(torch.from_numpy(src).to(dev), B)                                    # (B,3,N)
    wide = tile(torch.relu(torch.randn(2, N, C, generator=g)).to(dev), B)            # (B,N,C)
    feats = tile(torch.relu(torch.randn(2, D, N, generator=g)).to(dev), B)           # (B,D,N)
    o = tile(torch.sigmoid(torch.randn(2, N, generator=g)).to(dev), B)
    out = {"cfg": 4, "workload": "large-scale clouds", "clouds": B, "n_points": N, "J": J, "C": C}
    out["knn_xyz_ms"] = timed(lambda: ops.knn_graph(xyz.transpose(1, 2), xyz.transpose(1, 2), 20, want_edge=True), 3, 1)
    out["knn_wide_tensor_ms"] = timed(lambda: ops.knn_graph(wide, wide, 20), 3, 1)
    os.environ["OGMM_KNN_NO_TENSOR"] = "1"
    out["knn_wide_fp32_ms"] = timed(lambda: ops.knn_graph(wide, wide, 20), 2, 1)
    del os.environ["OGMM_KNN_NO_TENSOR"]
    out["gram_tflops_tensor"] = 2.0 * B * N * N * C / out["knn_wide_tensor_ms"] / 1e9
    res = {}

    def cl():
        res["g"] = ops.sinkhorn_cluster(xyz.transpose(1, 2), o, J)
    out["cluster_ms"] = timed(cl, 2, 1)
    gam = res["g"][0]
    out["feat_moments_ms"] = timed(lambda: ops.gmm_moments(gam, feats.transpose(1, 2)), 5, 1)
    out["feat_moments_gbs"] = 4.0 * B * (N * J + N * D + J * D) / out["feat_moments_ms"] / 1e6
    nf = ops.gmm_moments(gam, feats.transpose(1, 2))[1]
    mu = res["g"][2]
    out["procrustes_ms"] = timed(lambda: ops.soft_procrustes(mu, mu, nf, nf), 5, 1)
    total = 2 * (out["knn_xyz_ms"] + out["knn_wide_tensor_ms"] + out["cluster_ms"] + out["feat_moments_ms"]) + out["procrustes_ms"]
    out["pairs_per_s"] = B / total * 1e3
    return out


def cfg5(batches=(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192)):
    h = synth.hot_path_inputs(0, 16, 1024, 512)
    base = {k: torch.from_numpy(v).to(dev) for k, v in h.items()}
    rows = []
    for B in batches:
        d = {k: tile(v, B) for k, v in base.items() if k not in ("rot_gt", "t_gt")}
        reps = max(5, min(200, 8192 // B))
        # the step replayed from a CUDA graph (pipeline.GraphedHotPath): below ~64 pairs the eager call is bound by
        # its eleven host-side launches (~0.4 ms), not by the GPU
        g = pipeline.GraphedHotPath(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"], 16, 20, 10)
        ms = timed(g.replay, reps, warm=3)
        del g
        rows.append({"pairs": B, "ms_per_step": ms, "pairs_per_s": B / ms * 1e3})
        del d
        torch.cuda.empty_cache()
    return {"cfg": 5, "workload": "flagship hot path, batch sweep on 1 GPU, CUDA-graph replay", "sweep": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, nargs="+", default=[3, 4, 5])
    args = ap.parse_args()
    og._lib.load()
    for c in args.cfg:
        print(json.dumps({3: cfg3, 4: cfg4, 5: cfg5}[c]()), flush=True)
