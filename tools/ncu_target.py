"""Tiny launch targets for ncu captures (one stage of the hot path on the bench workload, few launches).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 2 -o gpurun_out/prof python tools/ncu_target.py knn
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogmm_b200 import ops, synth, _lib

what = sys.argv[1] if len(sys.argv) > 1 else "knn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
_lib.load()
h = synth.hot_path_inputs(0, 256, 1024, 512 if what in ("feat", "step") else 8, tile=32)
d = {k: torch.from_numpy(v).cuda() for k, v in h.items()}
pts = d["src"].transpose(1, 2)
for _ in range(reps):
    if what == "knn":
        ops.knn_graph(pts, pts, 20, want_edge=True)
    elif what == "cluster":
        ops.sinkhorn_cluster(pts, d["src_o"], 16)
    elif what == "feat":
        gam = ops.sinkhorn_cluster(pts, d["src_o"], 16)[0]
        ops.gmm_moments(gam, d["src_feats"].transpose(1, 2))
    elif what == "step":
        from ogmm_b200 import pipeline
        pipeline.register_hot_path(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"], 16, 20, overlap=False)
torch.cuda.synchronize()
if what == "wide":
    g = torch.Generator().manual_seed(64)
    xw = torch.relu(torch.randn(1, 16384, 64, generator=g)).cuda()
    for _ in range(reps):
        ops.knn_graph(xw, xw, 20)
    torch.cuda.synchronize()
if what in ("tiles", "dsmem"):
    big, _, _, _ = synth.modelnet_batch(0, 4, 16384)
    xb = torch.from_numpy(big).cuda().transpose(1, 2)                  # (4,16384,3) view, the cfg 4 clouds
    ob = torch.sigmoid(torch.randn(4, 16384, generator=torch.Generator().manual_seed(1))).cuda()
    for _ in range(reps):
        if what == "tiles":
            ops.knn_graph(xb, xb, 20, want_edge=True)
        else:
            ops.sinkhorn_cluster(xb, ob, 64)
    torch.cuda.synchronize()
