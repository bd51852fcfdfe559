"""Time the stand-alone edge gather (get_graph_feature with a caller-supplied idx) and the fused kNN+edge call."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ogmm_b200 as og
from ogmm_b200 import ops

B, N, k = 256, 1024, 20
x = torch.rand(B, 3, N, device="cuda")
pts = x.transpose(1, 2)
idx = ops.knn_graph(pts, pts, k)[0]
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_g = timed(lambda: ops.edge_gather(x, idx))
t_k = timed(lambda: ops.knn_graph(pts, pts, k))
t_f = timed(lambda: ops.knn_graph(pts, pts, k, want_edge=True))
byts = B * N * k * (24 + 8) + B * N * 12
print(json.dumps({"edge_gather_ms": round(t_g, 4), "edge_gather_gbs": round(byts / t_g / 1e6, 1), "knn_only_ms": round(t_k, 4),
                  "knn_fused_edge_ms": round(t_f, 4), "algorithmic_bytes": byts}))
for C, NN in ((64, 1024),):
    xf = torch.rand(32, C, NN, device="cuda")
    idf = idx[:32]
    t = timed(lambda: ops.edge_gather(xf, idf))
    b2 = 32 * NN * k * (8 * C + 8)
    print(json.dumps({"C": C, "edge_gather_ms": round(t, 4), "gbs": round(b2 / t / 1e6, 1)}))
