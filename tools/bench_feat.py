"""Time the feature M-step (ogmm_gmm_moments_feat) alone: python tools/bench_feat.py B N D [reps].
Two alternating input sets larger than L2 together; CUDA events on the launching stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogmm_b200 import ops

B, N, D = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda:0")
sets = []
for i in range(2):
    g = torch.softmax(torch.randn(B, N, 16, device=dev), -1)
    f = torch.randn(B, D, N, device=dev)
    sets.append((g, f.transpose(-1, -2)))
for i in range(4):
    ops.gmm_moments(*sets[i & 1])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    ops.gmm_moments(*sets[i & 1])
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
ms = ts[len(ts) // 2]
byts = 4.0 * B * (N * 16 + N * D + 16 * D)
print(json.dumps({"B": B, "N": N, "D": D, "no_tma": os.environ.get("OGMM_FEAT_NO_TMA", ""), "ms": round(ms, 4),
                  "GBs": round(byts / ms / 1e6, 1)}))
