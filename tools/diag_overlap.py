"""Diagnostic: overlapped hot path with per-stage CUDA events and host enqueue times."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogmm_b200 import pipeline, synth

dev = torch.device("cuda:0")
B = 256
h = synth.hot_path_inputs(0, B, 1024, 512, tile=8)
d = {k: torch.from_numpy(v).to(dev) for k, v in h.items()}
def step(t=None, ov=True):
    return pipeline.register_hot_path(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"], 16, 20, 10, t, ov)
for _ in range(5): step()
torch.cuda.synchronize()
for rep in range(4):
    timers = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record()
    t00 = time.perf_counter()
    for _ in range(50):
        t0 = time.perf_counter(); step(timers); host.append(time.perf_counter() - t0)
    e1.record()
    t_enq = time.perf_counter() - t00
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    st = {k: round(sum(a.elapsed_time(b) for a, b in v) / 50, 3) for k, v in timers.items()}
    print(json.dumps({"rep": rep, "gpu_ms_per_step": round(ms, 3), "host_enqueue_ms_per_step": round(1e3 * t_enq / 50, 3),
                      "host_max_ms": round(1e3 * max(host), 3), "stage_ms_sum_per_step": st}))
