#!/usr/bin/env python
"""Benchmark of the OGMM registration hot path (BASELINE.json metric: registration pairs/sec, 1024-pt, J=16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs B]

A step = one pass of the hot path (kNN graph + edge features for both clouds, overlap-guided
Sinkhorn clustering of both clouds, feature M-step, soft-correspondence Procrustes) over one batch
of B synthetic ModelNet40-shape partial-overlap pairs per GPU (BASELINE.json configs[1]: N=1024,
J=16, D=512, k=20, B=256).  One process per GPU; ranks own disjoint pairs and never communicate
inside the timed region (weak scaling).  Rank 0 prints ONE JSON line.

  value     pairs/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       same step through register_from_host: pinned host buffers in, (R, t) back on the host
  roofline  the dominant kernel's algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json
  kernels   the same numbers for every stage
  cpu_baseline  the oracle port (the reference's own PyTorch op sequence) on this box's host cores

``--impl reference`` times that CPU path as its own arm (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POINTS, N_CLUSTERS, EMB, KNN = 1024, 16, 512, 20
ITERS = 10


def algorithmic_bytes(n=N_POINTS, j=N_CLUSTERS, d=EMB, k=KNN, c=3):
    """Per CLOUD (a pair is two clouds), fp32 + int64 idx; SURVEY.md section 8(d), restated in DESIGN.md."""
    return {
        "knn_edge": 4 * c * n + 8 * n * k + 4 * 2 * c * n * k,            # xyz in, idx + edge out
        "cluster": 4 * (3 * n + n + n * j + 4 * j),                        # xyz + o in, gamma + pi + mu out
        "feat_moments": 4 * (n * j + n * d + j * d),                       # gamma re-read + feats in, node_feats out
        "procrustes": 4 * (2 * 3 * j + 2 * j * d) // 2 + 24,               # per cloud share of the per-pair bytes
        "em_step": 4 * (3 * n + n + n * j + n * d + j * d + 4 * j),       # E/M-step as one unit (gamma not re-read)
    }


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    """One ``nvidia-smi -lms`` child for the whole run.  It is started BEFORE the warm-up and the timed region only
    begins once it has delivered a sample: nvidia-smi's own start-up (NVML attach, a cold binary on a fresh box)
    holds driver locks for up to a few hundred ms and stalls kernel launches -- started right at the timed region it
    made the measured step time swing between 1.4 and 6.3 ms.  Only samples taken inside the timed region are kept.
    """

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None
        self.lo, self.hi = 0, None

    def wait_ready(self, timeout=10.0):
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def begin(self):
        self.lo = len(self.rows)

    def end(self):
        self.hi = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[self.lo:self.hi] or self.rows[-3:]     # a region shorter than the period: the closest samples
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic():
    """DRAM bytes per launch from the committed ncu --set full captures (profiles/traffic.json), or {}."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return {k: v.get("dram_bytes_per_launch") for k, v in json.load(f).items() if isinstance(v, dict)}
    return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_hot_path(orc, torch, t, n_clusters=N_CLUSTERS, k=KNN):
    """The reference's own op sequence for the hot path, on the CPU (oracle port)."""
    outs = {}
    for side in ("src", "tgt"):
        x = t[side]
        idx = orc.knn_indices(x.transpose(-1, -2), x.transpose(-1, -2), k)
        outs[side + "_edge"] = orc.edge_features(x, k, idx)
        outs[side] = orc.sinkhorn_kmeans(x.transpose(-1, -2), t[side + "_feats"].transpose(-1, -2), t[side + "_o"], n_clusters)
    _, pi_s, mu_s, nf_s = outs["src"]
    _, pi_t, mu_t, nf_t = outs["tgt"]
    return orc.soft_svd_head(mu_s, mu_t, nf_s, nf_t, pi_s, pi_t)[:2]


def time_cpu(pairs, steps, warmup):
    import torch
    from oracle import ogmm_oracle as orc
    from ogmm_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h = synth.hot_path_inputs(0, pairs, N_POINTS, EMB, tile=min(pairs, 8))
    t = {k: torch.from_numpy(v) for k, v in h.items()}
    with torch.no_grad():
        for _ in range(warmup):
            cpu_hot_path(orc, torch, t)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            cpu_hot_path(orc, torch, t)
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pairs = args.cpu_pairs or min(256, max(4, int(120 * 90 / (args.steps + args.warmup))))
    times, cores = time_cpu(pairs, args.steps, args.warmup)
    total = sum(times)
    value = pairs * len(times) / total
    line = {
        "impl": "reference", "metric": "registration pairs/sec (1024-pt, J=16)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(pairs, "host CPU; each step is a bounded sample of the workload"),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{pairs} pairs per step x {len(times)} steps (same synthetic pairs, hot path only)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(pairs, note=None):
    cfg = {"workload": "OGMM registration hot path: kNN(k=20)+edge features, overlap-guided Sinkhorn clustering "
                       "(10x10, J=16), feature M-step (D=512), soft-correspondence Procrustes; "
                       "ModelNet40-shape partial-overlap pairs, 1024 pts (BASELINE.json configs[1])",
           "pairs_per_gpu_per_step": pairs, "n_points": N_POINTS, "n_clusters": N_CLUSTERS, "emb_dims": EMB, "k": KNN,
           "sinkhorn": "10 outer x 10 inner, eps=1e-2", "parallelism": "pair-sharded, no hot-path collective",
           "l2_policy": "inputs larger than L2 (feature tensors are 1 GiB per step at 256 pairs; 126 MB L2)"}
    if note:
        cfg["note"] = note
    return cfg


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ogmm_b200 as og
    from ogmm_b200 import pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; ogmm_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the one JSON line only: NCCL prints its version banner with a bare printf when the communicator
        # is created, so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    og._lib.load()
    B = args.pairs

    # ---- synthetic inputs: this rank's shard of the global pair list ---------------------------------------
    h = synth.hot_path_inputs(rank * B, B, N_POINTS, EMB, tile=args.distinct)
    host = {k: torch.from_numpy(v) for k, v in h.items()}
    d = {k: v.to(dev) for k, v in host.items()}

    def eager_step(timers=None, overlap=not args.no_overlap):
        return pipeline.register_hot_path(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"],
                                          N_CLUSTERS, KNN, ITERS, timers, overlap)

    # The timed step replays the two-stream step captured into a CUDA graph over the resident inputs (one launch per
    # step, immune to host jitter); --no-graph / --no-overlap time the eager call instead.
    graphed, graph_note = None, None
    if not (args.no_graph or args.no_overlap):
        try:
            graphed = pipeline.GraphedHotPath(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"],
                                              N_CLUSTERS, KNN, ITERS)
        except Exception as e:                       # the number must still come out: time the eager two-stream call
            graph_note = f"graph capture failed ({type(e).__name__}: {e}); eager launches timed instead"
            print("bench.py: " + graph_note, file=sys.stderr)
            torch.cuda.synchronize()

    def step(timers=None, overlap=not args.no_overlap):
        if graphed is not None and timers is None and overlap:
            return graphed.replay()
        return eager_step(timers, overlap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = step()
    barrier()
    if rank == 0:
        sampler.wait_ready()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None          # samples taken inside the timed region; the poller is gone
    elapsed_ms = e0.elapsed_time(e1)                        # before the per-stage pass below

    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = B * world * args.steps / (elapsed_ms * 1e-3)

    # ---- per-stage times: a serial pass (one stream, src then tgt) of the same steps, so every stage's
    # CUDA-event time is that kernel alone (in the timed region above the two chains share the SMs) ---------
    for _ in range(3):                               # warm the allocator's main-stream pool for the serial order
        step(None, overlap=False)
    timers = {}
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        out = step(timers, overlap=False)
    s1.record()
    barrier()
    serial_ms = s0.elapsed_time(s1) / args.steps
    stage_ms = {s: sum(a.elapsed_time(b) for a, b in ev) / args.steps for s, ev in timers.items()}
    ab = algorithmic_bytes()
    peak, peak_src = measured_peaks()
    traffic = measured_traffic()
    launches = {"knn_edge": 2, "cluster": 4, "feat_moments": 2, "procrustes": 1}
    kernels = {}
    for s, ms in stage_ms.items():
        nbytes = ab[s] * 2 * B                       # two clouds per pair
        gbs = nbytes / (ms * 1e-3) / 1e9
        kernels[s] = {"ms_per_step": ms, "launches_per_step": launches[s], "algorithmic_bytes_per_step": nbytes,
                      "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                      "dram_traffic_bytes_per_launch_ncu": traffic.get(s)}
    em_ms = stage_ms.get("cluster", 0.0) + stage_ms.get("feat_moments", 0.0)
    em_bytes = ab["em_step"] * 2 * B
    em = {"ms_per_step": em_ms, "algorithmic_bytes_per_step": em_bytes, "achieved_gbs": em_bytes / (em_ms * 1e-3) / 1e9}
    em["frac_of_hbm_peak"] = em["achieved_gbs"] / peak
    em["frac_of_8tbs_nominal"] = em["achieved_gbs"] / 8000.0
    # The roofline object is for the HBM-bound kernel of the path, the feature M-step (the metric's
    # "E/M-step % of HBM peak"); kNN and the Sinkhorn loop are instruction-issue bound
    # and are listed with the same arithmetic under "kernels"; "dominant_by_time" names the longest stage.
    hb = "feat_moments"
    # Duration of the roofline kernel alone: its launches of `steps` steps back to back (src, tgt, src, ...; each feature
    # tensor is 537 MB = 4x L2, so nothing is reused), ONE CUDA-event pair around all of them on the launching stream.
    # The per-stage events of the serial pass also time the launch gap after a different kernel and the event records
    # themselves (~10 us on a ~100 us launch); that in-step figure stays in kernels["feat_moments"] and "frac_in_step".
    from ogmm_b200 import ops
    g_s, g_t = out["src_gamma"], out["tgt_gamma"]
    f_s, f_t = d["src_feats"].transpose(-1, -2), d["tgt_feats"].transpose(-1, -2)
    for _ in range(3):
        ops.gmm_moments(g_s, f_s); ops.gmm_moments(g_t, f_t)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        ops.gmm_moments(g_s, f_s); ops.gmm_moments(g_t, f_t)
    k1.record()
    barrier()
    launch_ms = k0.elapsed_time(k1) / (2 * args.steps)
    launch_gbs = ab[hb] * B / (launch_ms * 1e-3) / 1e9
    roofline = {"kernel": "gmm_moments_feat_tma_kernel", "bound": "hbm", "achieved": launch_gbs, "peak": peak,
                "unit": "GB/s", "frac": launch_gbs / peak, "traffic": traffic.get(hb), "peak_source": peak_src,
                "launches_per_step": launches[hb], "ms_per_launch": launch_ms, "launches_timed": 2 * args.steps,
                "frac_in_step": kernels[hb]["frac_of_hbm_peak"], "ms_per_launch_in_step": stage_ms[hb] / launches[hb],
                "dominant_by_time": max(stage_ms, key=stage_ms.get), "serial_ms_per_step": serial_ms,
                "note": "achieved = algorithmic bytes per launch (4(NJ+ND+JD) per cloud x 256 clouds) / average launch "
                        "duration over 2 x steps back-to-back launches of the kernel (CUDA events on its stream, inputs "
                        "4x L2); frac_in_step is the same arithmetic on the per-stage event time inside the serial pass "
                        "of the step, which includes the launch gap after the clustering kernel"}

    # ---- end to end: pinned host buffers in, (R, t) out ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pinned = {k: host[k].pin_memory() for k in ("src", "tgt", "src_feats", "tgt_feats", "src_o", "tgt_o")}
        for _ in range(2):
            pipeline.register_from_host(pinned, dev, N_CLUSTERS, KNN, ITERS)
        barrier()
        steps_e2e = max(3, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps_e2e):
            rot_h, trans_h, h2d, d2h = pipeline.register_from_host(pinned, dev, N_CLUSTERS, KNN, ITERS)
        a1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(a0.elapsed_time(a1), wall * 1e3)
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * steps_e2e / (float(tt.item()) * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": steps_e2e,
               "note": "every hot-path input (xyz, overlap scores AND the 512-d point features that the PyTorch DGCNN "
                       "produces on-device in the real model) is copied from pinned host memory each step"}

    # ---- evaluation metrics: the only collective (outside the timed region) --------------------------------------
    mvec = pipeline.local_metrics(out["rot"], out["trans"], d["rot_gt"], d["t_gt"])
    metrics = pipeline.reduce_metrics(mvec)

    cpu = None
    if rank == 0 and not args.no_cpu:
        args.cpu_pairs = args.cpu_pairs or 256           # the workload's own batch: ~3 s per pass on 16 cores
        times, cores = time_cpu(args.cpu_pairs, 3, 1)
        best = min(times)
        cpu = {"value": args.cpu_pairs / best, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_pairs} pairs of the same workload, best of 3 after 1 warm-up ({best:.2f} s)"}

    if rank == 0:
        line = {
            "metric": "registration pairs/sec (1024-pt, J=16)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(B), launch=("cuda_graph_replay (one two-stream step per graph)" if graphed is not None
                                                         else (graph_note or "eager, two streams") if not args.no_overlap
                                                         else "eager, one stream")),
            "roofline": roofline, "em_step": em, "kernels": kernels,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": pipeline.launches_per_step(ITERS) * args.steps,
            "clocks": clocks, "eval_metrics_allreduced": metrics,
            "eval_metrics_note": "synthetic relu(N(0,1)) point features carry no geometry, so the registration errors are "
                                 "meaningless here; the vector only exercises the path's one collective (4-float all-reduce)",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA-graph replay")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic pairs generated per rank (tiled to --pairs)")
    ap.add_argument("--cpu-pairs", type=int, default=None,
                    help="pairs per CPU step (default: 256 for the cpu_baseline leg = ~12 s of CPU work; for --impl "
                         "reference as many, up to 256, as keep steps + warmup within ~2 minutes at ~90 pairs/s)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the src and tgt chains back to back on one stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
