#!/usr/bin/env python
"""Benchmark of the OGMM registration hot path (BASELINE.json metric: registration pairs/sec, 1024-pt, J=16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--pairs B]

A step = one pass of the hot path over one batch of B synthetic pairs per GPU.  One process per GPU; ranks own
disjoint pairs and never communicate inside the timed region (weak scaling).  Rank 0 prints ONE JSON line.

  --config 2 (default)  BASELINE.json configs[1]: ModelNet40-shape partial-overlap pairs, N=1024, J=16, D=512, k=20,
                        B=256: kNN graph + edge features, overlap-guided Sinkhorn clustering, feature M-step,
                        soft-correspondence Procrustes (models/gmmreg.py:52-53,100-103)
  --config 3            configs[2]: the DeepGMR path (baseline/deepgmr.py:64-79) on ICL-NUIM-shape pairs with density
                        variation: kNN graph + edge features, fused softmax E-step + M-step with sigma, gmm_register
  --config 4            configs[3]: large-scale pairs, N=16384, J=64, the flagship path plus the feature-space kNN
                        graph on C=64 wide features (tensor-core kernel); pair-sharded over the GPUs
  --config 5            configs[4]: batch sweep 1..8192 pairs of the flagship path (one line, ``sweep`` holds the points)

  value         pairs/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e           same step through the host-buffer entry point: pinned host buffers in, (R, t) back on the host
  roofline      the dominant-by-time kernel's algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json
  roofline_hbm  the same for the HBM-bound kernel of the path (feature M-step / DeepGMR E+M), timed alone
  kernels       the same numbers for every stage;  em_step  the E/M-step as one unit
  parity        maxima of |GPU - CPU oracle| on THIS step's batch (same inputs on both arms)
  cpu_baseline  the oracle port (the reference's own PyTorch op sequence) on this box's host cores
  cuda_reference  the same op sequence on CUDA tensors (stock ATen: cuBLAS bmm, topk, logsumexp, host SVD): the
                  reference's stock-PyTorch GPU path on this B200, the kernel-vs-kernel bar
  gmmreg_forward  full GMMReg.forward of the unmodified reference (baseline/_ref) on this GPU, unpatched vs install()

``--impl reference`` times the CPU path as its own arm (rank 0 only under torchrun).
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

ITERS = 10
METRIC = "registration pairs/sec (1024-pt, J=16)"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle-reason samples taken INSIDE the timed region.

    NVML is polled from a thread of this process every few milliseconds (attached before the warm-up: the attach
    itself holds driver locks for up to a few hundred ms and used to stall kernel launches when it happened next to
    the timed region).  If NVML cannot be imported, one ``nvidia-smi -lms`` child is used instead."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu_index, self.period = gpu_index, period_s
        self.rows, self.lo, self.hi = [], 0, None
        self.proc, self.thread, self.stop_flag, self.nvml, self.handle, self.smax = None, None, False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                    idx = int(ids[self.gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                try:
                    power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    power = None
                self.rows.append((sm, self.smax, power, [name for name, bit in self.REASONS if mask & bit]))
            except Exception:
                pass
            time.sleep(self.period)

    def _pump_smi(self):
        names = [r[0] for r in self.REASONS]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                self.rows.append((float(f[1]), float(f[2]), float(f[3]),
                                  [nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")]))
            except ValueError:
                continue

    def wait_ready(self, timeout=10.0):
        t0 = time.time()
        while self.thread is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def begin(self):
        self.lo = len(self.rows)

    def end(self):
        self.hi = len(self.rows)

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"]}
        rows = self.rows[self.lo:self.hi]
        inside = len(rows)
        if inside < 3:                                   # a very short region: the closest samples around it
            rows = self.rows[max(0, self.lo - 2):(self.hi or 0) + 2]
        sm = sorted(r[0] for r in rows)
        reasons = sorted({x for r in rows for x in r[3]})
        power = [r[2] for r in rows if r[2] is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in rows), default=None),
                "power_w_max": max(power) if power else None, "samples": len(rows), "samples_inside_timed_region": inside,
                "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": reasons}


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


# ----------------------------------------------------------------------------------------- workloads
class Flagship:
    """configs[1] (cfg 2) and, with other sizes plus the wide kNN graph, configs[3] (cfg 4)."""

    def __init__(self, cfg=2):
        self.cfg = cfg
        if cfg == 4:
            self.N, self.J, self.D, self.K, self.C = 16384, 64, 512, 20, 64
            self.default_pairs, self.default_distinct = 4, 2
            self.what = ("large-scale pairs, 16384 pts, J=64: kNN(k=20)+edge features on xyz, feature-space kNN graph on 64-d wide "
                         "features (tensor cores), overlap-guided Sinkhorn clustering (10x10), feature M-step (D=512), "
                         "soft-correspondence Procrustes (BASELINE.json configs[3])")
        else:
            self.N, self.J, self.D, self.K, self.C = 1024, 16, 512, 20, 0
            self.default_pairs, self.default_distinct = 256, 32
            self.what = ("OGMM registration hot path: kNN(k=20)+edge features, overlap-guided Sinkhorn clustering "
                         "(10x10, J=16), feature M-step (D=512), soft-correspondence Procrustes; "
                         "ModelNet40-shape partial-overlap pairs, 1024 pts (BASELINE.json configs[1])")
        self.input_names = ("src", "tgt", "src_feats", "tgt_feats", "src_o", "tgt_o") + (("src_wide", "tgt_wide") if self.C else ())
        self.hbm_stage, self.hbm_kernel = "feat_moments", "gmm_moments_feat_tma_kernel" if cfg == 2 else "gmm_moments_feat kernels"
        self.stage_kernels = {"knn_edge": ("knn3_select_kernel<20> (distance + top-k + edge write)" if cfg == 2 else
                                           "knn3_presort_kernel + knn3_tile_sweep_kernel<20> (distance + top-k + edge write)"),
                              "knn_wide": "knn_wide2_kernel<20,2> (TMA + tcgen05)",
                              "cluster": ("sinkhorn_kernel<256,4,...> (FPS + Sinkhorn k-means, one CTA per cloud)" if cfg == 2 else
                                          "sinkhorn_cluster_dsmem_kernel (FPS + Sinkhorn k-means, one 16-CTA cluster per cloud)"),
                              "feat_moments": self.hbm_kernel, "procrustes": "soft_procrustes_full_kernel"}

    def host_inputs(self, first, pairs, distinct):
        import numpy as np
        import torch
        from ogmm_b200 import synth
        h = synth.hot_path_inputs(first, pairs, self.N, self.D, tile=distinct)
        if self.C:
            rng = np.random.default_rng(synth.BASE_SEED + 99 + int(first))
            d = min(distinct or pairs, pairs)
            w = np.maximum(rng.normal(size=(2, d, self.N, self.C)), 0.0).astype(np.float32)
            reps = -(-pairs // d)
            h["src_wide"] = np.ascontiguousarray(np.concatenate([w[0]] * reps, 0)[:pairs])
            h["tgt_wide"] = np.ascontiguousarray(np.concatenate([w[1]] * reps, 0)[:pairs])
        return {k: torch.from_numpy(v) for k, v in h.items()}

    def algorithmic_bytes(self):
        """Per CLOUD (a pair is two clouds), fp32 + int64 idx; SURVEY.md section 8(d), restated in DESIGN.md."""
        n, j, d, k, c = self.N, self.J, self.D, self.K, 3
        out = {"knn_edge": 4 * c * n + 8 * n * k + 4 * 2 * c * n * k,          # xyz in, idx + edge out
               "cluster": 4 * (3 * n + n + n * j + 4 * j),                      # xyz + o in, gamma + pi + mu out
               "feat_moments": 4 * (n * j + n * d + j * d),                     # gamma re-read + feats in, node_feats out
               "procrustes": 4 * (2 * 3 * j + 2 * j * d) // 2 + 24,             # per cloud share of the per-pair bytes
               "em_step": 4 * (3 * n + n + n * j + n * d + j * d + 4 * j)}      # E/M-step as one unit (gamma not re-read)
        if self.C:
            out["knn_wide"] = 4 * self.C * n + 8 * n * k                         # wide features in, idx out
        return out

    def launches(self):
        """Kernel launches per step and stage (both clouds).  cfg 4: the tiled kNN is a pre-pass + the sweep, the tensor-core
        kNN an operand pre-kernel + the pipeline kernel, the feature M-step in split mode the partial kernel + the fold."""
        if self.cfg == 4:
            return {"knn_edge": 4, "knn_wide": 4, "cluster": 2, "feat_moments": 4, "procrustes": 1}
        return {"knn_edge": 2, "cluster": 2, "feat_moments": 2, "procrustes": 1}

    def gpu_step(self, d, timers=None, overlap=True):
        from ogmm_b200 import pipeline
        wide = (d["src_wide"], d["tgt_wide"]) if self.C else (None, None)
        return pipeline.register_hot_path(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"],
                                          self.J, self.K, ITERS, timers, overlap, wide=wide)

    def graphed(self, d):
        from ogmm_b200 import pipeline
        return pipeline.GraphedHotPath(d["src"], d["tgt"], d["src_feats"], d["tgt_feats"], d["src_o"], d["tgt_o"],
                                       self.J, self.K, ITERS, fn=lambda: self.gpu_step(d))

    def oracle_step(self, orc, torch, t):
        """The reference's own op sequence for the hot path (oracle port); runs on whatever device ``t`` lives on."""
        outs = {}
        for side in ("src", "tgt"):
            x = t[side]
            idx = orc.knn_indices(x.transpose(-1, -2), x.transpose(-1, -2), self.K)
            outs[side + "_edge"] = orc.edge_features(x, self.K, idx)
            if self.C:
                outs[side + "_wide_idx"] = orc.knn_indices(t[side + "_wide"], t[side + "_wide"], self.K)
            gam, pi, mu, nf = orc.sinkhorn_kmeans(x.transpose(-1, -2), t[side + "_feats"].transpose(-1, -2), t[side + "_o"], self.J)
            outs.update({side + "_gamma": gam, side + "_pi": pi, side + "_mu": mu, side + "_node_feats": nf})
        rot, trans = orc.soft_svd_head(outs["src_mu"], outs["tgt_mu"], outs["src_node_feats"], outs["tgt_node_feats"],
                                       outs["src_pi"], outs["tgt_pi"])[:2]
        outs.update({"rot": rot, "trans": trans})
        return outs

    def time_hbm_kernel_alone(self, d, out, steps, barrier):
        """Duration of the HBM-bound kernel alone: its launches of `steps` steps back to back (src, tgt, src, ...; each
        feature tensor is larger than L2, so nothing is reused), ONE CUDA-event pair on the launching stream."""
        import torch
        from ogmm_b200 import ops
        g_s, g_t = out["src_gamma"], out["tgt_gamma"]
        f_s, f_t = d["src_feats"].transpose(-1, -2), d["tgt_feats"].transpose(-1, -2)
        for _ in range(3):
            ops.gmm_moments(g_s, f_s); ops.gmm_moments(g_t, f_t)
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(steps):
            ops.gmm_moments(g_s, f_s); ops.gmm_moments(g_t, f_t)
        k1.record()
        barrier()
        return k0.elapsed_time(k1) / (2 * steps), 2 * steps

    def e2e(self, pinned, dev, device_feats=None):
        from ogmm_b200 import pipeline
        if device_feats is not None and not self.C:
            # model boundary: static device buffers + one CUDA-graph launch per step (pipeline.HostBoundary)
            key = (str(dev), tuple(pinned["src"].shape), device_feats[0].data_ptr(), device_feats[1].data_ptr())
            if getattr(self, "_hb_key", None) != key:
                self._hb = pipeline.HostBoundary(pinned, dev, device_feats[0], device_feats[1], self.J, self.K, ITERS)
                self._hb_key = key
            rot, trans, h2d, d2h = self._hb(pinned)
            return h2d, d2h
        rot, trans, h2d, d2h = pipeline.register_from_host(pinned, dev, self.J, self.K, ITERS, device_feats=device_feats)
        return h2d, d2h

    def parity(self, torch, orc, out, ref, host, max_clouds=8):
        """Maxima of |GPU - CPU oracle| over this step's batch; per stage on identical stage inputs where the stage
        inputs are the raw inputs (kNN, clustering) and, for the head, on the GPU's own GMM parameters."""
        from tests_gpu_util import decidable_rows, rot_err_deg
        p = {}
        k = self.K
        same_rows, rows = 0, 0
        for side in ("src", "tgt"):
            got = out["edge_" + side].cpu()
            eq = (got == ref[side + "_edge"]).all(dim=1).all(dim=-1)               # (B,N): the whole k x 6 edge block of a query
            same_rows += int(eq.sum()); rows += eq.numel()
        p["edge_rows_identical_frac"] = same_rows / rows
        # decidable rows (fp64 gap at every one of the k boundaries above the fp32 error bound) must match exactly
        nb = min(max_clouds, host["src"].shape[0])
        x = host["src"][:nb].transpose(1, 2).contiguous()
        ok, _ = decidable_rows(x, x, k)
        got = out["edge_src"][:nb].cpu().permute(0, 2, 3, 1)
        want = ref["src_edge"][:nb].permute(0, 2, 3, 1)
        p["knn_decidable_rows_frac"] = float(ok.float().mean())
        p["knn_decidable_rows_mismatch"] = int((got[ok] != want[ok]).any(dim=-1).any(dim=-1).sum())
        for side in ("src", "tgt"):
            mu, rmu = out[side + "_mu"].cpu(), ref[side + "_mu"]
            sc = float(rmu.abs().max())
            p.setdefault("mu_scale_rel", 0.0); p.setdefault("pi_rel", 0.0); p.setdefault("node_feats_rel", 0.0); p.setdefault("gamma_abs", 0.0)
            p["mu_scale_rel"] = max(p["mu_scale_rel"], float((mu - rmu).abs().max()) / sc)
            p["pi_rel"] = max(p["pi_rel"], float((out[side + "_pi"].cpu() - ref[side + "_pi"]).abs().max() / ref[side + "_pi"].abs().max()))
            nf, rnf = out[side + "_node_feats"].cpu(), ref[side + "_node_feats"]
            p["node_feats_rel"] = max(p["node_feats_rel"], float((nf - rnf).abs().max() / rnf.abs().max()))
            p["gamma_abs"] = max(p["gamma_abs"], float((out[side + "_gamma"].cpu() - ref[side + "_gamma"]).abs().max()))
        # head on identical stage inputs: the oracle head fed with the GPU's GMM parameters
        rr, rt = orc.soft_svd_head(out["src_mu"].cpu(), out["tgt_mu"].cpu(), out["src_node_feats"].cpu(), out["tgt_node_feats"].cpu())[:2]
        r64, t64 = orc.soft_svd_head(out["src_mu"].cpu().double(), out["tgt_mu"].cpu().double(), out["src_node_feats"].cpu().double(),
                                     out["tgt_node_feats"].cpu().double())[:2]
        scale = float(torch.maximum(host["src"].abs().max(), host["tgt"].abs().max()))
        p["rot_deg_stage"] = float(rot_err_deg(out["rot"].cpu(), rr).max())
        p["trans_over_scale_stage"] = float((out["trans"].cpu() - rt).abs().max()) / scale
        p["rot_deg_stage_vs_fp64"] = float(rot_err_deg(out["rot"].cpu(), r64).max())
        p["trans_over_scale_stage_vs_fp64"] = float((out["trans"].cpu().double() - t64).abs().max()) / scale
        p["rot_deg_oracle_fp32_vs_fp64"] = float(rot_err_deg(rr, r64).max())
        p["rot_deg_end_to_end"] = float(rot_err_deg(out["rot"].cpu(), ref["rot"]).max())
        p["trans_over_scale_oracle_fp32_vs_fp64"] = float((rt.double() - t64).abs().max()) / scale
        # the head's conditioning depends on the descriptors: the bar is 1e-3 deg / 1e-5 x scale against the fp64 arbiter, or,
        # where the fp32 reference itself sits further than that from the arbiter on these inputs, 1.5 x the reference's distance
        rot_bar = max(1e-3, 1.5 * p["rot_deg_oracle_fp32_vs_fp64"])
        trans_bar = max(1e-5, 1.5 * p["trans_over_scale_oracle_fp32_vs_fp64"])
        p["bars"] = {"knn_decidable_rows_mismatch": 0, "mu_scale_rel": 1e-4, "pi_rel": 1e-4, "node_feats_rel": 1e-4,
                     "rot_deg_stage_vs_fp64": rot_bar, "trans_over_scale_stage_vs_fp64": trans_bar}
        p["ok"] = bool(p["knn_decidable_rows_mismatch"] == 0 and p["mu_scale_rel"] <= 1e-4 and p["pi_rel"] <= 1e-4 and
                       p["node_feats_rel"] <= 1e-4 and p["rot_deg_stage_vs_fp64"] <= rot_bar and
                       p["trans_over_scale_stage_vs_fp64"] <= trans_bar)
        p["note"] = ("GPU step vs the CPU oracle on the same batch; kNN: decidable rows of the first %d source clouds must be "
                     "identical; head: oracle head run on the GPU's own (mu, node_feats) = identical stage inputs; "
                     "rot_deg_end_to_end is reported, not gated.  The synthetic relu(N(0,1)) features carry no geometry: all "
                     "component descriptors are nearly parallel, the soft assignment is nearly uniform and the 3x3 covariance "
                     "nearly rank 0, so the fp32 reference itself is only defined to rot_deg_oracle_fp32_vs_fp64 here; smoke() "
                     "and tests/test_gpu_parity.py run the head on well-conditioned inputs against the absolute bars" % nb)
        return p

    def config(self, pairs, note=None):
        cfg = {"workload": self.what, "config_id": self.cfg, "pairs_per_gpu_per_step": pairs, "n_points": self.N,
               "n_clusters": self.J, "emb_dims": self.D, "k": self.K, "sinkhorn": "10 outer x 10 inner, eps=1e-2",
               "parallelism": "pair-sharded, no hot-path collective",
               "l2_policy": "inputs larger than L2 (feature tensors are %.0f MB per step; 126 MB L2)" % (2 * pairs * self.D * self.N * 4 / 1e6)}
        if self.C:
            cfg["wide_feature_dims"] = self.C
        if note:
            cfg["note"] = note
        return cfg


class DeepGMRPath:
    """configs[2] (cfg 3): DeepGMR path on ICL-NUIM-shape pairs with density variation."""

    def __init__(self):
        self.cfg, self.N, self.J, self.K, self.D, self.C = 3, 1024, 16, 20, 0, 0
        self.default_pairs, self.default_distinct = 256, 32
        self.what = ("DeepGMR path (baseline/deepgmr.py): kNN(k=20)+edge features, fused softmax E-step + M-step with sigma "
                     "(J=16), gmm_register; ICL-NUIM-shape pairs with density variation, 1024 pts (BASELINE.json configs[2])")
        self.input_names = ("src", "tgt", "src_logits", "tgt_logits")
        self.hbm_stage, self.hbm_kernel = "softmax_em", "softmax_moments16_kernel"
        self.stage_kernels = {"knn_edge": "knn3_select_kernel<20> (distance + top-k + edge write)", "softmax_em": self.hbm_kernel,
                              "gmm_register": "gmm_register_kernel"}

    def host_inputs(self, first, pairs, distinct):
        import numpy as np
        import torch
        from ogmm_b200 import synth
        d = min(distinct or pairs, pairs)
        src, tgt, rot, t = synth.icl_nuim_batch(first, d, self.N)
        rng = np.random.default_rng(synth.BASE_SEED + 3 + int(first))
        lg = (rng.normal(size=(2, d, self.J, self.N)) * 2).astype(np.float32)
        reps = -(-pairs // d)
        arrs = {"src": src, "tgt": tgt, "src_logits": lg[0], "tgt_logits": lg[1], "rot_gt": rot, "t_gt": t}
        return {k: torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * reps, 0)[:pairs], dtype=np.float32)) for k, a in arrs.items()}

    def algorithmic_bytes(self):
        n, j, k, c = self.N, self.J, self.K, 3
        return {"knn_edge": 4 * c * n + 8 * n * k + 4 * 2 * c * n * k,
                "softmax_em": 4 * (j * n + 3 * n + 13 * j),                      # logits + xyz in; pi, mu, sigma out
                "gmm_register": 4 * (j + 3 * j + 3 * j + 9 * j) // 2 + 32,       # per cloud share of the per-pair bytes
                "em_step": 4 * (j * n + 3 * n + 13 * j)}

    def launches(self):
        return {"knn_edge": 2, "softmax_em": 2, "gmm_register": 1}

    def gpu_step(self, d, timers=None, overlap=True):
        from ogmm_b200 import pipeline
        return pipeline.deepgmr_hot_path(d["src"], d["tgt"], d["src_logits"], d["tgt_logits"], self.K, timers, overlap)

    def graphed(self, d):
        from ogmm_b200 import pipeline
        return pipeline.GraphedHotPath(d["src"], d["tgt"], None, None, None, None, fn=lambda: self.gpu_step(d))

    def oracle_step(self, orc, torch, t):
        outs = {}
        for side in ("src", "tgt"):
            x = t[side]
            idx = orc.knn_indices(x.transpose(-1, -2), x.transpose(-1, -2), self.K)
            outs[side + "_edge"] = orc.edge_features(x, self.K, idx)
            _, pi, mu, sg = orc.deepgmr_em(t[side + "_logits"], x)
            outs.update({side + "_pi": pi, side + "_mu": mu, side + "_sigma": sg})
        outs["transform"] = orc.deepgmr_register(outs["src_pi"], outs["src_mu"], outs["tgt_mu"], outs["tgt_sigma"])
        return outs

    def time_hbm_kernel_alone(self, d, out, steps, barrier):
        import torch
        from ogmm_b200 import ops
        for _ in range(3):
            ops.softmax_moments(d["src_logits"], d["src"], want_gamma=False); ops.softmax_moments(d["tgt_logits"], d["tgt"], want_gamma=False)
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(steps):
            ops.softmax_moments(d["src_logits"], d["src"], want_gamma=False); ops.softmax_moments(d["tgt_logits"], d["tgt"], want_gamma=False)
        k1.record()
        barrier()
        return k0.elapsed_time(k1) / (2 * steps), 2 * steps

    def e2e(self, pinned, dev, device_feats=None):
        from ogmm_b200 import pipeline
        _, h2d, d2h = pipeline.deepgmr_from_host(pinned, dev, self.K)
        return h2d, d2h

    def parity(self, torch, orc, out, ref, host, max_clouds=8):
        from tests_gpu_util import decidable_rows, rot_err_deg
        p = {}
        same_rows, rows = 0, 0
        for side in ("src", "tgt"):
            eq = (out["edge_" + side].cpu() == ref[side + "_edge"]).all(dim=1).all(dim=-1)
            same_rows += int(eq.sum()); rows += eq.numel()
        p["edge_rows_identical_frac"] = same_rows / rows
        nb = min(max_clouds, host["src"].shape[0])
        x = host["src"][:nb].transpose(1, 2).contiguous()
        ok, _ = decidable_rows(x, x, self.K)
        got = out["edge_src"][:nb].cpu().permute(0, 2, 3, 1)
        want = ref["src_edge"][:nb].permute(0, 2, 3, 1)
        p["knn_decidable_rows_frac"] = float(ok.float().mean())
        p["knn_decidable_rows_mismatch"] = int((got[ok] != want[ok]).any(dim=-1).any(dim=-1).sum())
        rel = lambda a, b: float((a.cpu() - b).abs().max() / b.abs().max())
        p["pi_rel"] = max(rel(out["src_pi"], ref["src_pi"]), rel(out["tgt_pi"], ref["tgt_pi"]))
        p["mu_scale_rel"] = max(rel(out["src_mu"], ref["src_mu"]), rel(out["tgt_mu"], ref["tgt_mu"]))
        p["sigma_rel"] = max(rel(out["src_sigma"], ref["src_sigma"]), rel(out["tgt_sigma"], ref["tgt_sigma"]))
        c = lambda name: out[name].cpu()
        tf32 = orc.deepgmr_register(c("src_pi"), c("src_mu"), c("tgt_mu"), c("tgt_sigma"))
        tf64 = orc.deepgmr_register(c("src_pi").double(), c("src_mu").double(), c("tgt_mu").double(), c("tgt_sigma").double())
        tf = out["transform"].cpu()
        scale = float(torch.maximum(host["src"].abs().max(), host["tgt"].abs().max()))
        p["rot_deg_stage"] = float(rot_err_deg(tf[:, :3, :3], tf32[:, :3, :3]).max())
        p["rot_deg_stage_vs_fp64"] = float(rot_err_deg(tf[:, :3, :3], tf64[:, :3, :3]).max())
        p["rot_deg_oracle_fp32_vs_fp64"] = float(rot_err_deg(tf32[:, :3, :3], tf64[:, :3, :3]).max())
        p["trans_over_scale_stage_vs_fp64"] = float((tf[:, :3, 3].double() - tf64[:, :3, 3]).abs().max()) / scale
        p["ok"] = bool(p["knn_decidable_rows_mismatch"] == 0 and p["pi_rel"] <= 1e-4 and p["mu_scale_rel"] <= 1e-4 and
                       p["sigma_rel"] <= 1e-4 and p["rot_deg_stage_vs_fp64"] <= 1e-3 and p["trans_over_scale_stage_vs_fp64"] <= 1e-5)
        return p

    def config(self, pairs, note=None):
        cfg = {"workload": self.what, "config_id": 3, "pairs_per_gpu_per_step": pairs, "n_points": self.N, "n_clusters": self.J,
               "k": self.K, "parallelism": "pair-sharded, no hot-path collective",
               "l2_policy": "L2 flushed implicitly: the step writes %.0f MB of edge features, more than the 126 MB L2, between "
                            "consecutive reads of any input" % (2 * pairs * self.N * self.K * 24 / 1e6)}
        if note:
            cfg["note"] = note
        return cfg


def make_workload(cfg):
    if cfg == 3:
        return DeepGMRPath()
    return Flagship(4 if cfg == 4 else 2)


# ----------------------------------------------------------------------------------------- CPU arm
def time_cpu(wl, host, pairs, steps, warmup, budget_s=None):
    import torch
    from oracle import ogmm_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t = {k: v[:pairs] for k, v in host.items()}
    out = None
    with torch.no_grad():
        for _ in range(warmup):
            out = wl.oracle_step(orc, torch, t)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = wl.oracle_step(orc, torch, t)
            times.append(time.perf_counter() - t0)
            if budget_s is not None and sum(times) > budget_s:
                break
    return times, cores, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = make_workload(2 if args.config == 5 else args.config)
    # a bounded sample of the workload per step: as many pairs as keep steps + warmup within ~2 minutes
    rate = {2: 90.0, 3: 400.0, 4: 0.05}[wl.cfg]                       # rough pairs/s of the CPU path on ~16 cores
    pairs = args.cpu_pairs or int(min(wl.default_pairs, max(1 if wl.cfg == 4 else 4, 120 * rate / (args.steps + args.warmup))))
    host = wl.host_inputs(0, pairs, min(args.distinct or wl.default_distinct, pairs))
    times, cores, _ = time_cpu(wl, host, pairs, args.steps, args.warmup, budget_s=240.0)
    total = sum(times)
    value = pairs * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.config(pairs, "host CPU; each step is a bounded sample of the workload (the first %d pairs of rank 0's batch)" % pairs),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{pairs} pairs per step x {len(times)} steps (same synthetic pairs, hot path only)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "steps_timed": len(times),
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------- extra legs (rank 0, N = 1)
def time_cuda_reference(wl, host, dev, reps=3):
    """The reference's op sequence on CUDA tensors: stock ATen kernels (cuBLAS bmm, topk, logsumexp, index) plus the
    reference's own host round trips (SVD on the CPU, .item() per Sinkhorn iteration)."""
    import torch
    from oracle import ogmm_oracle as orc
    t = {k: v.to(dev) for k, v in host.items()}
    pairs = host["src"].shape[0]
    with torch.no_grad():
        wl.oracle_step(orc, torch, t)
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            wl.oracle_step(orc, torch, t)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return {"value": pairs / best, "unit": "pairs/s", "ms_per_step": best * 1e3, "pairs": pairs, "kind": "port on cuda tensors",
            "note": "oracle/ogmm_oracle.py (the reference's ATen op sequence, pinned to the reference by tests/golden) run on "
                    "CUDA tensors on this GPU: stock PyTorch kernels, host SVD and per-iteration .item() syncs as in the reference; "
                    f"best of {reps} after 1 warm-up, wall clock around a synchronize"}


def time_model_forward(dev, pairs=32, reps=3):
    """Figure (i) of BASELINE.md section 3: full GMMReg.forward of the unmodified reference, eval / no_grad, on this GPU,
    stock path vs after install()."""
    import torch
    from oracle import refload
    if refload.reference_path() is None:
        return {"unavailable": "no reference checkout (baseline/_ref is vendored by __graft_entry__.build() where /root/reference exists)"}
    import ogmm_b200.install as inst
    from ogmm_b200 import synth
    ref = refload.import_reference()
    torch.manual_seed(1234)
    model = ref["gmmreg"].GMMReg(512, 16, refload.model_config()).to(dev).eval()
    s, t, _, _ = synth.modelnet_batch(0, pairs, 1024)
    src, tgt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)

    def best_of():
        best = None
        # cuDNN off in both arms, as the reference's own scripts run (train.py:194-196): FP32 convolutions, no TF32
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
            for i in range(reps + 1):
                torch.manual_seed(7)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = model(src, tgt)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if i > 0:
                    best = dt if best is None else min(best, dt)
        return best, out

    t_ref, out_ref = best_of()
    inst.install(model=model)
    try:
        t_new, out_new = best_of()
    finally:
        inst.uninstall()
    train = time_model_train_step(dev, ref, refload, inst, synth)
    from tests_gpu_util import rot_err_deg
    return {"train_step": train,
            "pairs": pairs, "reference_ms": t_ref * 1e3, "patched_ms": t_new * 1e3, "reference_pairs_per_s": pairs / t_ref,
            "patched_pairs_per_s": pairs / t_new, "speedup": t_ref / t_new,
            "rot_deg_patched_vs_reference": float(rot_err_deg(out_new[0].cpu(), out_ref[0].cpu()).max()),
            "overlap_score_abs_diff": float((out_new[2] - out_ref[2]).abs().max()),
            "note": "whole model incl. the PyTorch DGCNN convolutions and transformer overlap detector that stay PyTorch in "
                    "both arms; random-init weights, eval(), no_grad, cuDNN disabled as in train.py:194-196, best of %d after 1 warm-up" % reps}


def time_model_train_step(dev, ref, refload, inst, synth, pairs=8, reps=3):
    """One training step of the unmodified GMMReg (train.py:53-75: forward in train() mode, registration + clustering +
    overlap loss, backward), stock path vs after install() -- where the differentiable hot-path pieces (feature M-step,
    soft-correspondence head, Procrustes / 3x3 SVD) run on the repo's forward and backward kernels."""
    import torch
    try:
        torch.manual_seed(99)
        model = ref["gmmreg"].GMMReg(512, 16, refload.model_config()).to(dev).train()
        s, t, R_gt, t_gt = synth.modelnet_batch(100, pairs, 1024)
        src, tgt = torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)
        rot_gt, trans_gt = torch.from_numpy(R_gt).to(dev).float(), torch.from_numpy(t_gt).to(dev).float().view(pairs, 3)
        o_gt = (torch.rand(pairs, 2048, generator=torch.Generator().manual_seed(1)) > 0.4).float().to(dev)
        loss_mod = ref["loss"]

        def best_of():
            best, gnorm, lossv = None, None, None
            with torch.backends.cudnn.flags(enabled=False):
                for i in range(reps + 1):
                    model.zero_grad(set_to_none=True)
                    torch.manual_seed(7)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    rot, trans, src_o, tgt_o, clu = model(src, tgt)
                    o_pred = torch.nan_to_num(torch.cat([src_o, tgt_o], dim=-1), nan=0.0).clip(min=0.0)
                    loss = 10 * loss_mod.dcp_loss(rot, rot_gt, trans, trans_gt) + clu + loss_mod.get_weighted_bce_loss(o_pred, o_gt)
                    loss.backward()
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    if i > 0:
                        best = dt if best is None else min(best, dt)
                gnorm = float(sum(p.grad.double().pow(2).sum() for p in model.parameters() if p.grad is not None) ** 0.5)
                lossv = float(loss.detach())
            return best, gnorm, lossv

        t_ref, g_ref, l_ref = best_of()
        inst.install(model=model)
        try:
            t_new, g_new, l_new = best_of()
        finally:
            inst.uninstall()
        return {"pairs": pairs, "reference_ms": t_ref * 1e3, "patched_ms": t_new * 1e3, "speedup": t_ref / t_new,
                "loss_reference": l_ref, "loss_patched": l_new, "grad_norm_reference": g_ref, "grad_norm_patched": g_new,
                "note": "forward (train mode) + loss + backward of the unmodified reference model, random-init weights, cuDNN "
                        "disabled as in train.py:194-196, best of %d after 1 warm-up; no optimizer step" % reps}
    except Exception as e:                                   # the forward figure must survive a failure here
        return {"error": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ogmm_b200 as og
    from ogmm_b200 import pipeline

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config == 5:
        line = run_sweep(args, torch, dist, dev, world, rank, local, barrier, allmax)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    wl = make_workload(args.config)
    B = args.pairs or wl.default_pairs
    distinct = args.distinct or wl.default_distinct

    # ---- synthetic inputs: this rank's shard of the global pair list ---------------------------------------
    host = wl.host_inputs(rank * B, B, distinct)
    d = {k: v.to(dev) for k, v in host.items()}

    # The timed step replays the two-stream step captured into a CUDA graph over the resident inputs (one launch per
    # step, immune to host jitter); --no-graph / --no-overlap time the eager call instead.
    graphed, graph_note = None, None
    if not (args.no_graph or args.no_overlap):
        try:
            graphed = wl.graphed(d)
        except Exception as e:                       # the number must still come out: time the eager two-stream call
            graph_note = f"graph capture failed ({type(e).__name__}: {e}); eager launches timed instead"
            print("bench.py: " + graph_note, file=sys.stderr)
            torch.cuda.synchronize()

    def step(timers=None, overlap=not args.no_overlap):
        if graphed is not None and timers is None and overlap:
            return graphed.replay()
        return wl.gpu_step(d, timers, overlap)

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
    elapsed_ms = allmax(e0.elapsed_time(e1))                # before the per-stage pass below
    value = B * world * args.steps / (elapsed_ms * 1e-3)
    final = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}     # the timed step's own results

    # ---- per-stage times: a serial pass (one stream, src then tgt) of the same steps, so every stage's
    # CUDA-event time is that kernel alone (in the timed region above the two chains share the SMs) ---------
    for _ in range(3):                               # warm the allocator's main-stream pool for the serial order
        step(None, overlap=False)
    timers = {}
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        step(timers, overlap=False)
    s1.record()
    barrier()
    serial_ms = s0.elapsed_time(s1) / args.steps
    stage_ms = {s: sum(a.elapsed_time(b) for a, b in ev) / args.steps for s, ev in timers.items()}
    ab = wl.algorithmic_bytes()
    peak, peak_src = measured_peaks()
    traffic = measured_traffic()
    launches = wl.launches()
    kernels = {}
    for s, ms in stage_ms.items():
        nbytes = ab[s] * 2 * B                       # two clouds per pair
        gbs = nbytes / (ms * 1e-3) / 1e9
        kernels[s] = {"kernel": wl.stage_kernels.get(s), "ms_per_step": ms, "launches_per_step": launches[s],
                      "algorithmic_bytes_per_step": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                      "share_of_serial_step": ms / serial_ms,
                      "dram_traffic_bytes_per_launch_ncu": traffic.get(s)}
    em_stages = [s for s in ("cluster", "feat_moments", "softmax_em") if s in stage_ms]
    em_ms = sum(stage_ms[s] for s in em_stages)
    em_bytes = ab["em_step"] * 2 * B
    em = {"stages": em_stages, "ms_per_step": em_ms, "algorithmic_bytes_per_step": em_bytes,
          "achieved_gbs": em_bytes / (em_ms * 1e-3) / 1e9}
    em["frac_of_hbm_peak"] = em["achieved_gbs"] / peak
    em["frac_of_8tbs_nominal"] = em["achieved_gbs"] / 8000.0

    # ---- roofline: the dominant-by-time stage (whatever bounds it), and the HBM-bound kernel timed alone ------------
    dom = max(stage_ms, key=stage_ms.get)
    dom_launches = {"procrustes": 1, "gmm_register": 1}.get(dom, 2)    # launches that do work (the clustering's redo rounds are empty)
    dom_ms = stage_ms[dom] / dom_launches
    dom_gbs = ab[dom] * B / (dom_ms * 1e-3) / 1e9
    bound_note = {"knn_edge": "selection (instruction issue) bound, not HBM bound: 1 M candidate distances per cloud",
                  "knn_wide": "tensor / selection bound", "cluster": "instruction-issue bound: up to 200 normalisation passes over "
                  "N x J on-chip entries per cloud, 82 KB of HBM traffic", "feat_moments": "HBM bound", "softmax_em": "HBM bound",
                  "procrustes": "latency bound", "gmm_register": "latency bound"}
    roofline = {"kernel": wl.stage_kernels.get(dom), "stage": dom, "bound": "hbm", "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                "frac": dom_gbs / peak, "traffic": traffic.get(dom), "peak_source": peak_src, "ms_per_launch": dom_ms,
                "launches_per_step": dom_launches, "share_of_serial_step": stage_ms[dom] / serial_ms,
                "serial_ms_per_step": serial_ms, "what_bounds_it": bound_note.get(dom),
                "note": "dominant stage by CUDA-event time in the serial pass of the step; achieved = algorithmic bytes per launch "
                        "(SURVEY.md section 8(d) per cloud x clouds per launch) / average launch duration"}
    hb = wl.hbm_stage
    launch_ms, n_timed = wl.time_hbm_kernel_alone(d, final, args.steps, barrier)
    launch_gbs = ab[hb] * B / (launch_ms * 1e-3) / 1e9
    roofline_hbm = {"kernel": wl.hbm_kernel, "stage": hb, "bound": "hbm", "achieved": launch_gbs, "peak": peak, "unit": "GB/s",
                    "frac": launch_gbs / peak, "traffic": traffic.get(hb), "peak_source": peak_src,
                    "launches_per_step": launches[hb], "ms_per_launch": launch_ms, "launches_timed": n_timed,
                    "frac_in_step": kernels[hb]["frac_of_hbm_peak"], "ms_per_launch_in_step": stage_ms[hb] / launches[hb],
                    "note": "the HBM-bound kernel of the path, its launches back to back inside one CUDA-event pair (inputs larger "
                            "than L2); frac_in_step is the same arithmetic on the per-stage event time inside the serial pass"}

    # ---- end to end: pinned host buffers in, (R, t) out ----------------------------------------------------------
    e2e, e2e_mb = None, None
    if not args.no_e2e:
        pinned = {k: host[k].pin_memory() for k in wl.input_names}

        def time_e2e(device_feats=None):
            for _ in range(2):
                wl.e2e(pinned, dev, device_feats)
            barrier()
            n = max(3, min(args.steps, args.e2e_steps))
            t0 = time.perf_counter()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                h2d, d2h = wl.e2e(pinned, dev, device_feats)
            a1.record()
            barrier()
            wall = time.perf_counter() - t0
            ms = allmax(max(a0.elapsed_time(a1), wall * 1e3))
            return {"value": B * world * n / (ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": n}

        e2e = time_e2e()
        if wl.cfg != 3:
            e2e["note"] = ("every hot-path input (xyz, overlap scores AND the 512-d point features that the PyTorch DGCNN "
                           "produces on-device in the real model) is copied from pinned host memory each step: PCIe bound by "
                           "construction (worst case); e2e_model_boundary is the same call with the features resident")
            e2e_mb = time_e2e((d["src_feats"], d["tgt_feats"]))
            e2e_mb["note"] = ("model boundary: xyz + overlap scores from pinned host memory each step, (R, t) back; the point "
                              "features stay on the device where models/gmmreg.py:52-97 produces them; device side = one "
                              "CUDA-graph launch over static buffers (pipeline.HostBoundary)")

    # ---- evaluation metrics: the only collective (outside the timed region) --------------------------------------
    mvec = pipeline.local_metrics(final["rot"], final["trans"], d["rot_gt"], d["t_gt"])
    metrics = pipeline.reduce_metrics(mvec)

    cpu = parity = cuda_ref = model_fwd = None
    if rank == 0 and not args.no_cpu:
        from oracle import ogmm_oracle as orc
        sys.modules.setdefault("tests_gpu_util", _load_gpu_util())
        cpu_pairs = args.cpu_pairs or (B if wl.cfg != 4 else 1)      # the workload's own batch: ~3 s per pass on 16 cores
        times, cores, ref_out = time_cpu(wl, host, cpu_pairs, 3 if wl.cfg != 4 else 1, 1 if wl.cfg != 4 else 0)
        best = min(times)
        cpu = {"value": cpu_pairs / best, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"the first {cpu_pairs} pairs of this rank's batch (same tensors as the GPU step), best of {len(times)} "
                         f"after {1 if wl.cfg != 4 else 0} warm-up ({best:.2f} s)"}
        if cpu_pairs == B:                               # same batch on both arms (the Sinkhorn exit test is a batch mean)
            try:
                parity = wl.parity(torch, orc, final, ref_out, host)
            except Exception as e:
                parity = {"error": f"{type(e).__name__}: {e}"}
        else:
            parity = {"skipped": f"the CPU leg ran {cpu_pairs} of the {B} pairs of the step: the Sinkhorn exit test is a mean over the "
                                 "batch of one call (lib/utils.py:99-102), so outputs are only comparable on identical batches; parity at "
                                 "this configuration's size is held by tests/test_gpu_parity.py (test_cluster_cfg4_size, "
                                 "test_knn_cfg4_size, test_knn_tiled_sweep_equals_exhaustive)"}
        if not args.no_cuda_ref and wl.cfg != 4:
            try:
                cuda_ref = time_cuda_reference(wl, host, dev)
                cuda_ref["our_value_over_it"] = (value / world) / cuda_ref["value"]
            except Exception as e:
                cuda_ref = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_model and wl.cfg == 2 and world == 1:
            try:
                model_fwd = time_model_forward(dev)
            except Exception as e:
                model_fwd = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(wl.config(B), launch=("cuda_graph_replay (one multi-stream step per graph: src / tgt chains on two streams, the kNN graphs on two more when a clustering launch cannot fill the GPU)" if graphed is not None
                                                 else (graph_note or "eager, two streams") if not args.no_overlap
                                                 else "eager, one stream")),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "em_step": em, "kernels": kernels,
            "cpu_baseline": cpu, "cuda_reference": cuda_ref, "gmmreg_forward": model_fwd, "parity": parity,
            "e2e": e2e, "e2e_model_boundary": e2e_mb,
            "gpu_launches": sum(launches.values()) * args.steps,
            "clocks": clocks, "eval_metrics_allreduced": metrics,
            "eval_metrics_note": "synthetic point features / logits carry no geometry, so the registration errors are "
                                 "meaningless here; the vector only exercises the path's one collective (4-float all-reduce)",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _load_gpu_util():
    """tests/gpu_util.py (decidable_rows, rot_err_deg): the parity helpers the tests use, shared with the parity leg."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("tests_gpu_util", os.path.join(ROOT, "tests", "gpu_util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_sweep(args, torch, dist, dev, world, rank, local, barrier, allmax):
    """configs[4]: B = 1 .. 8192 pairs per GPU of the flagship path, CUDA-graph replay, one JSON line."""
    from ogmm_b200 import pipeline
    wl = Flagship(2)
    base = wl.host_inputs(rank * 16, 16, 16)
    base = {k: v.to(dev) for k, v in base.items() if k in wl.input_names}
    batches = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
    if args.pairs:
        batches = [b for b in batches if b <= args.pairs]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    rows = []
    sampler.begin()
    for Bc in batches:
        d = {k: v.repeat(-(-Bc // 16), *([1] * (v.dim() - 1)))[:Bc].contiguous() for k, v in base.items()}
        g = wl.graphed(d)
        reps = max(3, min(args.steps, 16384 // Bc))
        for _ in range(max(3, args.warmup)):
            g.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        barrier()
        ms = allmax(e0.elapsed_time(e1)) / reps
        rows.append({"pairs_per_gpu": Bc, "steps": reps, "ms_per_step": ms, "pairs_per_s": Bc * world / ms * 1e3})
        del g, d
        torch.cuda.empty_cache()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and not args.no_cpu:
        host = wl.host_inputs(0, 64, 16)
        sweep_cpu = []
        for pairs in (1, 8, 64):
            times, cores, _ = time_cpu(wl, host, pairs, 2, 1)
            sweep_cpu.append({"pairs": pairs, "pairs_per_s": pairs / min(times)})
        cpu = {"value": sweep_cpu[-1]["pairs_per_s"], "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "1 / 8 / 64 pairs per call, best of 2 after 1 warm-up", "sweep": sweep_cpu}
    top = rows[-1]
    return {"metric": METRIC, "value": top["pairs_per_s"], "unit": "pairs/s", "n_gpus": world, "steps": top["steps"],
            "warmup": max(3, args.warmup), "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(wl.config(top["pairs_per_gpu"], "batch sweep (BASELINE.json configs[4]); value = the largest batch"),
                           config_id=5, launch="cuda_graph_replay"),
            "sweep": rows, "cpu_baseline": cpu, "clocks": clocks, "roofline": None, "e2e": None,
            "gpu_launches": sum(wl.launches().values()) * sum(r["steps"] for r in rows)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[config - 1]")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA-graph replay")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU per step (default: the config's own batch)")
    ap.add_argument("--distinct", type=int, default=None, help="distinct synthetic pairs generated per rank (tiled to --pairs)")
    ap.add_argument("--cpu-pairs", type=int, default=None,
                    help="pairs per CPU step (default: the whole batch for the cpu_baseline / parity leg; for --impl "
                         "reference as many as keep steps + warmup within ~2 minutes)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline, parity, cuda_reference and model legs")
    ap.add_argument("--no-cuda-ref", action="store_true")
    ap.add_argument("--no-model", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the src and tgt chains back to back on one stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
