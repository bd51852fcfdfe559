"""Numerical check (CPU, numpy float32) of the scaled-domain Sinkhorn used by the fast clustering kernel.

The kernel keeps G_ij = exp(-(c_ij - m_i)/eps), m_i = min_j c_ij, in registers for a whole Sinkhorn
call and iterates on scalings a_i, b_j with FMAs only (no exp per element per iteration):
    r_i = sum_j G_ij b_j ; a_i = (p_i + 1e-8) / r_i ; s_j = sum_i G_ij a_i ; b_j = (q + 1e-8) / s_j
    u_i = eps log a_i + m_i ; v_j = eps log b_j          (only for the change test)
This script runs that arithmetic in float32 next to the oracle (log domain, float32 and float64) on
the benchmark's synthetic clouds and prints the deviations, plus how often the safety monitor
(s_j, b_j out of range) would fire.  Run here; not imported by anything.
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ogmm_oracle as orc
from ogmm_b200 import synth

f32 = np.float32

def cdist_expanded(x, c):
    # ATen _euclidean_dist in fp32
    xn = (x * x).sum(-1, dtype=f32)[:, :, None]
    cn = (c * c).sum(-1, dtype=f32)[:, None, :]
    d2 = (f32(-2) * x) @ c.transpose(0, 2, 1) + xn + cn
    return np.sqrt(np.maximum(d2, f32(0)), dtype=f32)

def scaled_cluster(xyz, o, J, iters=10, inner=10, eps=f32(1e-2), thresh=1e-2):
    B, N, _ = xyz.shape
    ids = orc.fps_indices(torch.from_numpy(xyz), J, True).numpy()
    node = np.take_along_axis(xyz, ids[:, :, None], 1)
    p = (o / np.maximum(o.sum(-1, keepdims=True, dtype=f32), f32(1e-4))).astype(f32)
    P = p + f32(1e-8)
    Q = f32(1.0 / J) + f32(1e-8)
    trips = 0
    inv_eps = f32(1.0) / eps
    for _ in range(iters):
        c = cdist_expanded(xyz, node)
        m = c.min(-1, keepdims=True)
        G = np.exp((m - c) * inv_eps, dtype=f32)
        b = np.ones((B, 1, J), f32)
        u_old = np.zeros((B, N), f32); v_old = np.zeros((B, J), f32)
        for it in range(inner):
            r = (G * b).sum(-1, dtype=f32)
            a = P / r
            u = eps * np.log(a, dtype=f32) + m[..., 0]
            s = (G * a[..., None]).sum(1, dtype=f32)
            bn = Q / s
            v = eps * np.log(bn, dtype=f32)
            if not (np.all(s > 1e-30) and np.all(s < 1e30) and np.all(bn.max(-1) / bn.min(-1) < 1e24)):
                trips += 1
            diff = np.abs(u - u_old).sum(-1) + np.abs(v - v_old).sum(-1)
            u_old, v_old = u, v
            b = bn[:, None, :]
            if diff.mean() < thresh:
                break
        gam = a[..., None] * G * b
        gam = np.nan_to_num(gam, nan=0.0)
        gam = gam / np.maximum(gam.sum(-1, keepdims=True, dtype=f32), f32(1e-3))
        pi = gam.mean(1, dtype=f32)
        npi = pi * f32(N) + f32(1e-5)
        node = (gam.transpose(0, 2, 1) @ xyz) / npi[..., None]
    return gam, pi, node, trips

def report(tag, xyz, o, J):
    t = torch.from_numpy
    g32, p32, m32, _ = orc.sinkhorn_kmeans(t(xyz), t(xyz), t(o), J)
    g64, p64, m64, _ = orc.sinkhorn_kmeans(t(xyz).double(), t(xyz).double(), t(o).double(), J)
    gs, ps, ms, trips = scaled_cluster(xyz, o, J)
    sc = float(m64.abs().max())
    e = lambda a, b: float(np.abs(a - b.numpy()).max())
    print(f"{tag:28s} J={J:3d}  mu: scaled-vs-f32 {e(ms, m32)/sc:.2e}  scaled-vs-f64 {e(ms, m64)/sc:.2e}  "
          f"f32-vs-f64 {float((m32.double()-m64).abs().max())/sc:.2e} | pi {e(ps, p32)/float(p32.max()):.2e} | "
          f"gamma abs {e(gs, g32):.2e} (f32-vs-f64 {float((g32.double()-g64).abs().max()):.2e}) | monitor trips {trips}")

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (1024, 717):
        src, tgt, _, _ = synth.modelnet_batch(0, 8, n)
        xyz = np.ascontiguousarray(src.transpose(0, 2, 1))
        o = (1 / (1 + np.exp(-rng.normal(size=(8, n))))).astype(f32)
        report(f"modelnet N={n}", xyz, o, 16)
        report(f"modelnet N={n}", xyz, o, 64)
    src, _, _, _ = synth.icl_nuim_batch(0, 8, 1024)
    xyz = np.ascontiguousarray(src.transpose(0, 2, 1))
    o = (1 / (1 + np.exp(-rng.normal(size=(8, 1024))))).astype(f32)
    report("icl-nuim (metres) N=1024", xyz, o, 16)
    report("icl-nuim x10 scale", (xyz * 10).astype(f32), o, 16)
    o2 = o.copy(); o2[:, ::2] = 1e-6
    report("modelnet, half o ~ 0", np.ascontiguousarray(synth.modelnet_batch(3, 8, 1024)[0].transpose(0, 2, 1)), o2, 16)
