"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

The reference (pure Python / PyTorch) is imported from where it lies, with
empty stub modules for the four third-party imports that are absent here and
untouched by the hot path (``transforms3d``, ``open3d``, ``h5py``, ``plyfile``;
SURVEY.md section 8(c)).  Each hot-path function is run on seeded inputs and
the inputs + outputs are stored as small ``.npz`` files.  The fixtures travel
to the GPU box; the reference does not.

``baseline/deepgmr.py:30-31`` hard-codes ``.cuda()``; while that one function
runs, ``torch.Tensor.cuda`` is replaced by the identity (nothing in the
reference is edited).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("OGMM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """The unmodified reference from where it lies (never the vendored copy: fixtures come from the source)."""
    sys.path.insert(0, ROOT)
    from oracle import refload
    m = refload.import_reference(REF)
    return m["utils"], m["se3"], m["dgcnn"], m["deepgmr"]


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: npy(v) for k, v in arrs.items()})
    print(f"  wrote {name}.npz  ({', '.join(f'{k}{tuple(npy(v).shape)}' for k, v in arrs.items())})")


def main():
    sys.path.insert(0, ROOT)
    from ogmm_b200 import synth
    ru, rs, rd, rg = import_reference()
    torch.manual_seed(1234)
    torch.set_num_threads(1)   # keep ATen's reduction order independent of the host core count
    g = torch.Generator().manual_seed(20221211)

    def randn(*s):
        return torch.randn(*s, generator=g)

    def rand(*s):
        return torch.rand(*s, generator=g)

    # ---- a1/a2: square_distance + knn -------------------------------------------------
    src, tgt, _, _ = synth.modelnet_batch(0, 2, 128)
    xs = torch.from_numpy(src).transpose(1, 2).contiguous()        # (2,128,3)
    xt = torch.from_numpy(tgt).transpose(1, 2).contiguous()[:, :96]  # ragged: M != N
    save("knn_xyz", src=xs, dst=xt, k=np.int64(8),
         dist=ru.square_distance(xs, xt), idx=ru.knn(xs, xt, 8),
         idx_self=ru.knn(xs, xs, 20))
    fa = torch.nn.functional.normalize(randn(2, 96, 16), dim=-1)
    save("knn_cosine", src=fa, k=np.int64(5), dist=ru.square_distance(fa, fa, True),
         idx=ru.knn(fa, fa, 5, True))
    wide = torch.relu(randn(1, 256, 64))
    save("knn_wide", src=wide, k=np.int64(20), idx=ru.knn(wide, wide, 20))

    # ---- a3: get_graph_feature -----------------------------------------------------------
    x = torch.from_numpy(src)                                      # (2,3,128)
    idx = ru.knn(x.transpose(-1, -2), x.transpose(-1, -2), k=8)
    save("edge_xyz", x=x, k=np.int64(8), idx=idx.clone(), feat=ru.get_graph_feature(x, 8, idx.clone()),
         feat_auto=ru.get_graph_feature(x, 5))
    x9 = torch.cat([randn(2, 6, 128), x], dim=1)
    save("edge_extra", x=x9, k=np.int64(6), feat=ru.get_graph_feature(x9, 6, None, True))

    # ---- a5: farthest_point_sample / index_points -----------------------------------------
    pts = xs.clone()
    start = torch.tensor([5, 77])
    ids_c = ru.farthest_point_sample(pts, 16, True)
    torch.manual_seed(99)
    ids_r = ru.farthest_point_sample(pts, 12, False)
    torch.manual_seed(99)
    drawn = torch.randint(0, 128, (2,), dtype=torch.long)
    save("fps", xyz=pts, ids_center=ids_c, ids_random=ids_r, start_random=drawn,
         gathered=ru.index_points(pts, ids_c))
    del start

    # ---- a6: sinkhorn --------------------------------------------------------------------
    cost = torch.cdist(xs, ru.index_points(xs, ids_c[:, :8]))
    p = torch.sigmoid(randn(2, 128))
    p = p / p.sum(-1, keepdim=True)
    gam, loss = ru.sinkhorn(cost, p=p, q=None, max_iter=10)
    gam100, loss100 = ru.sinkhorn(cost, p=p, q=None, max_iter=100)
    gam_e, loss_e = ru.sinkhorn(cost, p=p, q=None, epsilon=0.5, thresh=1e-2, max_iter=50)   # exits early
    qq = torch.softmax(randn(2, 8), -1)
    gam_q, loss_q = ru.sinkhorn(cost, p=None, q=qq, epsilon=1e-2, thresh=1e-2, max_iter=30)
    save("sinkhorn", cost=cost, p=p, q=qq, gamma10=gam, loss10=loss, gamma100=gam100, loss100=loss100,
         gamma_early=gam_e, loss_early=loss_e, gamma_q=gam_q, loss_q=loss_q)

    # ---- a8/a9: gmm_params / og_params ---------------------------------------------------
    gamma = torch.softmax(randn(2, 128, 8) * 3, -1)
    feats = torch.relu(randn(2, 128, 32))
    pi, mu, sigma = ru.gmm_params(gamma, xs, True)
    pif, muf = ru.gmm_params(gamma, feats)
    o = torch.sigmoid(randn(2, 128))
    opi, omu, ofe = ru.og_params(xs, gamma, o, feats)
    save("moments", gamma=gamma, xyz=xs, feats=feats, o=o, pi=pi, mu=mu, sigma=sigma, mu_feats=muf,
         og_pi=opi, og_mu=omu, og_feats=ofe)

    # ---- a7: wkeans_plus -----------------------------------------------------------------
    for tag, n, j, d, nb in (("small", 256, 8, 32, 2), ("full", 1024, 16, 64, 1)):
        s3, _, _, _ = synth.modelnet_batch(10, nb, n)
        xyz = torch.from_numpy(s3).transpose(1, 2).contiguous()
        ft = torch.relu(randn(nb, d, n))
        osc = torch.sigmoid(randn(nb, n))
        gm, pk, nx, nf = ru.wkeans_plus(xyz, ft.transpose(-1, -2), osc, j, iters=10, tau=1.0)
        save(f"wkeans_{tag}", xyz=xyz, feats=ft, o=osc, J=np.int64(j), gamma=gm, pi=pk, node_xyz=nx, node_feats=nf)

    # ---- a12: compute_rigid_transformation ------------------------------------------------
    a = randn(6, 3, 16)
    ang = rand(6, 3) * 0.7
    rots = []
    for e in ang.numpy():
        cx, cy, cz, sx, sy, sz = *np.cos(e), *np.sin(e)
        rots.append(np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
                    @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
    rgt = torch.tensor(np.stack(rots), dtype=torch.float32)
    tgt_t = randn(6, 3, 1)
    b = rgt @ a + tgt_t + 0.01 * randn(6, 3, 16)
    b[4] = b[4] * torch.tensor([1.0, 1.0, -1.0]).view(3, 1)     # reflection: forces the det fix
    a[5, 2] = 0.0                                                   # planar source: rank-2 covariance
    b[5] = rgt[5] @ a[5] + tgt_t[5]
    w = rand(6, 1, 16)
    rr, tt = rs.compute_rigid_transformation(a, b, w)
    save("procrustes", src=a, corr=b, weight=w, rot=rr, t=tt, rot_gt=rgt, t_gt=tgt_t)

    # ---- a11: GMMSVD ---------------------------------------------------------------------
    mu_s = randn(3, 16, 3)
    mu_t = (rgt[:3] @ mu_s.transpose(1, 2) + tgt_t[:3]).transpose(1, 2).contiguous()
    ds = torch.relu(randn(3, 16, 64))
    perm = torch.randperm(16, generator=g)
    dt = ds[:, perm] + 0.05 * randn(3, 16, 64)
    mu_t = mu_t[:, perm].contiguous()
    pis = torch.softmax(randn(3, 16), -1)
    pit = torch.softmax(randn(3, 16), -1)
    r0, t0, c0, m0 = rd.GMMSVD(False)(mu_s, mu_t, ds, dt, pis, pit)
    r1, t1, c1, _ = rd.GMMSVD(True)(mu_s, mu_t, ds, dt, pis, pit)
    save("gmmsvd", src=mu_s, tgt=mu_t, src_desc=ds, tgt_desc=dt, src_pi=pis, tgt_pi=pit,
         rot=r0, t=t0, corr=c0, tgt_t=m0, rot_sk=r1, t_sk=t1, corr_sk=c1,
         sim=ru.cos_similarity(ds, dt))

    # ---- a13/a14: DeepGMR E+M and gmm_register ---------------------------------------------
    s4, t4, _, _ = synth.icl_nuim_batch(0, 2, 256)
    ps, pt = torch.from_numpy(s4), torch.from_numpy(t4)
    ls, lt = randn(2, 8, 256) * 2, randn(2, 8, 256) * 2
    gs, gt_ = torch.softmax(ls, 1), torch.softmax(lt, 1)
    pi_s, m_s, sg_s = ru.gmm_params(gs.transpose(-1, -2), ps.transpose(-1, -2), True)
    pi_t, m_t, sg_t = ru.gmm_params(gt_.transpose(-1, -2), pt.transpose(-1, -2), True)
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a_, **k_: self
    try:
        tf = rg.gmm_register(pi_s, m_s, m_t, sg_t)
    finally:
        torch.Tensor.cuda = saved_cuda
    save("deepgmr", src=ps, tgt=pt, src_logits=ls, tgt_logits=lt, src_gamma=gs, src_pi=pi_s, src_mu=m_s,
         src_sigma=sg_s, tgt_pi=pi_t, tgt_mu=m_t, tgt_sigma=sg_t, transform=tf)

    # ---- anchors: get_local_corrs / get_anchor_corrs ---------------------------------------
    xyz3 = torch.from_numpy(src)                                  # (2,3,128)
    f3 = randn(2, 24, 128)
    torch.manual_seed(7)
    anc, pos, amu = ru.get_anchor_corrs(xyz3, f3, 16, dst='eu', iters=10, is_fast=True)
    torch.manual_seed(7)
    st = torch.randint(0, 128, (2,), dtype=torch.long)
    save("anchors", xyz=xyz3, feats=f3, start=st, anchor=anc, pos=pos, mu=amu)

    # ---- a15: SE(3) helpers ------------------------------------------------------------------
    g1 = torch.cat([rgt[:2], tgt_t[:2]], dim=-1)
    g2 = torch.cat([rgt[2:4], tgt_t[2:4]], dim=-1)
    cloud = randn(2, 50, 3)
    save("se3", g1=g1, g2=g2, cloud=cloud, inv=rs.torch_inverse(g1), cat=rs.torch_concatenate(g1, g2),
         moved=rs.torch_transform(g1, cloud), integ=rs.integrate_trans(rgt[:2], tgt_t[:2]))
    print("done")


if __name__ == "__main__":
    main()
