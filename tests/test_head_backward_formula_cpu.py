"""The closed-form gradient of the soft-correspondence head that ``ogmm_b200/csrc/procrustes_bwd.cu`` implements,
written out in float64 torch and held against autograd through the oracle (= the reference's own op sequence,
models/dgcnn.py:96-115 + lib/se3.py:256-289).  CPU only: it pins the DERIVATION; the kernel itself is compared with the
same autograd on the GPU (tests/test_gpu_backward.py).

    M = cov + 1e-5 I = U S V^T,  R = V D U^T,  Q = R^T = U D V^T  (polar factor of M = Q P,  P = V (D S) V^T)
    dQ = U D Omega V^T,  Omega_ij = (B_ij - B_ji) / (s'_i + s'_j),  B = D U^T dM V,  s' = (s1, s2, d s3)
    =>  G = D U^T gR^T V,  H_ij = (G_ij - G_ji) / (s'_i + s'_j),  dL/dM = U D H V^T
"""
import torch

from oracle import ogmm_oracle as orc

DT = torch.float64


def rotation_grad(cov, g_rot):
    u, s, vh = torch.linalg.svd(cov)
    v = vh.T
    d = 1.0 if torch.det(v @ u.T) > 0 else -1.0
    dm = torch.diag(torch.tensor([1.0, 1.0, d], dtype=DT))
    sp = torch.stack([s[0], s[1], d * s[2]])
    g = dm @ u.T @ g_rot.T @ v
    h = torch.zeros(3, 3, dtype=DT)
    for i in range(3):
        for j in range(3):
            if i != j:
                h[i, j] = (g[i, j] - g[j, i]) / (sp[i] + sp[j])
    return u @ dm @ h @ v.T, v @ dm @ u.T


def head_backward(mus, mut, fs, ft, g_rot, g_t, g_corr, temp=0.05):
    nx, ny = fs.norm(dim=1).clamp(min=1e-12), ft.norm(dim=1).clamp(min=1e-12)
    xh, yh = fs / nx[:, None], ft / ny[:, None]
    sim = xh @ yh.T
    p = torch.softmax(sim / temp, dim=1)
    corr = p @ mut
    w = p.sum(1)
    wsum = w.sum()
    a = (w[:, None] * mus).sum(0) / wsum
    c = (w[:, None] * corr).sum(0) / wsum
    ds, dc = mus - a, corr - c
    cov = (w[:, None] * ds).T @ dc + 1e-5 * torch.eye(3, dtype=DT)
    # t = -R a + c:  dL/dR -= g_t a^T,  dL/da = -R^T g_t,  dL/dc = g_t
    g_m, rot = rotation_grad(cov, g_rot - torch.outer(g_t, a))
    ga, gc = -rot.T @ g_t, g_t
    # the covariance terms through the weighted centroids vanish: sum_n w_n (s_n - a) = 0
    gw = torch.einsum('ni,ij,nj->n', ds, g_m, dc) + ds @ ga / wsum + dc @ gc / wsum
    g_corr_n = w[:, None] * (ds @ g_m) + (w / wsum)[:, None] * gc + g_corr.T
    g_src = w[:, None] * (dc @ g_m.T) + (w / wsum)[:, None] * ga
    g_p = g_corr_n @ mut.T + gw[:, None]
    g_tgt = p.T @ g_corr_n
    # softmax backward in the cancellation-free form the kernel uses: gz_j = P_j sum_k P_k (gP_j - gP_k) / T
    g_sim = p * ((g_p[:, :, None] - g_p[:, None, :]) * p[:, None, :]).sum(-1) / temp
    # F.normalize backward with <x_hat_n, sum_m gsim_nm y_hat_m> = sum_m gsim_nm sim_nm
    g_fs = (g_sim @ yh - xh * (g_sim * sim).sum(1)[:, None]) / nx[:, None]
    g_ft = (g_sim.T @ xh - yh * (g_sim * sim).sum(0)[:, None]) / ny[:, None]
    return g_src, g_tgt, g_fs, g_ft


def _case(seed, js, jt, d, mirror=False):
    g = torch.Generator().manual_seed(seed)
    mus = torch.randn(js, 3, generator=g, dtype=DT)
    fs = torch.randn(js, d, generator=g, dtype=DT)
    pick = torch.randint(0, js, (jt,), generator=g)
    ft = fs[pick] + 0.6 * torch.randn(jt, d, generator=g, dtype=DT)
    mut = mus[pick] + 0.05 * torch.randn(jt, 3, generator=g, dtype=DT)
    if mirror:                                   # forces det(V U^T) < 0: the rot_neg branch of lib/se3.py:281-285
        mut = mut * torch.tensor([1.0, 1.0, -1.0], dtype=DT)
    ups = [torch.randn(s, generator=g, dtype=DT) for s in ((3, 3), (3,), (3, js))]
    return mus, mut, fs, ft, ups


def _check(mus, mut, fs, ft, ups):
    xs = [x.clone().requires_grad_(True) for x in (mus, mut, fs, ft)]
    rot, t, corr, _ = orc.soft_svd_head(*(x[None] for x in xs), is_sk=False)
    ((rot[0] * ups[0]).sum() + (t[0] * ups[1]).sum() + (corr[0] * ups[2]).sum()).backward()
    ours = head_backward(mus, mut, fs, ft, *ups)
    for mine, x in zip(ours, xs):
        assert float((mine - x.grad).abs().max()) <= 1e-11 * max(1.0, float(x.grad.abs().max()))
    return rot[0]


def test_closed_form_equals_autograd_of_the_reference_head():
    for seed, (js, jt, d) in enumerate(((16, 16, 64), (16, 12, 40), (5, 33, 24), (48, 48, 32))):
        _check(*_case(seed, js, jt, d))


def test_closed_form_on_the_reflection_branch():
    mus, mut, fs, ft, ups = _case(7, 16, 16, 32, mirror=True)
    p = torch.softmax((fs / fs.norm(dim=1, keepdim=True)) @ (ft / ft.norm(dim=1, keepdim=True)).T / 0.05, dim=1)
    corr, w = p @ mut, p.sum(1)
    cov = (w[:, None] * (mus - (w[:, None] * mus).sum(0) / w.sum())).T @ (corr - (w[:, None] * corr).sum(0) / w.sum())
    assert float(torch.det(cov)) < 0, "the case must exercise the det fix"
    rot = _check(mus, mut, fs, ft, ups)
    assert float(torch.det(rot.detach())) > 0.999   # the head still returns a proper rotation there
