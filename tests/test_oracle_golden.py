"""The oracle must reproduce the reference's outputs stored in tests/golden/ (CPU only).

The fixtures were produced by ``oracle/make_golden.py`` from the unmodified
reference.  Integer outputs must match exactly; float outputs must match to a
few ulp (they are bit-identical on the machine that generated them -- the
slack only absorbs a different host CPU's BLAS kernel choice).
"""
import torch

from oracle import ogmm_oracle as orc

RT, AT = 2e-6, 2e-7


def close(a, b, rtol=RT, atol=AT):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def setup_module(_):
    torch.set_num_threads(1)


def test_sqdist_knn(golden):
    g = golden("knn_xyz")
    close(orc.pairwise_sqdist(g["src"], g["dst"]), g["dist"])
    assert torch.equal(orc.knn_indices(g["src"], g["dst"], 8), g["idx"])
    assert torch.equal(orc.knn_indices(g["src"], g["src"], 20), g["idx_self"])
    assert g["idx"].dtype == torch.int64


def test_knn_cosine_and_wide(golden):
    g = golden("knn_cosine")
    close(orc.pairwise_sqdist(g["src"], g["src"], True), g["dist"])
    assert torch.equal(orc.knn_indices(g["src"], g["src"], 5, True), g["idx"])
    w = golden("knn_wide")
    assert torch.equal(orc.knn_indices(w["src"], w["src"], 20), w["idx"])


def test_edge_features(golden):
    g = golden("edge_xyz")
    idx = g["idx"].clone()
    out = orc.edge_features(g["x"], 8, idx)
    assert out.shape == g["feat"].shape and torch.equal(out, g["feat"])
    assert torch.equal(idx, g["idx"])                    # the oracle does not mutate idx
    assert torch.equal(orc.edge_features(g["x"], 5), g["feat_auto"])
    e = golden("edge_extra")
    assert torch.equal(orc.edge_features(e["x"], 6, None, True), e["feat"])


def test_fps(golden):
    g = golden("fps")
    assert torch.equal(orc.fps_indices(g["xyz"], 16, True), g["ids_center"])
    assert torch.equal(orc.fps_indices(g["xyz"], 12, False, start=g["start_random"]), g["ids_random"])
    torch.manual_seed(99)
    assert torch.equal(orc.fps_indices(g["xyz"], 12, False), g["ids_random"])
    assert torch.equal(orc.gather_points(g["xyz"], g["ids_center"]), g["gathered"])


def test_sinkhorn(golden):
    g = golden("sinkhorn")
    gam, loss, it = orc.sinkhorn_log(g["cost"], p=g["p"], q=None, max_iter=10, return_iters=True)
    close(gam, g["gamma10"]); close(loss, g["loss10"]); assert it == 10
    gam, loss = orc.sinkhorn_log(g["cost"], p=g["p"], q=None, max_iter=100)
    close(gam, g["gamma100"]); close(loss, g["loss100"])
    gam, loss, it = orc.sinkhorn_log(g["cost"], p=g["p"], q=None, epsilon=0.5, thresh=1e-2, max_iter=50,
                                     return_iters=True)
    close(gam, g["gamma_early"]); close(loss, g["loss_early"])
    assert it < 50                                       # this case exercises the early exit
    gam, loss = orc.sinkhorn_log(g["cost"], p=None, q=g["q"], epsilon=1e-2, thresh=1e-2, max_iter=30)
    close(gam, g["gamma_q"]); close(loss, g["loss_q"])


def test_moments(golden):
    g = golden("moments")
    pi, mu, sigma = orc.gmm_moments(g["gamma"], g["xyz"], True)
    close(pi, g["pi"]); close(mu, g["mu"]); close(sigma, g["sigma"])
    close(orc.gmm_moments(g["gamma"], g["feats"])[1], g["mu_feats"])
    opi, omu, ofe = orc.overlap_moments(g["xyz"], g["gamma"], g["o"], g["feats"])
    close(opi, g["og_pi"]); close(omu, g["og_mu"]); close(ofe, g["og_feats"])


def test_sinkhorn_kmeans(golden):
    for tag in ("small", "full"):
        g = golden(f"wkeans_{tag}")
        tr = []
        gam, pi, node, nf = orc.sinkhorn_kmeans(g["xyz"], g["feats"].transpose(-1, -2), g["o"], int(g["J"]),
                                                 iters=10, tau=1.0, trace=tr)
        close(gam, g["gamma"], rtol=1e-5, atol=1e-7); close(pi, g["pi"]); close(node, g["node_xyz"])
        close(nf, g["node_feats"], rtol=1e-5, atol=1e-6)
        assert len(tr) == 10


def test_procrustes(golden):
    g = golden("procrustes")
    rot, t = orc.rigid_from_corr(g["src"], g["corr"], g["weight"])
    close(rot, g["rot"], rtol=1e-5, atol=1e-6); close(t, g["t"], rtol=1e-5, atol=1e-6)
    assert torch.all(torch.det(rot[:4]) > 0.999)
    # noise-free-ish rows recover the ground-truth motion
    err = orc.rotation_error_deg(rot[:4], g["rot_gt"][:4])
    assert float(err.max()) < 1.5


def test_gmmsvd(golden):
    g = golden("gmmsvd")
    close(orc.cosine_similarity(g["src_desc"], g["tgt_desc"]), g["sim"])
    rot, t, corr, tt = orc.soft_svd_head(g["src"], g["tgt"], g["src_desc"], g["tgt_desc"], g["src_pi"], g["tgt_pi"])
    close(rot, g["rot"], rtol=1e-5, atol=1e-6); close(t, g["t"], rtol=1e-5, atol=1e-6)
    close(corr, g["corr"], rtol=1e-5, atol=1e-6); assert torch.equal(tt, g["tgt_t"])
    rot, t, corr, _ = orc.soft_svd_head(g["src"], g["tgt"], g["src_desc"], g["tgt_desc"], g["src_pi"], g["tgt_pi"],
                                        is_sk=True)
    close(rot, g["rot_sk"], rtol=1e-5, atol=1e-6); close(t, g["t_sk"], rtol=1e-5, atol=1e-6)
    close(corr, g["corr_sk"], rtol=1e-5, atol=1e-6)


def test_deepgmr(golden):
    g = golden("deepgmr")
    gam, pi, mu, sigma = orc.deepgmr_em(g["src_logits"], g["src"])
    close(gam, g["src_gamma"]); close(pi, g["src_pi"]); close(mu, g["src_mu"]); close(sigma, g["src_sigma"])
    tf = orc.deepgmr_register(g["src_pi"], g["src_mu"], g["tgt_mu"], g["tgt_sigma"])
    close(tf, g["transform"], rtol=1e-5, atol=1e-6)


def test_anchors(golden):
    g = golden("anchors")
    anc, pos, mu = orc.anchor_corrs(g["xyz"], g["feats"], 16, start=g["start"])
    assert torch.equal(anc, g["anchor"]) and torch.equal(pos, g["pos"]) and torch.equal(mu, g["mu"])


def test_se3(golden):
    g = golden("se3")
    close(orc.se3_inverse(g["g1"]), g["inv"]); close(orc.se3_concatenate(g["g1"], g["g2"]), g["cat"])
    close(orc.se3_transform(g["g1"], g["cloud"]), g["moved"])
    r, t = orc.se3_decompose(g["integ"])
    close(orc.se3_integrate(r, t), g["integ"])


def test_fp64_arbiter_runs(golden):
    """The same code in float64 is the arbiter for ill-conditioned cases."""
    g = golden("wkeans_small")
    out32 = orc.sinkhorn_kmeans(g["xyz"], g["feats"].transpose(-1, -2), g["o"], int(g["J"]))
    out64 = orc.sinkhorn_kmeans(g["xyz"].double(), g["feats"].transpose(-1, -2).double(), g["o"].double(), int(g["J"]))
    assert out64[0].dtype == torch.float64
    scale = out64[2].abs().max()
    assert float((out32[2].double() - out64[2]).abs().max() / scale) < 1e-4
