"""Backward kernels of the registration head against the reference's own gradients.

The oracle (oracle/ogmm_oracle.py) is the reference's torch op sequence, so autograd THROUGH it is what the reference
computes when train.py:69-75 back-propagates: run in float64 it is the arbiter, run in float32 it shows what the
reference itself can resolve (its spread to the arbiter is printed next to our error).  Our kernels take fp32 inputs
and differentiate in closed form (ogmm_b200/csrc/procrustes_bwd.cu).

Bar: 1e-4 relative to the largest gradient entry of the tensor, or 2x the reference's own fp32 spread if that is larger.
"""
import pytest
import torch

from gpu_util import within_bar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


def _ref_head_grads(orc, ms, mt, fs, ft, gR, gt, gc, dtype):
    xs = [x.detach().to(dtype).requires_grad_(True) for x in (ms, mt, fs, ft)]
    R, t, corr, _ = orc.soft_svd_head(*xs, is_sk=False)
    ((R * gR.to(dtype)).sum() + (t * gt.to(dtype)).sum() + (corr * gc.to(dtype)).sum()).backward()
    return [x.grad for x in xs]


@pytest.mark.parametrize("B,Js,Jt,D", [(6, 16, 16, 512), (3, 16, 12, 40), (2, 64, 64, 512), (2, 5, 33, 130), (1, 128, 96, 64)])
def test_soft_procrustes_backward_vs_reference_autograd(dev, B, Js, Jt, D):
    from oracle import ogmm_oracle as orc
    from ogmm_b200 import ops
    g = torch.Generator().manual_seed(100 + Js + D)
    ms = torch.randn(B, Js, 3, generator=g).to(dev)
    Rgt = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0].to(dev)
    # descriptors with real structure: target descriptors = noisy permuted source descriptors, so the softmax at
    # T = 0.05 is peaked but not one-hot (gradients flow) -- plain random descriptors give near-uniform rows
    fs = torch.randn(B, Js, D, generator=g).to(dev)
    pick = torch.randint(0, Js, (B, Jt), generator=g).to(dev)
    ft = torch.gather(fs, 1, pick[:, :, None].expand(B, Jt, D)) + 0.6 * torch.randn(B, Jt, D, generator=g).to(dev)
    mt = torch.gather(ms, 1, pick[:, :, None].expand(B, Jt, 3)) @ Rgt.transpose(1, 2) + 0.05 * torch.randn(B, Jt, 3, generator=g).to(dev)
    gR, gt, gc = (torch.randn(s, generator=g).to(dev) for s in ((B, 3, 3), (B, 3), (B, 3, Js)))

    # the forward at the same (also odd: Js * Jt % 4 != 0) shapes, against the fp64 arbiter
    from gpu_util import rot_err_deg
    R, t, corr, _ = ops.soft_procrustes(ms, mt, fs, ft, temperature=0.05)
    R64, t64, c64, _ = orc.soft_svd_head(ms.double(), mt.double(), fs.double(), ft.double(), is_sk=False)
    assert float(rot_err_deg(R.cpu(), R64.cpu()).max()) < 1e-3 and _rel(corr, c64) < 1e-4
    assert float((t.double() - t64).abs().max()) < 1e-5 * float(mt.abs().max())

    ours = ops.soft_procrustes_backward(ms, mt, fs, ft, gR, gt, gc, temperature=0.05)
    arb = _ref_head_grads(orc, ms, mt, fs, ft, gR, gt, gc, torch.float64)
    r32 = _ref_head_grads(orc, ms, mt, fs, ft, gR, gt, gc, torch.float32)
    print()
    for name, o, a, r in zip(("src_mu", "tgt_mu", "src_desc", "tgt_desc"), ours, arb, r32):
        assert o.shape == a.shape
        within_bar(_rel(o, a), 1e-4, _rel(r, a), f"d/d{name} (B={B}, Js={Js}, Jt={Jt}, D={D}) vs fp64 autograd")


def test_soft_procrustes_backward_partial_upstream(dev):
    """Missing upstream gradients (None) are zeros: only dL/dt, only dL/dR."""
    from oracle import ogmm_oracle as orc
    from ogmm_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, J, D = 4, 16, 64
    ms, mt = torch.randn(B, J, 3, generator=g).to(dev), torch.randn(B, J, 3, generator=g).to(dev)
    fs = torch.randn(B, J, D, generator=g).to(dev)
    ft = fs[:, torch.randperm(J, generator=g)] + 0.5 * torch.randn(B, J, D, generator=g).to(dev)
    gR, gt = torch.randn(B, 3, 3, generator=g).to(dev), torch.randn(B, 3, generator=g).to(dev)
    zero_c = torch.zeros(B, 3, J, device=dev)
    for up_R, up_t in ((None, gt), (gR, None)):
        ours = ops.soft_procrustes_backward(ms, mt, fs, ft, up_R, up_t, None)
        ups = (up_R if up_R is not None else torch.zeros_like(gR), up_t if up_t is not None else torch.zeros_like(gt), zero_c)
        arb = _ref_head_grads(orc, ms, mt, fs, ft, *ups, torch.float64)
        r32 = _ref_head_grads(orc, ms, mt, fs, ft, *ups, torch.float32)
        for name, o, a, r in zip(("src_mu", "tgt_mu", "src_desc", "tgt_desc"), ours, arb, r32):
            within_bar(_rel(o, a), 1e-4, _rel(r, a), f"d/d{name}, upstream {'dL/dt' if up_R is None else 'dL/dR'} only")


def test_gmmsvd_module_is_differentiable(dev):
    """modules.GMMSVD(is_sk=False) under autograd: loss.backward() reaches the descriptors through the backward kernel
    and equals the reference's autograd; the Sinkhorn variant still refuses (no backward)."""
    from oracle import ogmm_oracle as orc
    import ogmm_b200 as og
    g = torch.Generator().manual_seed(11)
    B, J, D = 8, 16, 512
    ms, mt = torch.randn(B, J, 3, generator=g).to(dev), torch.randn(B, J, 3, generator=g).to(dev)
    fs0 = torch.randn(B, J, D, generator=g).to(dev)
    ft0 = fs0[:, torch.randperm(J, generator=g)] + 0.7 * torch.randn(B, J, D, generator=g).to(dev)
    pi = torch.full((B, J), 1.0 / J, device=dev)
    Rt = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0].to(dev)
    tt = torch.randn(B, 3, generator=g).to(dev)

    def loss_of(R, t):
        return ((R - Rt) ** 2).sum() + ((t - tt) ** 2).sum()

    fs, ft = fs0.clone().requires_grad_(True), ft0.clone().requires_grad_(True)
    head = og.GMMSVD(is_sk=False)
    R, t, corr, tg = head(ms, mt, fs, ft, pi, pi)
    assert R.requires_grad and t.requires_grad and tuple(t.shape) == (B, 3)
    loss_of(R, t).backward()
    fs64, ft64 = fs0.double().requires_grad_(True), ft0.double().requires_grad_(True)
    R64, t64, _, _ = orc.soft_svd_head(ms.double(), mt.double(), fs64, ft64, is_sk=False)
    loss_of(R64, t64).backward()
    fs32, ft32 = fs0.clone().requires_grad_(True), ft0.clone().requires_grad_(True)
    R32, t32, _, _ = orc.soft_svd_head(ms, mt, fs32, ft32, is_sk=False)               # the reference's own fp32 autograd
    loss_of(R32, t32).backward()
    print()
    within_bar(_rel(fs.grad, fs64.grad), 1e-4, _rel(fs32.grad, fs64.grad), "GMMSVD(is_sk=False).backward d/dsrc_desc vs fp64 autograd")
    within_bar(_rel(ft.grad, ft64.grad), 1e-4, _rel(ft32.grad, ft64.grad), "GMMSVD(is_sk=False).backward d/dtgt_desc vs fp64 autograd")
    # no-grad call: same forward values, no history
    with torch.no_grad():
        R0, t0, _, _ = head(ms, mt, fs, ft, pi, pi)
    assert torch.equal(R0, R.detach()) and torch.equal(t0, t.detach()) and not R0.requires_grad
    with pytest.raises(RuntimeError, match="forward-only"):
        og.GMMSVD(is_sk=True)(ms, mt, fs, ft, pi, pi)


@pytest.mark.parametrize("B,n", [(5, 16), (3, 717), (2, 4096)])
def test_rigid_transform_backward_vs_reference_autograd(dev, B, n):
    from oracle import ogmm_oracle as orc
    import ogmm_b200 as og
    g = torch.Generator().manual_seed(n)
    src = torch.randn(B, 3, n, generator=g).to(dev)
    Rgt = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0].to(dev)
    corr = Rgt @ src + 0.1 * torch.randn(B, 3, n, generator=g).to(dev) + torch.randn(B, 3, 1, generator=g).to(dev)
    w = torch.rand(B, 1, n, generator=g).to(dev) + 0.05
    gR, gt = torch.randn(B, 3, 3, generator=g).to(dev), torch.randn(B, 3, 1, generator=g).to(dev)

    def ref_grads(dtype):
        xs = [x.detach().to(dtype).requires_grad_(True) for x in (src, corr, w)]
        R, t = orc.rigid_from_corr(*xs)
        ((R * gR.to(dtype)).sum() + (t * gt.to(dtype)).sum()).backward()
        return [x.grad for x in xs]

    xs = [x.clone().requires_grad_(True) for x in (src, corr, w)]
    R, t = og.compute_rigid_transformation(*xs)              # the public function records the call itself
    assert tuple(t.shape) == (B, 3, 1) and R.requires_grad
    ((R * gR).sum() + (t * gt).sum()).backward()
    arb, r32 = ref_grads(torch.float64), ref_grads(torch.float32)
    print()
    for name, x, a, r in zip(("src", "src_corr", "weight"), xs, arb, r32):
        within_bar(_rel(x.grad, a), 1e-4, _rel(r, a), f"d/d{name} (B={B}, n={n}) vs fp64 autograd")


def test_rigid_transform_backward_reflection_branch(dev):
    """det(V U^T) < 0 (lib/se3.py:281-285 takes rot_neg): correspondences are a mirror image of the source."""
    from oracle import ogmm_oracle as orc
    import ogmm_b200 as og
    g = torch.Generator().manual_seed(3)
    B, n = 4, 64
    src = torch.randn(B, 3, n, generator=g).to(dev) * torch.tensor([1.0, 0.7, 0.4], device=dev)[None, :, None]
    corr = src * torch.tensor([1.0, 1.0, -1.0], device=dev)[None, :, None] + 0.02 * torch.randn(B, 3, n, generator=g).to(dev)
    w = torch.rand(B, 1, n, generator=g).to(dev) + 0.1
    gR, gt = torch.randn(B, 3, 3, generator=g).to(dev), torch.randn(B, 3, 1, generator=g).to(dev)
    xs64 = [x.double().requires_grad_(True) for x in (src, corr, w)]
    R64, t64 = orc.rigid_from_corr(*xs64)
    cov = ((xs64[0] - (xs64[0] * xs64[2]).sum(2, keepdim=True) / xs64[2].sum(2, keepdim=True)) * xs64[2]) @ \
        (xs64[1] - (xs64[1] * xs64[2]).sum(2, keepdim=True) / xs64[2].sum(2, keepdim=True)).transpose(1, 2)
    assert bool((torch.det(cov.detach()) < 0).all()), "the case must exercise the reflection branch"
    ((R64 * gR.double()).sum() + (t64 * gt.double()).sum()).backward()
    xs = [x.clone().requires_grad_(True) for x in (src, corr, w)]
    R, t = og.compute_rigid_transformation(*xs)
    ((R * gR).sum() + (t * gt).sum()).backward()
    for x, a in zip(xs, xs64):
        assert _rel(x.grad, a.grad) < 2e-4


def test_backward_edge_shapes(dev):
    """Odd and degenerate shapes: descriptor length not a multiple of 4, a single component (the covariance is just the
    1e-5 I regulariser: equal singular values), an empty batch, and a descriptor row of zeros (F.normalize's clamp)."""
    from oracle import ogmm_oracle as orc
    from ogmm_b200 import ops
    g = torch.Generator().manual_seed(77)
    for (B, Js, Jt, D) in ((2, 7, 9, 3), (3, 1, 1, 8), (2, 1, 5, 6)):
        ms, mt = torch.randn(B, Js, 3, generator=g).to(dev), torch.randn(B, Jt, 3, generator=g).to(dev)
        fs, ft = torch.randn(B, Js, D, generator=g).to(dev), torch.randn(B, Jt, D, generator=g).to(dev)
        gR, gt, gc = (torch.randn(s, generator=g).to(dev) for s in ((B, 3, 3), (B, 3), (B, 3, Js)))
        ours = ops.soft_procrustes_backward(ms, mt, fs, ft, gR, gt, gc)
        assert all(bool(torch.isfinite(o).all()) for o in ours)
        if Js > 1:                                   # with one component R is the SVD of 1e-5 I: any gauge, gradients ill-defined
            arb = _ref_head_grads(orc, ms, mt, fs, ft, gR, gt, gc, torch.float64)
            r32 = _ref_head_grads(orc, ms, mt, fs, ft, gR, gt, gc, torch.float32)
            for name, o, a, r in zip(("src_mu", "tgt_mu", "src_desc", "tgt_desc"), ours, arb, r32):
                within_bar(_rel(o, a), 1e-4, _rel(r, a), f"d/d{name} (B={B}, Js={Js}, Jt={Jt}, D={D})")
    # empty batch
    e = ops.soft_procrustes_backward(torch.zeros(0, 4, 3, device=dev), torch.zeros(0, 4, 3, device=dev), torch.zeros(0, 4, 8, device=dev),
                                     torch.zeros(0, 4, 8, device=dev), None, None, None)
    assert tuple(e[2].shape) == (0, 4, 8)
    es = ops.rigid_transform_backward(torch.zeros(0, 3, 5, device=dev), torch.zeros(0, 3, 5, device=dev), torch.zeros(0, 1, 5, device=dev), None, None)
    assert tuple(es[0].shape) == (0, 3, 5)
    # a zero descriptor row: x / max(|x|, 1e-12) is linear there; gradients stay finite
    ms, mt = torch.randn(1, 4, 3, generator=g).to(dev), torch.randn(1, 4, 3, generator=g).to(dev)
    fs, ft = torch.randn(1, 4, 8, generator=g).to(dev), torch.randn(1, 4, 8, generator=g).to(dev)
    fs[0, 2] = 0.0
    out = ops.soft_procrustes_backward(ms, mt, fs, ft, torch.randn(1, 3, 3, generator=g).to(dev), torch.randn(1, 3, generator=g).to(dev), None)
    assert all(bool(torch.isfinite(o).all()) for o in out)


@pytest.mark.parametrize("B,N,J,with_sigma", [(3, 1024, 16, True), (2, 717, 5, True), (2, 300, 16, False), (1, 4096, 40, True)])
def test_narrow_moments_backward_vs_reference_autograd(dev, B, N, J, with_sigma):
    """gmm_params(gamma, xyz, return_sigma) differentiated with respect to gamma = softmax(logits), as DeepGMR trains it
    (baseline/deepgmr.py:71-74): through the public function, gamma held as the transposed view of a (B,J,N) tensor."""
    from oracle import ogmm_oracle as orc
    import ogmm_b200 as og
    g = torch.Generator().manual_seed(N + J)
    logits = torch.randn(B, J, N, generator=g).to(dev)
    pts = (torch.randn(B, 3, N, generator=g) * torch.tensor([1.0, 0.5, 0.25])[None, :, None]).to(dev)
    ups = [torch.randn(s, generator=g).to(dev) for s in ((B, J), (B, J, 3), (B, J, 3, 3))]

    def run(fn, dtype):
        lg = logits.detach().to(dtype).requires_grad_(True)
        gam = torch.softmax(lg, dim=1).transpose(-1, -2)                     # (B,N,J) view of (B,J,N)
        out = fn(gam, pts.to(dtype).transpose(-1, -2), with_sigma)
        loss = sum((o * u.to(dtype)).sum() for o, u in zip(out, ups))
        loss.backward()
        return lg.grad, out

    ours, out = run(og.gmm_params, torch.float32)
    assert out[1].grad_fn is not None and "NarrowMoments" in type(out[1].grad_fn).__name__
    arb, _ = run(orc.gmm_moments, torch.float64)
    r32, _ = run(orc.gmm_moments, torch.float32)
    print()
    within_bar(_rel(ours, arb), 1e-4, _rel(r32, arb), f"d/dlogits through the xyz moments (B={B}, N={N}, J={J}, sigma={with_sigma})")


def test_gmm_register_backward_vs_reference_autograd(dev):
    from oracle import ogmm_oracle as orc
    import ogmm_b200 as og
    g = torch.Generator().manual_seed(8)
    for (B, J) in ((5, 16), (3, 7), (2, 40)):
        pi = torch.softmax(torch.randn(B, J, generator=g), -1).to(dev)
        ms = torch.randn(B, J, 3, generator=g).to(dev)
        Rgt = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0].to(dev)
        mt = ms @ Rgt.transpose(1, 2) + 0.1 * torch.randn(B, J, 3, generator=g).to(dev) + torch.randn(B, 1, 3, generator=g).to(dev)
        a = torch.randn(B, J, 3, 3, generator=g) * 0.2
        sg = (a @ a.transpose(-1, -2) + 0.3 * torch.eye(3)).to(dev)
        gT = torch.randn(B, 4, 4, generator=g).to(dev)

        def run(fn, dtype):
            xs = [x.detach().to(dtype).requires_grad_(True) for x in (pi, ms, mt, sg)]
            T = fn(*xs)
            (T * gT.to(dtype)).sum().backward()
            return [x.grad for x in xs], T

        ours, T = run(og.gmm_register, torch.float32)
        assert "GmmRegister" in type(T.grad_fn).__name__
        arb, _ = run(orc.deepgmr_register, torch.float64)
        r32, _ = run(orc.deepgmr_register, torch.float32)
        print()
        for name, o, a_, r in zip(("pi_s", "mu_s", "mu_t", "sigma_t"), ours, arb, r32):
            within_bar(_rel(o, a_), 1e-4, _rel(r, a_), f"gmm_register d/d{name} (B={B}, J={J}) vs fp64 autograd")
