"""install()/uninstall() against the real reference modules, when the reference checkout is present (authoring
container only; skipped on the GPU box).  Name rebinding and the autograd fall-through are checked here on the CPU;
tests/test_gpu_model_dropin.py runs the patched reference models on a GPU."""
import os
import sys
import types

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_install_rebinds_every_import_site_and_uninstall_restores():
    for name in ("transforms3d", "transforms3d.quaternions", "open3d", "h5py", "plyfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["transforms3d"].quaternions = sys.modules["transforms3d.quaternions"]
    sys.modules["plyfile"].PlyData = getattr(sys.modules["plyfile"], "PlyData", object)
    sys.modules["plyfile"].PlyElement = getattr(sys.modules["plyfile"], "PlyElement", object)
    sys.path.insert(0, REF)
    try:
        import lib.utils, lib.se3, lib.loss, models.dgcnn, models.attn, models.gmmreg, baseline.deepgmr  # noqa: E401,F401
        import ogmm_b200 as og
        from ogmm_b200 import install
        orig_knn = models.dgcnn.knn
        orig_wk = models.gmmreg.wkeans_plus
        done = install.install()
        assert "models.dgcnn.knn" in done and "models.gmmreg.wkeans_plus" in done and "baseline.deepgmr.gmm_register" in done
        kern = lambda f: getattr(f, "ogmm_kernel", f)          # differentiable names are dispatchers around the kernel
        assert models.dgcnn.knn is og.utils.knn and lib.utils.knn is og.utils.knn          # index-only: bound directly
        assert kern(models.attn.get_graph_feature) is og.utils.get_graph_feature
        assert kern(models.gmmreg.wkeans_plus) is og.utils.wkeans_plus and issubclass(models.gmmreg.GMMSVD, og.modules.GMMSVD)
        assert kern(lib.loss.gmm_params) is og.utils.gmm_params and kern(baseline.deepgmr.gmm_register) is og.modules.gmm_register
        assert kern(models.dgcnn.compute_rigid_transformation) is og.se3.compute_rigid_transformation
        # a call that autograd has to record goes to the reference's own function (train.py keeps working) ...
        import torch
        g = torch.softmax(torch.rand(2, 32, 4), -1)
        f = torch.rand(2, 32, 3, requires_grad=True)           # narrow points with grad: not a call the kernels differentiate
        pi, mu = lib.loss.gmm_params(g, f)
        assert mu.grad_fn is not None
        mu.sum().backward()
        assert f.grad is not None and tuple(f.grad.shape) == (2, 32, 3)
        pi, mu, sg = lib.loss.gmm_params(g.requires_grad_(), torch.rand(2, 32, 8), True)       # gamma with grad, sigma: reference
        assert sg.grad_fn is not None
        g = g.detach()
        # ... except the feature M-step, which brings its own backward (ogmm_b200/autograd.py): wide features with grad
        # reach the kernels (here: their CUDA-only check)
        with pytest.raises(TypeError, match="CUDA"):
            lib.loss.gmm_params(g, torch.rand(2, 32, 8, requires_grad=True))
        f = torch.rand(2, 32, 8)
        # the softmax head (how models/gmmreg.py:41 builds it) has a backward kernel too; the Sinkhorn head does not and
        # stays with the reference's forward when autograd records
        d = torch.rand(2, 4, 8, requires_grad=True)
        with pytest.raises(TypeError, match="CUDA"):
            models.gmmreg.GMMSVD(False)(torch.rand(2, 4, 3), torch.rand(2, 4, 3), d, torch.rand(2, 4, 8), None, None)
        pi4 = torch.full((2, 4), 0.25)
        # Sinkhorn head under autograd: the reference's forward runs (its own sinkhorn on the graph) up to the Procrustes
        # solve, which is differentiable here and therefore a kernel call
        with pytest.raises(TypeError, match="weight|src|corr"):
            models.gmmreg.GMMSVD(True)(torch.rand(2, 4, 3), torch.rand(2, 4, 3), d, torch.rand(2, 4, 8), pi4, pi4)
        # ... and without a graph the same name reaches the kernels (here: their CUDA-only check)
        with torch.no_grad():
            with pytest.raises(TypeError, match="CUDA"):
                lib.loss.gmm_params(g, f)
        # install(model=...) switches existing instances to the kernel-backed class; uninstall switches them back
        ref_cls = install._saved[("models.gmmreg", "Clustering")]
        net = torch.nn.Sequential(ref_cls(16))
        assert "instance:Clustering" in install.install(model=net) and isinstance(net[0], og.modules.Clustering)
        # every patched name exists in the reference module it is patched into (no typos in the table)
        for mod_name, table in install.PATCH_TABLE.items():
            for attr in table:
                assert hasattr(sys.modules[mod_name], attr), f"{mod_name}.{attr} does not exist in the reference"
        install.uninstall()
        assert models.dgcnn.knn is orig_knn and models.gmmreg.wkeans_plus is orig_wk and type(net[0]) is ref_cls
    finally:
        sys.path.remove(REF)


def test_forward_only_guard_is_loud():
    """Differentiable reference functions refuse inputs that need a backward pass instead of dropping the graph;
    the check runs before anything touches CUDA, so it is testable here."""
    import pytest
    import torch
    from ogmm_b200 import utils, modules, se3
    g = torch.rand(2, 32, 4, requires_grad=True)
    x = torch.rand(2, 32, 8)
    with pytest.raises(RuntimeError, match="forward-only"):
        utils.gmm_params(g, x)                                         # a gamma that requires grad
    with pytest.raises(RuntimeError, match="forward-only"):
        utils.gmm_params(g.detach(), torch.rand(2, 32, 3, requires_grad=True))       # narrow points
    with pytest.raises(RuntimeError, match="forward-only"):
        utils.gmm_params(g.detach(), x.clone().requires_grad_(), True)               # sigma
    with pytest.raises(TypeError, match="CUDA"):                       # the feature M-step IS differentiable: reaches the kernels
        utils.gmm_params(g.detach(), x.clone().requires_grad_())
    # the Procrustes solve and the softmax head are differentiable (backward kernels): such calls reach the kernels
    with pytest.raises(TypeError, match="CUDA"):
        se3.compute_rigid_transformation(torch.rand(2, 3, 8, requires_grad=True), torch.rand(2, 3, 8), torch.rand(2, 1, 8))
    with pytest.raises(TypeError, match="CUDA"):
        modules.GMMSVD(is_sk=False)(torch.rand(2, 4, 3), torch.rand(2, 4, 3), torch.rand(2, 4, 8, requires_grad=True),
                                    torch.rand(2, 4, 8), torch.rand(2, 4), torch.rand(2, 4))
    with pytest.raises(RuntimeError, match="forward-only"):            # the Sinkhorn head has no backward
        modules.GMMSVD(is_sk=True)(torch.rand(2, 4, 3), torch.rand(2, 4, 3), torch.rand(2, 4, 8, requires_grad=True),
                                   torch.rand(2, 4, 8), torch.rand(2, 4), torch.rand(2, 4))
    # without a graph to record, the same call goes on to the kernels (and, on this CPU box, to the CUDA-only check)
    with torch.no_grad():
        with pytest.raises(TypeError, match="CUDA"):
            utils.gmm_params(g, x)
