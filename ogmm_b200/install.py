"""Patch the sm_100a kernels into an imported reference checkout.

The reference binds its helpers with ``from lib.utils import ...`` at import time, so replacing
``lib.utils.knn`` alone is not enough: every importing module holds its own name (SURVEY.md
section 8(b)).  ``install()`` rebinds the hot-path names in each of those namespaces and
``uninstall()`` restores the originals.

    import lib.utils, lib.se3, models.dgcnn, models.attn, models.gmmreg, lib.loss, baseline.deepgmr
    import ogmm_b200.install as inst
    inst.install()            # after the reference modules are imported
    inst.install(model=net)   # ... and re-class the GMMSVD / Clustering instances of an existing model

Two things to know:

* **Training keeps working.**  The kernels are forward / inference code.  A patched name is a dispatcher: when
  autograd is recording and an argument requires grad (``train.py:57-75``: the loss calls ``gmm_params``,
  ``get_local_corrs``; the model differentiates through ``gmm_params(gamma, feats)``, ``GMMSVD`` and
  ``compute_rigid_transformation``) the call goes to the reference's own function that the name held before -- exactly
  what would have run without ``install()``.  Replacements that carry their own backward (``ogmm_b200/autograd.py``:
  ``wkeans_plus`` and ``gmm_params`` on wide features = the feature M-step, ``GMMSVD(is_sk=False)`` = the
  soft-correspondence head, ``compute_rigid_transformation``) take those calls themselves.  Under ``torch.no_grad()`` /
  inference every call runs on the kernels.
* **Classes.**  ``models.gmmreg.Clustering`` and ``GMMSVD`` are replaced by subclasses of ``ogmm_b200.modules`` that
  dispatch the same way; they take effect for models constructed AFTER ``install()``.  A model built earlier still
  benefits (its modules call the patched function names), and ``install(model=net)`` additionally switches the class
  of its existing ``GMMSVD`` / ``Clustering`` instances so the fused registration-head kernel is used.
"""
from __future__ import annotations

import functools
import sys

import torch

from . import modules as _modules
from . import se3 as _se3
from . import utils as _utils

# reference module -> {attribute name -> replacement}
_UTIL_NAMES = ["knn", "get_graph_feature", "sinkhorn", "gmm_params", "og_params", "farthest_point_sample",
               "cos_similarity", "get_local_corrs", "get_anchor_corrs", "wkeans_plus"]
PATCH_TABLE = {
    "lib.utils": {n: getattr(_utils, n) for n in _UTIL_NAMES},
    "lib.se3": {"compute_rigid_transformation": _se3.compute_rigid_transformation},
    "models.dgcnn": {"knn": _utils.knn, "get_graph_feature": _utils.get_graph_feature, "sinkhorn": _utils.sinkhorn,
                     "cos_similarity": _utils.cos_similarity, "og_params": _utils.og_params,
                     "compute_rigid_transformation": _se3.compute_rigid_transformation, "GMMSVD": _modules.GMMSVD},
    "models.attn": {"get_graph_feature": _utils.get_graph_feature},
    "models.gmmreg": {"get_anchor_corrs": _utils.get_anchor_corrs, "wkeans_plus": _utils.wkeans_plus,
                      "GMMSVD": _modules.GMMSVD, "Clustering": _modules.Clustering},
    "lib.loss": {"gmm_params": _utils.gmm_params, "get_local_corrs": _utils.get_local_corrs},
    "baseline.deepgmr": {"gmm_params": _utils.gmm_params, "gmm_register": _modules.gmm_register},
}

_saved = {}
_reclassed = []
_saved_methods = {}


def _needs_grad(args, kwargs):
    if not torch.is_grad_enabled():
        return False
    return any(isinstance(a, torch.Tensor) and a.requires_grad for a in list(args) + list(kwargs.values()))


def _dispatcher(ours, ref):
    """ours under no_grad / on detached inputs; the reference's own function when autograd has to record the call
    (unless ``ours`` brings a backward of its own)."""
    if getattr(ours, "ogmm_autograd_safe", False) or not callable(ref):
        return ours          # index-only results (knn, FPS) or a replacement with its own backward

    can = getattr(ours, "ogmm_can_differentiate", None)   # a replacement that differentiates SOME calls says which

    @functools.wraps(ref)
    def call(*args, **kwargs):
        if _needs_grad(args, kwargs) and not (can is not None and can(*args, **kwargs)):
            return ref(*args, **kwargs)
        return ours(*args, **kwargs)

    call.ogmm_kernel = ours
    call.ogmm_reference = ref
    return call


def _class_dispatcher(ours_cls, ref_cls):
    class Patched(ours_cls):
        __doc__ = ours_cls.__doc__

        def forward(self, *args, **kwargs):
            can = getattr(ours_cls, "ogmm_can_differentiate", None)   # classes with a backward of their own say when
            if _needs_grad(args, kwargs) and not (can is not None and can(self, *args, **kwargs)):
                return ref_cls.forward(self, *args, **kwargs)        # same attributes (is_sk / n_clusters): duck-typed self
            return ours_cls.forward(self, *args, **kwargs)

    Patched.__name__ = ours_cls.__name__
    Patched.__qualname__ = ours_cls.__qualname__
    Patched.ogmm_reference = ref_cls
    return Patched


def install(modules=None, model=None):
    """Rebind the hot-path names in every already-imported reference module.  Returns the list patched.

    ``model``: an already constructed reference model whose ``GMMSVD`` / ``Clustering`` instances should switch to the
    kernel-backed classes as well (``uninstall()`` switches them back)."""
    done = []
    for mod_name, table in PATCH_TABLE.items():
        if modules is not None and mod_name not in modules:
            continue
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for attr, repl in table.items():
            if not hasattr(mod, attr):
                continue
            orig = _saved.setdefault((mod_name, attr), getattr(mod, attr))
            new = _class_dispatcher(repl, orig) if isinstance(repl, type) else _dispatcher(repl, orig)
            setattr(mod, attr, new)
            done.append(f"{mod_name}.{attr}")
    # DGCNN.forward: the kNN graph and the fused edge gather + conv1 + bn1 + ReLU + max (N3) when the module runs in
    # eval mode without autograd; the reference's own forward (with the patched knn / get_graph_feature) otherwise
    dg = sys.modules.get("models.dgcnn")
    if dg is not None and (modules is None or "models.dgcnn" in modules) and hasattr(dg, "DGCNN"):
        cls = dg.DGCNN
        orig_forward = _saved_methods.setdefault((cls, "forward"), cls.forward)

        def forward(self, x, _orig=orig_forward):
            if self.training or _needs_grad((x,), {}) or (torch.is_grad_enabled() and any(p.requires_grad for p in self.conv1.parameters())):
                return _orig(self, x)
            return _modules.dgcnn_forward(self, x)

        cls.forward = forward
        done.append("models.dgcnn.DGCNN.forward")
    # PositionEncoding.forward: the k = 5 angle feature (kNN graph + normalised offsets + conv_ang1 + max) in one kernel
    # under the same conditions
    at = sys.modules.get("models.attn")
    if at is not None and (modules is None or "models.attn" in modules) and hasattr(at, "PositionEncoding"):
        cls = at.PositionEncoding
        orig_pe = _saved_methods.setdefault((cls, "forward"), cls.forward)

        def pe_forward(self, points, k=5, _orig=orig_pe):
            if self.training or _needs_grad((points,), {}) or (torch.is_grad_enabled() and any(p.requires_grad for p in self.conv_ang1.parameters())):
                return _orig(self, points, k)
            return _modules.position_encoding_forward(self, points, k)

        cls.forward = pe_forward
        done.append("models.attn.PositionEncoding.forward")
    if model is not None:
        for m in model.modules():
            for (mod_name, attr), orig in _saved.items():
                if isinstance(orig, type) and type(m) is orig:
                    _reclassed.append((m, orig))
                    m.__class__ = getattr(sys.modules[mod_name], attr)
                    done.append(f"instance:{type(m).__name__}")
                    break
    return done


def uninstall():
    for (cls, name), orig in list(_saved_methods.items()):
        setattr(cls, name, orig)
        del _saved_methods[(cls, name)]
    for m, orig in _reclassed:
        m.__class__ = orig
    _reclassed.clear()
    for (mod_name, attr), orig in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, orig)
        del _saved[(mod_name, attr)]
