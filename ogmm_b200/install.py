"""Patch the sm_100a kernels into an imported reference checkout.

The reference binds its helpers with ``from lib.utils import ...`` at import time, so replacing
``lib.utils.knn`` alone is not enough: every importing module holds its own name (SURVEY.md
section 8(b)).  ``install()`` rebinds the hot-path names in each of those namespaces and
``uninstall()`` restores the originals.

    import lib.utils, lib.se3, models.dgcnn, models.attn, models.gmmreg, lib.loss, baseline.deepgmr
    import ogmm_b200.install as inst
    inst.install()            # after the reference modules are imported
"""
from __future__ import annotations

import sys

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


def install(modules=None):
    """Rebind the hot-path names in every already-imported reference module.  Returns the list patched."""
    done = []
    for mod_name, table in PATCH_TABLE.items():
        if modules is not None and mod_name not in modules:
            continue
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for attr, repl in table.items():
            if hasattr(mod, attr):
                _saved.setdefault((mod_name, attr), getattr(mod, attr))
                setattr(mod, attr, repl)
                done.append(f"{mod_name}.{attr}")
    return done


def uninstall():
    for (mod_name, attr), orig in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, orig)
        del _saved[(mod_name, attr)]
