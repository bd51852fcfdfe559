"""ogmm_b200 -- B200-native OGMM registration hot path.

Python mirror of the reference's function surface (``utils``, ``se3``, ``modules``) over hand-written
sm_100a CUDA kernels reached through the C ABI in ``include/ogmm_b200.h``.  There is no CPU path:
the first kernel call loads ``libogmm_b200.so`` and raises if it is missing.
"""
__version__ = "0.1.0"

from . import _lib, ops, utils, se3, modules, synth  # noqa: F401
from .utils import (square_distance, knn, get_graph_feature, sinkhorn, index_points, gmm_params, og_params,  # noqa: F401
                    farthest_point_sample, cos_similarity, get_local_corrs, get_anchor_corrs, wkeans_plus)
from .se3 import compute_rigid_transformation  # noqa: F401
from .modules import Clustering, GMMSVD, graph_features, gmm_register, deepgmr_em  # noqa: F401
