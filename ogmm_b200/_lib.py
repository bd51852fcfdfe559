"""ctypes binding of the C-ABI library ``libogmm_b200.so`` (include/ogmm_b200.h).

The library must exist: there is NO CPU or PyTorch fallback.  ``load()`` raises with build
instructions when the shared object is missing, and every op raises ``RuntimeError`` carrying
``ogmm_last_error()`` when a call returns a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libogmm_b200.so")
ABI_VERSION = 1

_lock = threading.Lock()
_lib = None

c_f = ctypes.c_void_p      # float*  (device pointers travel as integers)
c_i64p = ctypes.c_void_p   # int64_t*
c_i32p = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int
f32 = ctypes.c_float
vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/ogmm_b200.h one to one
SIGNATURES = {
    "ogmm_version": (i32, []),
    "ogmm_last_error": (ctypes.c_char_p, []),
    "ogmm_device_info": (i32, [ctypes.POINTER(i32)] * 3),
    "ogmm_knn_graph": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, i64, i32, c_i64p, c_f, c_f, vp]),
    "ogmm_knn3_select_stats": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, c_i64p, c_i32p, vp]),
    "ogmm_square_distance": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, i32, c_f, vp]),
    "ogmm_knn_wide": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, i64, i32, c_i64p, c_f, c_i32p, vp]),
    "ogmm_edge_gather": (i32, [c_f, i64, i64, i64, c_i64p, i64, i64, i64, i64, c_f, vp]),
    "ogmm_edge_conv_max": (i32, [c_f, i64, i64, i64, c_i64p, c_f, c_f, c_f, i64, i64, i64, i64, c_f, c_f, vp]),
    "ogmm_edge_angle_max": (i32, [c_f, i64, i64, i64, c_f, c_i64p, c_f, c_f, c_f, f32, i64, i64, i64, i64, c_f, c_f, vp]),
    "ogmm_fps": (i32, [c_f, i64, i64, i64, i64, i64, i64, c_i64p, c_i64p, c_f, vp]),
    "ogmm_sinkhorn_cluster_workspace": (i64, [i64, i64, i64, i64, i64]),
    "ogmm_sinkhorn_cluster": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, f32, f32, f32, i64,
                                    c_f, c_f, c_f, c_i32p, vp, i64, vp]),
    "ogmm_sinkhorn_workspace": (i64, [i64, i64, i64, i64]),
    "ogmm_sinkhorn": (i32, [c_f, c_f, c_f, i64, i64, i64, f32, f32, i64, c_f, c_f, c_i32p, vp, i64, vp]),
    "ogmm_gmm_moments": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, c_f, c_f, c_f, vp]),
    "ogmm_gmm_moments_feat": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, c_f, c_f, vp]),
    "ogmm_gmm_moments_feat_workspace": (i64, [i64, i64, i64, i64]),
    "ogmm_gmm_moments_feat_ws": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, i64, i64, i64, i64, c_f, c_f, vp, i64, vp]),
    "ogmm_gmm_moments_feat_backward": (i32, [c_f, i64, i64, i64, c_f, c_f, i64, i64, i64, i64, c_f, i64, i64, i64, vp]),
    "ogmm_softmax_moments": (i32, [c_f, c_f, i64, i64, i64, i64, i64, i64, c_f, c_f, c_f, c_f, vp]),
    "ogmm_gmm_moments_backward": (i32, [c_f, i64, i64, i64, c_f, c_f, c_f, c_f, c_f, c_f, i64, i64, i64, c_f, i64, i64, i64, vp]),
    "ogmm_gmm_register_backward": (i32, [c_f, c_f, c_f, c_f, i64, i64, c_f, c_f, c_f, c_f, c_f, vp]),
    "ogmm_rigid_transform": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, c_f, i64, i64, i64, i64, c_f, c_f, vp]),
    "ogmm_soft_procrustes": (i32, [c_f, c_f, c_f, c_f, i64, i64, i64, i64, f32, c_f, c_f, c_f, c_f, vp]),
    "ogmm_rigid_transform_backward": (i32, [c_f, i64, i64, i64, c_f, i64, i64, i64, c_f, i64, i64, i64, i64, c_f, c_f,
                                            c_f, c_f, c_f, vp]),
    "ogmm_soft_procrustes_backward": (i32, [c_f, c_f, c_f, c_f, i64, i64, i64, i64, f32, c_f, c_f, c_f, c_f, c_f, c_f,
                                            c_f, vp]),
    "ogmm_cos_similarity": (i32, [c_f, c_f, i64, i64, i64, i64, c_f, vp]),
    "ogmm_gmm_register": (i32, [c_f, c_f, c_f, c_f, i64, i64, c_f, vp]),
}


class OgmmError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the library is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OgmmError(
                f"{LIB_PATH} is missing. ogmm_b200 has no CPU fallback: build the CUDA library first with "
                "`python -m ogmm_b200._build` (needs nvcc; cross-compiles for sm_100a without a GPU).")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = library older than the header
            fn.restype = res
            fn.argtypes = args
        if lib.ogmm_version() != ABI_VERSION:
            raise OgmmError(f"{LIB_PATH}: ABI version {lib.ogmm_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return _lib


def check(status: int, what: str):
    if status != 0:
        msg = load().ogmm_last_error().decode("utf-8", "replace")
        names = {-1: "EINVAL", -2: "EUNSUPPORTED", -3: "ECUDA", -4: "EWORKSPACE"}
        raise OgmmError(f"{what} failed with {names.get(status, status)}: {msg}")
