"""Import the UNMODIFIED reference (gfmei/ogmm) as the checker / baseline.  TEST INFRASTRUCTURE ONLY.

Two places can hold the reference checkout:

  /root/reference        the authoring container (read-only); absent on the GPU box
  baseline/_ref/         a verbatim copy made by ``__graft_entry__.build()`` when /root/reference is present;
                         git-ignored (never committed), NOT gpurun-ignored, so it travels to the GPU box

``import_reference()`` puts the checkout on ``sys.path`` with empty stub modules for the four third-party
imports that are absent here and untouched by the ``is_test=False`` forward (``transforms3d``, ``open3d``,
``h5py``, ``plyfile``) and returns the imported modules.  Nothing in the reference is
edited.  Only ``tests/``, ``oracle/make_golden.py`` and the reference legs of ``bench.py`` call this.
"""
from __future__ import annotations

import os
import shutil
import stat
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCE = os.environ.get("OGMM_REFERENCE", "/root/reference")
VENDORED = os.path.join(ROOT, "baseline", "_ref")

_STUBS = ("transforms3d", "transforms3d.quaternions", "open3d", "h5py", "plyfile")


def reference_path():
    """The checkout to import from, or None: the vendored copy first (same bytes on both boxes), else the source."""
    for cand in (VENDORED, SOURCE):
        if os.path.isfile(os.path.join(cand, "lib", "utils.py")):
            return cand
    return None


def vendor(force=False):
    """Copy SOURCE -> baseline/_ref (when SOURCE exists).  Returns the vendored path or None."""
    if not os.path.isfile(os.path.join(SOURCE, "lib", "utils.py")):
        return VENDORED if os.path.isdir(VENDORED) else None
    marker = os.path.join(VENDORED, "lib", "utils.py")
    if os.path.exists(marker) and not force:
        return VENDORED
    if os.path.isdir(VENDORED):
        for d, _, _ in os.walk(VENDORED):
            os.chmod(d, os.stat(d).st_mode | stat.S_IWUSR)
        shutil.rmtree(VENDORED)
    os.makedirs(os.path.dirname(VENDORED), exist_ok=True)
    shutil.copytree(SOURCE, VENDORED, ignore=shutil.ignore_patterns(".git", ".idea", "__pycache__"),
                    copy_function=shutil.copyfile)
    for d, _, files in os.walk(VENDORED):                       # the source tree is read-only; the copy must be removable
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    return VENDORED


def import_reference(path=None):
    """-> dict of the reference modules (utils, se3, dgcnn, attn, gmmreg, loss, deepgmr), imported from ``path``."""
    path = path or reference_path()
    if path is None:
        raise ImportError("no reference checkout: neither baseline/_ref nor /root/reference exists")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["transforms3d"].quaternions = sys.modules["transforms3d.quaternions"]
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object
    if path not in sys.path:
        sys.path.insert(0, path)
    import lib.utils as ru
    import lib.se3 as rs
    import lib.loss as rl
    import models.dgcnn as rd
    import models.attn as ra
    import models.gmmreg as rm
    import baseline.deepgmr as rg
    return {"utils": ru, "se3": rs, "loss": rl, "dgcnn": rd, "attn": ra, "gmmreg": rm, "deepgmr": rg, "path": path}


def model_config(gnn_k=20, num_heads=4, km_clusters=128, overlap_radius=0.0375):
    """The four fields the models read from ``config`` (configs/cfgs.py:32-39,49)."""
    return types.SimpleNamespace(gnn_k=gnn_k, num_heads=num_heads, km_clusters=km_clusters, overlap_radius=overlap_radius)
