"""Seeded synthetic point-cloud pairs shaped like the reference's datasets.

There are no dataset files in this environment, so the benchmark and the
parity tests use synthetic pairs that follow the reference's own data
transforms:

* ``modelnet_pair``  -- the ``noise_type == "crop"`` pipeline of
  ``datasets/modelnet.py:73-88``: clone a unit-sphere-normalised surface into
  src/ref, crop each with a random half-space keeping 70 %
  (``datasets/transforms.py:428-453``), rigid-transform with per-axis Euler
  angles U(0,45 deg) and t ~ U(-0.5,0.5)^3 (``:152-190``), resample to N with
  top-up by replacement (``:297-328``), jitter N(0,0.01^2) clipped at 0.05
  (``:402-415``), shuffle (``:502-513``).
* ``icl_nuim_pair`` -- indoor-room surfaces in metres with non-uniform target
  density, cropped and posed as in ``datasets/realdata.py:138-193``.

Everything is numpy + ``np.random.default_rng(seed)``; seed = 1234 + pair index
(the reference's default ``--seed``, ``configs/cfgs.py:58``).
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 1234


def _unit_dir(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def _euler_se3(rng, rot_mag_deg=45.0, trans_mag=0.5):
    """datasets/transforms.py:158-190: R = Rx Ry Rz with angles U(0, rot_mag), t U(-m, m)."""
    ax, ay, az = rng.uniform(size=3) * np.pi * rot_mag_deg / 180.0
    cx, cy, cz, sx, sy, sz = np.cos(ax), np.cos(ay), np.cos(az), np.sin(ax), np.sin(ay), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rx @ ry @ rz, rng.uniform(-trans_mag, trans_mag, 3)


def _crop(rng, pts, p_keep):
    """datasets/transforms.py:441-453: half-space crop keeping ~p_keep of the points."""
    d = (pts - pts.mean(0)) @ _unit_dir(rng)
    return pts[d > np.percentile(d, (1.0 - p_keep) * 100)]


def _resample(rng, pts, k):
    """datasets/transforms.py:320-328: exactly k points; top-up by replacement when short."""
    n = pts.shape[0]
    if k <= n:
        sel = rng.choice(n, k, replace=False)
    else:
        sel = np.concatenate([rng.choice(n, n, replace=False), rng.choice(n, k - n, replace=True)])
    return pts[sel]


def _jitter(rng, pts, scale=0.01, clip=0.05):
    return pts + np.clip(rng.normal(0.0, scale, size=pts.shape), -clip, clip)


def _surface(rng, n):
    """A unit-sphere-normalised synthetic surface: ellipsoid shell + 2-3 planar patches."""
    n_shell = n // 2
    radii = rng.uniform(0.35, 1.0, 3)
    v = rng.normal(size=(n_shell, 3))
    shell = v / np.linalg.norm(v, axis=1, keepdims=True) * radii
    parts = [shell]
    n_left = n - n_shell
    n_patch = int(rng.integers(2, 4))
    for i in range(n_patch):
        m = n_left // n_patch + (1 if i < n_left % n_patch else 0)
        nrm = _unit_dir(rng)
        a = np.cross(nrm, _unit_dir(rng))
        a /= np.linalg.norm(a)
        b = np.cross(nrm, a)
        uv = rng.uniform(-0.6, 0.6, size=(m, 2))
        parts.append(rng.uniform(-0.5, 0.5) * nrm + uv[:, :1] * a + uv[:, 1:] * b)
    pts = np.concatenate(parts, 0)
    pts = pts - pts.mean(0)
    return pts / np.linalg.norm(pts, axis=1).max()


def modelnet_pair(index, n_points=1024, p_keep=0.7, rot_mag=45.0, trans_mag=0.5):
    """One ModelNet40-shape partial-overlap pair.

    Returns src (3,N), tgt (3,N) float32 and the ground-truth R (3,3), t (3,)
    taking src onto tgt (``tgt ~ R src + t`` on the overlap).
    """
    rng = np.random.default_rng(BASE_SEED + int(index))
    base = _surface(rng, n_points)
    src = _crop(rng, base.copy(), p_keep)
    ref = _crop(rng, base.copy(), p_keep)
    rot, t = _euler_se3(rng, rot_mag, trans_mag)
    # the reference transforms the source and asks for the inverse motion
    src_moved = src @ rot.T + t
    rot_gt, t_gt = rot.T, -rot.T @ t
    src_out = _jitter(rng, _resample(rng, src_moved, n_points))
    ref_out = _jitter(rng, _resample(rng, ref, n_points))
    src_out = src_out[rng.permutation(n_points)]
    ref_out = ref_out[rng.permutation(n_points)]
    return (np.ascontiguousarray(src_out.T, dtype=np.float32),
            np.ascontiguousarray(ref_out.T, dtype=np.float32),
            rot_gt.astype(np.float32), t_gt.astype(np.float32))


def _room(rng, n):
    """Points on 4-6 axis-aligned planes of a ~4x3x2.5 m box plus a few box objects (metres)."""
    size = np.array([4.0, 3.0, 2.5]) * rng.uniform(0.85, 1.15, 3)
    faces = [(ax, side) for ax in range(3) for side in (0, 1)]
    keep = rng.permutation(6)[: int(rng.integers(4, 7))]
    n_obj = int(rng.integers(2, 5))
    n_obj_pts = n // 4
    n_wall = n - n_obj_pts
    out = []
    for i, f in enumerate(keep):
        ax, side = faces[f]
        m = n_wall // len(keep) + (1 if i < n_wall % len(keep) else 0)
        p = rng.uniform(0, 1, size=(m, 3)) * size
        p[:, ax] = side * size[ax]
        out.append(p)
    for i in range(n_obj):
        m = n_obj_pts // n_obj + (1 if i < n_obj_pts % n_obj else 0)
        ext = rng.uniform(0.3, 0.9, 3)
        org = rng.uniform(0, 1, 3) * (size - ext)
        org[2] = 0.0
        p = rng.uniform(0, 1, size=(m, 3)) * ext
        ax = rng.integers(0, 3, size=m)
        p[np.arange(m), ax] = rng.integers(0, 2, size=m) * ext[ax]
        out.append(org + p)
    return np.concatenate(out, 0)


def _fps_np(pts, k):
    n = pts.shape[0]
    sel = np.zeros(k, dtype=np.int64)
    d = np.full(n, 1e10)
    far = 0
    for i in range(k):
        sel[i] = far
        d = np.minimum(d, ((pts - pts[far]) ** 2).sum(1))
        far = int(d.argmax())
    return pts[sel]


def icl_nuim_pair(index, n_points=1024, p_keep=0.7, tgt_factor=1):
    """One ICL-NUIM-shape pair with density variation (metres, un-normalised).

    The target is sampled with probability proportional to the distance from a
    random viewpoint; ``tgt_factor=2`` gives M = 2N as in
    ``datasets/modelnet.py:268-273``.  Crop + FPS-subsample to int(p_keep*n) as
    in ``datasets/realdata.py:171-176`` and then top up to the requested size so
    batches stay rectangular.  Returns src (3,N), tgt (3,M), R, t.
    """
    rng = np.random.default_rng(BASE_SEED + 100003 + int(index))
    dense = _room(rng, 8 * n_points)
    src = dense[rng.choice(dense.shape[0], n_points, replace=False)]
    view = dense.mean(0) + rng.uniform(-1.0, 1.0, 3)
    w = np.linalg.norm(dense - view, axis=1)
    m = n_points * tgt_factor
    tgt = dense[rng.choice(dense.shape[0], m, replace=False, p=w / w.sum())]
    src = _crop(rng, src, p_keep)
    tgt = _crop(rng, tgt, p_keep)
    ks, kt = int(p_keep * n_points), int(p_keep * m)
    if src.shape[0] > ks:
        src = _fps_np(src, ks)
    if tgt.shape[0] > kt:
        tgt = _fps_np(tgt, kt)
    ang = rng.uniform(-np.pi / 4, np.pi / 4, 3)
    cx, cy, cz, sx, sy, sz = *np.cos(ang), *np.sin(ang)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    rot, t = rx @ ry @ rz, rng.uniform(-0.5, 0.5, 3)
    tgt = tgt @ rot.T + t
    src = _resample(rng, src, n_points)
    tgt = _resample(rng, tgt, m)
    return (np.ascontiguousarray(src.T, dtype=np.float32), np.ascontiguousarray(tgt.T, dtype=np.float32),
            rot.astype(np.float32), t.astype(np.float32))


def modelnet_batch(first_index, count, n_points=1024):
    """Stack ``count`` ModelNet40-shape pairs -> src (B,3,N), tgt (B,3,N), R (B,3,3), t (B,3)."""
    items = [modelnet_pair(first_index + i, n_points) for i in range(count)]
    return tuple(np.stack([it[j] for it in items]) for j in range(4))


def icl_nuim_batch(first_index, count, n_points=1024, tgt_factor=1):
    items = [icl_nuim_pair(first_index + i, n_points, tgt_factor=tgt_factor) for i in range(count)]
    return tuple(np.stack([it[j] for it in items]) for j in range(4))


def hot_path_inputs(first_index, count, n_points=1024, emb_dims=512, tile=None):
    """Inputs of the isolated hot path for ``count`` pairs.

    xyz from ``modelnet_batch``; overlap scores o ~ sigmoid(N(0,1)); features
    relu(N(0,1)) of shape (B,emb_dims,N).  ``tile``: generate only that many
    distinct pairs and repeat them (generation is CPU-bound; the arithmetic the
    kernels do does not depend on the values being distinct).
    """
    distinct = count if tile is None else min(tile, count)
    src, tgt, rot, t = modelnet_batch(first_index, distinct, n_points)
    rng = np.random.default_rng(BASE_SEED + 7 + int(first_index))
    o = 1.0 / (1.0 + np.exp(-rng.normal(size=(2, distinct, n_points))))
    f = np.maximum(rng.normal(size=(2, distinct, emb_dims, n_points)), 0.0)
    arrs = [src, tgt, o[0], o[1], f[0], f[1], rot, t]
    if distinct < count:
        reps = -(-count // distinct)
        arrs = [np.concatenate([a] * reps, 0)[:count] for a in arrs]
    keys = ["src", "tgt", "src_o", "tgt_o", "src_feats", "tgt_feats", "rot_gt", "t_gt"]
    return {k: np.ascontiguousarray(a, dtype=np.float32) for k, a in zip(keys, arrs)}
