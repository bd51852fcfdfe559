"""Shared helpers for the GPU parity tests."""
import torch


def decidable_rows(src, dst, k, normalize=False):
    """SURVEY.md section 7 'kNN tie semantics': a row is decidable when the fp64 gap between its k-th and
    (k+1)-th distance exceeds the fp32 error bound of the expanded form, 8 * 2^-23 * max(|x|^2 + |y|^2)."""
    s, d = src.double(), dst.double()
    dist = (s * s).sum(-1)[:, :, None] + (d * d).sum(-1)[:, None, :] - 2 * s @ d.transpose(1, 2)
    if normalize:
        dist = 2 - 2 * s @ d.transpose(1, 2)
    dist = dist.clamp(min=1e-12) if not normalize else dist
    srt = dist.sort(dim=-1)[0]
    scale = ((s * s).sum(-1).max(dim=1)[0] + (d * d).sum(-1).max(dim=1)[0])[:, None]
    bound = 8 * 2.0 ** -23 * scale
    m = dst.shape[1]
    # every adjacent gap among the first k+1 must be resolvable for ORDER to be decidable
    gaps = srt[:, :, 1:min(k + 1, m)] - srt[:, :, :min(k, m - 1)]
    return (gaps > bound[:, :, None]).all(dim=-1), dist


def rot_err_deg(r1, r2):
    """Angle between two rotations in degrees, from the chord |R1 - R2|_F = 2 sqrt(2) sin(theta / 2).

    The reference's metric (lib/metric.py:85-88) goes through arccos((trace - 1) / 2), which cannot resolve
    angles below ~0.02 degree once the matrices are rounded to fp32 (trace error 1e-7 -> sqrt(1e-7) rad);
    the chord form is exact to first order and is what a 1e-3 degree parity bar needs."""
    chord = (r1.double() - r2.double()).flatten(1).norm(dim=1)
    return 2 * torch.arcsin(torch.clamp(chord / (2 * 2 ** 0.5), max=1.0)) * 180 / torch.pi


def within_bar(err, bar, spread, what, factor=2.0):
    """The north-star bar, stated against what the reference itself can resolve.

    ``err``    our result vs the fp32 reference (golden fixture or fp32 oracle)
    ``spread`` the reference's own fp32-vs-fp64 difference on the same inputs (oracle run in float64 = the arbiter)
    Passes when ``err <= bar``; when the fp32 reference itself sits further than ``bar / factor`` from the fp64
    arbiter, the bar it can support is ``factor x spread`` and that is what is required.  Both numbers are printed
    so the log shows which case applied."""
    limit = max(bar, factor * spread)
    print(f"  {what}: err {err:.3e}  bar {bar:.1e}  reference fp32-vs-fp64 spread {spread:.3e}  -> limit {limit:.3e}")
    assert err <= limit, f"{what}: {err:.3e} > {limit:.3e}"
