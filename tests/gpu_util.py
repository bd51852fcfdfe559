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
    c = torch.einsum('bij,bij->b', r1.double(), r2.double())
    return torch.arccos(torch.clamp((c - 1) / 2, -1.0, 1.0)) * 180 / torch.pi
