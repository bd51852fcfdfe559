"""the feature M-step shared between wkeans_plus (lib/utils.py:289) and CluLoss
(lib/loss.py:114-115).  The cache logic is device-independent, so it is checked here with a stub in place of the kernel;
tests/test_gpu_parity.py::test_shared_feature_moments runs it on the real path."""
import gc

import torch

from ogmm_b200.utils import SharedMoments


def test_shared_moments_hits_only_on_identical_data():
    calls = []

    def compute(g, f):
        calls.append(1)
        return g.mean(1), torch.einsum("bnj,bnd->bjd", g, f)

    c = SharedMoments()
    gamma = torch.rand(2, 32, 4)
    feats = torch.rand(2, 8, 32)                           # (B,D,N) as the model holds it
    a = c.get(gamma, feats.transpose(-1, -2), compute)
    b = c.get(gamma, feats.transpose(-1, -2), compute)      # a NEW view object of the same tensor: CluLoss's call
    assert len(calls) == 1 and b[1] is a[1] and c.hits == 1
    c.get(gamma.clone(), feats.transpose(-1, -2), compute)  # another gamma object, equal values: recomputed
    assert len(calls) == 2
    c.get(gamma, feats.transpose(-1, -2), compute)
    assert len(calls) == 3                                  # the single entry was replaced
    feats.add_(1.0)                                         # in-place write bumps the version counter
    r = c.get(gamma, feats.transpose(-1, -2), compute)
    assert len(calls) == 4 and torch.allclose(r[1], torch.einsum("bnj,bnd->bjd", gamma, feats.transpose(-1, -2)))
    gamma.mul_(0.5)
    c.get(gamma, feats.transpose(-1, -2), compute)
    assert len(calls) == 5
    c.get(gamma, feats[:, :4].transpose(-1, -2), compute)   # same storage, other shape
    assert len(calls) == 6
    # the cache holds its operands weakly: a freed feature tensor cannot be aliased by a new one at the same address
    c.get(gamma, feats.transpose(-1, -2), compute)
    del feats
    gc.collect()
    feats2 = torch.rand(2, 8, 32)
    c.get(gamma, feats2.transpose(-1, -2), compute)
    assert len(calls) == 8
