"""Row a15: the SE(3) helpers of ``ogmm_b200/se3.py`` (and ``index_points``) against the reference's outputs.

They are 3x4 / 4x4 tensor algebra kept as PyTorch with the reference's signatures (lib/se3.py:14-117,
lib/utils.py:111-127), device-generic, so the golden comparison runs on the CPU here and again on the GPU in
tests/test_gpu_parity.py::test_se3_helpers_on_device.
"""
import torch

from ogmm_b200 import se3, utils


def check_se3(golden, dev="cpu"):
    g = {k: v.to(dev) for k, v in golden("se3").items()}
    # bit-exact on the CPU (same ATen ops as the reference ran); a few ulp on the GPU, whose matmul sums in another order
    same = torch.equal if str(dev) == "cpu" else (lambda a, b: torch.allclose(a, b, rtol=1e-6, atol=1e-6))
    assert same(se3.torch_inverse(g["g1"]), g["inv"])
    assert same(se3.torch_concatenate(g["g1"], g["g2"]), g["cat"])
    assert same(se3.torch_transform(g["g1"], g["cloud"]), g["moved"])
    rot, t = se3.decompose_trans(g["integ"])
    assert tuple(rot.shape) == (2, 3, 3) and tuple(t.shape) == (2, 3, 1)
    assert torch.equal(rot, g["g1"][:, :, :3]) and torch.equal(t, g["g1"][:, :, 3:4])
    assert torch.equal(se3.integrate_trans(rot, t), g["integ"])
    single = se3.integrate_trans(rot[0], t[0])                       # un-batched branch (lib/se3.py:45-51)
    assert torch.equal(single, g["integ"][0])
    r1, t1 = se3.decompose_trans(single)
    assert torch.equal(r1, rot[0]) and torch.equal(t1, t[0])
    # group identities: g * g^-1 = identity, inverse of a product
    ident = se3.torch_concatenate(g["g1"], se3.torch_inverse(g["g1"]))
    assert torch.allclose(ident, se3.torch_identity(2).to(dev), atol=1e-6)
    moved, normals = se3.torch_transform(g["g1"], g["cloud"], g["cloud"])
    assert same(moved, g["moved"]) and torch.allclose(normals, g["moved"] - g["g1"][:, None, :, 3], atol=1e-5)
    f = {k: v.to(dev) for k, v in golden("fps").items()}
    assert torch.equal(utils.index_points(f["xyz"], f["ids_center"]), f["gathered"])


def test_se3_helpers_match_the_reference(golden):
    check_se3(golden)
