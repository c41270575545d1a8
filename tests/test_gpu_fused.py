"""-m gpu: torchgems.fused (csrc/bnrelu.cu) against the eager PyTorch modules it replaces inside spatial cells
(reference chain: nn.ReLU -> conv -> nn.BatchNorm2d, src/models/amoebanet.py:365-398; per-tile statistics, N4):
forward, input / gamma / beta gradients and the running-statistics update of training-mode BatchNorm2d, with and
without the fused ReLU.  fp32: tight (1e-5 relative to the tensor scale); bf16 storage: one bf16 ulp + floor."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("shape", [(2, 5, 8, 16), (1, 104, 64, 72), (3, 17, 130, 8)])
def test_bn_relu_matches_batchnorm2d(shape, relu, dtype, tol):
    from mpi4dl_b200.torchgems.fused import bn_relu

    torch.manual_seed(0)
    N, Cc, H, W = shape
    x = (torch.randn(shape, device=DEV) * 1.7 + 0.3).to(dtype)
    bn_a = nn.BatchNorm2d(Cc).to(DEV).to(dtype)
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-0.5, 0.5)
    bn_b = copy.deepcopy(bn_a)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    za = bn_relu(xa, bn_a, relu=relu)
    zb = bn_b(xb)
    zb = F.relu(zb) if relu else zb
    g = torch.randn(shape, device=DEV).to(dtype)
    za.backward(g)
    zb.backward(g)

    def close(a, b, name):
        a, b = a.detach().float(), b.detach().float()
        scale = float(b.abs().max()) + 1e-6
        assert float((a - b).abs().max()) <= tol * scale, (name, float((a - b).abs().max()), scale)

    close(za, zb, "z")
    close(xa.grad, xb.grad, "dx")
    close(bn_a.weight.grad, bn_b.weight.grad, "dgamma")
    close(bn_a.bias.grad, bn_b.bias.grad, "dbeta")
    close(bn_a.running_mean, bn_b.running_mean, "running_mean")
    close(bn_a.running_var, bn_b.running_var, "running_var")
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 1


def test_relu_conv_bn_chain_equals_the_eager_sequential():
    """The fused Sequential of a spatial cell (same children, same keys) computes what nn.Sequential computes."""
    from mpi4dl_b200.torchgems.fused import relu_conv_bn_chain
    from mpi4dl_b200.torchgems.spatial import local_conv2d

    torch.manual_seed(1)
    mods = []
    for ci, co, k in ((16, 8, 1), (8, 8, (1, 7)), (8, 8, (7, 1)), (8, 16, 1)):
        pad = ((k[0] - 1) // 2, (k[1] - 1) // 2) if isinstance(k, tuple) else 0
        mods += [nn.ReLU(), local_conv2d(ci, co, k, padding=pad, bias=False), nn.BatchNorm2d(co)]
    fused = relu_conv_bn_chain(*mods).to(DEV)
    eager = nn.Sequential(*copy.deepcopy(mods)).to(DEV)
    assert list(fused.state_dict().keys()) == list(eager.state_dict().keys())
    x = torch.randn(2, 16, 24, 64, device=DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = fused(xa), eager(xb)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(ya, yb, rtol=1e-4, atol=1e-4)
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-3, atol=1e-4)
    for (n1, p1), (n2, p2) in zip(fused.named_parameters(), eager.named_parameters()):
        assert n1 == n2 and torch.allclose(p1.grad, p2.grad, rtol=2e-3, atol=2e-4), n1
    fused.eval()
    eager.eval()
    with torch.no_grad():
        assert torch.allclose(fused(x), eager(x), rtol=1e-4, atol=1e-4)      # eval mode: the modules' own path
