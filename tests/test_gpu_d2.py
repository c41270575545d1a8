"""-m gpu: the D2 ("fused halo") building blocks -- local_conv2d / local_pool2d (valid conv / pool of a
tile that already carries its halo, on the libspconv kernels) and the AmoebaNet Cell_D2 built from
them -- against the PyTorch operators the reference's D2 cells call (nn.Conv2d(padding=0),
nn.AvgPool2d(3, padding=0), zero padding for a tile without neighbours; amoebanet_d2.py:88-117,
159-191, 569-676), forward and backward."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _pg():
    # the comparison side is cuDNN: keep it in true fp32 (TF32 convolutions are on by default and
    # are ~1e-3 off, far coarser than libspconv's fp32 path)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    made = False
    if not dist.is_initialized():
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29785")
        dist.init_process_group("gloo", rank=0, world_size=1)
        made = True
    yield
    if made:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [(1, 7), (7, 1), (3, 3), (1, 1)])
@pytest.mark.parametrize("dtype,C,K,tol", [(torch.float32, 5, 6, 1e-4), (torch.bfloat16, 64, 64, 3e-2)])
def test_local_conv2d_is_a_valid_convolution(k, dtype, C, K, tol):
    from mpi4dl_b200.torchgems.spatial import local_conv2d
    torch.manual_seed(3)
    m = local_conv2d(C, K, k, stride=1, padding=0, bias=(dtype == torch.float32)).cuda().to(dtype)
    x = torch.randn(2, C, 38, 70, device="cuda").to(dtype).requires_grad_()
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    got = (y.detach().float(), x.grad.float(), m.weight.grad.float())
    x2 = x.detach().float().requires_grad_()
    w2 = m.weight.detach().float().requires_grad_()
    y2 = F.conv2d(x2, w2, None if m.bias is None else m.bias.detach().float(), 1, 0)
    y2.backward(gy.float())
    for a, b, name in zip(got, (y2.detach(), x2.grad, w2.grad), ("y", "dx", "dw")):
        assert a.shape == b.shape, name
        assert torch.allclose(a, b, rtol=tol, atol=tol * max(1.0, b.abs().max().item())), name


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_local_pool2d_is_a_valid_pool(dtype, tol):
    from mpi4dl_b200.torchgems.spatial import local_pool2d
    torch.manual_seed(4)
    x = torch.randn(2, 8, 34, 72, device="cuda").to(dtype).requires_grad_()
    y = local_pool2d("AvgPool2d", 3, 1, 0)(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    x2 = x.detach().float().requires_grad_()
    y2 = F.avg_pool2d(x2, 3, 1, 0)
    y2.backward(gy.float())
    assert y.shape == y2.shape
    assert torch.allclose(y.float(), y2, rtol=tol, atol=tol) and torch.allclose(x.grad.float(), x2.grad, rtol=tol, atol=tol)


def test_cell_d2_matches_the_pytorch_operators_the_reference_calls():
    """One spatial part (no neighbours): the halo layers pad zeros, so the whole cell can be replayed
    with cuDNN / PyTorch ops on the same weights by calling the base-class forwards."""
    from mpi4dl_b200.models.amoebanet_d2 import Cell_D2
    from mpi4dl_b200.torchgems import spatial
    torch.manual_seed(5)
    sp = dict(local_rank=0, spatial_size=1, num_spatial_parts=1, slice_method="vertical")
    cell = Cell_D2(sp, 32, 32, 16, reduction_prev=False).cuda().train()
    x1 = torch.randn(2, 32, 24, 40, device="cuda", requires_grad=True)
    x2 = torch.randn(2, 32, 24, 40, device="cuda", requires_grad=True)

    def run():
        for p in cell.parameters():
            p.grad = None
        x1.grad = x2.grad = None
        y, skip = cell((x1, x2))
        y.square().mean().backward()
        return y.detach().clone(), x1.grad.clone(), x2.grad.clone(), [p.grad.clone() for p in cell.parameters()]

    ours = run()
    saved = (spatial.local_conv2d.forward, spatial.local_pool2d.forward, spatial.conv_spatial.forward,
             spatial.halo_exchange_layer.forward)
    try:
        spatial.local_conv2d.forward = lambda self, t: nn.Conv2d.forward(self, t)
        spatial.conv_spatial.forward = lambda self, t: nn.Conv2d.forward(self, t)          # 1x1, padding 0
        spatial.local_pool2d.forward = lambda self, t: F.avg_pool2d(t, self.kernel_size, self.stride, self.padding)
        spatial.halo_exchange_layer.forward = lambda self, t: F.pad(t, (self.halo_len,) * 4)
        ref = run()
    finally:
        (spatial.local_conv2d.forward, spatial.local_pool2d.forward, spatial.conv_spatial.forward,
         spatial.halo_exchange_layer.forward) = saved
    assert ours[0].shape == (2, 64, 24, 40)
    for a, b in zip(ours[:3], ref[:3]):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-4)
    for a, b in zip(ours[3], ref[3]):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-4 * max(1.0, b.abs().max().item()))


@pytest.mark.parametrize("strides,resblock", [(1, 1), (2, 0), (1, 0)])
def test_resnet_d2_cell_is_valid_convolutions_with_cropped_shortcut(strides, resblock):
    """make_cell_v2_spatial (models/resnet_spatial_d2.py; reference resnet_spatial_d2.py:396-480): every convolution is
    conv_spatial(halo_len=0, padding=0) = a valid convolution on libspconv -- incl. the strided 3x3 and 1x1 ones, whose
    sampling phase starts at the tile's first row / column -- and the shortcut is cropped by 2 (3 for stride 2).
    Replayed with nn.Conv2d.forward on the same weights (the base class holds padding=0)."""
    from mpi4dl_b200.models.resnet import _SpatialCtx  # noqa: F401
    from mpi4dl_b200.models.resnet_spatial_d2 import _ValidCtx, make_cell_v2_spatial
    from mpi4dl_b200.torchgems import spatial
    torch.manual_seed(6)
    ctx = _ValidCtx(0, 1, 4, "square")
    cin = 64 if resblock else 16
    cell = make_cell_v2_spatial(resblock, strides, cin, 16, 64, "relu", True, 2, ctx).cuda().train()
    x = torch.randn(2, cin, 39, 71, device="cuda", requires_grad=True)     # odd extents: the strided phase matters

    def run():
        for p in cell.parameters():
            p.grad = None
        x.grad = None
        y = cell(x)
        y.square().mean().backward()
        return y.detach().clone(), x.grad.clone(), [p.grad.clone() for p in cell.parameters() if p.grad is not None]

    ours = run()
    saved = spatial.conv_spatial.forward
    try:
        spatial.conv_spatial.forward = lambda self, t: nn.Conv2d.forward(self, t)
        ref = run()
    finally:
        spatial.conv_spatial.forward = saved
    assert ours[0].shape == ref[0].shape
    for a, b in zip(ours[:2], ref[:2]):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-4 * max(1.0, b.abs().max().item()))
    for a, b in zip(ours[2], ref[2]):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-4 * max(1.0, b.abs().max().item()))
