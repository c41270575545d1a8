"""TEST-ONLY hook (put on PYTHONPATH by tests/test_benchmark_scripts.py, never by the product): lets the
SP benchmark scripts run their control flow on CPU/gloo by swapping the libspconv-backed layers for
PyTorch ops WITHOUT halos (a tile is convolved with zero padding).  Numerically that is not the spatial
convolution -- it only exercises rank arithmetic, trainers, collectives and script plumbing."""
import os

if os.environ.get("SPCONV_TEST_CPU_SMOKE") == "1":
    import torch.nn as nn
    import torch.nn.functional as F

    from mpi4dl_b200.torchgems import spatial

    def _conv(self, t):
        if getattr(self, "_fused_valid", False):          # D2 ResNet cells: valid convolutions of the widened tile
            return F.conv2d(t, self.weight, self.bias, self.stride, 0)
        return F.conv2d(t, self.weight, self.bias, self.stride, (self.halo_len_height, self.halo_len_width))

    def _pool(self, t):
        if self.operation == "MaxPool2d":
            return F.max_pool2d(t, self.kernel_size, self.stride, self.padding)
        return F.avg_pool2d(t, self.kernel_size, self.stride, self.padding, count_include_pad=True)

    spatial.conv_spatial.forward = _conv
    spatial.Pool.forward = _pool
    spatial.halo_exchange_layer.forward = lambda self, t: F.pad(t, (self.halo_len,) * 4)
    spatial.local_conv2d.forward = lambda self, t: nn.Conv2d.forward(self, t)
    spatial.local_pool2d.forward = lambda self, t: F.avg_pool2d(t, self.kernel_size, self.stride, self.padding)
