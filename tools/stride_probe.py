"""Does the 104->208 pointwise GEMM go faster when channel planes are small (fewer 2 MB pages per tile)?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mpi4dl_b200 import _lib
from mpi4dl_b200.torchgems.spatial import _ConvSpatialFn
dev = "cuda:0"
for (N, C, K, H, W) in [(1, 104, 208, 4096, 4096), (16, 104, 208, 1024, 1024), (256, 104, 208, 256, 256), (4096, 104, 208, 64, 64),
                         (1, 208, 52, 4096, 4096), (256, 208, 52, 256, 256), (64, 104, 208, 512, 512), (1, 416, 104, 1024, 1024), (16, 416, 104, 256, 256),
                         (64, 416, 104, 128, 128), (1, 416, 416, 1024, 1024), (64, 416, 416, 128, 128)]:
    x = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16)
    w = (torch.randn(K, C, 1, 1, device=dev) / C ** 0.5).to(torch.bfloat16)
    desc = (N, C, H, W, K, 1, 1, 1, 1, 0, 0, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
    with torch.no_grad():
        for _ in range(3):
            y = _ConvSpatialFn.apply(x, w, None, desc, *([None] * 9))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = _ConvSpatialFn.apply(x, w, None, desc, *([None] * 9))
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = (C + K) * H * W * N * 2
    print("N=%d C=%d K=%d %dx%d  %.3f ms  %.0f GB/s" % (N, C, K, H, W, ms, byts / ms / 1e6), flush=True)
