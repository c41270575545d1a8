import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mpi4dl_b200 import _lib
from mpi4dl_b200.torchgems.spatial import _PoolFn
dev = "cuda:0"
for (C, H, W, k, s, mode) in [(208, 2048, 2048, 3, 1, 1), (416, 1024, 1024, 3, 1, 1), (208, 4096, 4096, 3, 2, 1), (208, 4096, 4096, 2, 2, 0)]:
    x = torch.randn(1, C, H, W, device=dev).to(torch.bfloat16).requires_grad_(True)
    desc = (1, C, H, W, k, s, (k - 1) // 2, mode, _lib.SPC_BF16)
    y = _PoolFn.apply(x, desc, *([None] * 9))
    gy = torch.randn_like(y)
    def f():
        return _PoolFn.apply(x, desc, *([None] * 9))
    def b():
        y.backward(gy, retain_graph=True); x.grad = None
    for nm, fn in (("fwd", f), ("bwd", b)):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byts = (x.numel() + y.numel()) * 2
        print("pool k%d s%d mode%d C=%d %dx%d %s %.3f ms %.0f GB/s" % (k, s, mode, C, H, W, nm, ms, byts / ms / 1e6), flush=True)
