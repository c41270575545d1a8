"""ncu target: a few launches of the dominant tcgen05 kernels at BASELINE shapes (fwd, dgrad, wgrad)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mpi4dl_b200 import _lib
dev = "cuda:0"
L = _lib.lib()
sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = [(104, 208, 1, 1, 4096, 4096), (1664, 416, 1, 1, 1024, 1024), (104, 104, 7, 1, 1024, 1024)]
for (Cc, K, R, S, H, W) in shapes:
    x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
    w = torch.randn(K, Cc, R, S, device=dev).to(torch.bfloat16)
    y = torch.empty(1, K, H, W, device=dev, dtype=torch.bfloat16)
    gy = torch.randn_like(y)
    dx = torch.empty_like(x)
    dw = torch.zeros(K, Cc, R, S, device=dev)
    d = _lib.ConvDesc(1, Cc, H, W, K, R, S, 1, 1, (R - 1) // 2, (S - 1) // 2, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
    nb = max(L.spc_conv_workspace_bytes(C.byref(d), i) for i in range(3))
    ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
    for _ in range(2):
        _lib.check(L.spc_conv2d_fwd(C.byref(d), x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(), ws.data_ptr(), nb, sp()), "fwd")
        _lib.check(L.spc_conv2d_dgrad(C.byref(d), gy.data_ptr(), w.data_ptr(), dx.data_ptr(), ws.data_ptr(), nb, sp()), "dgrad")
        _lib.check(L.spc_conv2d_wgrad(C.byref(d), x.data_ptr(), None, gy.data_ptr(), dw.data_ptr(), None, 0, ws.data_ptr(), nb, sp()), "wgrad")
    torch.cuda.synchronize()
# 3x3 stride-1 average pool (TMA-staged kernel), forward and backward, at the AmoebaNet cell shape
x = torch.randn(1, 416, 1024, 1024, device=dev).to(torch.bfloat16)
y = torch.empty_like(x)
dx = torch.empty_like(x)
pd = _lib.PoolDesc(1, 416, 1024, 1024, 3, 1, 1, _lib.SPC_POOL_AVG, _lib.SPC_BF16)
for _ in range(2):
    _lib.check(L.spc_pool2d_fwd(C.byref(pd), x.data_ptr(), None, y.data_ptr(), sp()), "pool fwd")
    _lib.check(L.spc_pool2d_bwd(C.byref(pd), x.data_ptr(), None, y.data_ptr(), dx.data_ptr(), sp()), "pool bwd")
torch.cuda.synchronize()
print("done")
