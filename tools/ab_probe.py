"""A/B of two builds of libspconv on the same GPU:  python tools/ab_probe.py <path to libspconv.so>
Times fprop of a few layers (the stem 3->104 3x3 s2 @8192^2 among them)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mpi4dl_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
L = _lib.lib()
dev = "cuda:0"
sp = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
for (Cc, K, R, S, st, H, W) in [(3, 104, 3, 3, 2, 8192, 8192), (52, 52, 3, 3, 2, 4096, 4096), (416, 416, 1, 1, 1, 1024, 1024),
                                (104, 104, 3, 3, 2, 2048, 2048), (208, 52, 1, 1, 1, 4096, 4096)]:
    x = torch.randn(1, Cc, H, W, device=dev).to(torch.bfloat16)
    w = (torch.randn(K, Cc, R, S, device=dev) / (Cc * R * S) ** 0.5).to(torch.bfloat16)
    y = torch.empty(1, K, H // st, W // st, device=dev, dtype=torch.bfloat16)
    d = _lib.ConvDesc(1, Cc, H, W, K, R, S, st, st, (R - 1) // 2, (S - 1) // 2, _lib.SPC_BF16, _lib.SPC_ALGO_TCGEN05)
    nb = max(L.spc_conv_workspace_bytes(C.byref(d), i) for i in range(3))
    ws = torch.empty(nb + 16, dtype=torch.uint8, device=dev)
    fn = lambda: _lib.check(L.spc_conv2d_fwd(C.byref(d), x.data_ptr(), None, w.data_ptr(), None, y.data_ptr(), ws.data_ptr(), nb, sp()), "f")  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%4d->%-4d %dx%d s%d @%dx%d fprop %.3f ms" % (Cc, K, R, S, st, H, W, e0.elapsed_time(e1) / 10), flush=True)
    del x, y, ws
